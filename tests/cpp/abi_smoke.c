/* Plain C (C99) client of the C ABI: proves include/diffeq_b200.h is C-clean and the library links from C.
 * gcc -std=c99 tests/cpp/abi_smoke.c -Ldiffeq_b200 -ldiffeq_b200 -Wl,-rpath,$PWD/diffeq_b200 -o /tmp/abi_smoke */
#include <stdio.h>
#include "../../include/diffeq_b200.h"
int main(void) {
    deq_rhs* rhs = NULL; deq_ensemble* ens = NULL; deq_result* res = NULL;
    double y0[2][3] = {{0.1, 0.0, 0.0}, {1.0, 2.0, 20.0}}, params[3] = {10.0, 28.0, 8.0 / 3.0}, tspan[2] = {0.0, 1.0}, yf[2][3];
    uint32_t acc[2], rej[2];
    deq_options o = deq_options_default();
    int S = 0, adaptive = 0;
    printf("ABI_VERSION %s\n", deq_version());
    if (deq_tableau_get(DEQ_ODE45, DEQ_TABLEAU_REFERENCE, &S, NULL, NULL, NULL, &adaptive, NULL, NULL) != DEQ_OK || S != 7 || !adaptive) return 2;
    if (deq_rhs_builtin(DEQ_SYS_LORENZ, &rhs) != DEQ_OK) return 3;
    if (deq_ensemble_create(rhs, 2, &y0[0][0], params, 0, NULL, 0, &ens) != DEQ_ERR_UNINITIALIZED) return 4;  /* build() without tspan */
    o.reltol = 1e-8; o.abstol = 1e-8;
    if (deq_ensemble_create(rhs, 2, &y0[0][0], params, 0, tspan, 2, &ens) != DEQ_OK) { printf("ABI_NO_GPU %s\n", deq_last_error_message()); deq_rhs_destroy(rhs); return 0; }
    if (deq_solve(ens, DEQ_ODE45, DEQ_TABLEAU_REFERENCE, &o, &res) != DEQ_OK) return 5;
    if (deq_result_final(res, &yf[0][0]) != DEQ_OK || deq_result_counts(res, acc, rej, NULL, NULL, NULL) != DEQ_OK) return 6;
    printf("ABI_OK %u %u %.17g\n", acc[0], acc[1], yf[0][0]);
    deq_result_destroy(res); deq_ensemble_destroy(ens); deq_rhs_destroy(rhs);
    return 0;
}
