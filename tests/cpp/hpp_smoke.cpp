// Compile-and-run check of include/diffeq_b200.hpp (C++ host mirror): Lorenz dopri5 on 1000 trajectories.
#include <cstdio>
#include "../../include/diffeq_b200.hpp"
int main() {
    try {
        auto rhs = deq::Rhs::builtin(DEQ_SYS_LORENZ);
        std::vector<std::vector<double>> y0(1000, {0.1, 0.0, 0.0});
        for (size_t i = 0; i < y0.size(); ++i) y0[i][0] += 1e-3 * i;
        auto prob = deq::OdeProblem::builder().fun(rhs).params({10.0, 28.0, 8.0 / 3.0}).init(y0).tspan({0.0, 1.0}).build();
        auto sol = prob.ode45(deq::OdeOptionMap().reltol(1e-8).abstol(1e-8));
        printf("HPP_OK %zu %u %.17g\n", sol.yout.size(), sol.accepted[0], sol.yout[0][0]);
        try { deq::OdeProblem::builder().fun(rhs).tspan({0.0, 1.0}).build(); } catch (const deq::OdeError& e) { printf("HPP_ERR %d\n", (int)e.status); }
    } catch (const deq::OdeError& e) { printf("HPP_FAIL %d %s\n", (int)e.status, e.what()); return 1; }
    return 0;
}
