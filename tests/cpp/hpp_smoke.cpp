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
        // reference-shaped per-problem solutions: the full Points::All stream of dopri5, RK4 on a grid, ode23s, oderosenbrock
        auto one = deq::OdeProblem::builder().fun(rhs).params({10.0, 28.0, 8.0 / 3.0}).init({{0.1, 0.0, 0.0}, {1.0, 2.0, 20.0}}).tspan({0.0, 0.5, 1.0}).build();
        auto a = one.solve_trajectories(deq::Ode::Ode45, deq::OdeOptionMap().reltol(1e-8).abstol(1e-8));
        auto f = one.solve_trajectories(deq::Ode::Ode4);
        auto s = one.solve_trajectories(deq::Ode::Ode23s);
        auto k = one.solve_trajectories(deq::Ode::Ode4ss);
        printf("HPP_TRAJ %zu %zu %u %.17g | %zu %.17g | %zu %u %.17g | %zu %zu\n", a[0].tout.size(), a[0].yout.size(), a[0].accepted[0], a[0].yout.back()[0],
               f[1].yout.size(), f[1].yout.back()[2], s[0].tout.size(), s[0].accepted[0], s[0].yout.back()[0], k[0].tout.size(), k[0].yout.size());
        try { deq::OdeProblem::builder().fun(rhs).tspan({0.0, 1.0}).build(); } catch (const deq::OdeError& e) { printf("HPP_ERR %d\n", (int)e.status); }
    } catch (const deq::OdeError& e) { printf("HPP_FAIL %d %s\n", (int)e.status, e.what()); return 1; }
    return 0;
}
