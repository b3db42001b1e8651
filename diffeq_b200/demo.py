"""The reference's demo entry point on the GPU backend.

`solve_lorenz_attractor(config_json)` mirrors diffeq-example-wasm/src/lib.rs:38-65: a JSON `Config{ode, init, start_time,
end_time, num_times}` (lib.rs:12-19) goes in, the serde form of `OdeSolution` (`{"tout": [...], "yout": [[x,y,z], ...]}`,
src/ode/solution.rs:22-29) comes out, so the demo's web front-end can talk to this backend unchanged.  Errors carry the
reference's messages.
"""
import json

from . import Ode, OdeOptionMap, OdeProblem, Output, Rhs, System, _traits
from .synth import LORENZ_PARAMS


def solution_to_json(tout, yout):
    """serde_json of `OdeSolution<f64, Vec<f64>>`: field order tout, yout (solution.rs:24-29)."""
    return json.dumps({"tout": [float(t) for t in tout], "yout": [[float(v) for v in y] for y in yout]})


def solve_lorenz_attractor(config_json):
    try:
        cfg = json.loads(config_json)
        ode_name, init = cfg["ode"], [float(v) for v in cfg["init"]]
        start, end, num = float(cfg["start_time"]), float(cfg["end_time"]), int(cfg["num_times"])
    except (ValueError, KeyError, TypeError):
        raise ValueError("Failed to serialize solution")            # lib.rs:40-42 (sic)
    ode = Ode.from_str(ode_name)                                     # lib.rs:44, mod.rs:28-46
    if len(init) != 3:
        raise ValueError(f"Can't solve for {len(init)} dimensional type")   # lib.rs:46-51
    problem = (OdeProblem.builder().tspan_linspace(start, end, num).fun(Rhs.builtin(System.Lorenz))
               .params(LORENZ_PARAMS).init(init).build())            # lib.rs:53-58
    adaptive = _traits(ode)[0]                                       # ode23s included; oderosenbrock is grid-stepped
    opts = OdeOptionMap().insert(Output.All if adaptive else Output.Tspan)   # Default::default() options, Points::All
    sol = problem.solve(ode, opts)
    if adaptive:
        tout, yout = sol.trajectory(0)
    else:
        tout, yout = sol.tout, sol.yout[0]
    return solution_to_json(tout, yout)
