"""Full-row colour pass probe (GPU box): times one finest-level sweep with the half-stored colour kernel and with the full-row copy
(relax_full), and the same for the matvec.   python scripts/probe_full.py 256 256 128 6 [reps]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from citcomcu_b200 import inputfile
from citcomcu_b200.problem import CartesianProblem
from citcomcu_b200.stokes import context_from_problem

elx, ely, elz, levels = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (256, 256, 128, 6))]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
prob = CartesianProblem(inputfile.tdepv_box(elx, ely, elz, levels, maxstep=1))
ctx = context_from_problem(prob)
T = prob.initial_temperature()
ctx.set_temperature(T)
ctx.assemble_forces(prob.buoyancy(T), want_host=False)
ctx.get_system_viscosity()
ctl = prob.control
ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
lm = prob.levmax
nno = prob.nno(lm)
rng = np.random.default_rng(1234)
f = ctx.strip_bcs_from_residual(rng.uniform(-1, 1, 3 * nno), lm)
ctx.vec_upload(lm, "RHS", f)


def timeit(fn, reps=reps, warm=3):
    import time
    for _ in range(warm):
        fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


tm = {}
res = {}
for name, opts in (("colour", dict(relax_col=0, matvec_col=0, relax_full=0)), ("full", dict(relax_col=0, matvec_col=0, full_nodes=100000, relax_full=1))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.vec_upload(lm, "VEL", np.zeros(3 * nno))
    tm[f"sweep_ms_{name}"] = round(timeit(lambda: ctx.dev_relax_sweeps(lm, "VEL", "RHS", 1)), 4)
    ctx.vec_upload(lm, "VEL", np.zeros(3 * nno))
    ctx.dev_relax_sweeps(lm, "VEL", "RHS", 4)
    res[name] = ctx.vec_download(lm, "VEL")
tm["full_vs_colour_after_4_sweeps"] = float(np.abs(res["full"] - res["colour"]).max() / np.abs(res["colour"]).max())
print(json.dumps(tm))
