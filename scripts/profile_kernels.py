"""Kernel probe for ncu / CUDA-event timing on the GPU box: builds the config-3 operator on the device for a
given mesh, then runs the finest-level smoother sweep and matvec in their kernel variants.

    python scripts/profile_kernels.py 256 256 128 6 [reps]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from citcomcu_b200 import inputfile
from citcomcu_b200.problem import CartesianProblem
from citcomcu_b200.stokes import context_from_problem

elx, ely, elz, levels = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (256, 256, 128, 6))]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
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
ctx.vec_upload(lm, "VEL", np.zeros(3 * nno))


def timeit(fn, reps=reps, warm=2):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


out = {"mesh": [elx, ely, elz], "nno": nno}
for name, opts in (("tab2", dict(matvec_tab=2, relax_tab=2)), ("base", dict(matvec_tab=0, relax_tab=0)),
                   ("tab4", dict(matvec_tab=4, relax_tab=4)), ("tab7", dict(matvec_tab=4, relax_tab=7))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ms = timeit(lambda: ctx.dev_relax_sweeps(lm, "VEL", "RHS", 1))
    out[f"gs_sweep_ms_{name}"] = round(ms, 4)
    out[f"gs_sweep_GBs_{name}"] = round(648 * nno / ms / 1e6, 1)
    ms = timeit(lambda: ctx.dev_matvec(lm, "VEL", "AU", 1))
    out[f"matvec_ms_{name}"] = round(ms, 4)
    out[f"matvec_GBs_{name}"] = round(552 * nno / ms / 1e6, 1)
print(json.dumps(out))
if reps == 1:
    sys.exit(0)

# ---- per-level smoother / matvec timings by kernel variant (threshold tuning)
variants = {"smem_or_cta": dict(small_nodes=10**9), "warp": dict(small_nodes=0, warp_nodes=10**9), "quad": dict(small_nodes=0, warp_nodes=0, quad_nodes=10**9),
            "tab2": dict(small_nodes=0, warp_nodes=0, quad_nodes=0, relax_tab=2, matvec_tab=4),
            "unrolled": dict(small_nodes=0, warp_nodes=0, quad_nodes=0, relax_tab=0, matvec_tab=0)}
lv = {}
for lev in range(prob.levmin, prob.levmax + 1):
    n = prob.nno(lev)
    row = {"nno": n}
    for name, opts in variants.items():
        if name == "smem_or_cta" and n > 3000:
            continue
        if name == "warp" and n > 300000:
            continue
        for k, v in opts.items():
            ctx.set_option(k, v)
        row[f"sweep3_us_{name}"] = round(timeit(lambda: ctx.dev_relax_sweeps(lev, "VEL", "RHS", 3), reps=10) * 1e3, 1)
        row[f"matvec_us_{name}"] = round(timeit(lambda: ctx.dev_matvec(lev, "VEL", "AU", 1), reps=10) * 1e3, 1)
    lv[lev] = row
print(json.dumps({"levels": lv}))
