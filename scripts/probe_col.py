"""Column-kernel probe (GPU box): builds the config-3 operator on the device for a given mesh, then on its finest level
  (1) checks the column-resident matvec against the colour-pass matvec on the same input (1e-12),
  (2) checks the fixed-point property of the column smoother (F = K x0  ->  a sweep leaves x0 alone),
  (3) times one smoother sweep and one matvec with CUDA events for the colour kernels and every column shape,
      with achieved GB/s against SURVEY.md 8(d)'s algorithmic bytes (600 B/node/sweep without Ad, 552 B/node matvec).

    python scripts/probe_col.py 256 256 128 6 [reps] [lev]
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
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
prob = CartesianProblem(inputfile.tdepv_box(elx, ely, elz, levels, maxstep=1))
ctx = context_from_problem(prob)
T = prob.initial_temperature()
ctx.set_temperature(T)
ctx.assemble_forces(prob.buoyancy(T), want_host=False)
ctx.get_system_viscosity()
ctl = prob.control
ctx.set_option("col_nodes", 100000)
ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
lm = int(sys.argv[6]) if len(sys.argv) > 6 else prob.levmax
nno = prob.nno(lm)
rng = np.random.default_rng(1234)
x0 = ctx.strip_bcs_from_residual(rng.uniform(-1, 1, 3 * nno), lm)


def timeit(fn, reps=reps, warm=3):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


out = {"mesh": [elx, ely, elz], "lev": lm, "nno": nno}
ctx.vec_upload(lm, "VEL", x0)
ctx.set_option("relax_col", 0); ctx.set_option("matvec_col", 0)
ctx.dev_matvec(lm, "VEL", "AU", 1)
Au_ref = ctx.vec_download(lm, "AU")
ctx.dev_matvec(lm, "VEL", "RHS", 0)          # F = K x0 (unstripped rows are what the sweep multiplies)
F0 = ctx.vec_download(lm, "RHS")
for shape in (0, 1, 2):
    ctx.set_option("col_shape", shape); ctx.set_option("relax_col", 1); ctx.set_option("matvec_col", 1)
    ctx.dev_matvec(lm, "VEL", "AU", 1)
    out[f"matvec_col{shape}_vs_colour"] = rel(ctx.vec_download(lm, "AU"), Au_ref)
    ctx.dev_matvec(lm, "VEL", "FL", 0)
    out[f"matvec_col{shape}_unstripped_vs_colour"] = rel(ctx.vec_download(lm, "FL"), F0)
    ctx.vec_upload(lm, "DEL_VEL", x0)
    ctx.dev_relax_sweeps(lm, "DEL_VEL", "RHS", 2)
    out[f"fixed_point_col{shape}"] = rel(ctx.vec_download(lm, "DEL_VEL"), x0)
print(json.dumps(out), flush=True)

f = ctx.strip_bcs_from_residual(rng.uniform(-1, 1, 3 * nno), lm)
ctx.vec_upload(lm, "RHS", f)
tm = {}
xfin = {}
for name, opts in (("colour", dict(relax_col=0, matvec_col=0)), ("col0", dict(relax_col=1, matvec_col=1, col_shape=0, col_wf=1)),
                   ("col0_4l", dict(relax_col=1, matvec_col=1, col_shape=0, col_wf=0)),
                   ("col1", dict(relax_col=1, matvec_col=1, col_shape=1, col_wf=1)), ("col1_4l", dict(relax_col=1, matvec_col=1, col_shape=1, col_wf=0)),
                   ("col2", dict(relax_col=1, matvec_col=1, col_shape=2, col_wf=1)), ("col2_4l", dict(relax_col=1, matvec_col=1, col_shape=2, col_wf=0))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.vec_upload(lm, "VEL", np.zeros(3 * nno))
    ms = timeit(lambda: ctx.dev_relax_sweeps(lm, "VEL", "RHS", 1))
    tm[f"sweep_ms_{name}"] = round(ms, 4)
    tm[f"sweep_GBs_{name}"] = round(600 * nno / ms / 1e6, 1)
    ms = timeit(lambda: ctx.dev_matvec(lm, "VEL", "AU", 1))
    tm[f"matvec_ms_{name}"] = round(ms, 4)
    tm[f"matvec_GBs_{name}"] = round(552 * nno / ms / 1e6, 1)
    # convergence sanity: residual norm after 6 sweeps from zero
    ctx.vec_upload(lm, "VEL", np.zeros(3 * nno))
    ctx.dev_relax_sweeps(lm, "VEL", "RHS", 6)
    xfin[name] = ctx.vec_download(lm, "VEL")
    if name.endswith("_4l"):      # one launch per sweep (progress words) must be bitwise the four-launch result
        tm[f"wf_bitwise_{name[:-3]}"] = bool(np.array_equal(xfin[name], xfin[name[:-3]]))
    ctx.set_option("matvec_col", 0)
    ctx.dev_matvec(lm, "VEL", "AU", 1)
    r = f - ctx.vec_download(lm, "AU")
    tm[f"res6_{name}"] = float(np.linalg.norm(r) / np.linalg.norm(f))
print(json.dumps(tm))
