"""Kernel timing probe (GPU box): operators from the reference's own setup on a Cartesian box,
CUDA-event timings of the hot kernels, achieved GB/s against the algorithmic bytes of SURVEY.md 8d."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po
from citcomcu_b200 import inputfile
from citcomcu_b200.stokes import context_from_dump

elx, ely, elz, levels = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 64, 32, 4))]
t0 = time.time()
dumps, err = po.run_harness(inputfile.tdepv_box(elx, ely, elz, levels, maxstep=1), tempfile.mkdtemp(), nsteps=0, setup_only=True, timeout=3000)
d = dumps[0]
print(f"reference setup {time.time() - t0:.1f}s", flush=True)
ctx = context_from_dump(d)
lm = d.levmax
nno, neq = d.dims(lm)["nno"], d.dims(lm)["neq"]
rng = np.random.default_rng(1234)
f = rng.uniform(-1, 1, neq)
ctx.vec_upload(lm, "RHS", ctx.strip_bcs_from_residual(f, lm))
ctx.vec_upload(lm, "VEL", np.zeros(neq))


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"mesh": [elx, ely, elz], "nno": nno}
ms = timeit(lambda: ctx.dev_matvec(lm, "VEL", "AU", 1))
out["matvec_ms"] = ms
out["matvec_GBs"] = 552 * nno / ms / 1e6
ms = timeit(lambda: ctx.dev_relax_sweeps(lm, "VEL", "RHS", 1))
out["gs_sweep_ms"] = ms
out["gs_sweep_GBs"] = 648 * nno / ms / 1e6
ctx.vec_upload(lm, "T0", ctx.strip_bcs_from_residual(f, lm))
ms = timeit(lambda: ctx.dev_multi_grid("T1", "T0"), reps=3, warm=1)
out["mg_cycle_ms"] = ms
n, npno = neq, d.dims(lm)["npno"]
t0 = time.time()
V, P, steps, res, hist = ctx.solve_Ahat_p_fhat(np.zeros(n), np.zeros(npno), d["s0_F"], d.control()["accuracy"], 375)
out["stokes_s"] = time.time() - t0
out["stokes_iters"] = steps
out["launches"] = ctx.launch_count
print(json.dumps(out))
