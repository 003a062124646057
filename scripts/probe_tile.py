"""Tile-resident kernels vs colour-pass kernels (GPU box): CUDA-event timings of one smoother sweep and one matvec on the
two finest levels of the config-3 operator, for every tile shape with and without the L2 eviction hints.

    python scripts/probe_tile.py 256 256 128 6 [reps]
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
rng = np.random.default_rng(1234)
for lev in (lm, lm - 1):
    n = prob.nno(lev)
    ctx.vec_upload(lev, "RHS", ctx.strip_bcs_from_residual(rng.uniform(-1, 1, 3 * n), lev))
    ctx.vec_upload(lev, "VEL", np.zeros(3 * n))


def timeit(fn, reps=reps, warm=2):
    import time
    for _ in range(warm):
        fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


variants = [("colour", dict(relax_tile=0, matvec_tile=0))]
for shape in (0, 1, 2):
    for hint in (1, 0):
        variants.append((f"tile{shape}_hint{hint}", dict(relax_tile=1, matvec_tile=1, tile_shape=shape, tile_hint=hint, tile_nodes=100000)))
out = {"mesh": [elx, ely, elz]}
for lev in (lm, lm - 1):
    nno = prob.nno(lev)
    row = {"nno": nno}
    for name, opts in variants:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ms = timeit(lambda: ctx.dev_relax_sweeps(lev, "VEL", "RHS", 3)) / 3
        mv = timeit(lambda: ctx.dev_matvec(lev, "VEL", "AU", 1))
        row[name] = dict(sweep_ms=round(ms, 4), sweep_GBs=round(648 * nno / ms / 1e6, 1), matvec_ms=round(mv, 4), matvec_GBs=round(552 * nno / mv / 1e6, 1))
    out[f"level{lev}"] = row
print(json.dumps(out))
