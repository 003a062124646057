"""ncu target: a few tile-kernel smoother sweeps / matvecs on the finest level of the config-3 operator.
    python scripts/prof_tile.py 256 256 128 6 <shape> <hint>"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from citcomcu_b200 import inputfile
from citcomcu_b200.problem import CartesianProblem
from citcomcu_b200.stokes import context_from_problem

elx, ely, elz, levels, shape, hint = [int(v) for v in sys.argv[1:7]]
prob = CartesianProblem(inputfile.tdepv_box(elx, ely, elz, levels, maxstep=1))
ctx = context_from_problem(prob)
T = prob.initial_temperature()
ctx.set_temperature(T)
ctx.assemble_forces(prob.buoyancy(T), want_host=False)
ctx.get_system_viscosity()
ctl = prob.control
ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
lm = prob.levmax
n = prob.nno(lm)
rng = np.random.default_rng(1234)
ctx.vec_upload(lm, "RHS", ctx.strip_bcs_from_residual(rng.uniform(-1, 1, 3 * n), lm))
ctx.vec_upload(lm, "VEL", np.zeros(3 * n))
for k, v in dict(relax_tile=1, matvec_tile=1, tile_shape=shape, tile_hint=hint, tile_nodes=100000).items():
    ctx.set_option(k, v)
ctx.dev_relax_sweeps(lm, "VEL", "RHS", 2)
ctx.dev_matvec(lm, "VEL", "AU", 1)
ctx.synchronize()

if len(sys.argv) > 7:
    ctx.set_option("tile_pad", int(sys.argv[7]))
    ctx.dev_relax_sweeps(lm, "VEL", "RHS", 2)
    ctx.synchronize()
