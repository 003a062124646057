"""Regional-spherical block driven from Python alone (SphericalProblem + StokesContext): setup, shell-averaged buoyancy and the
Stokes solve against the unmodified reference's step 0."""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import po
from citcomcu_b200 import inputfile
from citcomcu_b200.problem import SphericalProblem
from citcomcu_b200.stokes import context_from_problem

txt = inputfile.input1_rsphere(levels=3, maxstep=1, accuracy=1e-6, TDEPV="on", perturbmag=0.05)
d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rspy_"), nsteps=0, kat=True)[0][0]
prob = SphericalProblem(txt)
ctx = context_from_problem(prob)
T = prob.initial_temperature()
ctx.set_temperature(T)
adv = d["kat_adv_params"]
ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
b = ctx.thermal_buoyancy(float(adv[5]))
ctl = prob.control
U, P, its, res = ctx.general_stokes_solver(T, b, rebuild=1, augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"],
                                           precondition=ctl["precondition"], guess=0)
print("buoyancy", float(np.abs(b - d["s0_buoyancy"]).max() / np.abs(d["s0_buoyancy"]).max()),
      "U", float(np.linalg.norm(U - d["s0_U"]) / np.linalg.norm(d["s0_U"])), "its", its)
ctx.close()
