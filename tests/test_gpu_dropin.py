"""GPU test of the drop-in boundary: the UNMODIFIED reference time loop (oracle/_ref/ref_harness = main()'s
sequence, Citcom.c:54-175) with `general_stokes_solver` -- and, in the second variant, `PG_timestep` as well --
interposed by dropin/libcitcomcu_dropin.so, against the same run without the preload.  North-star tolerances: temperature after N steps, Nusselt numbers within
0.1 %; velocities within the solver tolerance of the input file."""
from pathlib import Path
import tempfile

import numpy as np
import pytest

from conftest import ROOT, po
from citcomcu_b200 import inputfile

pytestmark = pytest.mark.gpu
DROPIN = ROOT / "dropin" / "libcitcomcu_dropin.so"


# accuracy=1e-4: both arms stop at the solver tolerance, which has to sit well inside the 0.1 % being checked
@pytest.mark.parametrize("name,txt", [
    ("busse", inputfile.busse1a(levels=4, maxstep=6, accuracy=1e-4)),
    ("tdepv", inputfile.tdepv_box(32, 32, 16, 4, maxstep=6, accuracy=1e-4)),
    # extended-Boussinesq: adiabatic + viscous heating and both phase changes (latent heating, phase buoyancy on the host)
    ("eba", inputfile.tdepv_box(16, 16, 8, 3, maxstep=6, accuracy=1e-5, adi_heating=1, visc_heating=1, dissipation_number=0.5,
                                Ra_410=100.0, Ra_670=-100.0)),
], ids=["busse", "tdepv", "eba"])
@pytest.mark.parametrize("energy", [0, 1], ids=["stokes", "stokes+energy"])
def test_reference_time_loop_with_gpu_stokes(name, txt, energy, monkeypatch):
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    nsteps = 5
    monkeypatch.setenv("CCU_DROPIN_ENERGY", str(energy))
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix=f"ccu_ref_{name}_"), nsteps=nsteps)
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix=f"ccu_gpu_{name}_"), nsteps=nsteps, preload=str(DROPIN))
    assert "citcomcu_b200 drop-in: Stokes solve on CUDA device" in err
    assert ("citcomcu_b200 drop-in: energy step on the CUDA device" in err) == bool(energy)
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    for k in range(nsteps + 1):
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        assert np.linalg.norm(Ug - U) < 20 * acc * np.linalg.norm(U), k
        T, Tg = r[f"s{k}_T"], g[f"s{k}_T"]
        assert np.abs(Tg - T).max() < 1e-3 * np.abs(T).max(), k
        sr, sg = r[f"s{k}_scalars"], g[f"s{k}_scalars"]
        assert abs(sg[1] - sr[1]) <= 1e-3 * abs(sr[1]) + 1e-12, ("timestep", k)
        for q in (2, 3):                                   # Nut, Nub
            assert abs(sg[q] - sr[q]) <= 1e-3 * abs(sr[q]) + 1e-9, ("Nu", k, q)
    # Vrms (volume-weighted would need the mesh; nodal rms is the same statistic on both arms)
    vr = np.sqrt(sum((r[f"s{nsteps}_V{d}"].astype(np.float64) ** 2).mean() for d in (1, 2, 3)))
    vg = np.sqrt(sum((g[f"s{nsteps}_V{d}"].astype(np.float64) ** 2).mean() for d in (1, 2, 3)))
    assert abs(vg - vr) < 1e-3 * vr
