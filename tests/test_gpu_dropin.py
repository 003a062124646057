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
    # storage_spacing=1: the reference re-evaluates heat_flux / averages on every step (not only at step 0), so the Nu and
    # Vrms comparisons below are between values computed on the developed states of the two runs
    ("busse", inputfile.busse1a(levels=4, maxstep=6, accuracy=1e-4, storage_spacing=1)),
    # VISC_SMOOTH only smooths the nodal output array E->VI in the reference: the run is the plain one, the drop-in accepts it
    ("tdepv", inputfile.tdepv_box(32, 32, 16, 4, maxstep=6, accuracy=1e-4, storage_spacing=1, VISC_SMOOTH="on")),
    # extended-Boussinesq: adiabatic + viscous heating and both phase changes (latent heating, phase buoyancy on the host)
    ("eba", inputfile.tdepv_box(16, 16, 8, 3, maxstep=6, accuracy=1e-5, adi_heating=1, visc_heating=1, dissipation_number=0.5,
                                Ra_410=100.0, Ra_670=-100.0, storage_spacing=1)),
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
    # volume-weighted Vrms from the reference's own layer averages (E->Have.vrms, averages(): evaluated by the reference's host
    # code in both runs, from the CPU and from the device velocities) and the layers themselves
    from citcomcu_b200.stokes import StokesContext
    for k in range(1, nsteps + 1):
        z = r[f"s{k}_XP3"]
        vr, vg = StokesContext.volume_vrms(r[f"s{k}_Have_vrms"], z), StokesContext.volume_vrms(g[f"s{k}_Have_vrms"], z)
        assert abs(vg - vr) < 1e-3 * vr, k
        assert np.abs(g[f"s{k}_Have_vrms"] - r[f"s{k}_Have_vrms"]).max() < 1e-3 * r[f"s{k}_Have_vrms"].max(), k


FUNC_SETS = {
    "operators": "n_assemble_del2_u,assemble_del2_u,assemble_div_u,assemble_grad_p,global_vdot,global_pdot,strip_bcs_from_residual,"
                 "project_vector,interp_vector",
    "smoother": "gauss_seidel",
    "multigrid": "multi_grid",
    "solve_del2_u": "solve_del2_u",
    "uzawa": "solve_Ahat_p_fhat",
}


@pytest.mark.parametrize("which", list(FUNC_SETS))
def test_function_level_bindings(which, monkeypatch):
    """The reference's OWN driver (general_stokes_solver, the Uzawa loop, ...) with only the named inner functions bound to the
    device (dropin/citcom_dropin_funcs.c, reference signatures, the reference's operator arrays uploaded after each
    construct_stiffness_B_matrix): same solution and temperatures as the pure-CPU run within the solver tolerance."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=3, accuracy=1e-5)
    nsteps = 2
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_fref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_STOKES", "0")
    monkeypatch.setenv("CCU_DROPIN_ENERGY", "0")
    monkeypatch.setenv("CCU_DROPIN_FUNCS", FUNC_SETS[which])
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix=f"ccu_f{which}_"), nsteps=nsteps, preload=str(DROPIN))
    assert "Stokes solve on CUDA device" in err            # the context was created by a bound function
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    # same algorithm with another summation order (operators: the iterates agree far inside the solver tolerance), or another
    # smoother order / device control flow (agreement at the solver tolerance)
    tol = 1e-6 if which == "operators" else 20 * acc
    for k in range(nsteps + 1):
        U, Ug, P, Pg = r[f"s{k}_U"], g[f"s{k}_U"], r[f"s{k}_P"], g[f"s{k}_P"]
        assert np.linalg.norm(Ug - U) <= tol * np.linalg.norm(U), (which, k)
        assert np.linalg.norm(Pg - P) <= 10 * tol * np.linalg.norm(P), (which, k)
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3


def test_marker_bindings_in_the_reference_time_loop(monkeypatch):
    """PG_timestep_particle (with its on_off toggle), Euler and Runge_Kutta bound to the device inside the reference's own
    time loop (thermochemical run, host Stokes solve): marker positions, element assignment and the nodal composition follow
    the pure-CPU run."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=3, accuracy=1e-5, composition=1, rayleigh_comp=1e6, markers_per_ele=8, comp_depth=0.605)
    nsteps = 2
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_mref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_STOKES", "0")
    monkeypatch.setenv("CCU_DROPIN_FUNCS", "PG_timestep_particle,Euler,Runge_Kutta")
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_mgpu_"), nsteps=nsteps, preload=str(DROPIN))
    r, g = ref[0], gpu[0]
    for k in range(1, nsteps + 1):
        assert int(g[f"s{k}_nmarkers"][0]) == int(r[f"s{k}_nmarkers"][0])
        for d in (1, 2, 3):
            # the temperature step feeding the buoyancy differs in the last bits (device predictor-corrector), so do the velocities
            assert np.abs(g[f"s{k}_XMC{d}"] - r[f"s{k}_XMC{d}"]).max() < 1e-5, (k, d)
        assert (g[f"s{k}_CElement"] != r[f"s{k}_CElement"]).mean() < 1e-3
        assert np.abs(g[f"s{k}_C"] - r[f"s{k}_C"]).max() < 0.2 and np.abs(g[f"s{k}_C"] - r[f"s{k}_C"]).mean() < 1e-4
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("nproc,geometry", [((2, 1, 1), "cart3d"), ((2, 2, 2), "cart3d"), ((2, 1, 1), "Rsphere"), ((1, 1, 2), "Rsphere")],
                         ids=["2x1x1", "2x2x2", "2x1x1-rsphere", "1x1x2-rsphere"])
def test_multi_rank_reference_with_gpu_stokes(nproc, geometry, monkeypatch):
    """The unmodified reference on several MPI ranks (oracle/mpi_shim), one GPU per rank: every rank's general_stokes_solver and
    PG_timestep run on its device, halo sums and reductions over NCCL (the id travels through the reference's own MPI_Bcast)."""
    world = nproc[0] * nproc[1] * nproc[2]
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    if geometry == "Rsphere":       # BASELINE config 4's regional block split over two ranks (the shipped input1 runs 2x2x1)
        txt = inputfile.input1_rsphere(levels=3, maxstep=3, accuracy=1e-5, nproc=nproc, TDEPV="on", VISC_UPDATE="on", update_every_steps=1,
                                       perturbmag=0.05)
    else:
        txt = inputfile.tdepv_box(16, 16, 8, 3, nproc=nproc, maxstep=3, accuracy=1e-5)
    nsteps = 2
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_nref_"), nsteps=nsteps, nproc=world)
    monkeypatch.setenv("CCU_DROPIN_ENERGY", "1")
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_ngpu_"), nsteps=nsteps, nproc=world, preload=str(DROPIN))
    assert "citcomcu_b200 drop-in: Stokes solve on CUDA device" in err
    for r, g in zip(ref, gpu):
        acc = r.control()["accuracy"]
        for k in range(nsteps + 1):
            U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
            assert np.linalg.norm(Ug - U) < 20 * acc * max(np.linalg.norm(U), 1e-30), k
            assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k


def test_busse_case2_shipped_example(monkeypatch):
    """examples/Busse1993/case2.input as shipped (rheol=11, temperature-dependent viscosity, non-uniform z, restart from the
    shipped case2-tic temperature file; tests/golden holds copies of the two data files): the reference's time loop with the
    Stokes solve and the energy step on the device against the pure-CPU run."""
    import gzip
    import shutil
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    txt = (ROOT / "tests" / "golden" / "busse_case2.input").read_text()
    nsteps = 2
    runs = []
    for tag, preload in (("ref", None), ("gpu", str(DROPIN))):
        wd = Path(tempfile.mkdtemp(prefix=f"ccu_case2_{tag}_"))
        with gzip.open(ROOT / "tests" / "golden" / "busse_case2-tic.temp.0.0.gz", "rb") as src, open(wd / "case2-tic.temp.0.0", "wb") as dst:
            shutil.copyfileobj(src, dst)
        monkeypatch.setenv("CCU_DROPIN_ENERGY", "1")
        runs.append(po.run_harness(txt, wd, nsteps=nsteps, preload=preload, timeout=1200))
    (ref, _), (gpu, err) = runs
    assert "citcomcu_b200 drop-in: Stokes solve on CUDA device" in err
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    for k in range(nsteps + 1):
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        assert np.linalg.norm(Ug - U) < 20 * acc * np.linalg.norm(U), k
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k
        assert np.allclose(g[f"s{k}_EVI"], r[f"s{k}_EVI"], rtol=1e-3), k      # viscosity contrast of the law: exp(+-8)


@pytest.mark.parametrize("rheology,damp,geometry", [(1, 1.0, "cart3d"), (2, 1.0, "cart3d"), (1, 0.7, "cart3d"), (1, 1.0, "Rsphere")])
def test_stress_dependent_viscosity_loop(rheology, damp, geometry, monkeypatch):
    """SDEPV: visc_from_S (power law, sdepv_rheology 1 / 2) and the viscosity <-> velocity iteration of general_stokes_solver
    (Drive_solvers.c:120-159, with and without damping) on the device, inside the reference's own time loop."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    sd = dict(maxstep=3, accuracy=1e-6, viscE="4.6,4.6,4.6,4.6", SDEPV="on", sdepv_rheology=rheology, sdepv_expt="3,3,3,3",
              sdepv_trns="2e3,2e3,2e3,2e3", sdepv_misfit=1e-3, sdepv_iter_damp=damp, sdepv_trns_T=3000, sdepv_trns_c=2.0, storage_spacing=1)
    if geometry == "Rsphere":      # the strain rate of visc_from_S through the Rsphere branch of strain_rate_2_inv
        txt = inputfile.input1_rsphere(levels=3, TDEPV="on", VISC_UPDATE="on", update_every_steps=1, perturbmag=0.05, **sd)
    else:
        txt = inputfile.tdepv_box(16, 16, 8, 3, **sd)
    nsteps = 2
    ref, rerr = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_sref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_ENERGY", "1")
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_sgpu_"), nsteps=nsteps, preload=str(DROPIN))
    r, g = ref[0], gpu[0]
    for k in range(nsteps + 1):
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        # both arms stop the outer iteration at a relative velocity change of sdepv_misfit = 1e-3
        assert np.linalg.norm(Ug - U) < 5e-3 * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        assert np.allclose(g[f"s{k}_EVI"], r[f"s{k}_EVI"], rtol=2e-2), k
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k
    # the viscosity did become strain-rate dependent (not the Newtonian field)
    if geometry == "Rsphere":
        ntxt = inputfile.input1_rsphere(levels=3, maxstep=1, accuracy=1e-6, viscE="4.6,4.6,4.6,4.6", TDEPV="on", perturbmag=0.05)
    else:
        ntxt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=1, accuracy=1e-6, viscE="4.6,4.6,4.6,4.6")
    newt = po.run_harness(ntxt, tempfile.mkdtemp(prefix="ccu_snewt_"), nsteps=0)[0][0]
    assert np.abs(r["s0_EVI"] / newt["s0_EVI"] - 1).max() > 0.05


@pytest.mark.parametrize("energy,geometry", [(0, "cart3d"), (1, "cart3d"), (1, "Rsphere")], ids=["stokes", "stokes+energy", "rsphere"])
def test_imposed_plate_velocity(energy, geometry, monkeypatch):
    """topvbc=1 with a non-zero plate velocity: E->VB enters U (velocities_conform_bcs) and F (the K.VB term of get_elt_f,
    Element_calculations.c:1038-1063, evaluated with the viscosity of the previous update) on the device."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    # accuracy: the hot, weak bottom layer under a driven lid is poorly conditioned -- the two arms (and the reference against itself
    # at a tighter tolerance) differ there by about 1000 x accuracy (measured: 5.5e-3, 4e-4, 3e-5 of |U| at 1e-5, 1e-6, 1e-7), so the
    # comparison runs at 1e-7 and allows 3e-4
    vb = dict(maxstep=4, accuracy=1e-7, topvbc=1, plate_velocity=40.0, topvbyval=-15.0, storage_spacing=1)
    mk = (lambda **kw: inputfile.input1_rsphere(levels=3, TDEPV="on", VISC_UPDATE="on", update_every_steps=1, perturbmag=0.05, **kw)) \
        if geometry == "Rsphere" else (lambda **kw: inputfile.tdepv_box(16, 16, 8, 3, **kw))
    txt = mk(**vb)
    nsteps = 2
    monkeypatch.setenv("CCU_DROPIN_ENERGY", str(energy))
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_vbref_"), nsteps=nsteps)
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_vbgpu_"), nsteps=nsteps, preload=str(DROPIN))
    assert "citcomcu_b200 drop-in: Stokes solve on CUDA device" in err
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    free = po.run_harness(mk(maxstep=1, accuracy=1e-5), tempfile.mkdtemp(prefix="ccu_vbfree_"), nsteps=0)[0][0]
    assert np.abs(r["VB1"]).max() == 40.0 and np.abs(r["VB2"]).max() == 15.0
    # the plate drives the flow: nothing like the free-slip solution of the same state
    assert np.linalg.norm(r["s0_U"] - free["s0_U"]) > 0.5 * np.linalg.norm(r["s0_U"])
    for k in range(nsteps + 1):
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        assert np.linalg.norm(Ug - U) < 3e-4 * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k
        sr, sg = r[f"s{k}_scalars"], g[f"s{k}_scalars"]
        assert abs(sg[1] - sr[1]) <= 1e-3 * abs(sr[1]) + 1e-12, ("timestep", k)
    assert acc == 1e-7


@pytest.mark.parametrize("funcs", ["solve_Ahat_p_fhat", "n_assemble_del2_u,assemble_div_u,assemble_grad_p,gauss_seidel,global_vdot,global_pdot"])
def test_regional_sphere_solver_on_device(funcs, monkeypatch):
    """BASELINE config 4 geometry (examples/input1's regional-spherical block): the operator is assembled by the reference's host code
    (the device has no Rsphere element routines yet), the SOLVER -- the Uzawa iteration with its multigrid velocity solves, or its
    operator / smoother functions one by one -- runs on the device on the uploaded node-stored operator: same U, P as the pure-CPU run."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    txt = inputfile.input1_rsphere(levels=3, maxstep=3, accuracy=1e-5)
    nsteps = 2
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rsref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_STOKES", "0")
    monkeypatch.setenv("CCU_DROPIN_ENERGY", "0")
    monkeypatch.setenv("CCU_DROPIN_FUNCS", funcs)
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rsgpu_"), nsteps=nsteps, preload=str(DROPIN))
    assert "Rsphere geometry: solver functions on CUDA device" in err
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    for k in range(nsteps + 1):
        U, Ug, P, Pg = r[f"s{k}_U"], g[f"s{k}_U"], r[f"s{k}_P"], g[f"s{k}_P"]
        assert np.linalg.norm(Ug - U) <= 20 * acc * np.linalg.norm(U), k
        assert np.linalg.norm(Pg - P) <= 200 * acc * np.linalg.norm(P), k
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k


@pytest.mark.parametrize("energy", [0, 1], ids=["stokes", "stokes+energy"])
@pytest.mark.parametrize("tdepv", ["off", "on", "on-eba", "on-rheol2"])
def test_regional_sphere_stokes_assembled_and_solved_on_device(tdepv, energy, monkeypatch):
    """BASELINE config 4 geometry through the whole-step bindings: operator assembly (Rsphere get_elt_k / get_elt_g / get_elt_f) and the
    Stokes solve on the device; in the second variant also the SUPG energy step with the Rsphere branches of pg_shape_fn /
    element_residual (Advection_diffusion.c:506-528, 620-640, 668-674)."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    # "on-eba": BASELINE config 4 proper, the extended-Boussinesq terms of the shipped input1 (adiabatic + viscous heating, computed by
    # the reference's host process_heating and uploaded before each device energy step)
    eba = dict(adi_heating=1, visc_heating=1, surf_temp=0.078947) if tdepv == "on-eba" else {}
    if tdepv == "on-rheol2":        # a depth-dependent law: the depth coordinate is r (E->SX[3]) on the sphere
        eba = dict(rheol=2, viscE="3.0,3.0,3.0,3.0", viscT="0.5,0.5,0.5,0.5", viscZ="1.0,1.0,1.0,1.0")
    txt = inputfile.input1_rsphere(levels=3, maxstep=5, accuracy=1e-5, TDEPV=tdepv.split("-")[0], VISC_UPDATE="on", update_every_steps=1,
                                   storage_spacing=1, perturbmag=0.05, **eba)
    nsteps = 4
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rswref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_ENERGY", str(energy))
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rswgpu_"), nsteps=nsteps, preload=str(DROPIN))
    assert "Stokes solve on CUDA device" in err and "regional-spherical element routines" in err
    assert ("energy step on the CUDA device" in err) == bool(energy)
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    for k in range(nsteps + 1):
        U, Ug, P, Pg = r[f"s{k}_U"], g[f"s{k}_U"], r[f"s{k}_P"], g[f"s{k}_P"]
        assert np.linalg.norm(Ug - U) <= 20 * acc * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        assert np.linalg.norm(Pg - P) <= 200 * acc * np.linalg.norm(P), k
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k
        sr, sg = r[f"s{k}_scalars"], g[f"s{k}_scalars"]
        assert abs(sg[1] - sr[1]) <= 1e-3 * abs(sr[1]) + 1e-12, ("timestep", k)
        for q in (2, 3):                                   # Nut, Nub (the reference's host heat_flux on both sides)
            assert abs(sg[q] - sr[q]) <= 1e-3 * abs(sr[q]) + 1e-9, ("Nu", k, q)
    # the temperature did move over the run (the comparison above is not between two copies of the initial field)
    assert np.abs(r[f"s{nsteps}_T"] - r["s0_T"]).max() > 1e-3


def test_regional_sphere_thermochemical_loop_on_device(monkeypatch):
    """Regional-spherical thermochemical run: Stokes assembly + solve, the energy step and the marker advection (Euler / Runge_Kutta
    with the 1/r, 1/(r sin theta) position update of Composition_adv.c:79-92, 124-135) on the device inside the reference's loop."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    txt = inputfile.input1_rsphere(levels=3, maxstep=5, accuracy=1e-5, composition=1, rayleigh_comp=1e6, markers_per_ele=8, comp_depth=0.3,
                                   perturbmag=0.05)
    nsteps = 4
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rsmref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_FUNCS", "PG_timestep_particle,Euler,Runge_Kutta")
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rsmgpu_"), nsteps=nsteps, preload=str(DROPIN))
    assert "regional-spherical element routines" in err
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    assert r[f"s{nsteps}_C"].max() > 0.9 and r[f"s{nsteps}_C"].min() < 0.1          # a two-layer composition
    assert np.abs(r[f"s{nsteps}_XMC1"] - r["s1_XMC1"]).max() > 1e-3                 # markers did move in colatitude
    for k in range(1, nsteps + 1):
        assert int(g[f"s{k}_nmarkers"][0]) == int(r[f"s{k}_nmarkers"][0])
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        # both arms stop each solve at the solver tolerance; over coupled thermochemical steps the differences add up (2.8e-4 at step 4)
        assert np.linalg.norm(Ug - U) <= 50 * acc * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        for d in (1, 2, 3):
            assert np.abs(g[f"s{k}_XMC{d}"] - r[f"s{k}_XMC{d}"]).max() < 1e-5, (k, d)
        assert (g[f"s{k}_CElement"] != r[f"s{k}_CElement"]).mean() < 1e-3
        assert np.abs(g[f"s{k}_C"] - r[f"s{k}_C"]).max() < 0.2 and np.abs(g[f"s{k}_C"] - r[f"s{k}_C"]).mean() < 1e-4
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3


@pytest.mark.parametrize("name,txt", [
    ("config1_busse1a", inputfile.busse1a(levels=5, maxstep=4, accuracy=1e-4, storage_spacing=1)),           # examples/Busse1993 case 1a: 32x16x32
    ("config2_input1_cart", inputfile.input1_cart(levels=4, maxstep=4, storage_spacing=1)),                  # examples/input1 mesh as Cartesian: 48^3
    ("config4_input1_rsphere", inputfile.input1_rsphere(levels=4, maxstep=4, storage_spacing=1)),            # examples/input1: 48^3 regional sphere
], ids=["config1_busse1a", "config2_input1_cart", "config4_input1_rsphere"])
def test_baseline_configs_at_shipped_sizes(name, txt, monkeypatch):
    """The BASELINE.json parity configurations at the mesh sizes and level counts of the shipped input files (not scaled down), Stokes
    and energy step on the device inside the reference's own time loop."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    nsteps = 3
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix=f"ccu_full_ref_{name}_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_ENERGY", "1")
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix=f"ccu_full_gpu_{name}_"), nsteps=nsteps, preload=str(DROPIN))
    assert "Stokes solve on CUDA device" in err and "energy step on the CUDA device" in err
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    for k in range(nsteps + 1):
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        assert np.linalg.norm(Ug - U) < 20 * acc * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3 * np.abs(r[f"s{k}_T"]).max(), k
        sr, sg = r[f"s{k}_scalars"], g[f"s{k}_scalars"]
        assert abs(sg[1] - sr[1]) <= 1e-3 * abs(sr[1]) + 1e-12, ("timestep", k)
        for q in (2, 3):
            assert abs(sg[q] - sr[q]) <= 1e-3 * abs(sr[q]) + 1e-9, ("Nu", k, q)


@pytest.mark.parametrize("opts", [dict(pre_comp="1.0,50.0"), dict(pre_comp="1.0,0.05", cdepv_absolute="on"),
                                  dict(pre_comp="1.0,20.0,1.0,30.0,1.0,40.0,1.0,50.0", layer_pre_comp="on")],
                         ids=["prefactor", "absolute", "per_layer"])
def test_composition_dependent_viscosity(opts, monkeypatch):
    """CDEPV: visc_from_C (Viscosity_structures.c:1784-1935, prefactor mode / cdepv_absolute / per-layer factors) on the device in a
    thermochemical run through the drop-in; the composition reaches the device as the host's nodal field E->C."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    base = dict(maxstep=4, accuracy=1e-5, composition=1, rayleigh_comp=1e6, markers_per_ele=8, comp_depth=0.605, storage_spacing=1)
    txt = inputfile.tdepv_box(16, 16, 8, 3, CDEPV="on", **opts, **base)
    nsteps = 2
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_cdref_"), nsteps=nsteps)
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_cdgpu_"), nsteps=nsteps, preload=str(DROPIN))
    assert "Stokes solve on CUDA device" in err
    plain = po.run_harness(inputfile.tdepv_box(16, 16, 8, 3, **base), tempfile.mkdtemp(prefix="ccu_cd0_"), nsteps=0)[0][0]
    r, g = ref[0], gpu[0]
    acc = r.control()["accuracy"]
    ratio = r["s0_EVI"] / plain["s0_EVI"]
    assert max(ratio.max(), 1.0 / ratio.min()) > 10.0                       # the composition did change the viscosity
    for k in range(nsteps + 1):
        assert np.allclose(g[f"s{k}_EVI"], r[f"s{k}_EVI"], rtol=1e-5 if k == 0 else 1e-2), k
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        assert np.linalg.norm(Ug - U) <= 50 * acc * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k


@pytest.mark.parametrize("trans", ["on", "off"])
def test_plastic_yielding_loop(trans, monkeypatch):
    """BDEPV: visc_from_B (Viscosity_structures.c:1470-1755, regular branch: yield stress a depth + b, harmonic or minimum combination with the
    viscosity so far) and the viscosity <-> velocity iteration it shares with SDEPV, on the device inside the reference's own time loop."""
    if not po.have_ref() or not DROPIN.exists():
        pytest.skip("needs the prebuilt reference (oracle/_ref) and dropin/libcitcomcu_dropin.so")
    four = lambda v: ",".join([v] * 4)      # noqa: E731
    base = dict(maxstep=3, accuracy=1e-6, viscE=four("4.6"), storage_spacing=1, sdepv_misfit=1e-3)
    txt = inputfile.tdepv_box(16, 16, 8, 3, BDEPV="on", plasticity_dimensional="off", abyerlee=four("1e5"), bbyerlee=four("2e4"), lbyerlee=four("1e20"),
                              plasticity_trans=trans, plasticity_viscosity_offset=1e-2, **base)
    nsteps = 2
    ref, _ = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_bdref_"), nsteps=nsteps)
    monkeypatch.setenv("CCU_DROPIN_ENERGY", "1")
    gpu, err = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_bdgpu_"), nsteps=nsteps, preload=str(DROPIN))
    assert "Stokes solve on CUDA device" in err
    r, g = ref[0], gpu[0]
    for k in range(nsteps + 1):
        U, Ug = r[f"s{k}_U"], g[f"s{k}_U"]
        # both arms stop the outer iteration at a relative velocity change of sdepv_misfit = 1e-3
        assert np.linalg.norm(Ug - U) < 5e-3 * np.linalg.norm(U), (k, np.linalg.norm(Ug - U) / np.linalg.norm(U))
        assert np.allclose(g[f"s{k}_EVI"], r[f"s{k}_EVI"], rtol=2e-2), k
        assert np.abs(g[f"s{k}_T"] - r[f"s{k}_T"]).max() < 1e-3, k
    plain = po.run_harness(inputfile.tdepv_box(16, 16, 8, 3, **base), tempfile.mkdtemp(prefix="ccu_bd0_"), nsteps=0)[0][0]
    assert (r["s1_EVI"] / plain["s0_EVI"]).min() < 0.2                  # the material did yield (from the second solve on: the first call sees unit strain rate)
