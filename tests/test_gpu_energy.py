"""GPU parity of the energy step against the unmodified reference:
std_timestep, pg_solver, PG_timestep, thermal_buoyancy on the reference's own state after its step-0 Stokes solve,
then the coupled timestep loop (energy -> buoyancy -> Stokes) against the reference's T after N steps (north star:
temperature within 0.1 %).  The kernels restate the reference's operand types (ccu_build_exact.cu, -fmad=false), so
the single-kernel checks are bit-exact or within 1 ulp(fp32)."""
import tempfile
from pathlib import Path

import numpy as np
import pytest

from conftest import po
from citcomcu_b200 import inputfile

pytestmark = pytest.mark.gpu


def ulps(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) / np.finfo(np.float32).eps


@pytest.fixture(scope="module", params=["tdepv", "busse"])
def state(request):
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    text = inputfile.tdepv_box(16, 16, 8, 3, maxstep=4) if request.param == "tdepv" else inputfile.busse1a(levels=3, maxstep=4)
    dumps, err = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_energy_")), nsteps=3, kat=True)
    d = dumps[0]
    prob = CartesianProblem(text)
    ctx = context_from_problem(prob)
    adv = d["kat_adv_params"]
    ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
    yield d, prob, ctx, float(adv[5])
    ctx.close()


def load_s0(d, ctx):
    ctx.set_temperature(d["s0_T"])
    ctx.set_tdot(d["s0_Tdot"])
    ctx.set_velocity(d["s0_V1"], d["s0_V2"], d["s0_V3"])


def test_v_from_vector(state):
    d, prob, ctx, _ = state
    ctx.vec_upload(prob.levmax, "U", d["s0_U"])
    V = ctx.v_from_vector()
    for a in range(3):
        assert np.array_equal(V[a], d[f"s0_V{a + 1}"])


def test_std_timestep_bit_exact(state):
    d, prob, ctx, _ = state
    load_s0(d, ctx)
    assert ctx.std_timestep() == np.float32(d["kat_dt"][0])


def test_pg_solver(state):
    d, prob, ctx, _ = state
    load_s0(d, ctx)
    assert ulps(ctx.pg_solver(), d["kat_pg_DTdot"]) <= 2.0


def test_PG_timestep_and_buoyancy(state):
    d, prob, ctx, Atemp = state
    load_s0(d, ctx)
    T, Tdot, dt, Tint = ctx.PG_timestep(d["s0_T"], d["s0_Tdot"])
    assert dt == np.float32(d["s1_scalars"][1])
    assert ulps(T, d["s1_T"]) <= 4.0
    assert ulps(Tdot, d["s1_Tdot"]) <= 8.0
    b = ctx.thermal_buoyancy(Atemp)
    assert np.abs(b - d["s1_buoyancy"]).max() <= 4e-7 * np.abs(d["s1_buoyancy"]).max()


def test_heat_flux_nusselt_numbers(state):
    """heat_flux on the reference's step-0 state (the harness, like the input files' storage_spacing, evaluates it at step 0
    only): Nut, Nub within 1e-4 (north star: 0.1 %); the nodal fluxes are float sums and the surface value is the
    extrapolation 2 f(top) - f(top-1), so a few 1e-5 is rounding."""
    d, prob, ctx, _ = state
    ctx.set_temperature(d["s0_T"])
    ctx.set_velocity(d["s0_V1"], d["s0_V2"], d["s0_V3"])
    nut, nub = ctx.heat_flux()
    sc = d["s0_scalars"]
    assert abs(nut - sc[2]) <= 1e-4 * abs(sc[2])
    assert abs(nub - sc[3]) <= 1e-4 * abs(sc[3])


def test_coupled_timesteps_match_reference(state):
    """main()'s loop (Citcom.c:111-161) on the device: PG_timestep, thermal_buoyancy, general_stokes_solver,
    v_from_vector.  T after 3 steps within 0.1 % (relative L2) of the reference run with the same input."""
    d, prob, ctx, Atemp = state
    ctl = prob.control
    kw = dict(augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"], precondition=ctl["precondition"])
    T0 = d["s0_T"]
    ctx.set_temperature(T0)
    ctx.set_tdot(None)
    ctx.assemble_forces(d["s0_buoyancy"], want_host=False)
    ctx.general_stokes_solver(None, None, rebuild=1, guess=0, want_host=False, **kw)
    ctx.v_from_vector(want_host=False)
    rebuild = 1 if prob.visc["tdepv"] else 0
    for step in (1, 2, 3):
        _, _, dt, _ = ctx.PG_timestep()
        ctx.thermal_buoyancy(Atemp, want_host=False)
        ctx.general_stokes_solver(None, None, rebuild=rebuild, guess=2, want_host=False, **kw)
        ctx.v_from_vector(want_host=False)
        T = ctx.get_temperature()
        ref = d[f"s{step}_T"]
        assert abs(dt - d[f"s{step}_scalars"][1]) <= 2e-3 * d[f"s{step}_scalars"][1]
        assert np.linalg.norm(T - ref) <= 1e-3 * np.linalg.norm(ref)


# ---------------------------------------------------------------- extended-Boussinesq heating
@pytest.fixture(scope="module")
def eba_state():
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    text = inputfile.tdepv_box(16, 16, 8, 3, maxstep=4, adi_heating=1, visc_heating=1, dissipation_number=0.5, accuracy=1e-6)
    dumps, err = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_eba_")), nsteps=3, kat=True)
    d = dumps[0]
    prob = CartesianProblem(text)
    ctx = context_from_problem(prob)
    adv = d["kat_adv_params"]
    ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
    eb = d["s1_eba"]
    ctx.set_heating_params(1, 1, eb[0], eb[1], eb[2])
    yield d, prob, ctx, float(adv[5])
    ctx.close()


def test_process_heating_and_heated_timestep(eba_state):
    """process_heating (adiabatic + viscous heating, strain_rate_2_inv) on the reference's step-0 state, then the
    PG_timestep that consumes it: element heating terms within 2 ulp(fp32), T after the step within 4 ulp."""
    d, prob, ctx, Atemp = eba_state
    load_s0(d, ctx)
    ctx.set_element_viscosity(prob.levmax, d["s0_EVI"])
    adi, visc = ctx.process_heating()
    assert np.abs(d["s1_heating_adi"]).max() > 0 and np.abs(d["s1_heating_visc"]).max() > 0
    assert ulps(adi, d["s1_heating_adi"]) <= 2.0
    assert ulps(visc, d["s1_heating_visc"]) <= 2.0
    T, Tdot, dt, Tint = ctx.PG_timestep(d["s0_T"], d["s0_Tdot"])
    assert dt == np.float32(d["s1_scalars"][1])
    assert ulps(T, d["s1_T"]) <= 4.0


def test_coupled_timesteps_with_heating(eba_state):
    """main()'s loop with process_heating at the head of every step: T after 3 steps within 0.1 % of the reference."""
    d, prob, ctx, Atemp = eba_state
    ctl = prob.control
    kw = dict(augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"], precondition=ctl["precondition"])
    ctx.set_temperature(d["s0_T"])
    ctx.set_tdot(None)
    ctx.assemble_forces(d["s0_buoyancy"], want_host=False)
    ctx.general_stokes_solver(None, None, rebuild=1, guess=0, want_host=False, **kw)
    ctx.v_from_vector(want_host=False)
    for step in (1, 2, 3):
        dt, its = ctx.advance(Atemp, rebuild=1, **kw)
        T = ctx.get_temperature()
        ref = d[f"s{step}_T"]
        assert abs(dt - d[f"s{step}_scalars"][1]) <= 2e-3 * d[f"s{step}_scalars"][1]
        assert np.linalg.norm(T - ref) <= 1e-3 * np.linalg.norm(ref)
        adi, visc = ctx.process_heating() if step < 3 else (None, None)
        if adi is not None:   # the terms the NEXT step will use, against the reference's next-step dump
            assert np.linalg.norm(adi - d[f"s{step + 1}_heating_adi"]) <= 2e-3 * np.linalg.norm(d[f"s{step + 1}_heating_adi"])


@pytest.fixture(scope="module")
def phase_state():
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    text = inputfile.tdepv_box(16, 16, 8, 3, maxstep=4, adi_heating=1, visc_heating=1, dissipation_number=0.5, accuracy=1e-6,
                               Ra_410=100.0, Ra_670=-100.0)
    dumps, err = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_phase_")), nsteps=3, kat=True)
    d = dumps[0]
    prob = CartesianProblem(text)
    ctx = context_from_problem(prob)
    adv = d["kat_adv_params"]
    ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
    eb, ph = d["s1_eba"], d["s1_phase"]
    ctx.set_heating_params(1, 1, eb[0], eb[1], eb[2])
    ctx.set_phase_params(ph[0], ph[1], ph[2], ph[3], ph[4], ph[6], ph[7], ph[8])
    yield d, prob, ctx, float(adv[5])
    ctx.close()


def test_phase_change_and_latent_heating(phase_state):
    """phase_change on the reference's step-0 temperature (phase functions, transition temperatures), the latent-heating
    terms process_heating derives from them, and the buoyancy with the phase term."""
    d, prob, ctx, Atemp = phase_state
    load_s0(d, ctx)
    ctx.set_element_viscosity(prob.levmax, d["s0_EVI"])
    ph = d["s0_phase"]
    F6, F4, tT = ctx.phase_change(update_transT=True)
    assert abs(tT[0] - ph[5]) <= 1e-5 * abs(ph[5]) and abs(tT[1] - ph[9]) <= 1e-5 * abs(ph[9])
    assert np.abs(F6 - d["s0_Fas670"]).max() <= 1e-5 and np.abs(F4 - d["s0_Fas410"]).max() <= 1e-5
    assert d["s0_Fas670"].min() < 0.1 and d["s0_Fas670"].max() > 0.9            # the transition is inside the box
    adi, visc = ctx.process_heating()
    lat = ctx.get_heating_latent()
    assert np.abs(lat - d["s1_heating_latent"]).max() <= 1e-5
    assert np.linalg.norm(adi - d["s1_heating_adi"]) <= 1e-4 * np.linalg.norm(d["s1_heating_adi"])
    ctx.set_step(0)
    b = ctx.thermal_buoyancy(Atemp)
    assert np.abs(b - d["s0_buoyancy"]).max() <= 1e-5 * np.abs(d["s0_buoyancy"]).max()


def test_coupled_timesteps_with_phase_changes(phase_state):
    d, prob, ctx, Atemp = phase_state
    ctl = prob.control
    kw = dict(augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"], precondition=ctl["precondition"])
    ctx.set_temperature(d["s0_T"])
    ctx.set_tdot(None)
    ctx.set_step(0)
    ctx.thermal_buoyancy(Atemp, want_host=False)           # also leaves the step-0 phase functions resident
    ctx.general_stokes_solver(None, None, rebuild=1, guess=0, want_host=False, **kw)
    ctx.v_from_vector(want_host=False)
    for step in (1, 2, 3):
        dt, its = ctx.advance(Atemp, rebuild=1, **kw)
        T = ctx.get_temperature()
        ref = d[f"s{step}_T"]
        assert abs(dt - d[f"s{step}_scalars"][1]) <= 2e-3 * d[f"s{step}_scalars"][1]
        assert np.linalg.norm(T - ref) <= 1e-3 * np.linalg.norm(ref)


def test_observables_on_a_developed_state():
    """Nusselt numbers, layer vrms (averages, Process_velocity.c:179) and the volume-weighted Vrms on a DEVELOPED state: the
    reference runs 30 timesteps from a strong perturbation with storage_spacing=1 (so that it re-evaluates heat_flux and averages on every step, not
    only at step 0), the device evaluates the same observables from the reference's T and V of each step: 0.1 % (north star)."""
    import tempfile
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import StokesContext, context_from_problem
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=40, accuracy=1e-4, storage_spacing=1, rayleigh=1e6, perturbmag=0.3,
                              viscE="4.6,4.6,4.6,4.6")
    nsteps = 30
    d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_obs_"), nsteps=nsteps)[0][0]
    prob = CartesianProblem(txt)
    ctx = context_from_problem(prob)
    lm = prob.levmax
    noz = prob.dims(lm)[2]
    ctx.set_energy_params(0.75, 0.0, 0.5, 2, np.ones(noz, np.float32), np.ones(noz, np.float32), 0.0)
    ctl = prob.control
    seen = []
    for k in (0, 10, 20, 30):
        sc = d[f"s{k}_scalars"]
        ctx.set_temperature(d[f"s{k}_T"])
        # main()'s order (Citcom.c:111-161): energy step, process_temp_field (heat_flux with the NEW temperature and the velocity
        # of the previous solve), Stokes solve, process_new_velocity (averages with the new velocity)
        kv = max(k - 1, 0)
        ctx.set_velocity(d[f"s{kv}_V1"], d[f"s{kv}_V2"], d[f"s{kv}_V3"])
        nut, nub = ctx.heat_flux()
        assert abs(nut - sc[2]) <= 1e-3 * abs(sc[2]) and abs(nub - sc[3]) <= 1e-3 * abs(sc[3]), (k, nut, nub, sc[2], sc[3])
        ctx.set_velocity(d[f"s{k}_V1"], d[f"s{k}_V2"], d[f"s{k}_V3"])
        ctx.set_element_viscosity(lm, d[f"s{k}_EVI"])
        vr, vi = ctx.averages()
        ref_vr, ref_vi = d[f"s{k}_Have_vrms"], d[f"s{k}_Have_Vi"]
        assert np.abs(vr - ref_vr).max() <= 1e-3 * np.abs(ref_vr).max(), k
        assert np.abs(vi - ref_vi).max() <= 1e-3 * np.abs(ref_vi).max(), k
        z = d[f"s{k}_XP3"]
        V, Vr = StokesContext.volume_vrms(vr, z), StokesContext.volume_vrms(ref_vr, z)
        assert abs(V - Vr) <= 1e-3 * Vr
        seen.append((float(sc[2]), Vr))
    ctx.close()
    # the state did develop: Nu and Vrms moved away from their step-0 values
    assert abs(seen[-1][0] - seen[0][0]) > 1.0 and abs(seen[-1][1] - seen[0][1]) > 0.1 * seen[0][1]


@pytest.mark.parametrize("geometry", ["cart3d", "Rsphere"])
def test_stress_and_dynamic_topography(geometry):
    """get_stress / get_STD_topo (Topo_gravity.c:352,307) on the reference's state after a Stokes solve: the six nodal stress
    fields and the top / bottom dynamic topography within 1e-4 of the reference's own functions (float sums)."""
    import tempfile
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    if geometry == "Rsphere":          # the Rsphere branch of get_stress (Topo_gravity.c:429-439)
        from test_gpu_build import build_ctx
        txt = inputfile.input1_rsphere(levels=3, maxstep=2, accuracy=1e-5, TDEPV="on", perturbmag=0.05)
        d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_stress_"), nsteps=0, kat=True)[0][0]
        ctx = build_ctx(d, 0, 0.0)
        lm = d.levmax
    else:
        txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=2, accuracy=1e-5, viscE="4.6,4.6,4.6,4.6")
        d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_stress_"), nsteps=0, kat=True)[0][0]      # the known answers are taken on the step-0 state
        prob = CartesianProblem(txt)
        ctx = context_from_problem(prob)
        lm = prob.levmax
    ctx.set_temperature(d["s0_T"])
    ctx.set_velocity(d["s0_V1"], d["s0_V2"], d["s0_V3"])
    ctx.set_element_viscosity(lm, d["s0_EVI"])
    ctx.pvec_upload(d["s0_P"])
    S, tpg, tpgb = ctx.get_stress_topo()
    ctx.close()
    for q in range(6):
        ref = d[f"kat_stress{q}"]
        assert np.abs(S[q] - ref).max() <= 1e-4 * np.abs(ref).max(), q
    assert np.abs(tpg - d["kat_tpg"]).max() <= 1e-4 * np.abs(d["kat_tpg"]).max()
    assert np.abs(tpgb - d["kat_tpgb"]).max() <= 1e-4 * np.abs(d["kat_tpgb"]).max()


def test_output_staging_writes_the_reference_files():
    """Output staging (ccu_output_stage / _write / _wait): the .velo file written by the background thread from the staged device
    fields is byte-identical to the file the reference's output_velo_related wrote for the same state; the .temp file agrees in
    its T and Vz columns (its third column is a host-only diagnostic of the reference)."""
    import glob
    import tempfile
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=1, storage_spacing=1)
    wd = tempfile.mkdtemp(prefix="ccu_out_")
    d = po.run_harness(txt, wd, nsteps=0)[0][0]
    ref_velo = open(glob.glob(wd + "/out/*.velo.0.0")[0]).read()
    ref_temp = open(glob.glob(wd + "/out/*.temp.0.0")[0]).read().split("\n")
    ctx = context_from_problem(CartesianProblem(txt))
    ctx.set_temperature(d["s0_T"])
    ctx.set_velocity(d["s0_V1"], d["s0_V2"], d["s0_V3"])
    ctx.output_stage()
    pre = tempfile.mkdtemp(prefix="ccu_outdev_") + "/run"
    ctx.output_write(pre, 0, 0, 0, 0.0)
    ctx.output_wait()
    ctx.close()
    assert open(pre + ".velo.0.0").read() == ref_velo
    mine = open(pre + ".temp.0.0").read().split("\n")
    assert mine[0] == ref_temp[0] and len(mine) == len(ref_temp)
    for a, b in zip(mine[1:-1], ref_temp[1:-1]):
        assert a.split()[:2] == b.split()[:2]


def test_regional_sphere_heating_and_energy_step():
    """Regional-spherical geometry, kernel level: process_heating (Rsphere branch of strain_rate_2_inv, Viscosity_structures.c:996-1040,
    adiabatic heating from the radial velocity) and PG_timestep (Rsphere branches of pg_shape_fn / element_residual) on the
    reference's step-0 state against its step-1 arrays."""
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    from test_gpu_build import build_ctx
    text = inputfile.input1_rsphere(levels=3, maxstep=3, accuracy=1e-6, TDEPV="on", perturbmag=0.05, adi_heating=1, visc_heating=1,
                                    surf_temp=0.078947, storage_spacing=1)
    d = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_rsheat_")), nsteps=2, kat=True)[0][0]
    ctx = build_ctx(d, 0, 0.0)
    adv = d["kat_adv_params"]
    ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
    eb = d["s1_eba"]
    ctx.set_heating_params(1, 1, eb[0], eb[1], eb[2])
    load_s0(d, ctx)
    ctx.set_element_viscosity(d.levmax, d["s0_EVI"])
    adi, visc = ctx.process_heating()
    assert np.abs(d["s1_heating_adi"]).max() > 0 and np.abs(d["s1_heating_visc"]).max() > 0
    assert ulps(adi, d["s1_heating_adi"]) <= 2.0
    assert ulps(visc, d["s1_heating_visc"]) <= 8.0, ulps(visc, d["s1_heating_visc"])
    T, Tdot, dt, Tint = ctx.PG_timestep(d["s0_T"], d["s0_Tdot"])
    assert abs(dt - d["s1_scalars"][1]) <= 1e-6 * d["s1_scalars"][1]
    assert ulps(T, d["s1_T"]) <= 8.0, ulps(T, d["s1_T"])
    # thermal_buoyancy: the layer averages over spherical shells (return_horiz_ave with the Rsphere surface Jacobian of
    # get_global_1d_shape_fn) removed from Ra T alpha
    ctx.set_temperature(d["s0_T"])
    b = ctx.thermal_buoyancy(float(adv[5]))
    assert np.abs(b - d["s0_buoyancy"]).max() <= 1e-5 * np.abs(d["s0_buoyancy"]).max()
    # averages: layer vrms and viscosity of the reference's step-1 state
    ctx.set_velocity(d["s1_V1"], d["s1_V2"], d["s1_V3"])
    ctx.set_element_viscosity(d.levmax, d["s1_EVI"])
    vr, vi = ctx.averages()
    assert np.abs(vr - d["s1_Have_vrms"]).max() <= 1e-4 * np.abs(d["s1_Have_vrms"]).max()
    assert np.abs(vi - d["s1_Have_Vi"]).max() <= 1e-4 * np.abs(d["s1_Have_Vi"]).max()
    # heat_flux: Nusselt numbers from the radial derivative (main()'s order: the new temperature with the velocity of the solve before)
    for k in (1, 2):
        ctx.set_temperature(d[f"s{k}_T"])
        ctx.set_velocity(d[f"s{k - 1}_V1"], d[f"s{k - 1}_V2"], d[f"s{k - 1}_V3"])
        nut, nub = ctx.heat_flux()
        sc = d[f"s{k}_scalars"]
        assert abs(nut - sc[2]) <= 1e-4 * abs(sc[2]) and abs(nub - sc[3]) <= 1e-4 * abs(sc[3]), (k, nut, nub, sc[2], sc[3])
    ctx.close()


def test_regional_sphere_phase_changes():
    """phase_change on the regional sphere (the depth coordinate is r = E->SX[3], the layer averages are shell averages): phase functions,
    transition temperatures, latent heating and the buoyancy with the phase terms against the reference."""
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    from test_gpu_build import build_ctx
    text = inputfile.input1_rsphere(levels=3, maxstep=3, accuracy=1e-6, TDEPV="on", perturbmag=0.05, adi_heating=1, visc_heating=1,
                                    surf_temp=0.078947, Ra_410=100.0, Ra_670=-100.0, storage_spacing=1)
    d = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_rsphase_")), nsteps=2, kat=True)[0][0]
    ctx = build_ctx(d, 0, 0.0)
    adv = d["kat_adv_params"]
    ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
    eb, ph = d["s1_eba"], d["s1_phase"]
    ctx.set_heating_params(1, 1, eb[0], eb[1], eb[2])
    ctx.set_phase_params(ph[0], ph[1], ph[2], ph[3], ph[4], ph[6], ph[7], ph[8])
    load_s0(d, ctx)
    ctx.set_element_viscosity(d.levmax, d["s0_EVI"])
    p0 = d["s0_phase"]
    F6, F4, tT = ctx.phase_change(update_transT=True)
    assert abs(tT[0] - p0[5]) <= 1e-5 * abs(p0[5]) and abs(tT[1] - p0[9]) <= 1e-5 * abs(p0[9])
    assert np.abs(F6 - d["s0_Fas670"]).max() <= 1e-5 and np.abs(F4 - d["s0_Fas410"]).max() <= 1e-5
    assert d["s0_Fas670"].min() < 0.1 and d["s0_Fas670"].max() > 0.9            # the transition is inside the shell
    adi, visc = ctx.process_heating()
    lat = ctx.get_heating_latent()
    assert np.abs(d["s1_heating_latent"] - 1).max() > 1e-3
    assert np.abs(lat - d["s1_heating_latent"]).max() <= 1e-5
    ctx.set_step(0)
    b = ctx.thermal_buoyancy(float(adv[5]))
    assert np.abs(b - d["s0_buoyancy"]).max() <= 1e-5 * np.abs(d["s0_buoyancy"]).max()
    ctx.close()
