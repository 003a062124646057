"""GPU tests: operator construction on the device vs the arrays the reference's host code builds
(mass_matrix, get_elt_g, get_system_viscosity, project_viscosity, construct_node_ks,
build_diagonal_of_Ahat, assemble_forces).  Bar: bit-exact (integer-like comparison of the fp32
coefficient arrays); `exp` in the viscosity law may differ from glibc in the last double bit, so
EVI / K allow a vanishing fraction of 1-ulp fp32 differences and report it."""
import numpy as np
import pytest

from conftest import get_case, CASES

pytestmark = pytest.mark.gpu


def build_ctx(d, tdepv, viscE, N0=None):
    from citcomcu_b200.stokes import StokesContext
    ctl = d.control()
    nox, noy, noz = {}, {}, {}
    for lev in range(ctl["levmin"], ctl["levmax"] + 1):
        dm = d.dims(lev)
        nox[lev], noy[lev], noz[lev] = dm["nox"], dm["noy"], dm["noz"]
    ctx = StokesContext(ctl["levmin"], ctl["levmax"], nox, noy, noz, v_steps_low=ctl["v_steps_low"],
                        v_steps_high=ctl["v_steps_high"], down_heavy=ctl["down_heavy"], up_heavy=ctl["up_heavy"],
                        mg_cycle=ctl["mg_cycle"], p_iterations=ctl["p_iterations"], accuracy=ctl["accuracy"])
    for lev in range(ctl["levmin"], ctl["levmax"] + 1):
        ctx.set_node_flags(lev, d[f"L{lev}_NODE"])
        ctx.set_coordinates(lev, d[f"L{lev}_XX1"], d[f"L{lev}_XX2"], d[f"L{lev}_XX3"])
        if f"L{lev}_SXX1" in d:                      # regional-spherical run: XX above are the Cartesian node positions
            ctx.set_spherical_coordinates(lev, d[f"L{lev}_SXX1"], d[f"L{lev}_SXX2"], d[f"L{lev}_SXX3"])
    ctx.build_geometry()
    ctx.set_viscosity_law(tdepv, 0, N0 or [1, 1, 1, 1], [viscE] * 4, [273] * 4, [5e-6] * 4)
    ctx.set_material(d["s0_mat"])
    ctx.set_temperature(d["s0_T"])
    ctx.get_system_viscosity()
    ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
    return ctx


def frac_diff(a, b):
    return float(np.mean(a != b))


@pytest.mark.parametrize("name,tdepv,viscE", [("busse_l3", 0, 0.0), ("tdepv_l3", 1, 11.512925), ("input1_cart_l3", 0, 0.0)])
def test_device_operator_construction_bit_exact(name, tdepv, viscE):
    d = get_case(name)[0]
    ctx = build_ctx(d, tdepv, viscE)
    rep = {}
    for lev in range(d.levmin, d.levmax + 1):
        for arr in ("TWW", "MASS", "eco_size", "elt_del"):
            assert np.array_equal(ctx.get_level_array(lev, arr), d[f"L{lev}_{arr}"]), (lev, arr)
        evi = ctx.get_level_array(lev, "EVI")
        ref = d[f"L{lev}_EVI"]
        rep[f"EVI{lev}"] = frac_diff(evi, ref)
        assert np.allclose(evi, ref, rtol=3e-7, atol=0) and rep[f"EVI{lev}"] < 1e-3
        k1, k2, k3, BI = ctx.get_stiffness(lev)
        for k, nm in ((k1, "Eqn_k1"), (k2, "Eqn_k2"), (k3, "Eqn_k3")):
            ref = d[f"L{lev}_{nm}"]
            rep[f"{nm}_{lev}"] = frac_diff(k, ref)
            assert np.allclose(k, ref, rtol=1e-6, atol=1e-6 * np.abs(ref).max())
            if not tdepv:
                assert np.array_equal(k, ref), (lev, nm)
        refBI = d[f"L{lev}_BI"]
        assert np.allclose(BI, refBI, rtol=1e-6 if tdepv else 1e-14, atol=0)
        bpi = ctx.get_level_array(lev, "BPI")
        assert np.allclose(bpi, d[f"L{lev}_BPI"], rtol=1e-6 if tdepv else 1e-13, atol=0)
    F = ctx.assemble_forces(d["s0_buoyancy"])
    assert np.allclose(F, d["s0_F"], rtol=1e-13, atol=1e-15 * np.abs(d["s0_F"]).max())
    print("fraction of entries differing:", {k: v for k, v in rep.items() if v})
    # and the solve on the device-built operator lands on the reference's solution
    lm = d.levmax
    n, npno = d.dims(lm)["neq"], d.dims(lm)["npno"]
    acc = d.control()["accuracy"]
    V, P, steps, res, hist = ctx.solve_Ahat_p_fhat(np.zeros(n), np.zeros(npno), F, acc, 375)
    assert np.linalg.norm(V - d["s0_U"]) < 20 * acc * np.linalg.norm(d["s0_U"])
    ctx.close()


RHEOL_CASES = {
    # rheol: (viscE, viscT, viscZ) -- Viscosity_structures.c:568-742
    2: (3.0, 0.5, 1.0),              # eta0 exp((E + (1-z) Z) / (T + T0))
    4: (5.0, 0.5, 2.0),              # eta0 exp(E (Tc - T) + (1-z) Z)
    10: (3.0, 0.5, 1.0),             # eta0 exp(E / (T + T0) + Z)
    11: (74.357912, 4.507123, 5e-6), # eta0 exp(E/(T+T0) - E/(0.5+T0)): examples/Busse1993/case2.input
    0: (11.512925, 273.0, 5e-6),
}


@pytest.mark.parametrize("rheol,smooth", [(2, 1), (4, 1), (10, 1), (11, 1), (0, 0), (0, 2), (0, 3), (11, 3), (4, 0)])
def test_rheologies_and_viscosity_coarsening_modes(rheol, smooth):
    """get_system_viscosity for every temperature-dependent law of the reference (RHEOL 0, 1, 2, 3, 4, 10, 11) and
    project_viscosity for the four visc_smooth_cycles modes (Solver_multigrid.c:398-474), against the arrays the unmodified
    reference builds from the same input file: Gauss-point viscosity on every level within one fp32 ulp (`exp`), stiffness 1e-6."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    E, T0, Z = RHEOL_CASES[rheol]
    four = lambda v: ",".join([repr(float(v))] * 4)      # noqa: E731
    txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=1, rheol=rheol, visc_smooth_cycles=smooth, viscE=four(E), viscT=four(T0), viscZ=four(Z))
    d = po.run_harness(txt, tempfile.mkdtemp(prefix=f"ccu_rheol{rheol}_{smooth}_"), nsteps=0)[0][0]
    prob = CartesianProblem(txt)
    ctx = context_from_problem(prob)
    ctl = prob.control
    ctx.set_temperature(d["s0_T"])
    ctx.get_system_viscosity()
    ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
    for lev in range(d.levmin, d.levmax + 1):
        evi, ref = ctx.get_level_array(lev, "EVI"), d[f"L{lev}_EVI"]
        assert np.allclose(evi, ref, rtol=3e-7, atol=0), (lev, float(np.abs(evi / ref - 1).max()))
        assert frac_diff(evi, ref) < 2e-2
        k1, k2, k3, BI = ctx.get_stiffness(lev)
        for k, nm in ((k1, "Eqn_k1"), (k2, "Eqn_k2"), (k3, "Eqn_k3")):
            ref = d[f"L{lev}_{nm}"]
            assert np.allclose(k, ref, rtol=1e-6, atol=1e-6 * np.abs(ref).max()), (lev, nm)
    ctx.close()


@pytest.mark.parametrize("geometry", ["cart3d", "Rsphere"])
def test_imposed_velocity_force_term(geometry):
    """assemble_forces with non-zero E->VB: the -K.VB term of get_elt_f (Element_calculations.c:1038-1063) against the reference's
    F at steps 0 and 1 (step k is assembled with the viscosity of the update before it)."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    if geometry == "Rsphere":      # the element matrix of the Rsphere branch, point bases of the element itself (get_elt_k with iconv = 1)
        txt = inputfile.input1_rsphere(levels=3, maxstep=3, accuracy=1e-5, TDEPV="on", VISC_UPDATE="on", update_every_steps=1, perturbmag=0.05,
                                       topvbc=1, plate_velocity=40.0, topvbyval=-15.0, storage_spacing=1)
    else:
        txt = inputfile.tdepv_box(16, 16, 8, 3, maxstep=3, accuracy=1e-5, topvbc=1, plate_velocity=40.0, topvbyval=-15.0, storage_spacing=1)
    d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_vbF_"), nsteps=1)[0][0]
    assert np.abs(d["VB1"]).max() == 40.0 and np.abs(d["VB2"]).max() == 15.0
    ctx = build_ctx(d, 0, 0.0)
    lm = d.levmax
    ctx.set_velocity_bcs(d["VB1"], d["VB2"], d["VB3"])
    plain = ctx.assemble_forces(d["s0_buoyancy"])              # viscosity of build_ctx (isoviscous): not the reference's
    ctx.set_element_viscosity(lm, d["s0_EVI"])
    for k in (0, 1):
        F = ctx.assemble_forces(d[f"s{k}_buoyancy"])
        ref = d[f"s{k}_F"]
        assert np.abs(F - ref).max() < 1e-12 * np.abs(ref).max(), (k, np.abs(F - ref).max() / np.abs(ref).max())
    assert np.abs(plain - d["s0_F"]).max() > 1e-3 * np.abs(d["s0_F"]).max()
    ctx.close()


@pytest.mark.parametrize("tdepv", ["off", "on"])
def test_regional_sphere_operator_construction(tdepv):
    """BASELINE config 4 geometry (examples/input1's regional-spherical block): the Rsphere branches of mass_matrix (ECO.size),
    get_elt_g, get_elt_k and get_elt_f on the device against the arrays the reference's host code builds.  The reference evaluates
    the basis-projection matrices once per radial column and reuses them; the device evaluates them per element, so the fp32 arrays
    agree to rounding (a vanishing fraction of 1-ulp differences), not bit for bit."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    txt = inputfile.input1_rsphere(levels=3, maxstep=1, accuracy=1e-6, TDEPV=tdepv, perturbmag=0.05)
    d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rsbuild_"), nsteps=0)[0][0]
    if tdepv == "on":
        assert d[f"L{d.levmax}_EVI"].max() > 10 * d[f"L{d.levmax}_EVI"].min()
    ctx = build_ctx(d, 0, 0.0)
    ctx.set_element_viscosity(d.levmax, d[f"L{d.levmax}_EVI"])         # the reference's viscosity (visc_from_mat + VMIN / VMAX)
    ctl = d.control()
    ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
    rep = {}
    for lev in range(d.levmin, d.levmax + 1):
        for arr in ("TWW", "MASS"):
            assert np.array_equal(ctx.get_level_array(lev, arr), d[f"L{lev}_{arr}"]), (lev, arr)
        for arr, tol in (("eco_size", 1e-6), ("elt_del", 2e-6)):
            got, ref = ctx.get_level_array(lev, arr), d[f"L{lev}_{arr}"]
            rep[f"{arr}{lev}"] = frac_diff(got, ref)
            assert np.abs(got - ref).max() <= tol * np.abs(ref).max(), (lev, arr, np.abs(got - ref).max() / np.abs(ref).max())
        assert np.allclose(ctx.get_level_array(lev, "EVI"), d[f"L{lev}_EVI"], rtol=3e-7, atol=0), lev
        k1, k2, k3, BI = ctx.get_stiffness(lev)
        for k, nm in ((k1, "Eqn_k1"), (k2, "Eqn_k2"), (k3, "Eqn_k3")):
            ref = d[f"L{lev}_{nm}"]
            rep[f"{nm}_{lev}"] = frac_diff(k, ref)
            assert np.abs(k - ref).max() <= 2e-6 * np.abs(ref).max(), (lev, nm, np.abs(k - ref).max() / np.abs(ref).max())
        assert np.allclose(BI, d[f"L{lev}_BI"], rtol=1e-5, atol=0), lev
        assert np.allclose(ctx.get_level_array(lev, "BPI"), d[f"L{lev}_BPI"], rtol=1e-5, atol=0), lev
    F = ctx.assemble_forces(d["s0_buoyancy"])
    assert np.abs(F - d["s0_F"]).max() <= 1e-12 * np.abs(d["s0_F"]).max()
    print("fraction of entries differing:", {k: round(v, 4) for k, v in rep.items() if v})
    lm = d.levmax
    n, npno = d.dims(lm)["neq"], d.dims(lm)["npno"]
    V, P, steps, res, hist = ctx.solve_Ahat_p_fhat(np.zeros(n), np.zeros(npno), F, ctl["accuracy"], 375)
    assert np.linalg.norm(V - d["s0_U"]) < 20 * ctl["accuracy"] * np.linalg.norm(d["s0_U"])
    ctx.close()


def test_regional_sphere_from_the_python_mirror():
    """A regional-spherical block driven from Python alone: SphericalProblem (mesh, flags, initial temperature, material groups) ->
    context_from_problem -> shell-averaged thermal_buoyancy and general_stokes_solver on the device, against the reference's step 0."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import SphericalProblem
    from citcomcu_b200.stokes import context_from_problem
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    txt = inputfile.input1_rsphere(levels=3, maxstep=1, accuracy=1e-6, TDEPV="on", perturbmag=0.05)
    d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rspy_"), nsteps=0, kat=True)[0][0]
    prob = SphericalProblem(txt)
    ctx = context_from_problem(prob)
    T = prob.initial_temperature()
    assert np.array_equal(T, d["s0_T"])
    ctx.set_temperature(T)
    adv = d["kat_adv_params"]
    ctx.set_energy_params(adv[0], adv[1], adv[2], int(adv[3]), d["kat_diffusivity"], d["kat_expansivity"], adv[4])
    b = ctx.thermal_buoyancy(float(adv[5]))
    assert np.abs(b - d["s0_buoyancy"]).max() <= 1e-5 * np.abs(d["s0_buoyancy"]).max()
    ctl = prob.control
    U, P, its, res = ctx.general_stokes_solver(T, b, rebuild=1, augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"],
                                               precondition=ctl["precondition"], guess=0)
    assert np.linalg.norm(U - d["s0_U"]) <= 20 * ctl["accuracy"] * np.linalg.norm(d["s0_U"])
    ctx.close()
