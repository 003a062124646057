"""CPU test: the host-side problem setup (citcomcu_b200.problem) against the arrays the reference's own
read_instructions builds (oracle dumps): coordinates, BC flag bits, initial temperature and material
groups bit-exact; buoyancy to float rounding (its layer average is a float64 quadrature here)."""
import numpy as np
import pytest

from conftest import get_case, CASES
from citcomcu_b200.problem import CartesianProblem, BC_MASK, INTX, INTY, INTZ


@pytest.mark.parametrize("name", ["busse_l3", "tdepv_l3", "input1_cart_l3"])
def test_setup_matches_reference(name):
    d = get_case(name)[0]
    txt = CASES[name]()[0]
    P = CartesianProblem(txt)
    assert (P.levmin, P.levmax) == (d.levmin, d.levmax)
    ctl = d.control()
    for k in ("v_steps_low", "v_steps_high", "down_heavy", "up_heavy", "mg_cycle", "p_iterations", "precondition", "augmented_Lagr"):
        assert P.control[k] == ctl[k], k
    assert P.control["accuracy"] == ctl["accuracy"] and P.control["augmented"] == ctl["augmented"]
    for lev in range(d.levmin, d.levmax + 1):
        dm = d.dims(lev)
        assert P.dims(lev) == (dm["nox"], dm["noy"], dm["noz"])
        if f"L{lev}_XX1" in d:
            for X, nm in zip(P.coordinates(lev), ("XX1", "XX2", "XX3")):
                assert np.array_equal(X, d[f"L{lev}_{nm}"]), (lev, nm)
        mask = np.uint32(BC_MASK | (INTX | INTY | INTZ if lev == d.levmax else 0))
        assert np.array_equal(P.node_flags(lev) & mask, d[f"L{lev}_NODE"] & mask), lev
    T = P.initial_temperature()
    assert np.array_equal(T, d["s0_T"])
    assert np.array_equal(P.material(), d["s0_mat"])
    b = P.buoyancy(d["s0_T"])
    assert np.allclose(b, d["s0_buoyancy"], rtol=0, atol=2e-6 * np.abs(d["s0_buoyancy"]).max())


def test_imposed_velocities_match_reference(oracle_built):
    """E->VB (no-slip lid and base with non-zero values) where a velocity flag reads it, and the flags of that configuration."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import VBX, VBY, VBZ
    if not po.have_ref():
        pytest.skip("needs the reference build (oracle/_ref)")
    txt = inputfile.tdepv_box(8, 8, 8, 2, maxstep=1, topvbc=1, plate_velocity=40.0, topvbyval=-15.0, botvbc=1, botvbxval=3.0, botvbyval=7.0)
    d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_vbsetup_"), nsteps=0)[0][0]
    P = CartesianProblem(txt)
    lm = d.levmax
    node = d[f"L{lm}_NODE"]
    assert np.array_equal(P.node_flags(lm) & np.uint32(BC_MASK), node & np.uint32(BC_MASK))
    for vb, nm, bit in zip(P.velocity_bcs(), ("VB1", "VB2", "VB3"), (VBX, VBY, VBZ)):
        m = (node & np.uint32(bit)) != 0
        assert np.array_equal(vb[m], d[nm][m]), nm
        assert np.abs(d[nm][m]).max() > 0 or nm == "VB3"


@pytest.mark.parametrize("nproc", [(1, 1, 1), (2, 1, 2)])
def test_regional_sphere_mesh_and_flags_match_reference(nproc, oracle_built):
    """SphericalProblem: E->SXX (theta, phi, r), the Cartesian node positions E->XX and the boundary flags of every level and rank of a
    regional-spherical block with refined radial boundary layers, the initial temperature and the material groups, bit for bit
    against the reference's own setup."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import SphericalProblem
    if not po.have_ref():
        pytest.skip("needs the reference build (oracle/_ref)")
    txt = inputfile.input1_rsphere(levels=3, maxstep=1, nproc=nproc)
    world = nproc[0] * nproc[1] * nproc[2]
    dumps = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rssetup_"), nsteps=0, nproc=world)[0]
    for d in dumps:
        P = SphericalProblem(txt, me_loc=d.control()["me_loc"])
        for lev in range(d.levmin, d.levmax + 1):
            dm = d.dims(lev)
            assert P.dims(lev) == (dm["nox"], dm["noy"], dm["noz"])
            if f"L{lev}_SXX1" not in d:
                continue
            for A, nm in zip(P.spherical_coordinates(lev), ("SXX1", "SXX2", "SXX3")):
                assert np.array_equal(A, d[f"L{lev}_{nm}"]), (lev, nm)
            for A, nm in zip(P.coordinates(lev), ("XX1", "XX2", "XX3")):
                assert np.array_equal(A, d[f"L{lev}_{nm}"]), (lev, nm)
            mask = np.uint32(BC_MASK | (INTX | INTY | INTZ if lev == d.levmax else 0))
            assert np.array_equal(P.node_flags(lev) & mask, d[f"L{lev}_NODE"] & mask), lev
        assert np.array_equal(P.initial_temperature(), d["s0_T"])
        assert np.array_equal(P.material(), d["s0_mat"])


def test_regional_sphere_imposed_velocities_match_reference(oracle_built):
    """E->VB of a regional-spherical block with a moving lid, where a velocity flag reads it, and the flags of that configuration."""
    import tempfile
    from conftest import po
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import SphericalProblem, VBX, VBY, VBZ
    if not po.have_ref():
        pytest.skip("needs the reference build (oracle/_ref)")
    txt = inputfile.input1_rsphere(levels=2, maxstep=1, topvbc=1, plate_velocity=40.0, topvbyval=-15.0)
    d = po.run_harness(txt, tempfile.mkdtemp(prefix="ccu_rsvbsetup_"), nsteps=0)[0][0]
    P = SphericalProblem(txt)
    lm = d.levmax
    node = d[f"L{lm}_NODE"]
    assert np.array_equal(P.node_flags(lm) & np.uint32(BC_MASK), node & np.uint32(BC_MASK))
    for vb, nm, bit in zip(P.velocity_bcs(), ("VB1", "VB2", "VB3"), (VBX, VBY, VBZ)):
        m = (node & np.uint32(bit)) != 0
        assert np.array_equal(vb[m], d[nm][m]), nm
    assert np.abs(d["VB1"]).max() == 40.0
