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
