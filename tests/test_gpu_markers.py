"""GPU parity of the marker kernels against the unmodified reference's Euler / Runge_Kutta on its
own marker state: element assignment bit-exact (north star), positions and interpolated velocities bit-exact (the
kernels restate the reference's operand types, no FMA contraction), nodal composition within 1 ulp(fp32)."""
import tempfile
from pathlib import Path

import numpy as np
import pytest

from conftest import po
from citcomcu_b200 import inputfile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def state():
    if not po.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem
    text = inputfile.tdepv_box(16, 16, 8, 3, maxstep=3, composition=1, rayleigh_comp=5e6, markers_per_ele=8, comp_depth=0.4)
    dumps, err = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_markers_")), nsteps=2, kat=True)
    d = dumps[0]
    prob = CartesianProblem(text)
    ctx = context_from_problem(prob)
    ip, dp = d["mk_ints"], d["mk_doubles"]
    ctx.markers_setup(int(ip[3]), int(ip[1]), int(ip[0]), d["mk_XP1"], d["mk_XP2"], d["mk_XP3"], d["mk_RG3"], dp[0:3], dp[3:6],
                      d["mk_Element"], Acomp=float(dp[7]))
    yield d, prob, ctx, np.float32(dp[6])
    ctx.close()


def test_euler_then_runge_kutta(state):
    d, prob, ctx, dt = state
    ctx.markers_upload(d["mk_in_XMC1"], d["mk_in_XMC2"], d["mk_in_XMC3"], d["mk_in_C12"], d["mk_in_CElement"], d["mk_in_CE"])
    ctx.set_velocity(d["mk_in_V1"], d["mk_in_V2"], d["mk_in_V3"])
    ctx.Euler(dt)
    m = ctx.markers_download()
    for a in range(3):
        assert np.array_equal(m["VO"][a], d[f"mk_euler_VO{a + 1}"])
        assert np.array_equal(m["XMCpred"][a], d[f"mk_euler_XMCpred{a + 1}"])
    assert np.array_equal(m["CElement"], d["mk_euler_CElement"])           # marker -> element assignment: bit-exact
    assert np.array_equal(m["CE"], d["mk_euler_CE"])
    assert np.abs(m["C"] - d["mk_euler_C"]).max() <= 2 * np.finfo(np.float32).eps
    ctx.Runge_Kutta(dt)
    m = ctx.markers_download()
    for a in range(3):
        assert np.array_equal(m["Vpred"][a], d[f"mk_rk_Vpred{a + 1}"])
        assert np.array_equal(m["XMC"][a], d[f"mk_rk_XMC{a + 1}"])
    assert np.array_equal(m["CElement"], d["mk_rk_CElement"])
    assert np.array_equal(m["CE"], d["mk_rk_CE"])
    assert np.abs(m["C"] - d["mk_rk_C"]).max() <= 2 * np.finfo(np.float32).eps


def test_counts_are_conserved(state):
    """Size-independent property: every marker lands in exactly one element, so the per-element counts behind CE sum to
    the number of markers; CE stays in [0, 1]."""
    d, prob, ctx, dt = state
    ctx.markers_upload(d["mk_in_XMC1"], d["mk_in_XMC2"], d["mk_in_XMC3"], d["mk_in_C12"], d["mk_in_CElement"], d["mk_in_CE"])
    ctx.set_velocity(d["mk_in_V1"], d["mk_in_V2"], d["mk_in_V3"])
    ctx.Euler(dt)
    m = ctx.markers_download()
    nel = prob.nel(prob.levmax)
    assert m["CElement"].min() >= 1 and m["CElement"].max() <= nel
    assert 0.0 <= m["CE"].min() and m["CE"].max() <= 1.0
    dense = np.bincount(m["CElement"][d["mk_in_C12"] == 1] - 1, minlength=nel).astype(np.float32)
    total = np.bincount(m["CElement"] - 1, minlength=nel).astype(np.float32)
    has = total > 0
    assert np.array_equal(m["CE"][has], dense[has] / total[has])
