"""CPU model of the overlapped smoother sweep (d_relax_sweeps in csrc/ccu_stokes.cu, option halo_overlap): an 8-colour Gauss-Seidel
sweep on a small 27-point-stencil problem whose subdomain faces hold duplicated nodes.  The serial order is: partial rows of the
duplicated nodes from the current solution, their damped-Jacobi update, then the colour passes 7 .. 0 over the other nodes.  The
overlapped order runs the "far" part of pass p (nodes more than p + 1 away from a duplicated node) BEFORE the duplicated nodes are
updated -- and, worst case for the hazard analysis, reads the partial rows only after all far passes have written -- then the
update, then the "near" parts.  Both must give the same numbers bit for bit for every combination of faces that can hold
duplicated nodes (the GPU test checks the 2-subdomain splits on hardware; this covers edges and corners of 2x2x2 and inner blocks)."""
import itertools

import numpy as np
import pytest


def build(n, faces, seed):
    rng = np.random.default_rng(seed)
    ny, nx, nz = n
    idx = np.arange(ny * nx * nz).reshape(ny, nx, nz)
    I, J, K = np.meshgrid(np.arange(ny), np.arange(nx), np.arange(nz), indexing="ij")
    d = np.full(n, 15)
    for has, dist in zip(faces, (I, ny - 1 - I, J, nx - 1 - J, K, nz - 1 - K)):
        if has:
            d = np.minimum(d, dist)
    colour = ((I & 1) << 2) | ((J & 1) << 1) | (K & 1)
    # 27-point stencil with random off-diagonal weights and a dominant diagonal
    nbrs = []
    for di, dj, dk in itertools.product((-1, 0, 1), repeat=3):
        if (di, dj, dk) == (0, 0, 0):
            continue
        ii, jj, kk = I + di, J + dj, K + dk
        ok = (ii >= 0) & (ii < ny) & (jj >= 0) & (jj < nx) & (kk >= 0) & (kk < nz)
        tgt = np.where(ok, idx[np.clip(ii, 0, ny - 1), np.clip(jj, 0, nx - 1), np.clip(kk, 0, nz - 1)], -1)
        nbrs.append((tgt.reshape(-1), rng.uniform(-1, 1, ny * nx * nz)))
    diag = 30.0 + rng.uniform(0, 1, ny * nx * nz)
    f = rng.uniform(-1, 1, ny * nx * nz)
    return d.reshape(-1), colour.reshape(-1), nbrs, diag, f


def row(nodes, x, nbrs):
    """sum over the neighbours of `nodes`, in a fixed neighbour order (what one GPU thread does)."""
    r = np.zeros(len(nodes))
    for tgt, w in nbrs:
        t = tgt[nodes]
        r = r + np.where(t >= 0, w[nodes] * x[np.maximum(t, 0)], 0.0)
    return r


def relax(nodes, x, nbrs, diag, f):
    """all `nodes` are of one colour: independent of each other, updated from the current x"""
    x[nodes] = x[nodes] + (f[nodes] - diag[nodes] * x[nodes] - row(nodes, x, nbrs)) / diag[nodes]


def face_update(shared, rows, x, diag, f):
    x[shared] = x[shared] + 0.5 * (f[shared] - diag[shared] * x[shared] - rows) / diag[shared]


def sweep_serial(x, d, colour, nbrs, diag, f):
    shared = np.flatnonzero(d == 0)
    rows = row(shared, x, nbrs)
    face_update(shared, rows, x, diag, f)
    for c in range(7, -1, -1):
        relax(np.flatnonzero((colour == c) & (d >= 1)), x, nbrs, diag, f)


def sweep_overlapped(x, d, colour, nbrs, diag, f, rows_late):
    shared = np.flatnonzero(d == 0)
    rows = None if rows_late else row(shared, x, nbrs)
    for c in range(7, -1, -1):
        p = 7 - c
        relax(np.flatnonzero((colour == c) & (d >= p + 2)), x, nbrs, diag, f)          # far: CCU_D_RANGE(p + 2, 15)
    if rows_late:
        rows = row(shared, x, nbrs)             # the partial-row kernel may run at any time during the far passes
    face_update(shared, rows, x, diag, f)
    for c in range(7, -1, -1):
        p = 7 - c
        relax(np.flatnonzero((colour == c) & (d >= 1) & (d <= p + 1)), x, nbrs, diag, f)  # near: CCU_D_RANGE(1, p + 1)


@pytest.mark.parametrize("faces", [(0, 1, 0, 0, 0, 0), (1, 0, 1, 0, 0, 1), (0, 1, 0, 1, 0, 1), (1, 1, 1, 1, 1, 1), (0, 0, 0, 0, 0, 0)],
                         ids=["one_face", "corner_lo_lo_hi", "corner_hi_hi_hi", "inner_block", "no_neighbours"])
@pytest.mark.parametrize("rows_late", [False, True])
def test_overlapped_sweep_is_the_serial_sweep(faces, rows_late):
    n = (23, 21, 25)                            # more than 2 x 9 nodes per direction: every far window is non-empty
    d, colour, nbrs, diag, f = build(n, faces, seed=11)
    xs = np.random.default_rng(5).uniform(-1, 1, d.size)
    xo = xs.copy()
    for _ in range(3):
        sweep_serial(xs, d, colour, nbrs, diag, f)
        sweep_overlapped(xo, d, colour, nbrs, diag, f, rows_late)
    assert np.array_equal(xs, xo)
    # and the windows matter: shifting the far window by one node breaks the equality when there are duplicated nodes
    if any(faces):
        xb = np.random.default_rng(5).uniform(-1, 1, d.size)
        shared = np.flatnonzero(d == 0)
        rows = row(shared, xb, nbrs)
        for c in range(7, -1, -1):
            relax(np.flatnonzero((colour == c) & (d >= (7 - c) + 1)), xb, nbrs, diag, f)     # one node too close
        face_update(shared, rows, xb, diag, f)
        for c in range(7, -1, -1):
            relax(np.flatnonzero((colour == c) & (d >= 1) & (d <= 7 - c)), xb, nbrs, diag, f)
        xr = np.random.default_rng(5).uniform(-1, 1, d.size)
        sweep_serial(xr, d, colour, nbrs, diag, f)
        assert not np.array_equal(xb, xr)
