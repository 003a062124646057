"""Block domain decomposition of the reference (Parallel_related.c:80-173) and the duplicated-node
tables the CUDA library derives from it (csrc/ccu_comm.cu), for one process per GPU.

rank = z + nprocz*x + nprocz*nprocx*y  (Parallel_related.c:108-121); `me_loc` = (x, y, z).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def rank_of(me_loc, nproc) -> int:
    x, y, z = me_loc
    return z + nproc[2] * x + nproc[2] * nproc[0] * y


def me_loc_of(rank: int, nproc):
    z = rank % nproc[2]
    x = (rank // nproc[2]) % nproc[0]
    y = rank // (nproc[2] * nproc[0])
    return (x, y, z)


def nproc_for(n_ranks: int, mgunit=(8, 8, 4)):
    """Processor grid for `n_ranks` GPUs the way SURVEY.md 8d lays out config 3: 1 -> 1x1x1, 2 -> 2x1x1,
    4 -> 2x2x1, 8 -> 2x2x2; every factor must divide mgunit (README:149-151)."""
    grid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n_ranks)
    if grid is None:
        raise ValueError(f"no processor grid defined for {n_ranks} ranks")
    for m, n in zip(mgunit, grid):
        if m % n:
            raise ValueError("mgunit must be divisible by nproc in each direction")
    return grid


def halo_tables(nproc, me_loc, nox, noy, noz):
    """The duplicated-node tables of one level exactly as ccu_comm_init builds them (host code of the CUDA
    library, no GPU involved).  Returns a dict of numpy arrays:
      nb_rank/nb_off/nb_cnt  neighbour segments in ascending rank order (offsets / counts in nodes)
      send_n                 natural index of every packed node, by segment
      sh_n                   natural index of every duplicated node
      sh_ptr, sh_src         per duplicated node, its owners' contributions in ascending rank order:
                             -1 = the local value, else node offset into the concatenated receive buffer
      owned                  [nno] 1 where this rank counts the node in global dot products
    """
    lib = _lib.lib()
    np3 = (C.c_int * 3)(*nproc)
    me3 = (C.c_int * 3)(*me_loc)
    sizes = (C.c_int * 4)()
    _lib.check(lib.ccu_halo_sizes(np3, me3, nox, noy, noz, sizes))
    n_nb, n_send, n_sh, n_ent = list(sizes)
    i32 = lambda n: np.zeros(max(n, 1), dtype=np.int32)  # noqa: E731
    t = dict(nb_rank=i32(n_nb), nb_off=i32(n_nb), nb_cnt=i32(n_nb), send_n=i32(n_send), sh_n=i32(n_sh),
             sh_ptr=i32(n_sh + 1), sh_src=i32(n_ent), owned=np.zeros(nox * noy * noz, dtype=np.uint8))
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _lib.check(lib.ccu_halo_tables(np3, me3, nox, noy, noz, p(t["nb_rank"]), p(t["nb_off"]), p(t["nb_cnt"]), p(t["send_n"]),
                                   p(t["sh_n"]), p(t["sh_ptr"]), p(t["sh_src"]), p(t["owned"])))
    for k, n in (("nb_rank", n_nb), ("nb_off", n_nb), ("nb_cnt", n_nb), ("send_n", n_send), ("sh_n", n_sh), ("sh_src", n_ent)):
        t[k] = t[k][:n]
    return t


def global_node_ids(nproc, me_loc, nox, noy, noz):
    """Global natural index of every local node (local order n = k + noz*(j + nox*i))."""
    ex, ey, ez = nox - 1, noy - 1, noz - 1
    GX, GZ = ex * nproc[0] + 1, ez * nproc[2] + 1
    i = np.arange(noy)[:, None, None] + me_loc[1] * ey
    j = np.arange(nox)[None, :, None] + me_loc[0] * ex
    k = np.arange(noz)[None, None, :] + me_loc[2] * ez
    return (k + GZ * (j + GX * i)).reshape(-1)
