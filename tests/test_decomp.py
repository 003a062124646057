"""CPU tests of the subdomain-per-GPU host logic: the duplicated-node tables ccu_comm_init builds
(pure index arithmetic inside the CUDA library, exercised here without a GPU), single-process and with two
real ranks over torch.distributed/gloo (world_size 2) emulating the NCCL send/recv round."""
import os
import socket

import numpy as np
import pytest

from citcomcu_b200 import _lib, decomp


@pytest.fixture(scope="module", autouse=True)
def built():
    _lib.build_library()


def all_ranks(nproc):
    return [decomp.me_loc_of(r, nproc) for r in range(nproc[0] * nproc[1] * nproc[2])]


def test_rank_numbering_roundtrip():
    nproc = (2, 3, 2)
    for r, me in enumerate(all_ranks(nproc)):
        assert decomp.rank_of(me, nproc) == r          # rank = z + nprocz*x + nprocz*nprocx*y (Parallel_related.c:108-121)


@pytest.mark.parametrize("nproc,dims", [((2, 1, 1), (5, 3, 3)), ((2, 2, 1), (5, 5, 3)), ((2, 2, 2), (5, 5, 3)), ((1, 3, 2), (3, 5, 9))])
def test_halo_sum_emulated(nproc, dims):
    """Emulate the exchange with numpy: every duplicated node must end up with the sum over ALL its owners, added in
    ascending rank order (bitwise identical on every owner); ownership masks must partition the global nodes."""
    nox, noy, noz = dims
    ranks = all_ranks(nproc)
    T = [decomp.halo_tables(nproc, me, nox, noy, noz) for me in ranks]
    G = [decomp.global_node_ids(nproc, me, nox, noy, noz) for me in ranks]
    rng = np.random.default_rng(3)
    vals = [rng.standard_normal(nox * noy * noz) for _ in ranks]
    nglob = max(g.max() for g in G) + 1
    # reference result: per global node, contributions added in ascending rank order
    expect = np.zeros(nglob)
    seen = np.zeros(nglob, dtype=bool)
    for r in range(len(ranks)):
        first = ~seen[G[r]]
        expect[G[r][first]] = vals[r][first]
        expect[G[r][~first]] = expect[G[r][~first]] + vals[r][~first]
        seen[G[r]] = True
    owned_count = np.zeros(nglob, dtype=int)
    for r, t in enumerate(T):
        np.add.at(owned_count, G[r][t["owned"] == 1], 1)
        # "receive": the segment for neighbour q holds what q packed for us
        recv = np.zeros(len(t["send_n"]))
        for q, (nb, off, cnt) in enumerate(zip(t["nb_rank"], t["nb_off"], t["nb_cnt"])):
            tq = T[nb]
            back = list(tq["nb_rank"]).index(r)
            o2, c2 = tq["nb_off"][back], tq["nb_cnt"][back]
            assert c2 == cnt
            sent_nodes = tq["send_n"][o2:o2 + c2]
            # both sides enumerate the same physical nodes in the same order
            assert np.array_equal(G[nb][sent_nodes], G[r][t["send_n"][off:off + cnt]])
            recv[off:off + cnt] = vals[nb][sent_nodes]
        out = vals[r].copy()
        for i, n in enumerate(t["sh_n"]):
            acc = None
            for e in range(t["sh_ptr"][i], t["sh_ptr"][i + 1]):
                v = vals[r][n] if t["sh_src"][e] < 0 else recv[t["sh_src"][e]]
                acc = v if acc is None else acc + v
            out[n] = acc
        assert np.array_equal(out, expect[G[r]])        # bitwise, also on the nodes that are not duplicated
    assert np.all(owned_count == 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nproc, (nox, noy, noz) = (1, 1, 2), (5, 3, 4)
        me = decomp.me_loc_of(rank, nproc)
        t = decomp.halo_tables(nproc, me, nox, noy, noz)
        val = torch.arange(nox * noy * noz, dtype=torch.float64) * (rank + 1) + 0.25 * rank
        send = val[torch.from_numpy(t["send_n"].astype(np.int64))].contiguous()
        recv = torch.zeros_like(send)
        ops = []
        for nb, off, cnt in zip(t["nb_rank"], t["nb_off"], t["nb_cnt"]):
            ops.append(dist.P2POp(dist.isend, send[off:off + cnt], int(nb)))
            ops.append(dist.P2POp(dist.irecv, recv[off:off + cnt], int(nb)))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        out = val.clone()
        for i, n in enumerate(t["sh_n"]):
            acc = None
            for e in range(t["sh_ptr"][i], t["sh_ptr"][i + 1]):
                v = val[n] if t["sh_src"][e] < 0 else recv[t["sh_src"][e]]
                acc = v if acc is None else acc + v
            out[n] = acc
        # masked global dot product (global_vdot): allreduce of the owned parts == dot over the unique global nodes
        own = torch.from_numpy(t["owned"].astype(np.float64))
        local = (out * out * own).sum().reshape(1)
        dist.all_reduce(local)
        gids = decomp.global_node_ids(nproc, me, nox, noy, noz)
        allv = [torch.zeros_like(out) for _ in range(world)]
        dist.all_gather(allv, out)
        q.put((rank, float(local[0]), out.numpy(), gids, [a.numpy() for a in allv]))
    finally:
        dist.destroy_process_group()


def test_halo_sum_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, dot0, out0, g0, _), (r1, dot1, out1, g1, _) = res
    assert dot0 == dot1
    common, i0, i1 = np.intersect1d(g0, g1, return_indices=True)
    assert len(common) == 5 * 3                                  # one shared z-face of nox*noy nodes
    assert np.array_equal(out0[i0], out1[i1])                    # duplicated nodes agree bitwise after the halo sum
    uniq = {}
    for g, o in ((g0, out0), (g1, out1)):
        for gi, v in zip(g, o):
            uniq[gi] = v
    assert abs(dot0 - sum(v * v for v in uniq.values())) < 1e-9 * dot0


@pytest.mark.parametrize("nproc", [(2, 1, 1), (2, 2, 2), (3, 2, 2), (1, 3, 1)])
def test_marker_exchange_routes(nproc):
    """Routing table of the marker exchange (ccu_marker_routes, the host half of ccu_marker_exchange): every record a
    subdomain sends under direction code c = (ox+1) + 3 (oy+1) + 9 (oz+1) is expected by exactly the neighbour at that
    offset, under the opposite code; nothing is expected from outside the processor grid."""
    import ctypes as C
    lib = C.CDLL(str(_lib.LIB_PATH))
    ranks = all_ranks(nproc)
    n = len(ranks)
    rng = np.random.default_rng(11)
    send = np.zeros((n, 27), dtype=np.int32)
    for r, me in enumerate(ranks):
        for code in range(27):
            o = (code % 3 - 1, (code // 3) % 3 - 1, code // 9 - 1)
            tgt = tuple(me[d] + o[d] for d in range(3))
            if code != 13 and all(0 <= tgt[d] < nproc[d] for d in range(3)):
                send[r, code] = rng.integers(0, 50)
    total_recv = 0
    for r, me in enumerate(ranks):
        nb, rc = (C.c_int * 27)(), (C.c_int * 27)()
        assert lib.ccu_marker_routes((C.c_int * 3)(*nproc), (C.c_int * 3)(*me), send.ctypes.data_as(C.c_void_p), nb, rc) == 0
        for code in range(27):
            o = (code % 3 - 1, (code // 3) % 3 - 1, code // 9 - 1)
            src = tuple(me[d] + o[d] for d in range(3))
            inside = code != 13 and all(0 <= src[d] < nproc[d] for d in range(3))
            if not inside:
                assert nb[code] == -1 and rc[code] == 0
                continue
            rs = decomp.rank_of(src, nproc)
            assert nb[code] == rs
            back = sum((-o[d] + 1) * 3 ** d for d in range(3))         # the sender's code of the offset that points here
            assert rc[code] == send[rs, back]
            total_recv += rc[code]
    assert total_recv == int(send.sum())


@pytest.mark.parametrize("nproc", [(2, 1, 1), (1, 2, 1), (1, 1, 2), (2, 2, 1), (2, 2, 2), (3, 2, 2)])
def test_duplicated_nodes_are_the_faces_with_a_neighbour(nproc):
    """ccu_comm_init derives the distance-to-duplicated-node byte of the overlapped sweep from the subdomain faces that have a
    neighbour (y faces <-> i, x faces <-> j, z faces <-> k) and refuses to start when that set differs from the duplicated-node
    table: the two agree for every rank of every processor grid bench.py and the tests use."""
    nox, noy, noz = 5, 7, 4
    for me in all_ranks(nproc):
        t = decomp.halo_tables(nproc, me, nox, noy, noz)
        shared = np.zeros(nox * noy * noz, dtype=bool)
        shared[t["sh_n"]] = True
        i, j, k = np.meshgrid(np.arange(noy), np.arange(nox), np.arange(noz), indexing="ij")
        d = np.full(i.shape, 15)
        for has, dist in ((me[1] > 0, i), (me[1] < nproc[1] - 1, noy - 1 - i), (me[0] > 0, j), (me[0] < nproc[0] - 1, nox - 1 - j),
                          (me[2] > 0, k), (me[2] < nproc[2] - 1, noz - 1 - k)):
            if has:
                d = np.minimum(d, dist)
        assert np.array_equal((d == 0).reshape(-1), shared), me
