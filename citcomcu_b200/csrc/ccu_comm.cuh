// ccu_comm.cuh -- subdomain-per-GPU communication layer.
//
// The mesh is partitioned exactly as the reference does it (Parallel_related.c:80-173): an
// nprocx x nprocy x nprocz grid of blocks, rank = z + nprocz*x + nprocz*nprocx*y, neighbouring blocks
// DUPLICATE their common face nodes.  Two exchange primitives replace the reference's MPI layer:
//
//   halo sum   (exchange_id_d20 / exchange_node_f20, Parallel_related.c:1181,1270): every duplicated node ends up
//              with the sum of the partial values all its owners hold.  The reference runs three sequential
//              per-dimension passes; here ONE grouped ncclSend/ncclRecv round goes to all (<= 26) face, edge and
//              corner neighbours and the unpack kernel adds the contributions in ascending rank order, the local one
//              included at its rank's position -- so every owner of a node computes the bitwise identical sum.
//   allreduce  (global_vdot / global_pdot, Global_operations.c:339,359): ownership-masked local dot + ncclAllReduce
//              of up to three scalars at a time.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch ships or the system one) so the library
// loads, and its single-GPU path runs, on machines without NCCL.
#pragma once
#include <vector>

// host description of one level's duplicated-node tables (pure index arithmetic, no GPU needed)
struct CcuHaloHost
{
    std::vector<int> nb_rank, nb_off, nb_cnt;   // neighbour segments, ascending rank; offsets/counts in nodes
    std::vector<int> send_t;                    // [n_send] compact shared-node index of every packed node, by segment
    std::vector<int> sh_n;                      // [n_shared] natural node index n = k + noz*(j + nox*i)
    std::vector<int> sh_ptr;                    // [n_shared+1] CSR into sh_src
    std::vector<int> sh_src;                    // per contribution in ascending rank order: -1 = local value, else node offset in the receive buffer
    std::vector<unsigned char> owned;           // [nno] 1 if this rank owns the node for dot products (lowest rank among its owners)
};

// nproc = (x, y, z) block counts, me = this block's (x, y, z) position
void ccu_build_halo_host(const int nproc[3], const int me[3], int nox, int noy, int noz, CcuHaloHost &h);
static inline int ccu_rank_of(const int nproc[3], int x, int y, int z) { return z + nproc[2] * x + nproc[2] * nproc[0] * y; }
