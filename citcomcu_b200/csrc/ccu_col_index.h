// ccu_col_index.h -- index arithmetic of the column-resident smoother / matvec (ccu_col.cuh).
//
// Plain C++ (no CUDA types) so that the very same functions drive the device kernels, the re-layout kernel that
// builds their stiffness copy, and the host emulation in tests/col_emul.cpp that checks both against the oracle
// without a GPU.
//
// A COLUMN is TI (y) x TJ (x) nodes spanning all z layers of a multigrid level; columns at the upper mesh edge are
// clipped to ti x tj.  The stiffness copy `Kc` holds, per column and per z layer k = 0 .. noz-1, one contiguous CHUNK
// laid out exactly as the kernel wants it in shared memory:
//
//     [ BI      : 3 x nt doubles  ]   inverse diagonal (0 for nodes the sweep must skip), nt = ti*tj, p = li*tj + lj
//     [ blocks  : nb = 14 nt + nh 3x3 blocks ]  block id = p*14 + slot for the half-matrix row of the node p (slot 0
//                                            self, slot b+1 = lower neighbour CCU_LO[b]), then nh HALO blocks; the nine
//                                            coefficients e = 3 a + bb of a block sit in three arrays so that a lane reads
//                                            a whole block with two 128-bit and one 32-bit shared-memory loads:
//                                            float4 A[nb] (e = 0..3), float4 B[nb] (e = 4..7), float C[nb] (e = 8)
//     [ flags   : nt bytes ]           boundary-condition flag byte per node (matvec strip)
//
// Which LAYER a block of chunk k comes from: the blocks that couple two nodes of the same layer (self and the four
// in-plane lower neighbours) and those that reach UP (dz = +1) are the ones stored at the nodes of layer k; the blocks
// that reach DOWN (dz = -1) are those stored at the nodes of layer k+1.  So chunk k = "layer k in-plane" + "everything
// between layers k and k+1", and the rows of layer k need exactly the chunks k-1 and k (not three layers of blocks):
// a ring of three chunks in shared memory holds two in use and one in flight.
//
// HALO blocks are the blocks stored at nodes OUTSIDE the column (its upper neighbours in y and x) that couple to
// nodes inside it; the row product of a boundary node needs them transposed.  They are duplicated into the chunk of
// the layer of the node that stores them (nh = 9 ti + 9 tj per layer, ~24 % of a chunk for 8 x 4 columns), which is what
// lets a CTA stream a whole column with one bulk copy per layer and touch every stiffness byte once per sweep.
#pragma once

#if !defined(__CUDACC__) && !defined(__host__)
#define __host__
#define __device__
#endif
#include <cstddef>
#include "ccu_layout.cuh"

struct CcuColDims
{
    int ti, tj, nt, nh;     // clipped column extent, nodes per layer, halo blocks per layer
    int nb;                 // blocks per chunk
    int kofs, flofs, cb;    // byte offsets of the block arrays and the flags inside a chunk, chunk bytes (multiple of 16)
};
__host__ __device__ constexpr inline CcuColDims ccu_col_dims(int ti, int tj)
{
    CcuColDims d = { ti, tj, ti * tj, 9 * ti + 9 * tj, 0, 0, 0, 0 };
    d.nb = 14 * d.nt + d.nh;
    d.kofs = (24 * d.nt + 15) & ~15;
    d.flofs = d.kofs + 36 * d.nb;
    d.cb = (d.flofs + d.nt + 15) & ~15;
    return d;
}
// byte offset inside a chunk of coefficient e (= 3 a + bb) of block `id`
__host__ __device__ inline int ccu_col_coef_ofs(const CcuColDims &cd, int id, int e)
{
    if(e < 4) return cd.kofs + 16 * id + 4 * e;
    if(e < 8) return cd.kofs + 16 * cd.nb + 16 * id + 4 * (e - 4);
    return cd.kofs + 32 * cd.nb + 4 * id;
}

// index of a lower-neighbour offset (dy, dx, dz) in CCU_LO (ccu_layout.cuh), -1 if the offset is not a lower neighbour
__host__ __device__ inline int ccu_lo_index(int di, int dj, int dk)
{
    if(di == -1) return 3 * (dj + 1) + (dk + 1);
    if(di == 0 && dj == -1) return 9 + (dk + 1);
    if(di == 0 && dj == 0 && dk == -1) return 12;
    return -1;
}
__host__ __device__ inline void ccu_lo_offset(int b, int &di, int &dj, int &dk)
{
    if(b < 9) { di = -1; dj = b / 3 - 1; dk = b % 3 - 1; }
    else if(b < 12) { di = 0; dj = -1; dk = b - 10; }
    else { di = 0; dj = 0; dk = -1; }
}

// Halo block index h (0 .. nh-1) of the block stored at the out-of-column node (sli, slj) [local y, x] in its slot b+1
// (lower-neighbour offset CCU_LO[b]); -1 when that block does not couple to a node of the column.
//   row sli == ti          : slj = -1 (b = 6..8), slj = 0..tj-1 (b = 0..8), slj = tj (b = 0..2)          -> 9 tj + 6
//   face slj == -1         : sli = 1..ti-1 (b = 6..8)                                                   -> 3 (ti - 1)
//   face slj == tj         : sli = 0 (b = 9..11), sli = 1..ti-1 (b = 0..2 and 9..11)                    -> 6 ti - 3
__host__ __device__ inline int ccu_col_halo_id(int ti, int tj, int sli, int slj, int b)
{
    int di, dj, dk;
    ccu_lo_offset(b, di, dj, dk);
    const int li = sli + di, lj = slj + dj;                    // the node the block couples to
    if(li < 0 || li >= ti || lj < 0 || lj >= tj) return -1;
    if(sli >= 0 && sli < ti && slj >= 0 && slj < tj) return -1; // stored inside the column: an own block, not a halo block
    if(sli == ti)
    {
        if(slj == -1) return b - 6;
        if(slj < tj) return 3 + 9 * slj + b;
        return 3 + 9 * tj + b;
    }
    if(slj == -1) return 9 * tj + 6 + 3 * (sli - 1) + (b - 6);
    const int base = 9 * tj + 3 * ti + 3;                      // slj == tj
    if(sli == 0) return base + (b - 9);
    return base + 3 + 6 * (sli - 1) + (b < 3 ? b : b - 6);
}
// inverse of ccu_col_halo_id
__host__ __device__ inline void ccu_col_halo_decode(int ti, int tj, int h, int &sli, int &slj, int &b)
{
    if(h < 9 * tj + 6)
    {
        sli = ti;
        if(h < 3) { slj = -1; b = h + 6; return; }
        h -= 3;
        if(h < 9 * tj) { slj = h / 9; b = h % 9; return; }
        slj = tj; b = h - 9 * tj; return;
    }
    h -= 9 * tj + 6;
    if(h < 3 * (ti - 1)) { slj = -1; sli = 1 + h / 3; b = 6 + h % 3; return; }
    h -= 3 * (ti - 1);
    slj = tj;
    if(h < 3) { sli = 0; b = 9 + h; return; }
    h -= 3;
    sli = 1 + h / 6;
    const int r = h % 6;
    b = r < 3 ? r : r + 6;
}

// What lane q (0..8) of the nine lanes that relax node (li, lj) of layer k reads for the stencil direction
// (di, dj, dk) = (q / 3 - 1, q % 3 - 1, t - 1):
//   id   the block (ccu_col_coef_ofs), inside chunk k - 1 when dk = -1 and inside chunk k otherwise
//   tr   0: the block is the node's own (rows = the node's dofs)
//        1: the block is stored at the upper neighbour (rows = the neighbour's dofs: use it transposed)
//   xof  byte offset of the neighbour inside one dof plane of a layer of the solution window (box of (TI+2) x (TJ+2))
struct CcuColDesc { int id, xof, tr; };
__host__ __device__ inline CcuColDesc ccu_col_desc(const CcuColDims &cd, int TJ, int li, int lj, int q, int t)
{
    const int di = q / 3 - 1, dj = q % 3 - 1, dk = t - 1;
    CcuColDesc r;
    r.xof = 8 * ((li + di + 1) * (TJ + 2) + (lj + dj + 1));
    const bool self = di == 0 && dj == 0 && dk == 0;
    const int lo = ccu_lo_index(di, dj, dk);
    if(self || lo >= 0)
    {
        r.id = (li * cd.tj + lj) * 14 + (self ? 0 : lo + 1);
        r.tr = 0;
        return r;
    }
    const int b = ccu_lo_index(-di, -dj, -dk);                  // the upper neighbour sees this node at CCU_LO[b]
    const int sli = li + di, slj = lj + dj;
    if(sli >= 0 && sli < cd.ti && slj >= 0 && slj < cd.tj) r.id = (sli * cd.tj + slj) * 14 + b + 1;
    else r.id = 14 * cd.nt + ccu_col_halo_id(cd.ti, cd.tj, sli, slj, b);
    r.tr = 1;
    return r;
}

// One chunk of the column stiffness copy: chunk k (0 .. noz-1) of the column at (i0, j0), filled from the level's
// coefficient-major arrays.  Work items first, first + stride, ... (a device thread group or a host loop).
// `bits` (multi-subdomain runs, else null): nodes duplicated on a neighbouring subdomain (bit 1) get BI = 0 in the chunk --
// the sweep leaves them alone, they are relaxed from their summed rows (ccu_k_face_update).
__host__ __device__ inline void ccu_col_fill_chunk(const CcuGeom &g, const CcuColDims &cd, int i0, int j0, int k, const float *K,
                                                   const double *BI, const unsigned char *flags, const unsigned char *bits,
                                                   unsigned char *chunk, int first, int stride)
{
    const size_t NS = (size_t)g.NS;
    for(int w = first; w < cd.kofs / 8; w += stride)
    {
        double v = 0.0;
        if(w < 3 * cd.nt)
        {
            const int dd = w / cd.nt, p = w % cd.nt;
            const int s = ccu_sidx(g, i0 + p / cd.tj, j0 + p % cd.tj, k);
            v = BI[dd * NS + s];
            if(bits && (bits[s] & 2)) v = 0.0;
        }
        ((double *)chunk)[w] = v;
    }
    for(int id = first; id < cd.nb; id += stride)
    {
        int sli, slj, slot;
        if(id < 14 * cd.nt) { const int p = id / 14; slot = id % 14; sli = p / cd.tj; slj = p % cd.tj; }
        else { int b; ccu_col_halo_decode(cd.ti, cd.tj, id - 14 * cd.nt, sli, slj, b); slot = b + 1; }
        int di = 0, dj = 0, dk = 0;
        if(slot) ccu_lo_offset(slot - 1, di, dj, dk);
        const int ks = dk < 0 ? k + 1 : k;                     // blocks that reach down belong to the layer above
        const int gi = i0 + sli, gj = j0 + slj;
        const bool in = ks < g.noz && gi >= 0 && gi < g.noy && gj >= 0 && gj < g.nox;
        const int s = in ? ccu_sidx(g, gi, gj, ks) : 0;
        for(int e = 0; e < 9; e++) *(float *)(chunk + ccu_col_coef_ofs(cd, id, e)) = in ? K[(size_t)(slot * 9 + e) * NS + s] : 0.0f;
    }
    for(int p = first; p < cd.cb - cd.flofs; p += stride)
        chunk[cd.flofs + p] = p < cd.nt ? flags[ccu_sidx(g, i0 + p / cd.tj, j0 + p % cd.tj, k)] : (unsigned char)0;
}
