// ccu_col_index.h -- index arithmetic of the column-resident smoother / matvec (ccu_col.cuh).
//
// Plain C++ (no CUDA types) so that the very same functions drive the device kernels, the re-layout kernel that
// builds their stiffness copy, and the host emulation in tests/col_emul.cpp that checks both against the oracle
// without a GPU.
//
// A COLUMN is TI (y) x TJ (x) nodes spanning all z layers of a multigrid level; columns at the upper mesh edge are
// clipped to ti x tj.  The stiffness copy `Kc` holds, per column and per z layer k = -1 .. noz (the two out-of-mesh
// layers are all-zero chunks), one contiguous CHUNK laid out exactly as the kernel wants it in shared memory:
//
//     [ BI      : 3 x nt doubles  ]   inverse diagonal (0 for nodes the sweep must skip), nt = ti*tj, p = li*tj + lj
//     [ blocks  : (14 nt + nh) x 9 floats ]   AoS 3x3 blocks: block id = p*14 + slot for the node's own half-matrix row
//                                            (slot 0 self, slot b+1 = lower neighbour CCU_LO[b]), then nh HALO blocks
//     [ flags   : nt bytes ]           boundary-condition flag byte per node (matvec strip)
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
    int kofs, flofs, cb;    // byte offsets of the block array and the flags inside a chunk, chunk bytes (multiple of 16)
};
__host__ __device__ inline CcuColDims ccu_col_dims(int ti, int tj)
{
    CcuColDims d;
    d.ti = ti; d.tj = tj; d.nt = ti * tj; d.nh = 9 * ti + 9 * tj;
    d.kofs = 24 * d.nt;
    d.flofs = d.kofs + 36 * (14 * d.nt + d.nh);
    d.cb = (d.flofs + d.nt + 15) & ~15;
    return d;
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

// What lane (d, q) of the warp that relaxes node (li, lj) reads for the stencil direction (q / 3 - 1, q % 3 - 1, t - 1):
//   kof  byte offset, inside the chunk of the layer that stores the block, of the first of its three coefficients
//   tr   0: the block is the node's own (row d: the coefficients are 4 bytes apart, chunk of the node's layer)
//        1: the block is stored at the upper neighbour (column d of it: 12 bytes apart, chunk of the neighbour's layer)
//   xof  byte offset of the neighbour inside one dof plane of a layer of the solution window (box of (TI+2) x (TJ+2))
struct CcuColDesc { int kof, xof, tr; };
__host__ __device__ inline CcuColDesc ccu_col_desc(const CcuColDims &cd, int TJ, int li, int lj, int d, int q, int t)
{
    const int di = q / 3 - 1, dj = q % 3 - 1, dk = t - 1;
    CcuColDesc r;
    r.xof = 8 * ((li + di + 1) * (TJ + 2) + (lj + dj + 1));
    const bool self = di == 0 && dj == 0 && dk == 0;
    const int lo = ccu_lo_index(di, dj, dk);
    if(self || lo >= 0)
    {
        const int id = (li * cd.tj + lj) * 14 + (self ? 0 : lo + 1);
        r.tr = 0;
        r.kof = cd.kofs + 4 * (9 * id + 3 * d);
        return r;
    }
    const int b = ccu_lo_index(-di, -dj, -dk);                  // the upper neighbour sees this node at CCU_LO[b]
    const int sli = li + di, slj = lj + dj;
    int id;
    if(sli >= 0 && sli < cd.ti && slj >= 0 && slj < cd.tj) id = (sli * cd.tj + slj) * 14 + b + 1;
    else id = 14 * cd.nt + ccu_col_halo_id(cd.ti, cd.tj, sli, slj, b);
    r.tr = 1;
    r.kof = cd.kofs + 4 * (9 * id + d);
    return r;
}

// One chunk of the column stiffness copy: layer k (-1 .. noz) of the column at (i0, j0), filled from the level's
// coefficient-major arrays.  Work items first, first + stride, ... (a device thread group or a host loop).
// `bits` (multi-subdomain runs, else null): nodes duplicated on a neighbouring subdomain (bit 1) get BI = 0 in the chunk --
// the sweep leaves them alone, they are relaxed from their summed rows (ccu_k_face_update).
__host__ __device__ inline void ccu_col_fill_chunk(const CcuGeom &g, const CcuColDims &cd, int i0, int j0, int k, const float *K,
                                                   const double *BI, const unsigned char *flags, const unsigned char *bits,
                                                   unsigned char *chunk, int first, int stride)
{
    const bool inz = k >= 0 && k < g.noz;
    const size_t NS = (size_t)g.NS;
    for(int w = first; w < 3 * cd.nt; w += stride)
    {
        const int dd = w / cd.nt, p = w % cd.nt;
        double v = 0.0;
        if(inz)
        {
            const int s = ccu_sidx(g, i0 + p / cd.tj, j0 + p % cd.tj, k);
            v = BI[dd * NS + s];
            if(bits && (bits[s] & 2)) v = 0.0;
        }
        ((double *)chunk)[w] = v;
    }
    float *kb = (float *)(chunk + cd.kofs);
    for(int id = first; id < 14 * cd.nt + cd.nh; id += stride)
    {
        int sli, slj, slot;
        if(id < 14 * cd.nt) { const int p = id / 14; slot = id % 14; sli = p / cd.tj; slj = p % cd.tj; }
        else { int b; ccu_col_halo_decode(cd.ti, cd.tj, id - 14 * cd.nt, sli, slj, b); slot = b + 1; }
        const int gi = i0 + sli, gj = j0 + slj;
        const bool in = inz && gi >= 0 && gi < g.noy && gj >= 0 && gj < g.nox;
        const int s = in ? ccu_sidx(g, gi, gj, k) : 0;
        for(int e = 0; e < 9; e++) kb[9 * id + e] = in ? K[(size_t)(slot * 9 + e) * NS + s] : 0.0f;
    }
    for(int p = first; p < cd.cb - cd.flofs; p += stride)
        chunk[cd.flofs + p] = (inz && p < cd.nt) ? flags[ccu_sidx(g, i0 + p / cd.tj, j0 + p % cd.tj, k)] : (unsigned char)0;
}
