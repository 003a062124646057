// ccu_kernels.cuh -- sm_100a kernels of the Stokes hot path (velocity block).
// All are HBM-bound stencil / streaming kernels in the reference's mixed precision
// (fp32 coefficients, fp64 vectors and accumulation); no tensor-core work here.
#pragma once
#include "ccu_layout.cuh"
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#define CCU_VBX 0x2u
#define CCU_VBZ 0x4u
#define CCU_VBY 0x8u
// device flag byte per storage slot
#define CCU_F_VBX 1
#define CCU_F_VBY 2
#define CCU_F_VBZ 4
#define CCU_F_VALID 128

struct CcuLevelDev
{
    CcuGeom g;
    const float *K;            // [14*9][NS]
    const double *BI;          // [3][NS]
    const unsigned char *flags;// [NS]
    const float *MASS;         // [nno] natural order
    const float *TWW;          // [nel*8]
    const float *eco;          // [nel*3]
    const float *elt_del;      // [nel*24]
    const double *BPI;         // [npno]
};

// ------------------------------------------------------------------------------------------
// One row-block of K*x for a node of colour C: the 14 blocks the node owns (self + 13 lower
// neighbours) and the 13 transposed blocks owned by its upper neighbours.  Every load is a unit
// -stride, fully coalesced 128-byte warp access: colour-blocked layout makes "neighbour of my
// neighbour thread" = "my neighbour + 1".
// Replaces the gather+scatter loops of n_assemble_del2_u (Element_calculations.c:592-605) and
// the gather/scatter of gauss_seidel (General_matrix_functions.c:1241-1255) with a pure gather.
// ------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void ccu_row_product(const CcuGeom &g, const float *__restrict__ K, const double *x,
                                                const int cell, double &a0, double &a1, double &a2)
{
    constexpr int LO[13][3] = CCU_LO_INIT;
    constexpr int pi = (C >> 2) & 1, pj = (C >> 1) & 1, pk = C & 1;
    const size_t NS = (size_t)g.NS;
    const int s = C * g.NC + cell;
    double r0, r1, r2;
    {   // self block
        const double x0 = x[s], x1 = x[NS + s], x2 = x[2 * NS + s];
        const float *Kp = K + s;
        r0 = (double)__ldg(Kp) * x0 + (double)__ldg(Kp + NS) * x1 + (double)__ldg(Kp + 2 * NS) * x2;
        r1 = (double)__ldg(Kp + 3 * NS) * x0 + (double)__ldg(Kp + 4 * NS) * x1 + (double)__ldg(Kp + 5 * NS) * x2;
        r2 = (double)__ldg(Kp + 6 * NS) * x0 + (double)__ldg(Kp + 7 * NS) * x1 + (double)__ldg(Kp + 8 * NS) * x2;
    }
#pragma unroll
    for(int b = 0; b < 13; b++)
    {   // own blocks: lower neighbour m = n + LO[b], block stored at n in slot b+1
        constexpr int dummy = 0; (void)dummy;
        const int di = LO[b][0], dj = LO[b][1], dk = LO[b][2];
        const int cm = C ^ (((di != 0) << 2) | ((dj != 0) << 1) | (dk != 0));
        const int sm = cm * g.NC + cell + ccu_shift(pi, di) * g.JK + ccu_shift(pj, dj) * g.Kd + ccu_shift(pk, dk);
        const double x0 = x[sm], x1 = x[NS + sm], x2 = x[2 * NS + sm];
        const float *Kp = K + (size_t)((b + 1) * 9) * NS + s;
        r0 += (double)__ldg(Kp) * x0 + (double)__ldg(Kp + NS) * x1 + (double)__ldg(Kp + 2 * NS) * x2;
        r1 += (double)__ldg(Kp + 3 * NS) * x0 + (double)__ldg(Kp + 4 * NS) * x1 + (double)__ldg(Kp + 5 * NS) * x2;
        r2 += (double)__ldg(Kp + 6 * NS) * x0 + (double)__ldg(Kp + 7 * NS) * x1 + (double)__ldg(Kp + 8 * NS) * x2;
    }
#pragma unroll
    for(int b = 0; b < 13; b++)
    {   // transposed blocks: upper neighbour m = n - LO[b] owns K_mn in its slot b+1
        const int di = -LO[b][0], dj = -LO[b][1], dk = -LO[b][2];
        const int cm = C ^ (((di != 0) << 2) | ((dj != 0) << 1) | (dk != 0));
        const int sm = cm * g.NC + cell + ccu_shift(pi, di) * g.JK + ccu_shift(pj, dj) * g.Kd + ccu_shift(pk, dk);
        const double x0 = x[sm], x1 = x[NS + sm], x2 = x[2 * NS + sm];
        const float *Kp = K + (size_t)((b + 1) * 9) * NS + sm;
        r0 += (double)__ldg(Kp) * x0 + (double)__ldg(Kp + 3 * NS) * x1 + (double)__ldg(Kp + 6 * NS) * x2;
        r1 += (double)__ldg(Kp + NS) * x0 + (double)__ldg(Kp + 4 * NS) * x1 + (double)__ldg(Kp + 7 * NS) * x2;
        r2 += (double)__ldg(Kp + 2 * NS) * x0 + (double)__ldg(Kp + 5 * NS) * x1 + (double)__ldg(Kp + 8 * NS) * x2;
    }
    a0 = r0; a1 = r1; a2 = r2;
}

// ------------------------------------------------------------------------------------------
// The same row product split over T lanes per node (T = 4 or 32): lane q takes blocks q, q+T, ... of the 27.
// One thread per node walks 27 blocks x 12 loads as a chain of dependent round trips; on the coarse levels
// (too few nodes to hide that latency with occupancy) the chain IS the kernel time.  T lanes cut the chain T-fold;
// a warp of T=4 covers 8 consecutive cells so every 32-byte sector it touches is still fully used.
// Block numbering: 0 self, 1..13 own blocks (lower neighbour CCU_LO[b-1]), 14..26 the transposed blocks stored at
// the upper neighbours (-CCU_LO[b-14]).  The T partial sums are folded with an xor-shuffle tree (deterministic).
// ------------------------------------------------------------------------------------------
template <int T, int C>
__device__ __forceinline__ void ccu_row_product_lanes(const CcuGeom &g, const float *__restrict__ K, const double *x,
                                                      const int cell, const int q, const bool valid, double &a0, double &a1, double &a2)
{   // every lane of the warp must arrive here (the shuffles below name the full mask); `valid` gates the loads
    constexpr int pi = (C >> 2) & 1, pj = (C >> 1) & 1, pk = C & 1;
    const size_t NS = (size_t)g.NS;
    const int s = C * g.NC + cell;
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
    for(int it = 0; it < (27 + T - 1) / T; it++)
    {
        const int b = q + it * T;
        if(valid && b < 27)
        {
            const bool tr = b >= 14;
            const int t = tr ? b - 14 : b - 1;                 // index into CCU_LO, -1 for the self block
            int di = 0, dj = 0, dk = 0;
            if(t >= 0)
            {
                if(t < 9) { di = -1; dj = t / 3 - 1; dk = t % 3 - 1; }
                else if(t < 12) { dj = -1; dk = t - 10; }
                else dk = -1;
            }
            if(tr) { di = -di; dj = -dj; dk = -dk; }
            const int slot = (b == 0) ? 0 : t + 1;
            const int cm = C ^ (((di != 0) << 2) | ((dj != 0) << 1) | (dk != 0));
            const int sm = cm * g.NC + cell + ccu_shift(pi, di) * g.JK + ccu_shift(pj, dj) * g.Kd + ccu_shift(pk, dk);
            const float *Kp = K + (size_t)(slot * 9) * NS + (tr ? sm : s);
            float k[9];
#pragma unroll
            for(int e = 0; e < 9; e++) k[e] = __ldg(Kp + (size_t)e * NS);
            const double x0 = x[sm], x1 = x[NS + sm], x2 = x[2 * NS + sm];
            if(!tr)
            {
                r0 += (double)k[0] * x0 + (double)k[1] * x1 + (double)k[2] * x2;
                r1 += (double)k[3] * x0 + (double)k[4] * x1 + (double)k[5] * x2;
                r2 += (double)k[6] * x0 + (double)k[7] * x1 + (double)k[8] * x2;
            }
            else
            {
                r0 += (double)k[0] * x0 + (double)k[3] * x1 + (double)k[6] * x2;
                r1 += (double)k[1] * x0 + (double)k[4] * x1 + (double)k[7] * x2;
                r2 += (double)k[2] * x0 + (double)k[5] * x1 + (double)k[8] * x2;
            }
        }
    }
#pragma unroll
    for(int o = T / 2; o > 0; o >>= 1)
    {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    }
    a0 = r0; a1 = r1; a2 = r2;
}

// the per-node update of the smoother (General_matrix_functions.c:1250-1259): scalar BI per equation, correction
// rounded to fp32 (`higher_precision *temp`, :1172)
__device__ __forceinline__ void ccu_relax_update(const CcuGeom &g, const double *__restrict__ BI, const double *__restrict__ F,
                                                 double *x, const int s, const double a0, const double a1, const double a2)
{
    const size_t NS = (size_t)g.NS;
    const float t0 = (float)((F[s] - a0) * BI[s]);
    const float t1 = (float)((F[NS + s] - a1) * BI[NS + s]);
    const float t2 = (float)((F[2 * NS + s] - a2) * BI[2 * NS + s]);
    x[s] += (double)t0;
    x[NS + s] += (double)t1;
    x[2 * NS + s] += (double)t2;
}

// `bits` (multi-subdomain runs only, else null): bit 1 marks nodes duplicated on a neighbouring subdomain (the
// reference's OFFSIDE flag); those are relaxed separately from their summed rows (ccu_k_face_update), not here.
#define CCU_B_OWNED 1
#define CCU_B_SHARED 2
// bits 4-7: Chebyshev distance (in nodes, capped at 15) to the nearest duplicated node.  A colour pass takes the nodes whose distance
// lies in [drange & 0xff, drange >> 8]: CCU_D_ALL = every node that is not duplicated; the split passes of the overlapped sweep
// (d_relax_sweeps) take "far" and "near" shells separately.
#define CCU_D_RANGE(lo, hi) ((lo) | ((hi) << 8))
#define CCU_D_ALL CCU_D_RANGE(1, 15)
__device__ __forceinline__ bool ccu_skip_node(const unsigned char *__restrict__ bits, const int s, const int drange)
{
    if(!bits) return false;
    const int d = bits[s] >> 4;
    return d < (drange & 0xff) || d > (drange >> 8);
}
template <int C>
__device__ __forceinline__ void ccu_relax_cell(const CcuGeom &g, const float *__restrict__ K, const double *__restrict__ BI,
                                               const double *__restrict__ F, double *x, const int cell,
                                               const unsigned char *__restrict__ bits, const int drange)
{
    int i, j, k;
    if(!ccu_decode(g, C, cell, i, j, k)) return;
    if(ccu_skip_node(bits, C * g.NC + cell, drange)) return;
    double a0, a1, a2;
    ccu_row_product<C>(g, K, x, cell, a0, a1, a2);
    ccu_relax_update(g, BI, F, x, C * g.NC + cell, a0, a1, a2);
}

// One colour pass of the 8-colour Gauss-Seidel smoother, one thread per node (replaces the lexicographic node
// loop of gauss_seidel, General_matrix_functions.c:1231-1260).
template <int C>
__global__ void __launch_bounds__(128) ccu_k_relax(const CcuGeom g, const float *__restrict__ K,
                                                    const double *__restrict__ BI, const double *__restrict__ F, double *x,
                                                    const unsigned char *__restrict__ bits, const int drange)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if(cell >= g.NC) return;
    ccu_relax_cell<C>(g, K, BI, F, x, cell, bits, drange);
}

// T lanes per node, one colour per launch (mid-size levels)
template <int T, int C>
__global__ void __launch_bounds__(128) ccu_k_relax_lanes(const CcuGeom g, const float *__restrict__ K, const double *__restrict__ BI,
                                                          const double *__restrict__ F, double *x, const unsigned char *__restrict__ bits,
                                                          const int drange)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int cell = tid / T, q = tid % T;
    int i, j, k;
    const bool valid = cell < g.NC && ccu_decode(g, C, cell, i, j, k) && !ccu_skip_node(bits, C * g.NC + cell, drange);
    double a0, a1, a2;
    ccu_row_product_lanes<T, C>(g, K, x, cell, q, valid, a0, a1, a2);
    if(valid && q == 0) ccu_relax_update(g, BI, F, x, C * g.NC + cell, a0, a1, a2);
}

// Coarsest levels (a few thousand nodes at most): all sweeps and all eight colour passes in ONE launch of one
// CTA, a warp per node (a lane per stencil block), __syncthreads() between colours -- instead of 8*cycles
// launches (v_steps_low = 20 sweeps at the bottom of every V-cycle, General_matrix_functions.c:572-574, 611-613).
template <int C>
__device__ __forceinline__ void ccu_relax_pass_cta(const CcuGeom &g, const float *__restrict__ K, const double *__restrict__ BI,
                                                   const double *__restrict__ F, double *x)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for(int cell = warp; cell < g.NC; cell += nwarps)
    {
        int i, j, k;
        if(!ccu_decode(g, C, cell, i, j, k)) continue;
        double a0, a1, a2;
        ccu_row_product_lanes<32, C>(g, K, x, cell, lane, true, a0, a1, a2);
        if(lane == 0) ccu_relax_update(g, BI, F, x, C * g.NC + cell, a0, a1, a2);
    }
    __syncthreads();
}
__global__ void __launch_bounds__(1024) ccu_k_relax_small(const CcuGeom g, const float *__restrict__ K, const double *__restrict__ BI,
                                                           const double *__restrict__ F, double *x, const int cycles, const int zero_first)
{
    if(zero_first)
    {
        for(int s = threadIdx.x; s < 3 * g.NS; s += blockDim.x) x[s] = 0.0;
        __syncthreads();
    }
    for(int sw = 0; sw < cycles; sw++)
    {
        ccu_relax_pass_cta<7>(g, K, BI, F, x); ccu_relax_pass_cta<6>(g, K, BI, F, x);
        ccu_relax_pass_cta<5>(g, K, BI, F, x); ccu_relax_pass_cta<4>(g, K, BI, F, x);
        ccu_relax_pass_cta<3>(g, K, BI, F, x); ccu_relax_pass_cta<2>(g, K, BI, F, x);
        ccu_relax_pass_cta<1>(g, K, BI, F, x); ccu_relax_pass_cta<0>(g, K, BI, F, x);
    }
}

// ---------------------------------------------------------------- duplicated (inter-subdomain face) nodes
// Row product of one node with run-time colour, one warp per node, lane b = stencil block b (0 self, 1..13 own,
// 14..26 transposed).  ABS = 1 gives the absolute row sums sum_j |K_ij| (rebuild_BI_on_boundary, Construct_arrays.c:892).
template <int ABS>
__device__ __forceinline__ void ccu_row_product_warp(const CcuGeom &g, const float *__restrict__ K, const double *x, const int s,
                                                     const int lane, double &a0, double &a1, double &a2)
{
    const size_t NS = (size_t)g.NS;
    const int c = s / g.NC, cell = s - c * g.NC;
    const int pi = (c >> 2) & 1, pj = (c >> 1) & 1, pk = c & 1;
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    const int b = lane;
    if(b < 27)
    {
        const bool tr = b >= 14;
        const int t = tr ? b - 14 : b - 1;
        int di = 0, dj = 0, dk = 0;
        if(t >= 0)
        {
            if(t < 9) { di = -1; dj = t / 3 - 1; dk = t % 3 - 1; }
            else if(t < 12) { dj = -1; dk = t - 10; }
            else dk = -1;
        }
        if(tr) { di = -di; dj = -dj; dk = -dk; }
        const int slot = (b == 0) ? 0 : t + 1;
        const int cm = c ^ (((di != 0) << 2) | ((dj != 0) << 1) | (dk != 0));
        const int sm = cm * g.NC + cell + ccu_shift(pi, di) * g.JK + ccu_shift(pj, dj) * g.Kd + ccu_shift(pk, dk);
        const float *Kp = K + (size_t)(slot * 9) * NS + (tr ? sm : s);
        double k[9];
#pragma unroll
        for(int e = 0; e < 9; e++) { const double v = (double)__ldg(Kp + (size_t)e * NS); k[e] = ABS ? fabs(v) : v; }
        const double x0 = ABS ? 1.0 : x[sm], x1 = ABS ? 1.0 : x[NS + sm], x2 = ABS ? 1.0 : x[2 * NS + sm];
        if(!tr) { r0 = k[0] * x0 + k[1] * x1 + k[2] * x2; r1 = k[3] * x0 + k[4] * x1 + k[5] * x2; r2 = k[6] * x0 + k[7] * x1 + k[8] * x2; }
        else { r0 = k[0] * x0 + k[3] * x1 + k[6] * x2; r1 = k[1] * x0 + k[4] * x1 + k[7] * x2; r2 = k[2] * x0 + k[5] * x1 + k[8] * x2; }
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    }
    a0 = r0; a1 = r1; a2 = r2;
}
// this subdomain's part of (K x) [or of sum|K|] at every duplicated node -> face[d*n + t]
template <int ABS>
__global__ void __launch_bounds__(128) ccu_k_face_rows(const CcuGeom g, const float *__restrict__ K, const double *x, const int n,
                                                        const int *__restrict__ sh_s, double *face)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(t >= n) return;
    double a0, a1, a2;
    ccu_row_product_warp<ABS>(g, K, x, sh_s[t], lane, a0, a1, a2);
    if(lane == 0) { face[t] = a0; face[(size_t)n + t] = a1; face[2 * (size_t)n + t] = a2; }
}
// all owners' parts of a duplicated node's row, added in ascending rank order (bitwise the same on every owner)
__device__ __forceinline__ double ccu_face_total(const int t, const int d, const int n, const int *__restrict__ ptr, const int *__restrict__ src,
                                                 const double *__restrict__ face, const double *__restrict__ recv)
{
    double acc = 0.0;
    for(int e = ptr[t]; e < ptr[t + 1]; e++)
    {
        const int q = src[e];
        const double v = (q < 0) ? face[(size_t)d * n + t] : recv[(size_t)q * 3 + d];
        acc = (e == ptr[t]) ? v : acc + v;
    }
    return acc;
}
// Jacobi update of the duplicated nodes from their summed rows with the damped BI (the reference's treatment of
// OFFSIDE nodes, General_matrix_functions.c:1218-1230; BI from rebuild_BI_on_boundary)
__global__ void __launch_bounds__(128) ccu_k_face_update(const CcuGeom g, const int n, const int *__restrict__ sh_s, const int *__restrict__ ptr,
                                                          const int *__restrict__ src, const double *__restrict__ face,
                                                          const double *__restrict__ recv, const double *__restrict__ BI,
                                                          const double *__restrict__ F, double *x)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n) return;
    const int s = sh_s[t];
    const size_t NS = (size_t)g.NS;
#pragma unroll
    for(int d = 0; d < 3; d++)
    {
        const double a = ccu_face_total(t, d, n, ptr, src, face, recv);
        const float tt = (float)((F[d * NS + s] - a) * BI[d * NS + s]);
        x[d * NS + s] += (double)tt;
    }
}
// rebuild_BI_on_boundary (Construct_arrays.c:892-952): BI of a duplicated node = 1 / (sum_j |K_ij| - K_ii)
__global__ void __launch_bounds__(128) ccu_k_face_damp_BI(const CcuGeom g, const int n, const int *__restrict__ sh_s, const int *__restrict__ ptr,
                                                           const int *__restrict__ src, const double *__restrict__ face,
                                                           const double *__restrict__ recv, double *BI)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n) return;
    const int s = sh_s[t];
    const size_t NS = (size_t)g.NS;
#pragma unroll
    for(int d = 0; d < 3; d++)
    {
        const double a = ccu_face_total(t, d, n, ptr, src, face, recv);
        BI[d * NS + s] = 1.0 / (a - 1.0 / BI[d * NS + s]);
    }
}

// ---------------------------------------------------------------- bottom level, shared-memory resident
// The coarsest level (a few hundred nodes: 9x9x5 at 256x256x128 / 6 levels) is smoothed v_steps_low = 20 times at the
// bottom of every V-cycle (General_matrix_functions.c:572-574, 611-613).  Its whole half-matrix fits in ONE SM's shared
// memory (126 floats * n <= 227 KB for n <= 439), so a single CTA loads it once, keeps the solution in shared memory
// and runs every sweep and colour phase out of it: one thread per node, nodes sorted by colour so that a phase is a
// contiguous thread range reading consecutive shared-memory words.  Column n of the tables is a zero dummy that
// out-of-grid neighbours point at (the reference's "equation neq+1", Construct_arrays.c:305-309).
struct CcuSmemLevel
{
    int n = 0;                       // nodes
    int cstart[9] = { 0 };           // thread range of each colour
    int *s = nullptr;                // [n] storage slot of compact node t
    unsigned short *nbr = nullptr;   // [27][n] compact index of block b's neighbour (n = none)
};
__global__ void __launch_bounds__(1024) ccu_k_relax_smem(const CcuGeom g, const CcuSmemLevel sl, const float *__restrict__ K,
                                                          const double *__restrict__ BI, const double *__restrict__ F, double *x,
                                                          const int cycles, const int zero_first)
{
    extern __shared__ double smem_d[];
    const int n = sl.n, n1 = n + 1, t = threadIdx.x;
    const size_t NS = (size_t)g.NS;
    double *xs = smem_d;                                   // [3][n1]
    float *Ks = (float *)(smem_d + 3 * n1);                // [126][n1]
    unsigned *pk = (unsigned *)(Ks + 126 * n1);            // [n] node indices i | j << 8 | k << 16 of compact node t
    unsigned short *cmp = (unsigned short *)(pk + n);      // [nno] natural index -> compact index
    int s = 0;
    if(t < n)
    {   // every thread streams its own node's 126 coefficients (independent loads, consecutive threads ~ consecutive slots)
        s = sl.s[t];
#pragma unroll 18
        for(int q = 0; q < 126; q++) Ks[q * n1 + t] = __ldg(K + (size_t)q * NS + s);
        int i, j, k;
        ccu_decode(g, s / g.NC, s % g.NC, i, j, k);
        pk[t] = (unsigned)i | ((unsigned)j << 8) | ((unsigned)k << 16);
        cmp[k + g.noz * (j + g.nox * i)] = (unsigned short)t;
#pragma unroll
        for(int d = 0; d < 3; d++) xs[d * n1 + t] = zero_first ? 0.0 : x[d * NS + s];
    }
    for(int q = t; q < 126; q += blockDim.x) Ks[q * n1 + n] = 0.0f;
    if(t < 3) xs[t * n1 + n] = 0.0;
    __syncthreads();
    // A phase = one colour.  8 lanes per node (lane q takes stencil blocks q, q+8, q+16, q+24) and 128 node groups per
    // CTA keep all ~50-75 nodes of a phase in flight at once.  Nothing in a phase's dependent chain leaves the SM:
    // neighbours are found through the natural-index table in shared memory, and the right-hand side and inverse
    // diagonal of the eight nodes a group relaxes (one per colour) sit in the registers of its three updating lanes --
    // loaded once per call, not once per phase (a phase is ~1000 cycles; an L2 round trip per phase doubled it).
    const int q = t & 7, grp = t >> 3;
    double fr[8], br[8];
#pragma unroll
    for(int c = 0; c < 8; c++)
    {
        const int tt = sl.cstart[c] + grp;
        fr[c] = 0.0; br[c] = 0.0;
        if(q < 3 && tt < sl.cstart[c + 1]) { const int sn = sl.s[tt]; fr[c] = F[q * NS + sn]; br[c] = BI[q * NS + sn]; }
    }
    for(int sw = 0; sw < cycles; sw++)
#pragma unroll
        for(int c = 7; c >= 0; c--)
        {
            const int tt = sl.cstart[c] + grp;
            const bool act = tt < sl.cstart[c + 1];
            const double fq = fr[c], bq = br[c];
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            if(act)
            {
                const unsigned p = pk[tt];
                const int i = p & 255, j = (p >> 8) & 255, k = (p >> 16) & 255;
#pragma unroll
                for(int it = 0; it < 4; it++)
                {
                    const int b = q + 8 * it;
                    if(b < 27)
                    {
                        const bool tr = b >= 14;
                        const int u = tr ? b - 14 : b - 1;         // index into CCU_LO, -1 for the self block
                        int di = 0, dj = 0, dk = 0;
                        if(u >= 0)
                        {
                            if(u < 9) { di = -1; dj = u / 3 - 1; dk = u % 3 - 1; }
                            else if(u < 12) { dj = -1; dk = u - 10; }
                            else dk = -1;
                        }
                        if(tr) { di = -di; dj = -dj; dk = -dk; }
                        const int ii = i + di, jj = j + dj, kk = k + dk;
                        const bool in = ii >= 0 && ii < g.noy && jj >= 0 && jj < g.nox && kk >= 0 && kk < g.noz;
                        const int m = in ? (int)cmp[kk + g.noz * (jj + g.nox * ii)] : n;
                        const double x0 = xs[m], x1 = xs[n1 + m], x2 = xs[2 * n1 + m];
                        if(!tr)
                        {
                            const float *kp = Ks + (b * 9) * n1 + tt;
                            r0 += (double)kp[0] * x0 + (double)kp[n1] * x1 + (double)kp[2 * n1] * x2;
                            r1 += (double)kp[3 * n1] * x0 + (double)kp[4 * n1] * x1 + (double)kp[5 * n1] * x2;
                            r2 += (double)kp[6 * n1] * x0 + (double)kp[7 * n1] * x1 + (double)kp[8 * n1] * x2;
                        }
                        else
                        {
                            const float *kp = Ks + ((b - 13) * 9) * n1 + m;
                            r0 += (double)kp[0] * x0 + (double)kp[3 * n1] * x1 + (double)kp[6 * n1] * x2;
                            r1 += (double)kp[n1] * x0 + (double)kp[4 * n1] * x1 + (double)kp[7 * n1] * x2;
                            r2 += (double)kp[2 * n1] * x0 + (double)kp[5 * n1] * x1 + (double)kp[8 * n1] * x2;
                        }
                    }
                }
            }
#pragma unroll
            for(int o = 4; o > 0; o >>= 1)
            {
                r0 += __shfl_xor_sync(0xffffffffu, r0, o);
                r1 += __shfl_xor_sync(0xffffffffu, r1, o);
                r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            }
            if(act && q < 3)
            {   // lanes 0..2 of the group update one equation each
                const double r = q == 0 ? r0 : (q == 1 ? r1 : r2);
                xs[q * n1 + tt] += (double)(float)((fq - r) * bq);
            }
            __syncthreads();
        }
    if(t < n) { x[s] = xs[t]; x[NS + s] = xs[n1 + t]; x[2 * NS + s] = xs[2 * n1 + t]; }
}

// ---------------------------------------------------------------- bottom level, one 8-CTA cluster, fp64 rows in shared memory
// ccu_k_relax_smem above runs on ONE SM and re-converts every fp32 coefficient to fp64 at each use (40 times per call:
// 20 sweeps x own + transposed use); ncu/timing showed it bound by that SM's conversion pipe (~350 us per call, 126 calls
// per Stokes solve at 256x256x128 = 7 % of the step).  Here the level is spread over the eight SMs of a thread-block
// cluster: CTA r owns every eighth node of each colour and keeps their FULL rows (27 blocks, transposed blocks already
// transposed) in its shared memory as fp64 -- converted once per call --, every CTA holds a replica of the solution,
// a warp relaxes one node per phase (lane = stencil block), and the three updating lanes publish the new values to all
// eight replicas through distributed shared memory before the cluster barrier that ends the phase.
#define CCU_BOT_CTAS 8
#define CCU_BOT_THREADS 512
#define CCU_BOT_LD 28            // 27 blocks padded
__global__ void __cluster_dims__(CCU_BOT_CTAS, 1, 1) __launch_bounds__(CCU_BOT_THREADS)
ccu_k_relax_bottom(const CcuGeom g, const CcuSmemLevel sl, const int maxrows, const float *__restrict__ K, const double *__restrict__ BI,
                   const double *__restrict__ F, double *x, const int cycles, const int zero_first)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int r = (int)cl.block_rank();
    extern __shared__ double smem_d[];
    const int n = sl.n, n1 = n + 1, t = threadIdx.x;
    const size_t NS = (size_t)g.NS;
    double *xs = smem_d;                                         // [3][n1] replica of the solution (+ zero dummy column n)
    double *Kd = xs + 3 * n1;                                    // [maxrows][9][CCU_BOT_LD] full rows of this CTA's nodes
    double *Fl = Kd + (size_t)maxrows * 9 * CCU_BOT_LD;          // [maxrows][3]
    double *Bl = Fl + maxrows * 3;                               // [maxrows][3]
    unsigned short *nb = (unsigned short *)(Bl + maxrows * 3);   // [maxrows][CCU_BOT_LD] compact index of block b's neighbour
    int cnt[8], base[8], rows = 0;
#pragma unroll
    for(int c = 0; c < 8; c++)
    {
        const int nc = sl.cstart[c + 1] - sl.cstart[c];
        cnt[c] = (nc - r + CCU_BOT_CTAS - 1) / CCU_BOT_CTAS;     // nodes idx = r, r+8, ... of colour c
        if(cnt[c] < 0) cnt[c] = 0;
        base[c] = rows; rows += cnt[c];
    }
    // ---- once per call: rows -> fp64, neighbour table, F, BI, solution replica
    for(int w = t; w < rows * CCU_BOT_LD; w += CCU_BOT_THREADS)
    {
        const int row = w / CCU_BOT_LD, b = w - row * CCU_BOT_LD;
        int c = 0;
#pragma unroll
        for(int q = 1; q < 8; q++) if(row >= base[q]) c = q;
        const int tt = sl.cstart[c] + r + CCU_BOT_CTAS * (row - base[c]);
        if(b < 27)
        {
            const int m = sl.nbr[b * n + tt];
            nb[row * CCU_BOT_LD + b] = (unsigned short)m;
            const bool tr = b >= 14;
            const int slot = tr ? b - 13 : b;
            const int sn = tr ? (m < n ? sl.s[m] : -1) : sl.s[tt];
#pragma unroll
            for(int e = 0; e < 9; e++)
            {
                const int es = tr ? 3 * (e % 3) + e / 3 : e;        // transposed block: K_nm = (K_mn)^T
                Kd[((size_t)row * 9 + e) * CCU_BOT_LD + b] = sn >= 0 ? (double)__ldg(K + (size_t)(slot * 9 + es) * NS + sn) : 0.0;
            }
        }
        else
        {
            nb[row * CCU_BOT_LD + b] = (unsigned short)n;
            for(int e = 0; e < 9; e++) Kd[((size_t)row * 9 + e) * CCU_BOT_LD + b] = 0.0;
        }
        if(b < 3) { const int sn = sl.s[tt]; Fl[row * 3 + b] = F[b * NS + sn]; Bl[row * 3 + b] = BI[b * NS + sn]; }
    }
    for(int w = t; w < 3 * n1; w += CCU_BOT_THREADS)
    {
        const int d = w / n1, m = w - d * n1;
        xs[w] = (m < n && !zero_first) ? x[d * NS + sl.s[m]] : 0.0;
    }
    cl.sync();
    // ---- phases: one colour each; warp w relaxes the CTA's w-th node of the colour, lane = stencil block
    const int warp = t >> 5, lane = t & 31;
    for(int sw = 0; sw < cycles; sw++)
#pragma unroll
        for(int c = 7; c >= 0; c--)
        {
            if(warp < cnt[c])
            {
                const int row = base[c] + warp, tt = sl.cstart[c] + r + CCU_BOT_CTAS * warp;
                double r0 = 0.0, r1 = 0.0, r2 = 0.0;
                if(lane < 27)
                {
                    const int m = nb[row * CCU_BOT_LD + lane];
                    const double x0 = xs[m], x1 = xs[n1 + m], x2 = xs[2 * n1 + m];
                    const double *kp = Kd + (size_t)row * 9 * CCU_BOT_LD + lane;
                    r0 = kp[0] * x0 + kp[CCU_BOT_LD] * x1 + kp[2 * CCU_BOT_LD] * x2;
                    r1 = kp[3 * CCU_BOT_LD] * x0 + kp[4 * CCU_BOT_LD] * x1 + kp[5 * CCU_BOT_LD] * x2;
                    r2 = kp[6 * CCU_BOT_LD] * x0 + kp[7 * CCU_BOT_LD] * x1 + kp[8 * CCU_BOT_LD] * x2;
                }
#pragma unroll
                for(int o = 16; o > 0; o >>= 1)
                {
                    r0 += __shfl_xor_sync(0xffffffffu, r0, o);
                    r1 += __shfl_xor_sync(0xffffffffu, r1, o);
                    r2 += __shfl_xor_sync(0xffffffffu, r2, o);
                }
                if(lane < 3)
                {   // General_matrix_functions.c:1250-1259: scalar BI per equation, correction rounded to fp32
                    const double rr = lane == 0 ? r0 : (lane == 1 ? r1 : r2);
                    const double xn = xs[lane * n1 + tt] + (double)(float)((Fl[row * 3 + lane] - rr) * Bl[row * 3 + lane]);
#pragma unroll
                    for(int q = 0; q < CCU_BOT_CTAS; q++) cl.map_shared_rank(xs, q)[lane * n1 + tt] = xn;
                }
            }
            cl.sync();
        }
    // every CTA writes back the nodes it owns
    for(int w = t; w < rows * 3; w += CCU_BOT_THREADS)
    {
        const int row = w / 3, d = w - row * 3;
        int c = 0;
#pragma unroll
        for(int q = 1; q < 8; q++) if(row >= base[q]) c = q;
        const int tt = sl.cstart[c] + r + CCU_BOT_CTAS * (row - base[c]);
        x[d * NS + sl.s[tt]] = xs[d * n1 + tt];
    }
}

// Au = K*u for all nodes (n_assemble_del2_u, Element_calculations.c:552).  One warp per colour,
// the eight warps of a block cover the same 32 cells, so the transposed reads of a block hit
// lines its sibling warps stream in at the same time: each coefficient crosses HBM once.
// MODE 0: Au = K u (optionally stripped); MODE 1: out = rhs - K u (residual, stripped rows give rhs)
template <int MODE>
__global__ void __launch_bounds__(256) ccu_k_matvec(const CcuGeom g, const float *__restrict__ K,
                                                     const unsigned char *__restrict__ flags, const double *u,
                                                     const double *rhs, double *out, const int strip)
{
    const int c = threadIdx.x >> 5;
    const int cell = blockIdx.x * 32 + (threadIdx.x & 31);
    if(cell >= g.NC) return;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    double a0, a1, a2;
    switch(c)
    {
    case 0: ccu_row_product<0>(g, K, u, cell, a0, a1, a2); break;
    case 1: ccu_row_product<1>(g, K, u, cell, a0, a1, a2); break;
    case 2: ccu_row_product<2>(g, K, u, cell, a0, a1, a2); break;
    case 3: ccu_row_product<3>(g, K, u, cell, a0, a1, a2); break;
    case 4: ccu_row_product<4>(g, K, u, cell, a0, a1, a2); break;
    case 5: ccu_row_product<5>(g, K, u, cell, a0, a1, a2); break;
    case 6: ccu_row_product<6>(g, K, u, cell, a0, a1, a2); break;
    default: ccu_row_product<7>(g, K, u, cell, a0, a1, a2); break;
    }
    const size_t NS = (size_t)g.NS;
    const int s = c * g.NC + cell;
    if(strip)
    {
        const unsigned char f = flags[s];
        if(f & CCU_F_VBX) a0 = 0.0;
        if(f & CCU_F_VBY) a1 = 0.0;
        if(f & CCU_F_VBZ) a2 = 0.0;
    }
    if(MODE == 0) { out[s] = a0; out[NS + s] = a1; out[2 * NS + s] = a2; }
    else { out[s] = rhs[s] - a0; out[NS + s] = rhs[NS + s] - a1; out[2 * NS + s] = rhs[2 * NS + s] - a2; }
}

// ---------------------------------------------------------------- table-driven row product (large levels)
// The fully unrolled ccu_row_product<C> is ~2000 instructions per colour; a kernel that runs all eight colours side
// by side (the matvec: one warp per colour so that sibling warps share the transposed lines) then thrashes the
// instruction cache (ncu r01: `no_instruction` was its top stall, 17 % of DRAM peak).  Here the 27 neighbour offsets
// of every colour come from a per-level table in the kernel parameters (constant bank, warp-uniform index) and one
// short loop serves all colours.
struct CcuStencil
{
    int off[8][27];      // storage-slot offset of block b's neighbour relative to the node's own slot, by colour
};
__host__ inline CcuStencil ccu_make_stencil(const CcuGeom &g)
{
    const int LO[13][3] = CCU_LO_INIT;
    CcuStencil st;
    for(int c = 0; c < 8; c++)
    {
        const int pi = (c >> 2) & 1, pj = (c >> 1) & 1, pk = c & 1;
        for(int b = 0; b < 27; b++)
        {
            int di = 0, dj = 0, dk = 0;
            if(b >= 1 && b <= 13) { di = LO[b - 1][0]; dj = LO[b - 1][1]; dk = LO[b - 1][2]; }
            if(b >= 14) { di = -LO[b - 14][0]; dj = -LO[b - 14][1]; dk = -LO[b - 14][2]; }
            const int cm = c ^ (((di != 0) << 2) | ((dj != 0) << 1) | (dk != 0));
            st.off[c][b] = (cm - c) * g.NC + ccu_shift(pi, di) * g.JK + ccu_shift(pj, dj) * g.Kd + ccu_shift(pk, dk);
        }
    }
    return st;
}
// streaming load of a coefficient that this SM will not touch again: do not let it displace the neighbour values in L1
__device__ __forceinline__ float ccu_ldg_na(const float *p)
{
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// NA: 0 = every coefficient load allocates in L1 (__ldg); 1 = the node's own blocks stream past L1 (each is read once per
// pass by this SM; the transposed blocks, which sibling warps fetch as their own, still allocate); 2 = all stream past L1
template <int U, int NA = 0>
__device__ __forceinline__ void ccu_row_product_tab(const size_t NS, const int *__restrict__ off, const float *__restrict__ K,
                                                    const double *x, const int s, double &a0, double &a1, double &a2)
{
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll U
    for(int b = 0; b < 14; b++)
    {   // self + own blocks, stored at this node
        const int sm = s + off[b];
        const float *Kp = K + (size_t)(b * 9) * NS + s;
        float k[9];
#pragma unroll
        for(int e = 0; e < 9; e++) k[e] = NA >= 1 ? ccu_ldg_na(Kp + (size_t)e * NS) : __ldg(Kp + (size_t)e * NS);
        const double x0 = x[sm], x1 = x[NS + sm], x2 = x[2 * NS + sm];
        r0 += (double)k[0] * x0 + (double)k[1] * x1 + (double)k[2] * x2;
        r1 += (double)k[3] * x0 + (double)k[4] * x1 + (double)k[5] * x2;
        r2 += (double)k[6] * x0 + (double)k[7] * x1 + (double)k[8] * x2;
    }
#pragma unroll U
    for(int b = 14; b < 27; b++)
    {   // transposed blocks, stored at the upper neighbours
        const int sm = s + off[b];
        const float *Kp = K + (size_t)((b - 13) * 9) * NS + sm;
        float k[9];
#pragma unroll
        for(int e = 0; e < 9; e++) k[e] = NA >= 2 ? ccu_ldg_na(Kp + (size_t)e * NS) : __ldg(Kp + (size_t)e * NS);
        const double x0 = x[sm], x1 = x[NS + sm], x2 = x[2 * NS + sm];
        r0 += (double)k[0] * x0 + (double)k[3] * x1 + (double)k[6] * x2;
        r1 += (double)k[1] * x0 + (double)k[4] * x1 + (double)k[7] * x2;
        r2 += (double)k[2] * x0 + (double)k[5] * x1 + (double)k[8] * x2;
    }
    a0 = r0; a1 = r1; a2 = r2;
}
// The same row from FULL storage: the thirteen blocks towards the upper neighbours were copied (transposed) to the node's own
// slot (KT, ccu_k_build_KT), so every coefficient of the row streams with unit stride and is read from HBM once per pass --
// the gather above reads each stored block twice per sweep (own + transposed use), 2.5x the algorithmic bytes (ncu r01).
template <int U>
__device__ __forceinline__ void ccu_row_product_full(const size_t NS, const int *__restrict__ off, const float *__restrict__ K,
                                                     const float *__restrict__ KT, const double *x, const int s, double &a0, double &a1, double &a2)
{
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll U
    for(int b = 0; b < 27; b++)
    {
        const int sm = s + off[b];
        const float *Kp = (b < 14 ? K + (size_t)(b * 9) * NS : KT + (size_t)((b - 14) * 9) * NS) + s;
        float k[9];
#pragma unroll
        for(int e = 0; e < 9; e++) k[e] = ccu_ldg_na(Kp + (size_t)e * NS);
        const double x0 = x[sm], x1 = x[NS + sm], x2 = x[2 * NS + sm];
        r0 += (double)k[0] * x0 + (double)k[1] * x1 + (double)k[2] * x2;
        r1 += (double)k[3] * x0 + (double)k[4] * x1 + (double)k[5] * x2;
        r2 += (double)k[6] * x0 + (double)k[7] * x1 + (double)k[8] * x2;
    }
    a0 = r0; a1 = r1; a2 = r2;
}
// KT[(t*9 + 3a + bb)][s] = K[((t+1)*9 + 3bb + a)][s + off[14+t]]: block t+1 of the upper neighbour, transposed
__global__ void __launch_bounds__(256) ccu_k_build_KT(const CcuGeom g, const __grid_constant__ CcuStencil st, const float *__restrict__ K, float *__restrict__ KT)
{
    const int c = blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if(cell >= g.NC) return;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    const size_t NS = (size_t)g.NS;
    const int s = c * g.NC + cell;
    for(int t = 0; t < 13; t++)
    {
        const int sm = s + st.off[c][14 + t];
        const float *Kp = K + (size_t)((t + 1) * 9) * NS + sm;
        float *o = KT + (size_t)(t * 9) * NS + s;
#pragma unroll
        for(int a = 0; a < 3; a++)
#pragma unroll
            for(int bb = 0; bb < 3; bb++) o[(size_t)(3 * a + bb) * NS] = __ldg(Kp + (size_t)(3 * bb + a) * NS);
    }
}
template <int U>
__global__ void __launch_bounds__(128) ccu_k_relax_full(const CcuGeom g, const __grid_constant__ CcuStencil st, const int c,
                                                         const float *__restrict__ K, const float *__restrict__ KT, const double *__restrict__ BI,
                                                         const double *__restrict__ F, double *x, const unsigned char *__restrict__ bits,
                                                         const int drange)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if(cell >= g.NC) return;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    const int s = c * g.NC + cell;
    if(ccu_skip_node(bits, s, drange)) return;
    double a0, a1, a2;
    ccu_row_product_full<U>((size_t)g.NS, st.off[c], K, KT, x, s, a0, a1, a2);
    ccu_relax_update(g, BI, F, x, s, a0, a1, a2);
}
template <int MODE, int U>
__global__ void __launch_bounds__(256) ccu_k_matvec_full(const CcuGeom g, const __grid_constant__ CcuStencil st, const float *__restrict__ K,
                                                          const float *__restrict__ KT, const unsigned char *__restrict__ flags, const double *u,
                                                          const double *rhs, double *out, const int strip)
{
    const int c = threadIdx.x >> 5;
    const int cell = blockIdx.x * 32 + (threadIdx.x & 31);
    if(cell >= g.NC) return;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    const size_t NS = (size_t)g.NS;
    const int s = c * g.NC + cell;
    double a0, a1, a2;
    ccu_row_product_full<U>(NS, st.off[c], K, KT, u, s, a0, a1, a2);
    if(strip)
    {
        const unsigned char f = flags[s];
        if(f & CCU_F_VBX) a0 = 0.0;
        if(f & CCU_F_VBY) a1 = 0.0;
        if(f & CCU_F_VBZ) a2 = 0.0;
    }
    if(MODE == 0) { out[s] = a0; out[NS + s] = a1; out[2 * NS + s] = a2; }
    else { out[s] = rhs[s] - a0; out[NS + s] = rhs[NS + s] - a1; out[2 * NS + s] = rhs[2 * NS + s] - a2; }
}
template <int MODE, int U, int NA = 0>
__global__ void __launch_bounds__(256) ccu_k_matvec_tab(const CcuGeom g, const __grid_constant__ CcuStencil st, const float *__restrict__ K,
                                                         const unsigned char *__restrict__ flags, const double *u,
                                                         const double *rhs, double *out, const int strip)
{
    const int c = threadIdx.x >> 5;
    const int cell = blockIdx.x * 32 + (threadIdx.x & 31);
    if(cell >= g.NC) return;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    const size_t NS = (size_t)g.NS;
    const int s = c * g.NC + cell;
    double a0, a1, a2;
    ccu_row_product_tab<U, NA>(NS, st.off[c], K, u, s, a0, a1, a2);
    if(strip)
    {
        const unsigned char f = flags[s];
        if(f & CCU_F_VBX) a0 = 0.0;
        if(f & CCU_F_VBY) a1 = 0.0;
        if(f & CCU_F_VBZ) a2 = 0.0;
    }
    if(MODE == 0) { out[s] = a0; out[NS + s] = a1; out[2 * NS + s] = a2; }
    else { out[s] = rhs[s] - a0; out[NS + s] = rhs[NS + s] - a1; out[2 * NS + s] = rhs[2 * NS + s] - a2; }
}
// one colour pass of the smoother with the same table-driven row (colour = kernel argument)
template <int U, int NA = 0>
__global__ void __launch_bounds__(128) ccu_k_relax_tab(const CcuGeom g, const __grid_constant__ CcuStencil st, const int c,
                                                        const float *__restrict__ K, const double *__restrict__ BI,
                                                        const double *__restrict__ F, double *x, const unsigned char *__restrict__ bits,
                                                        const int drange)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if(cell >= g.NC) return;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    const int s = c * g.NC + cell;
    if(ccu_skip_node(bits, s, drange)) return;
    double a0, a1, a2;
    ccu_row_product_tab<U, NA>((size_t)g.NS, st.off[c], K, x, s, a0, a1, a2);
    ccu_relax_update(g, BI, F, x, s, a0, a1, a2);
}

// T lanes per node variant of the matvec (coarse and mid levels): thread -> (colour, cell, lane)
template <int T, int MODE>
__global__ void __launch_bounds__(256) ccu_k_matvec_lanes(const CcuGeom g, const float *__restrict__ K,
                                                           const unsigned char *__restrict__ flags, const double *u,
                                                           const double *rhs, double *out, const int strip)
{
    const int c = threadIdx.x >> 5;                               // one warp per colour, 32/T cells per warp
    const int lane = threadIdx.x & 31;
    const int cell = blockIdx.x * (32 / T) + lane / T, q = lane % T;
    int i, j, k;
    const bool valid = cell < g.NC && ccu_decode(g, c, cell, i, j, k);
    double a0, a1, a2;
    switch(c)
    {
    case 0: ccu_row_product_lanes<T, 0>(g, K, u, cell, q, valid, a0, a1, a2); break;
    case 1: ccu_row_product_lanes<T, 1>(g, K, u, cell, q, valid, a0, a1, a2); break;
    case 2: ccu_row_product_lanes<T, 2>(g, K, u, cell, q, valid, a0, a1, a2); break;
    case 3: ccu_row_product_lanes<T, 3>(g, K, u, cell, q, valid, a0, a1, a2); break;
    case 4: ccu_row_product_lanes<T, 4>(g, K, u, cell, q, valid, a0, a1, a2); break;
    case 5: ccu_row_product_lanes<T, 5>(g, K, u, cell, q, valid, a0, a1, a2); break;
    case 6: ccu_row_product_lanes<T, 6>(g, K, u, cell, q, valid, a0, a1, a2); break;
    default: ccu_row_product_lanes<T, 7>(g, K, u, cell, q, valid, a0, a1, a2); break;
    }
    if(!valid || q != 0) return;
    const size_t NS = (size_t)g.NS;
    const int s = c * g.NC + cell;
    if(strip)
    {
        const unsigned char f = flags[s];
        if(f & CCU_F_VBX) a0 = 0.0;
        if(f & CCU_F_VBY) a1 = 0.0;
        if(f & CCU_F_VBZ) a2 = 0.0;
    }
    if(MODE == 0) { out[s] = a0; out[NS + s] = a1; out[2 * NS + s] = a2; }
    else { out[s] = rhs[s] - a0; out[NS + s] = rhs[NS + s] - a1; out[2 * NS + s] = rhs[2 * NS + s] - a2; }
}

// ---------------------------------------------------------------- layout conversion
// reference vector double[neq] (equation 3n+d) <-> colour-blocked SoA
__global__ void ccu_k_vec_to_dev(const CcuGeom g, const double *__restrict__ nat, double *dev)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    dev[s] = nat[3 * (size_t)n]; dev[(size_t)g.NS + s] = nat[3 * (size_t)n + 1]; dev[2 * (size_t)g.NS + s] = nat[3 * (size_t)n + 2];
}
__global__ void ccu_k_vec_to_nat(const CcuGeom g, const double *__restrict__ dev, double *nat)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    nat[3 * (size_t)n] = dev[s]; nat[3 * (size_t)n + 1] = dev[(size_t)g.NS + s]; nat[3 * (size_t)n + 2] = dev[2 * (size_t)g.NS + s];
}
__global__ void ccu_k_flags_to_dev(const CcuGeom g, const unsigned *__restrict__ node, unsigned char *flags)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const unsigned f = node[n];
    flags[ccu_sidx(g, i, j, k)] = (unsigned char)(CCU_F_VALID | ((f & CCU_VBX) ? CCU_F_VBX : 0) | ((f & CCU_VBY) ? CCU_F_VBY : 0) |
                                                  ((f & CCU_VBZ) ? CCU_F_VBZ : 0));
}
// Eqn_k1/2/3 in the reference's layout (node-major, 42 per node, slots only for in-grid lower
// neighbours, Construct_arrays.c:320-347) -> fixed 13-neighbour slots, coefficient-major SoA.
__global__ void ccu_k_stiffness_to_dev(const CcuGeom g, const float *__restrict__ k1, const float *__restrict__ k2,
                                       const float *__restrict__ k3, float *K)
{
    constexpr int LO[13][3] = CCU_LO_INIT;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    const size_t NS = (size_t)g.NS, base = (size_t)n * 42;
    const float *kk[3] = { k1 + base, k2 + base, k3 + base };
    for(int a = 0; a < 3; a++)
        for(int b = 0; b < 3; b++) K[(size_t)(a * 3 + b) * NS + s] = kk[a][b];
    int rs = 0;
    for(int q = 0; q < 13; q++)
    {
        const int ii = i + LO[q][0], jj = j + LO[q][1], kz = k + LO[q][2];
        const bool in = ii >= 0 && jj >= 0 && jj < g.nox && kz >= 0 && kz < g.noz;
        if(in) rs++;
        for(int a = 0; a < 3; a++)
            for(int b = 0; b < 3; b++) K[(size_t)((q + 1) * 9 + a * 3 + b) * NS + s] = in ? kk[a][3 * rs + b] : 0.0f;
    }
}

// ---------------------------------------------------------------- vector algebra on whole padded arrays
__global__ void ccu_k_strip(const CcuGeom g, const unsigned char *__restrict__ flags, double *v)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if(s >= g.NS) return;
    const unsigned char f = flags[s];
    if(f & CCU_F_VBX) v[s] = 0.0;
    if(f & CCU_F_VBY) v[(size_t)g.NS + s] = 0.0;
    if(f & CCU_F_VBZ) v[2 * (size_t)g.NS + s] = 0.0;
}
// y = a*x + b*y with a, b read from device scalars: a = sa*num_a/den_a (den null -> 1)
struct CcuCoef { const double *num, *den; double scale; };
__device__ __forceinline__ double ccu_coef(const CcuCoef &c)
{
    double v = c.scale;
    if(c.num) v *= *c.num;
    if(c.den) v /= *c.den;
    return v;
}
__global__ void ccu_k_axpby(const size_t n, double *y, const double *__restrict__ x, const CcuCoef ca, const CcuCoef cb)
{
    const double a = ccu_coef(ca), b = ccu_coef(cb);
    for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = a * x[i] + b * y[i];
}
// z = a*x + b*y
__global__ void ccu_k_waxpby(const size_t n, double *z, const double *__restrict__ x, const double *__restrict__ y,
                             const CcuCoef ca, const CcuCoef cb)
{
    const double a = ccu_coef(ca), b = ccu_coef(cb);
    for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        z[i] = a * x[i] + b * y[i];
}
__global__ void ccu_k_mul(const size_t n, double *z, const double *__restrict__ x, const double *__restrict__ y)
{
    for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        z[i] = x[i] * y[i];
}

// deterministic dot: fixed grid, fixed per-thread stride order, shuffle tree, then one block folds
// the partials (global_vdot / global_pdot, Global_operations.c:339-375).  Up to 3 dots per pass so
// the pairs the callers need together share one read of the vectors.
#define CCU_DOT_BLOCKS 592
// `own` (multi-subdomain runs, nodal vectors only): bit 0 of own[i % ns] says this subdomain counts the node
// (the reference's IDD ownership mask, Construct_arrays.c:169-191)
__global__ void __launch_bounds__(256) ccu_k_dot_partial(const size_t n, const double *__restrict__ a0, const double *__restrict__ b0,
                                                          const double *__restrict__ a1, const double *__restrict__ b1,
                                                          const double *__restrict__ a2, const double *__restrict__ b2,
                                                          double *partial, const unsigned char *__restrict__ own, const size_t ns)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        if(own && !(own[i % ns] & CCU_B_OWNED)) continue;
        s0 += a0[i] * b0[i];
        if(a1) s1 += a1[i] * b1[i];
        if(a2) s2 += a2[i] * b2[i];
    }
    __shared__ double sh[3][8];
    for(int o = 16; o > 0; o >>= 1)
    {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; sh[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if(threadIdx.x < 3)
    {
        double t = 0.0;
        for(int w = 0; w < 8; w++) t += sh[threadIdx.x][w];
        partial[threadIdx.x * CCU_DOT_BLOCKS + blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) ccu_k_dot_final(const double *__restrict__ partial, const int nblocks, double *out0, double *out1, double *out2)
{
    __shared__ double sh[8];
    double *outs[3] = { out0, out1, out2 };
    for(int q = 0; q < 3; q++)
    {
        if(!outs[q]) continue;
        double s = 0.0;
        for(int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partial[q * CCU_DOT_BLOCKS + i];
        for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if(threadIdx.x == 0)
        {
            double t = 0.0;
            for(int w = 0; w < 8; w++) t += sh[w];
            *outs[q] = t;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- multigrid transfers
// project_vector (Solver_multigrid.c:72-159): one thread per coarse node gathers the eight coarse
// elements around it in the reference's accumulation order (ascending element number), each
// contributing TWW * (sum of the 8 nodes of the fine sub-element in that octant); then * MASS.
__global__ void __launch_bounds__(128) ccu_k_project(const CcuGeom gc, const CcuGeom gf, const float *__restrict__ TWW,
                                                      const float *__restrict__ MASS, const double *__restrict__ fine, double *coarse,
                                                      const int apply_mass)
{
    constexpr int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };   // [dz][dx][dy] -> local node
    // threads run over the coarse nodes in NATURAL order (z fastest): the fine node at a fixed offset from (2I, 2J, 2Kz) then sits
    // at consecutive slots of one colour block of the fine vector, so all 81 loads of a warp are contiguous (the coarse-colour order
    // made them stride-2); the coarse stores are the strided side, an eighth of the data
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= gc.nno) return;
    const int Kz = t % gc.noz, J = (t / gc.noz) % gc.nox, I = t / (gc.noz * gc.nox);
    // The eight sub-elements around fine node (2I, 2J, 2Kz) overlap: their 64 nodal values are 27 distinct fine nodes.
    // W[ay][ax][az] = TWW of the coarse element in octant (ay, ax, az) (0 outside the mesh); fine node at offset
    // (dy, dx, dz) collects the weights of every octant whose sub-element holds it, so each fine value is read once.
    float W[2][2][2];
#pragma unroll
    for(int ay = 0; ay < 2; ay++)
#pragma unroll
        for(int ax = 0; ax < 2; ax++)
#pragma unroll
            for(int az = 0; az < 2; az++)
            {
                const int ey = I - 1 + ay, ex = J - 1 + ax, ez = Kz - 1 + az;
                const bool in = ey >= 0 && ey < gc.ely && ex >= 0 && ex < gc.elx && ez >= 0 && ez < gc.elz;
                W[ay][ax][az] = in ? TWW[(size_t)(ez + gc.elz * (ex + gc.elx * ey)) * 8 + LUT[1 - az][1 - ax][1 - ay] - 1] : 0.0f;
            }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for(int dy = -1; dy <= 1; dy++)
#pragma unroll
        for(int dx = -1; dx <= 1; dx++)
#pragma unroll
            for(int dz = -1; dz <= 1; dz++)
            {
                double w = 0.0;
#pragma unroll
                for(int ay = (dy > 0); ay <= (dy >= 0); ay++)
#pragma unroll
                    for(int ax = (dx > 0); ax <= (dx >= 0); ax++)
#pragma unroll
                        for(int az = (dz > 0); az <= (dz >= 0); az++) w += (double)W[ay][ax][az];
                const int fy = 2 * I + dy, fx = 2 * J + dx, fz = 2 * Kz + dz;
                if(w != 0.0 && fy >= 0 && fy < gf.noy && fx >= 0 && fx < gf.nox && fz >= 0 && fz < gf.noz)
                {
                    const int sf = ccu_sidx(gf, fy, fx, fz);
                    s0 += w * fine[sf]; s1 += w * fine[(size_t)gf.NS + sf]; s2 += w * fine[2 * (size_t)gf.NS + sf];
                }
            }
    const double m = apply_mass ? (double)MASS[t] : 1.0;
    const int sc = ccu_sidx(gc, I, J, Kz);
    coarse[sc] = s0 * m; coarse[(size_t)gc.NS + sc] = s1 * m; coarse[2 * (size_t)gc.NS + sc] = s2 * m;
}
// second half of project_vector when the halo sum sits between the gather and the mass factor (Solver_multigrid.c:150-157)
__global__ void __launch_bounds__(128) ccu_k_mass_mul(const CcuGeom g, const float *__restrict__ MASS, double *v)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= 8 * g.NC) return;
    const int c = t / g.NC, cell = t - c * g.NC;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    const double m = (double)MASS[k + g.noz * (j + g.nox * i)];
    const int s = c * g.NC + cell;
    v[s] *= m; v[(size_t)g.NS + s] *= m; v[2 * (size_t)g.NS + s] *= m;
}

__global__ void __launch_bounds__(256) ccu_k_eco_transpose(const int nel, const float *__restrict__ eco, float *__restrict__ ecoT)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if(t >= (size_t)nel * 3) return;
    const int d = (int)(t / nel), e = (int)(t - (size_t)d * nel);
    ecoT[t] = eco[(size_t)e * 3 + d];
}
// interp_vector + un_inject_vector (Solver_multigrid.c:173-298, 581-634): the reference fills x,
// then z, then y gaps in place; evaluated here per fine node as the same nested two-point
// interpolations (fp32 weights from the element sizes the reference looks up, Appendix A #8):
// along x between the coarse nodes, then along z between those results, then along y.
__device__ __forceinline__ int ccu_first_elt(const CcuGeom &g, int i, int j, int k)
{
    return (k > 0 ? k - 1 : 0) + g.elz * ((j > 0 ? j - 1 : 0) + g.elx * (i > 0 ? i - 1 : 0));
}
// eco: element sizes direction-major, [3][nel] (Level::ecoT), so that neighbouring threads read neighbouring floats
__device__ __forceinline__ void ccu_w12(const float *__restrict__ eco, const int nel, int e1, int e2, int dir, float &w1, float &w2)
{
    const float x1 = eco[(size_t)dir * nel + e1], x2 = eco[(size_t)dir * nel + e2];
    w1 = x2 / (x1 + x2); w2 = x1 / (x1 + x2);
}
// weights are evaluated once per node and applied to the three dofs (they are the same for each)
struct CcuInterpW { float w1, w2; };
__device__ __forceinline__ CcuInterpW ccu_wpair(const float *__restrict__ eco, const int nel, int e1, int e2, int dir)
{
    CcuInterpW w; ccu_w12(eco, nel, e1, e2, dir, w.w1, w.w2); return w;
}
__global__ void __launch_bounds__(128) ccu_k_interp(const CcuGeom gc, const CcuGeom gf, const float *__restrict__ eco_f,
                                                     const unsigned char *__restrict__ flags_f, const double *__restrict__ coarse,
                                                     double *fine, const int strip)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= 8 * gf.NC) return;
    const int c = t / gf.NC, cell = t - c * gf.NC;
    int i, j, k;
    if(!ccu_decode(gf, c, cell, i, j, k)) return;
    const int s = c * gf.NC + cell;
    const unsigned char f = strip ? flags_f[s] : 0;
    const bool oi = i & 1, oj = j & 1, ok = k & 1;
    // rows i0 (and i1 when i is odd), columns k0 (k1), x positions j0 (j1): up to eight coarse nodes
    const int i0 = oi ? i - 1 : i, i1 = i + 1, k0 = ok ? k - 1 : k, k1 = k + 1, j0 = oj ? j - 1 : j, j1 = j + 1;
    CcuInterpW wy = { 1.0f, 0.0f }, wz[2] = { { 1.0f, 0.0f }, { 1.0f, 0.0f } }, wx[2][2] = { { { 1.0f, 0.0f }, { 1.0f, 0.0f } }, { { 1.0f, 0.0f }, { 1.0f, 0.0f } } };
    if(oi) wy = ccu_wpair(eco_f, gf.nel, ccu_first_elt(gf, i - 1, j, k), ccu_first_elt(gf, i + 1, j, k), 1);
#pragma unroll
    for(int a = 0; a < 2; a++)
    {
        if(a && !oi) break;
        const int ii = a ? i1 : i0;
        if(ok) wz[a] = ccu_wpair(eco_f, gf.nel, ccu_first_elt(gf, ii, j, k - 1), ccu_first_elt(gf, ii, j, k + 1), 2);
#pragma unroll
        for(int b = 0; b < 2; b++)
        {
            if(b && !ok) break;
            const int kk = b ? k1 : k0;
            if(oj) wx[a][b] = ccu_wpair(eco_f, gf.nel, ccu_first_elt(gf, ii, j - 1, kk), ccu_first_elt(gf, ii, j + 1, kk), 0);
        }
    }
    // coarse storage slots of the (up to) eight corners
    int sc[2][2][2];
#pragma unroll
    for(int a = 0; a < 2; a++)
#pragma unroll
        for(int b = 0; b < 2; b++)
#pragma unroll
            for(int e = 0; e < 2; e++)
                sc[a][b][e] = ccu_sidx(gc, (a && oi ? i1 : i0) >> 1, (e && oj ? j1 : j0) >> 1, (b && ok ? k1 : k0) >> 1);
#pragma unroll
    for(int d = 0; d < 3; d++)
    {
        const double *cv = coarse + (size_t)d * gc.NS;
        double vy[2];
#pragma unroll
        for(int a = 0; a < 2; a++)
        {
            if(a && !oi) { vy[a] = 0.0; break; }
            double vz[2];
#pragma unroll
            for(int b = 0; b < 2; b++)
            {
                if(b && !ok) { vz[b] = 0.0; break; }
                // interp along x (Solver_multigrid.c:195-226): even j copies the coarse value
                vz[b] = oj ? (double)wx[a][b].w1 * cv[sc[a][b][0]] + (double)wx[a][b].w2 * cv[sc[a][b][1]] : cv[sc[a][b][0]];
            }
            vy[a] = ok ? (double)wz[a].w1 * vz[0] + (double)wz[a].w2 * vz[1] : vz[0];
        }
        double v = oi ? (double)wy.w1 * vy[0] + (double)wy.w2 * vy[1] : vy[0];
        if(f & (d == 0 ? CCU_F_VBX : (d == 1 ? CCU_F_VBY : CCU_F_VBZ))) v = 0.0;
        fine[(size_t)d * gf.NS + s] = v;
    }
}

// ---------------------------------------------------------------- pressure coupling
// elt_del[nel][24] (the reference's element-major layout) -> eltT[24][nel]: consecutive elements of one coefficient are
// contiguous, so the element / node threads of div_u / grad_p read whole sectors (the element-major array costs a
// 96-byte stride per thread: 24 wavefronts per load instruction, which made both kernels LSU-bound)
__global__ void __launch_bounds__(256) ccu_k_elt_del_transpose(const int nel, const float *__restrict__ elt_del, float *__restrict__ eltT)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if(t >= (size_t)nel * 24) return;
    const int p = (int)(t / nel), e = (int)(t - (size_t)p * nel);
    eltT[t] = elt_del[(size_t)e * 24 + p];
}
// assemble_div_u (Element_calculations.c:691-720): one thread per element, local nodes in order.
__global__ void __launch_bounds__(128) ccu_k_div_u(const CcuGeom g, const float *__restrict__ eltT, const double *__restrict__ U, double *divU)
{
    constexpr int OFFS[9][3] = CCU_OFFS_INIT;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    const float *gg = eltT + e;
    const size_t nel = (size_t)g.nel;
    double s = 0.0;
#pragma unroll
    for(int a = 1; a <= 8; a++)
    {
        const int sn = ccu_sidx(g, ey + OFFS[a][2], ex + OFFS[a][1], ez + OFFS[a][0]);
        s += (double)__ldg(gg + (3 * (a - 1)) * nel) * U[sn] + (double)__ldg(gg + (3 * (a - 1) + 1) * nel) * U[(size_t)g.NS + sn] +
             (double)__ldg(gg + (3 * (a - 1) + 2) * nel) * U[2 * (size_t)g.NS + sn];
    }
    divU[e] = s;
}
// assemble_grad_p (Element_calculations.c:727-769) as a gather: one thread per node sums its <= 8
// elements in ascending element order (the order the reference's element loop adds them), then
// strips the boundary dofs.
__global__ void __launch_bounds__(128) ccu_k_grad_p(const CcuGeom g, const float *__restrict__ eltT, const unsigned char *__restrict__ flags,
                                                     const double *__restrict__ P, double *gradP)
{
    constexpr int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= 8 * g.NC) return;
    const int c = t / g.NC, cell = t - c * g.NC;
    int i, j, k;
    if(!ccu_decode(g, c, cell, i, j, k)) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                const int a = LUT[k - ez][j - ex][i - ey];
                const double p = P[e];
                const float *gg = eltT + (size_t)(3 * (a - 1)) * g.nel + e;
                s0 += (double)__ldg(gg) * p; s1 += (double)__ldg(gg + g.nel) * p; s2 += (double)__ldg(gg + 2 * (size_t)g.nel) * p;
            }
        }
    }
    const int s = c * g.NC + cell;
    const unsigned char f = flags[s];
    gradP[s] = (f & CCU_F_VBX) ? 0.0 : s0;
    gradP[(size_t)g.NS + s] = (f & CCU_F_VBY) ? 0.0 : s1;
    gradP[2 * (size_t)g.NS + s] = (f & CCU_F_VBZ) ? 0.0 : s2;
}

// ---------------------------------------------------------------- replicated coarse levels (multi-subdomain runs)
// Block decomposition arithmetic (Parallel_related.c:139-170): subdomain (px, py, pz) holds elements
// [p*el_local, (p+1)*el_local) per axis; rank = pz + npz*px + npz*npx*py.
struct CcuAgg { int npx, npy, npz, mex, mey, mez; };
// all subdomains' level vectors (colour layout, gathered in rank order) -> the global level vector
__global__ void __launch_bounds__(128) ccu_k_agg_scatter_vec(const CcuGeom gl, const CcuGeom gg, const CcuAgg a, const double *__restrict__ buf, double *gvec)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= gg.nno) return;
    const int K = n % gg.noz, J = (n / gg.noz) % gg.nox, I = n / (gg.noz * gg.nox);
    const int px = min(J / gl.elx, a.npx - 1), py = min(I / gl.ely, a.npy - 1), pz = min(K / gl.elz, a.npz - 1);
    const int r = pz + a.npz * px + a.npz * a.npx * py;
    const int sl = ccu_sidx(gl, I - py * gl.ely, J - px * gl.elx, K - pz * gl.elz), sg = ccu_sidx(gg, I, J, K);
    const double *src = buf + (size_t)r * 3 * gl.NS;
#pragma unroll
    for(int d = 0; d < 3; d++) gvec[(size_t)d * gg.NS + sg] = src[(size_t)d * gl.NS + sl];
}
// this subdomain's piece of a global level vector
__global__ void __launch_bounds__(128) ccu_k_agg_extract_vec(const CcuGeom gl, const CcuGeom gg, const CcuAgg a, const double *__restrict__ gvec, double *lvec)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= gl.nno) return;
    const int k = n % gl.noz, j = (n / gl.noz) % gl.nox, i = n / (gl.noz * gl.nox);
    const int sl = ccu_sidx(gl, i, j, k), sg = ccu_sidx(gg, i + a.mey * gl.ely, j + a.mex * gl.elx, k + a.mez * gl.elz);
#pragma unroll
    for(int d = 0; d < 3; d++) lvec[(size_t)d * gl.NS + sl] = gvec[(size_t)d * gg.NS + sg];
}
// element viscosities EVI[nel*8] of all subdomains -> global element order
__global__ void __launch_bounds__(128) ccu_k_agg_scatter_evi(const CcuGeom gl, const CcuGeom gg, const CcuAgg a, const float *__restrict__ buf, float *gEVI)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= gg.nel) return;
    const int ez = e % gg.elz, ex = (e / gg.elz) % gg.elx, ey = e / (gg.elz * gg.elx);
    const int px = ex / gl.elx, py = ey / gl.ely, pz = ez / gl.elz;
    const int r = pz + a.npz * px + a.npz * a.npx * py;
    const int el = (ez - pz * gl.elz) + gl.elz * ((ex - px * gl.elx) + gl.elx * (ey - py * gl.ely));
    const float *src = buf + ((size_t)r * gl.nel + el) * 8;
#pragma unroll
    for(int q = 0; q < 8; q++) gEVI[(size_t)e * 8 + q] = src[q];
}
