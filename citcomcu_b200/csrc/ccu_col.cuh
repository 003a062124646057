// ccu_col.cuh -- column-resident Gauss-Seidel smoother and matvec for the large multigrid levels (sm_100a).
//
// Why.  The colour-pass kernels (ccu_k_relax_tab, ccu_kernels.cuh) keep the reference's half-stored stiffness
// (Eqn_k1-3: own 3x3 block + 13 lower-neighbour blocks per node, 504 B/node) and gather every row, so each stored
// block crosses HBM twice per sweep -- once as an "own" block of the node that stores it, once transposed for the lower
// neighbour -- plus uncached neighbour values: 2.5x the 600 B/node a sweep has to move (ncu, profiles/r01_*).
// Here a CTA owns a COLUMN of TI x TJ nodes over all z layers and marches through it layer by layer:
//   * the stiffness of a layer arrives as ONE bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA engine) of a
//     chunk that was laid out in HBM exactly as it is used in shared memory (ccu_col_index.h); a ring of S chunks
//     holds the layers k-1, k, k+1 the row products of layer k read and S-3 layers in flight;
//   * the solution values of the column plus a one-node rim, and the right-hand side, live in two more rings of S
//     layers in shared memory, filled two layers ahead by plain loads (their sectors are shared by 8 layers: L2 hits);
//   * per layer, the products with the layers below and above (2/3 of a row, independent of this layer's colour phases)
//     are formed for all four colours at once; a colour phase then only adds the same-layer blocks;
//   * a warp relaxes one node: lane (d, q) multiplies row d of the three 3x3 blocks in the directions
//     (q/3-1, q%3-1, -1|0|+1), nine lanes fold their sums with four shuffles, lane q = 0 applies the reference's
//     update (scalar BI per equation, correction rounded to fp32, General_matrix_functions.c:1250-1259);
//   * every stiffness byte is read from HBM once per sweep (+ the duplicated halo blocks, 24 % for 8 x 4 columns).
// Gauss-Seidel order: columns are 4-coloured by the parity of their column indices, one launch per column colour
// (3..0); inside a column z ascends; inside a layer the four (y, x)-parity colours 3..0.  All nodes relaxed
// concurrently share no stencil neighbour, so this is a Gauss-Seidel ordering of the same point-block smoother
// (General_matrix_functions.c:1231-1260); oracle/restate.c `ccu_r_ordered_gs` mode 10 states it on the CPU and
// contracts like the lexicographic order inside the multigrid cycle (tests/test_oracle_restate.py).
// MODE 1 / 2 run the same march without colours: Au = K u, or rhs - K u with boundary rows stripped
// (n_assemble_del2_u, Element_calculations.c:552).
#pragma once
#include "ccu_kernels.cuh"
#include "ccu_col_index.h"

template <int TI_, int TJ_, int S_>
struct CcuColShape
{
    static constexpr int TI = TI_, TJ = TJ_, S = S_;
    static constexpr int NT = TI * TJ;                    // nodes of one layer of a full column
    static constexpr int NW = NT / 4;                     // warps = nodes of one in-plane colour
    static constexpr int THREADS = NW * 32;
    static constexpr int BJ = TJ + 2, BOX = (TI + 2) * BJ; // solution window of one layer (column + rim)
    static constexpr int NH = 9 * TI + 9 * TJ;
    static constexpr int CHUNK = (24 * NT + 36 * (14 * NT + NH) + NT + 15) & ~15;   // bytes of a full column's chunk
    static constexpr int XLAYER = 3 * BOX * 8;            // bytes of one layer of the solution ring
    static constexpr int FLAYER = 3 * NT * 8;             // bytes of one layer of the right-hand-side ring
    static constexpr size_t SMEM = (size_t)S * CHUNK + (size_t)S * XLAYER + (size_t)S * FLAYER + (size_t)S * 8;
    static_assert(TI_ % 2 == 0 && TJ_ % 2 == 0, "column extents must be even (in-plane colours)");
    static_assert(3 * BOX <= THREADS, "one thread per solution-window entry");
    static_assert(S_ >= 4, "ring: three layers in use and at least one in flight");
};

struct CcuColArgs
{
    CcuGeom g;
    const unsigned char *Kc;      // chunks: column-major, per column layers -1 .. noz
    const size_t *colofs;         // [nI * nJ] byte offset of a column's first chunk
    const double *F;              // MODE 0: right-hand side; MODE 2: rhs of the residual
    double *x;                    // MODE 0: solution (in/out); MODE 1, 2: the vector to multiply
    double *out;                  // MODE 1, 2: result
    int nI, nJ;                   // columns along y and x
    int cc;                       // MODE 0: column colour of this launch
    int strip;                    // MODE 1: zero the boundary rows of the product
};

__device__ __forceinline__ unsigned ccu_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ccu_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ccu_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ccu_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Returns v, but the compiler may not assume the value is the same as last time: keeps it from hoisting the dozens of
// loop-invariant shared-memory addresses derived from v out of the layer loop (they spilled to local memory, and every
// reload was an L2 round trip in front of an LDS -- ncu r02: 35 % long-scoreboard stalls).
__device__ __forceinline__ int ccu_opaque(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ void ccu_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do
    {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while(!ok);
}

// sum of r over the nine direction lanes q = 0..8 of a row, valid on lane q = 0 (fixed order: deterministic)
__device__ __forceinline__ double ccu_col_fold9(double r, const int q)
{
    const double e = __shfl_down_sync(0xffffffffu, r, 8);          // lane 0 <- lane 8, off the critical path of the tree below
    double o;
    o = __shfl_down_sync(0xffffffffu, r, 4); if(q < 4) r += o;
    o = __shfl_down_sync(0xffffffffu, r, 2); if(q < 2) r += o;
    o = __shfl_down_sync(0xffffffffu, r, 1); if(q < 1) r += o;
    return r + e;
}

template <class SH, int MODE>
__global__ void __launch_bounds__(SH::THREADS, (SH::SMEM <= 113 * 1024 ? 2 : 1)) ccu_k_col(const __grid_constant__ CcuColArgs A)
{
    constexpr int TI = SH::TI, TJ = SH::TJ, S = SH::S, BJ = SH::BJ, BOX = SH::BOX, CH = SH::CHUNK, XL = SH::XLAYER, FL = SH::FLAYER;
    extern __shared__ __align__(128) unsigned char ccu_col_smem[];
    unsigned char *stg = ccu_col_smem;                                   // [S][CH] stiffness chunks
    unsigned char *xrb = stg + (size_t)S * CH;                           // [S][3][BOX] doubles, solution ring
    unsigned char *frb = xrb + (size_t)S * XL;                           // [S][3][NT] doubles, right-hand-side ring
    const unsigned bar0 = ccu_smem_u32(frb + (size_t)S * FL);            // [S] mbarriers
    const CcuGeom &g = A.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int I, J;
    if(MODE == 0)
    {
        const int ci = A.cc >> 1, cj = A.cc & 1, nJc = (A.nJ - cj + 1) / 2;
        I = 2 * ((int)blockIdx.x / nJc) + ci; J = 2 * ((int)blockIdx.x % nJc) + cj;
    }
    else { I = (int)blockIdx.x / A.nJ; J = (int)blockIdx.x % A.nJ; }
    const int i0 = I * TI, j0 = J * TJ, noz = g.noz;
    const CcuColDims cd = ccu_col_dims(min(TI, g.noy - i0), min(TJ, g.nox - j0));
    const unsigned char *chunks = A.Kc + A.colofs[I * A.nJ + J];
    const size_t NS = (size_t)g.NS;

    // ---- per-thread constants: what this lane reads in each in-plane colour (ccu_col_index.h)
    const bool act = lane < 27;
    const int d = act ? lane / 9 : 0, q = act ? lane % 9 : 0;
    const bool upd = act && q == 0;
    const int wa = warp / (TJ / 2), wb = warp % (TJ / 2);
    int kof[4][3], xof[4], nodeA[4], xself[4], pidx[4];
    bool nv[4];
    bool tr[3];
#pragma unroll
    for(int t = 0; t < 3; t++)
    {
        const int di = q / 3 - 1, dj = q % 3 - 1, dk = t - 1;
        tr[t] = !((di == 0 && dj == 0 && dk == 0) || ccu_lo_index(di, dj, dk) >= 0);
    }
#pragma unroll
    for(int c2 = 0; c2 < 4; c2++)
    {
        const int li = 2 * wa + (c2 >> 1), lj = 2 * wb + (c2 & 1);
        nv[c2] = li < cd.ti && lj < cd.tj;
#pragma unroll
        for(int t = 0; t < 3; t++)
        {
            CcuColDesc ds = { 0, 0, 0 };
            if(nv[c2]) ds = ccu_col_desc(cd, TJ, li, lj, d, q, t);
            kof[c2][t] = ds.kof; xof[c2] = ds.xof;
        }
        const int gi = i0 + li, gj = j0 + lj;
        nodeA[c2] = (4 * (gi & 1) + 2 * (gj & 1)) * g.NC + ((gi >> 1) + 1) * g.JK + ((gj >> 1) + 1) * g.Kd + 1;
        xself[c2] = (li + 1) * BJ + (lj + 1);
        pidx[c2] = li * cd.tj + lj;
    }
    // ring loaders: thread `tid` owns entry (dx, bn) of every layer of the solution window and entry (fd, fp) of the rhs ring
    const bool xl = tid < 3 * BOX;
    const int dx = xl ? tid / BOX : 0, bn = xl ? tid % BOX : 0;
    const int xgi = i0 + bn / BJ - 1, xgj = j0 + bn % BJ - 1;
    const bool xin = xl && xgi >= 0 && xgi < g.noy && xgj >= 0 && xgj < g.nox;
    const size_t xA = xin ? (size_t)dx * NS + (size_t)((4 * (xgi & 1) + 2 * (xgj & 1)) * g.NC + ((xgi >> 1) + 1) * g.JK + ((xgj >> 1) + 1) * g.Kd + 1) : 0;
    const bool fl = MODE != 1 && tid < 3 * cd.nt;
    const int fd = fl ? tid / cd.nt : 0, fp = fl ? tid % cd.nt : 0;
    const int fgi = i0 + fp / cd.tj, fgj = j0 + fp % cd.tj;
    const size_t fA = fl ? (size_t)fd * NS + (size_t)((4 * (fgi & 1) + 2 * (fgj & 1)) * g.NC + ((fgi >> 1) + 1) * g.JK + ((fgj >> 1) + 1) * g.Kd + 1) : 0;
    auto xload = [&](int k) -> double { return (xin && k >= 0 && k < noz) ? A.x[xA + (size_t)((k & 1) * g.NC + (k >> 1))] : 0.0; };
    auto fload = [&](int k) -> double { return (fl && k < noz) ? A.F[fA + (size_t)((k & 1) * g.NC + (k >> 1))] : 0.0; };
    auto xslot = [&](int slot) -> double * { return (double *)(xrb + (size_t)slot * XL) + dx * BOX + bn; };
    auto fslot = [&](int slot) -> double * { return (double *)(frb + (size_t)slot * FL) + fd * cd.nt + fp; };
    auto issue = [&](int layer)        // thread 0: bulk copy of the chunk of `layer` (-1 .. noz) into its ring stage
    {
        const int s = (layer + S) % S;
        ccu_mbar_expect_tx(bar0 + 8 * s, (unsigned)cd.cb);
        ccu_bulk_g2s(ccu_smem_u32(stg + (size_t)s * CH), chunks + (size_t)(layer + 1) * cd.cb, (unsigned)cd.cb, bar0 + 8 * s);
    };

    // ---- prologue: barriers, the first S chunks in flight, layers -1, 0, 1 of the solution ring, 0, 1 of the rhs ring
    if(tid == 0)
    {
        for(int s = 0; s < S; s++) ccu_mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if(tid == 0)
        for(int layer = -1; layer <= S - 2 && layer <= noz; layer++) issue(layer);
    if(xl)
    {
        *xslot(S - 1) = xload(-1);
        *xslot(0) = xload(0);
        *xslot(1 % S) = xload(1);
    }
    if(fl)
    {
        *fslot(0) = fload(0);
        *fslot(1 % S) = fload(1);
    }
    ccu_mbar_wait(bar0 + 8 * (S - 1), 0);        // layer -1
    ccu_mbar_wait(bar0, 0);                      // layer 0
    __syncthreads();

    for(int k0 = 0; k0 < noz; k0 += S)
    {
#pragma unroll
        for(int JJ = 0; JJ < S; JJ++)
        {
            const int k = k0 + JJ;               // k % S == JJ: ring positions are compile-time constants below
            if(k >= noz) break;
            const double xpre = xload(k + 2), fpre = fload(k + 2);   // land in the rings before this layer's last barrier
            ccu_mbar_wait(bar0 + 8 * ((JJ + 1) % S), (unsigned)(((k + 2) / S) & 1));     // chunk of layer k + 1
            const int zoff = (k & 1) * g.NC + (k >> 1);
            const unsigned char *own = stg + (size_t)JJ * CH;
            // product of this lane's block in direction layer t - 1 with the neighbour's values, for a node of colour c2
            auto prod = [&](const int c2, const int t) -> double
            {
                const int ring = (JJ + t - 1 + S) % S;               // layer k + t - 1
                const unsigned char *kb = stg + (ccu_opaque(kof[c2][t]) + (tr[t] ? ring * CH : JJ * CH));
                const int st = ccu_opaque(tr[t] ? 12 : 4);
                const float c0 = *(const float *)kb, c1 = *(const float *)(kb + st), c2f = *(const float *)(kb + 2 * st);
                const double *xp = (const double *)(xrb + (ring * XL + ccu_opaque(xof[c2])));
                return (double)c0 * xp[0] + (double)c1 * xp[BOX] + (double)c2f * xp[2 * BOX];
            };
            // Everything a row needs from the layers below and above is independent of this layer's colour phases:
            // all four colours at once, before the phases (MODE 1, 2: the whole row)
            double acc[4];
#pragma unroll
            for(int c2 = 0; c2 < 4; c2++)
            {
                acc[c2] = 0.0;
                if(nv[c2] && act) acc[c2] = MODE == 0 ? prod(c2, 0) + prod(c2, 2) : (prod(c2, 0) + prod(c2, 2)) + prod(c2, 1);
            }
            if(MODE == 0)
            {
#pragma unroll
                for(int ph = 0; ph < 4; ph++)
                {
                    const int c2 = 3 - ph;
                    const bool v = nv[c2];                         // warp-uniform
                    if(v)
                    {
                        double r = acc[c2];
                        if(act) r += prod(c2, 1);
                        r = ccu_col_fold9(r, q);
                        if(upd)
                        {   // General_matrix_functions.c:1250-1259: scalar BI per equation, correction rounded to fp32
                            const double bi = ((const double *)own)[d * cd.nt + pidx[c2]];
                            const double Fv = ((const double *)(frb + (size_t)JJ * FL))[d * cd.nt + pidx[c2]];
                            double *xs = (double *)(xrb + (size_t)JJ * XL) + d * BOX + xself[c2];
                            const double xn = *xs + (double)(float)((Fv - r) * bi);
                            *xs = xn;
                            A.x[(size_t)d * NS + (size_t)(nodeA[c2] + zoff)] = xn;
                        }
                    }
                    if(ph == 3)
                    {
                        if(xl) *xslot((JJ + 2) % S) = xpre;
                        if(fl) *fslot((JJ + 2) % S) = fpre;
                    }
                    __syncthreads();
                }
            }
            else
            {
#pragma unroll
                for(int c2 = 0; c2 < 4; c2++)
                {
                    if(!nv[c2]) continue;                          // warp-uniform
                    double r = ccu_col_fold9(acc[c2], q);
                    if(upd)
                    {
                        const unsigned char flg = own[cd.flofs + pidx[c2]];
                        if((MODE == 2 || A.strip) && ((flg >> d) & 1)) r = 0.0;
                        if(MODE == 2) r = ((const double *)(frb + (size_t)JJ * FL))[d * cd.nt + pidx[c2]] - r;
                        A.out[(size_t)d * NS + (size_t)(nodeA[c2] + zoff)] = r;
                    }
                }
                if(xl) *xslot((JJ + 2) % S) = xpre;
                if(fl) *fslot((JJ + 2) % S) = fpre;
                __syncthreads();
            }
            if(tid == 0 && k + S - 1 <= noz) issue(k + S - 1);       // the stage of layer k - 1 is free now
        }
    }
}

// K (coefficient-major, colour-blocked), BI, flags -> column chunks (ccu_col_fill_chunk).  grid = (columns, groups of 32
// layers); a warp takes 32 consecutive layers of one block so that its reads of K run along z (unit stride per colour).
template <int TI, int TJ>
__global__ void __launch_bounds__(256) ccu_k_col_relayout(const CcuGeom g, const int nJ, const size_t *__restrict__ colofs,
                                                           const float *__restrict__ K, const double *__restrict__ BI,
                                                           const unsigned char *__restrict__ flags, const unsigned char *__restrict__ bits,
                                                           unsigned char *Kc)
{
    const int col = blockIdx.x, I = col / nJ, J = col % nJ;
    const int kk = (int)blockIdx.y * 32 + (int)(threadIdx.x & 31);                   // chunk index 0 .. noz + 1 = layer -1 .. noz
    if(kk > g.noz + 1) return;
    const int i0 = I * TI, j0 = J * TJ;
    const CcuColDims cd = ccu_col_dims(min(TI, g.noy - i0), min(TJ, g.nox - j0));
    ccu_col_fill_chunk(g, cd, i0, j0, kk - 1, K, BI, flags, bits, Kc + colofs[col] + (size_t)kk * cd.cb, threadIdx.x >> 5, blockDim.x >> 5);
}
