// ccu_tile.cuh -- tile-resident smoother / matvec for the large multigrid levels (sm_100a).
//
// Why: the colour-pass kernels (ccu_k_relax_tab, ccu_kernels.cuh) stream the half-stored stiffness twice per sweep --
// block K_nm is needed by the pass that relaxes n (as an "own" block) and by the pass that relaxes m (transposed) -- and
// re-fetch every neighbour value from L2: ncu r01 shows 1.58 GB DRAM and 2.4 GB L2->SM per colour launch at
// 256x256x128 against 0.69 GB algorithmic, DRAM saturated.  Here one CTA owns a TILE of TI x TJ x TK colour cells
// (all eight node colours: 2TI x 2TJ x 2TK nodes) and
//   * keeps the tile's solution values plus a one-cell halo in shared memory (every neighbour read is an LDS),
//   * runs the eight node-colour passes of the tile back to back, so the second use of a stiffness block follows the
//     first by microseconds and is served by L2 (first use loaded with L2::evict_last, second with L2::evict_first),
//   * splits a row's 27 blocks over Q warps-groups (partial sums folded through shared memory in a fixed order), which
//     keeps 1024 threads per CTA busy although a colour pass touches only TI*TJ*TK nodes.
// Gauss-Seidel ordering: tiles are 8-coloured by the parity of their tile indices; one launch relaxes all tiles of one
// tile colour (they share no stencil neighbour), tile colours 7..0, node colours 7..0 inside a tile.  This is a valid
// Gauss-Seidel ordering of the same point-block smoother (General_matrix_functions.c:1231-1260); oracle/restate.c
// `ccu_r_ordered_gs` mode 9 is its CPU statement, and converges slightly faster than the plain 8-colour order.
#pragma once
#include "ccu_kernels.cuh"

template <int TI_, int TJ_, int TK_, int Q_>
struct CcuTileShape
{
    static constexpr int TI = TI_, TJ = TJ_, TK = TK_, Q = Q_;
    static constexpr int CT = TI * TJ * TK;                      // cells (= nodes of one colour) per tile
    static constexpr int SJ = TJ + 2, SK = TK + 2, SJK = SJ * SK;
    static constexpr int BOX = (TI + 2) * SJK;                   // halo box of one colour in shared memory
    static constexpr int THREADS = CT * Q;
    static constexpr size_t SMEM = sizeof(double) * (3 * 8 * BOX + Q * 3 * CT);
};

struct CcuTileTab
{
    int goff[8][27];            // storage-slot offset of block b's neighbour (as CcuStencil)
    int soff[8][27];            // the same neighbour in the shared-memory box: cm*BOX + si*SJK + sj*SK + sk
    unsigned char cm[8][27];    // its colour
};
template <class S>
__host__ inline CcuTileTab ccu_make_tile_tab(const CcuGeom &g)
{
    const int LO[13][3] = CCU_LO_INIT;
    CcuTileTab t;
    for(int c = 0; c < 8; c++)
    {
        const int pi = (c >> 2) & 1, pj = (c >> 1) & 1, pk = c & 1;
        for(int b = 0; b < 27; b++)
        {
            int di = 0, dj = 0, dk = 0;
            if(b >= 1 && b <= 13) { di = LO[b - 1][0]; dj = LO[b - 1][1]; dk = LO[b - 1][2]; }
            if(b >= 14) { di = -LO[b - 14][0]; dj = -LO[b - 14][1]; dk = -LO[b - 14][2]; }
            const int cm = c ^ (((di != 0) << 2) | ((dj != 0) << 1) | (dk != 0));
            const int si = ccu_shift(pi, di), sj = ccu_shift(pj, dj), sk = ccu_shift(pk, dk);
            t.goff[c][b] = (cm - c) * g.NC + si * g.JK + sj * g.Kd + sk;
            t.soff[c][b] = cm * S::BOX + si * S::SJK + sj * S::SK + sk;
            t.cm[c][b] = (unsigned char)cm;
        }
    }
    return t;
}

__device__ __forceinline__ float ccu_ldk(const float *p, const unsigned long long pol)
{
    float v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

// MODE 0: relax the tiles of tile colour `tcol` (x updated in place; F, BI read)
// MODE 1: out = K x over all tiles (rows flagged in `fl` zeroed when strip)
// MODE 2: out = F - K x with the flagged rows of K x zeroed first (res = rhs - AU, AU stripped)
// `fl`: MODE 0 the duplicated-node bits of multi-subdomain runs (or null), MODE 1/2 the boundary-condition flag bytes.
template <class S, int MODE>
__global__ void __launch_bounds__(S::THREADS, (S::SMEM <= 110 * 1024 && S::THREADS <= 1024) ? 2 : 1)
ccu_k_tile(const CcuGeom g, const __grid_constant__ CcuTileTab tab, const int tcol, const float *__restrict__ K,
           const double *__restrict__ BI, const double *__restrict__ F, double *x, double *out,
           const unsigned char *__restrict__ fl, const int strip, const int hint)
{
    constexpr int TI = S::TI, TJ = S::TJ, TK = S::TK, Q = S::Q, CT = S::CT, SJK = S::SJK, SK = S::SK, BOX = S::BOX;
    extern __shared__ double ccu_tile_smem[];
    double *xs = ccu_tile_smem;                  // [3][8][BOX]
    double *ps = ccu_tile_smem + 3 * 8 * BOX;    // [Q][3][CT]
    const int tid = threadIdx.x;
    const size_t NS = (size_t)g.NS;

    int tk = blockIdx.x, tj = blockIdx.y, ti = blockIdx.z;
    if(MODE == 0) { tk = 2 * tk + (tcol & 1); tj = 2 * tj + ((tcol >> 1) & 1); ti = 2 * ti + ((tcol >> 2) & 1); }
    const int ic0 = ti * TI, jc0 = tj * TJ, kc0 = tk * TK;

    unsigned long long pol_keep, pol_drop;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    if(!hint) { asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep)); pol_drop = pol_keep; }

    // the tile's values and a one-cell halo, all eight colours, into shared memory (cells outside the storage box: 0)
    for(int idx = tid; idx < 3 * 8 * BOX; idx += S::THREADS)
    {
        const int d = idx / (8 * BOX), r = idx - d * (8 * BOX);
        const int cm = r / BOX, r2 = r - cm * BOX;
        const int a = r2 / SJK, r3 = r2 - a * SJK, b = r3 / SK, cc = r3 - b * SK;
        const int ic = ic0 - 1 + a, jc = jc0 - 1 + b, kc = kc0 - 1 + cc;
        double v = 0.0;
        if(ic >= 0 && ic < g.Id && jc >= 0 && jc < g.Jd && kc >= 0 && kc < g.Kd)
            v = x[d * NS + (size_t)cm * g.NC + (size_t)ic * g.JK + jc * g.Kd + kc];
        xs[idx] = v;
    }
    __syncthreads();

    const int q = tid / CT, lc = tid - q * CT;
    const int li = lc / (TJ * TK), lr = lc - li * (TJ * TK), lj = lr / TK, lk = lr - lj * TK;
    int cell;
    unsigned vmask = 0;                          // bit c: this thread's cell holds a node of colour c
    {
        const int ic = ic0 + li, jc = jc0 + lj, kc = kc0 + lk;
        cell = ic * g.JK + jc * g.Kd + kc;
        if(ic >= 1 && jc >= 1 && kc >= 1 && ic < g.Id && jc < g.Jd && kc < g.Kd)
            for(int c = 0; c < 8; c++)
                if(2 * (ic - 1) + ((c >> 2) & 1) < g.noy && 2 * (jc - 1) + ((c >> 1) & 1) < g.nox && 2 * (kc - 1) + (c & 1) < g.noz) vmask |= 1u << c;
    }
    const int sbase = (li + 1) * SJK + (lj + 1) * SK + lk + 1;
    const int qq = Q - 1 - q;                    // block group: the updating threads (q < 3) get the short groups

#pragma unroll 1
    for(int pass = 0; pass < 8; pass++)
    {
        const int c = (MODE == 0) ? 7 - pass : pass;
        const int s = c * g.NC + cell;
        bool valid = (vmask >> c) & 1u;
        if(MODE == 0 && fl && valid) valid = !(fl[s] & CCU_B_SHARED);
        double fq = 0.0, bq = 0.0;
        if(q < 3 && valid)
        {
            if(MODE != 1) fq = F[q * NS + s];
            if(MODE == 0) bq = BI[q * NS + s];
        }
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        if(valid)
        {
#pragma unroll 1
            for(int b = qq; b < 27; b += Q)
            {
                const bool tr = b >= 14;
                const int cm = tab.cm[c][b];
                // first of the two uses of this block inside the tile -> keep it in L2; second (or only) use -> let it go
                const bool first = (b != 0) && ((MODE == 0) ? (c > cm) : (c < cm));
                const unsigned long long pol = first ? pol_keep : pol_drop;
                const float *Kp = K + (size_t)((tr ? b - 13 : b) * 9) * NS + s + (tr ? tab.goff[c][b] : 0);
                float k[9];
#pragma unroll
                for(int e = 0; e < 9; e++) k[e] = ccu_ldk(Kp + (size_t)e * NS, pol);
                const int sm = sbase + tab.soff[c][b];
                const double x0 = xs[sm], x1 = xs[8 * BOX + sm], x2 = xs[16 * BOX + sm];
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
        ps[(q * 3 + 0) * CT + lc] = r0;
        ps[(q * 3 + 1) * CT + lc] = r1;
        ps[(q * 3 + 2) * CT + lc] = r2;
        __syncthreads();
        if(q < 3 && valid)
        {   // thread (q, node) owns equation q of the node: fold the Q partial rows in a fixed order
            double a = 0.0;
#pragma unroll
            for(int w = 0; w < Q; w++) a += ps[(w * 3 + q) * CT + lc];
            if(MODE == 0)
            {   // General_matrix_functions.c:1250-1259: scalar BI per equation, correction rounded to fp32
                const float t = (float)((fq - a) * bq);
                const int sx = q * 8 * BOX + c * BOX + sbase;
                const double xn = xs[sx] + (double)t;
                xs[sx] = xn;
                x[q * NS + s] = xn;
            }
            else
            {
                if(strip && (fl[s] & (q == 0 ? CCU_F_VBX : (q == 1 ? CCU_F_VBY : CCU_F_VBZ)))) a = 0.0;
                out[q * NS + s] = (MODE == 1) ? a : fq - a;
            }
        }
        __syncthreads();
    }
}
