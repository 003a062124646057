// ccu_tile.cuh -- tile-resident smoother / matvec for the large multigrid levels (sm_100a).
//
// Why: the colour-pass kernels (ccu_k_relax_tab, ccu_kernels.cuh) stream the half-stored stiffness twice per sweep --
// block K_nm is needed by the pass that relaxes n (as an "own" block) and by the pass that relaxes m (transposed) -- and
// re-fetch every neighbour value from L2: ncu r01 shows 1.58 GB DRAM and 2.4 GB L2->SM per colour launch at
// 256x256x128 against 0.69 GB algorithmic, DRAM saturated.  Here one CTA owns a TILE of TI x TJ x TK colour cells
// (all eight node colours: 2TI x 2TJ x 2TK nodes) and
//   * keeps the tile's solution values plus a one-cell halo in shared memory (every neighbour read is an LDS; the box
//     arrives by cp.async, all requests in flight at once),
//   * runs the eight node-colour passes of the tile back to back, so the second use of a stiffness block follows the
//     first by microseconds and is served by L2 (first use loaded with L2::evict_last, second with L2::evict_first),
//   * reads the stiffness from a TILE-MAJOR copy Kt[tile][colour][14*9 coefficient planes][cell]: what one pass needs
//     is one contiguous 14*9*CT*4-byte run plus 512-byte runs of the neighbours' blocks, all inside one 2 MB page --
//     the coefficient-major array K[plane][slot] scatters a tile over 126 planes 40 MB apart (64-byte pieces, a TLB miss
//     and a DRAM row miss each; measured 3x SLOWER than the colour passes, profiles/r01_tile_v1_probe.json),
//   * gives every thread V consecutive z cells (128-/64-bit loads), splits a row's 27 blocks over Q groups of warps
//     (partial sums folded through shared memory in a fixed order), and double-buffers the blocks in registers ACROSS
//     the colour passes: the first block of the next pass does not depend on this pass's result, so it is in flight
//     during the two barriers of the fold.
// Gauss-Seidel ordering: tiles are 8-coloured by the parity of their tile indices; one launch relaxes all tiles of one
// tile colour (they share no stencil neighbour), tile colours 7..0, node colours 7..0 inside a tile.  This is a valid
// Gauss-Seidel ordering of the same point-block smoother (General_matrix_functions.c:1231-1260); oracle/restate.c
// `ccu_r_ordered_gs` mode 9 is its CPU statement; per multigrid cycle it contracts within +-10 % of the plain 8-colour order.
#pragma once
#include "ccu_kernels.cuh"

template <int TI_, int TJ_, int TK_, int Q_, int V_>
struct CcuTileShape
{
    static constexpr int TI = TI_, TJ = TJ_, TK = TK_, Q = Q_, V = V_;
    static constexpr int CT = TI * TJ * TK;                      // cells (= nodes of one colour) per tile
    static constexpr int SJ = TJ + 2, SK = TK + 2, SJK = SJ * SK;
    static constexpr int BOX = (TI + 2) * SJK;                   // halo box of one colour in shared memory
    static constexpr int GT = CT / V;                            // threads of one block group
    static constexpr int THREADS = GT * Q;
    static constexpr size_t SMEM = sizeof(double) * (3 * 8 * BOX + Q * 3 * CT);
    static_assert(TK_ % V_ == 0 && GT % 32 == 0, "a block group must be whole warps of V-cell threads");
};

struct CcuTileTab
{
    int soff[8][27];            // block b's neighbour in the shared-memory box: cm*BOX + si*SJK + sj*SK + sk
    signed char sh[8][27][3];   // its cell shift (si, sj, sk)
    unsigned char cm[8][27];    // its colour
    int ntj, ntk;               // tiles along x and z (tile id = (ti*ntj + tj)*ntk + tk)
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
            t.soff[c][b] = cm * S::BOX + si * S::SJK + sj * S::SK + sk;
            t.sh[c][b][0] = (signed char)si; t.sh[c][b][1] = (signed char)sj; t.sh[c][b][2] = (signed char)sk;
            t.cm[c][b] = (unsigned char)cm;
        }
    }
    t.ntj = (g.Jd + S::TJ - 1) / S::TJ;
    t.ntk = (g.Kd + S::TK - 1) / S::TK;
    return t;
}
template <class S>
__host__ inline size_t ccu_tile_elems(const CcuGeom &g)
{
    const size_t nt = (size_t)((g.Id + S::TI - 1) / S::TI) * ((g.Jd + S::TJ - 1) / S::TJ) * ((g.Kd + S::TK - 1) / S::TK);
    return nt * 8 * 126 * S::CT;
}

// K[plane][slot] -> Kt[tile][colour][plane][cell]; cells of a tile that lie outside the colour box hold zeros
template <class S>
__global__ void __launch_bounds__(S::CT) ccu_k_tile_relayout(const CcuGeom g, const int ntj, const int ntk, const float *__restrict__ K, float *__restrict__ Kt)
{
    const int tile = blockIdx.x >> 3, c = blockIdx.x & 7, lc = threadIdx.x;
    const int tk = tile % ntk, tj = (tile / ntk) % ntj, ti = tile / (ntk * ntj);
    const int li = lc / (S::TJ * S::TK), lr = lc - li * (S::TJ * S::TK), lj = lr / S::TK, lk = lr - lj * S::TK;
    const int ic = ti * S::TI + li, jc = tj * S::TJ + lj, kc = tk * S::TK + lk;
    const bool in = ic < g.Id && jc < g.Jd && kc < g.Kd;
    const size_t s = (size_t)c * g.NC + (size_t)ic * g.JK + jc * g.Kd + kc;
    float *dst = Kt + ((size_t)blockIdx.x * 126) * S::CT + lc;
#pragma unroll 6
    for(int p = 0; p < 126; p++) dst[(size_t)p * S::CT] = in ? __ldg(K + (size_t)p * g.NS + s) : 0.0f;
}

__device__ __forceinline__ float ccu_ldk(const float *p, const unsigned long long pol)
{
    float v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
template <int V>
__device__ __forceinline__ void ccu_ldkv(float (&o)[V], const float *p, const unsigned long long pol)
{
    if constexpr(V == 1)
        o[0] = ccu_ldk(p, pol);
    else if constexpr(V == 4)
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
            : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p), "l"(pol));
    else
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(o[0]), "=f"(o[1]) : "l"(p), "l"(pol));
}

// MODE 0: relax the tiles of tile colour `tcol` (x updated in place; F, BI read)
// MODE 1: out = K x over all tiles (rows flagged in `fl` zeroed when strip)
// MODE 2: out = F - K x with the flagged rows of K x zeroed first (res = rhs - AU, AU stripped)
// `fl`: MODE 0 the duplicated-node bits of multi-subdomain runs (or null), MODE 1/2 the boundary-condition flag bytes.
template <class S, int MODE>
__global__ void __launch_bounds__(S::THREADS, (S::SMEM <= 74 * 1024 && S::THREADS <= 341) ? 3 : ((S::SMEM <= 110 * 1024 && S::THREADS <= 512) ? 2 : 1))
ccu_k_tile(const CcuGeom g, const __grid_constant__ CcuTileTab tab, const int tcol, const float *__restrict__ Kt,
           const double *__restrict__ BI, const double *__restrict__ F, double *x, double *out,
           const unsigned char *__restrict__ fl, const int strip, const int hint)
{
    constexpr int TI = S::TI, TJ = S::TJ, TK = S::TK, Q = S::Q, CT = S::CT, SJK = S::SJK, SK = S::SK, BOX = S::BOX, V = S::V, GT = S::GT;
    constexpr int NST = (27 + Q - 1) / Q;        // blocks of the longest group = pipeline stages per pass
    extern __shared__ double ccu_tile_smem[];
    double *xs = ccu_tile_smem;                  // [3][8][BOX]
    double *ps = ccu_tile_smem + 3 * 8 * BOX;    // [Q][3][CT]
    const int tid = threadIdx.x;
    const size_t NS = (size_t)g.NS;

    int tk = blockIdx.x, tj = blockIdx.y, ti = blockIdx.z;
    if(MODE == 0) { tk = 2 * tk + (tcol & 1); tj = 2 * tj + ((tcol >> 1) & 1); ti = 2 * ti + ((tcol >> 2) & 1); }
    const int ic0 = ti * TI, jc0 = tj * TJ, kc0 = tk * TK;
    const int tile = (ti * tab.ntj + tj) * tab.ntk + tk;

    unsigned long long pol_keep, pol_drop;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    if(!hint) { asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep)); pol_drop = pol_keep; }

    // the tile's values and a one-cell halo, all eight colours -> shared memory; cells outside the storage box read as 0
    {
        const unsigned xs_s = (unsigned)__cvta_generic_to_shared(xs);
        for(int idx = tid; idx < 3 * 8 * BOX; idx += S::THREADS)
        {
            const int d = idx / (8 * BOX), r = idx - d * (8 * BOX);
            const int cm = r / BOX, r2 = r - cm * BOX;
            const int a = r2 / SJK, r3 = r2 - a * SJK, b = r3 / SK, cc = r3 - b * SK;
            const int ic = ic0 - 1 + a, jc = jc0 - 1 + b, kc = kc0 - 1 + cc;
            const bool ok = ic >= 0 && ic < g.Id && jc >= 0 && jc < g.Jd && kc >= 0 && kc < g.Kd;
            const double *src = ok ? x + (d * NS + (size_t)cm * g.NC + (size_t)ic * g.JK + jc * g.Kd + kc) : x;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(xs_s + 8u * idx), "l"(src), "r"(ok ? 8 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    const int q = tid / GT, lg = tid - q * GT;
    const int lc = lg * V;                       // first of this thread's V cells (one z row: TK % V == 0)
    const int li = lc / (TJ * TK), lr = lc - li * (TJ * TK), lj = lr / TK, lk = lr - lj * TK;
    int cell;
    unsigned vmask = 0;                          // bit V*c + m: cell m of this thread holds a node of colour c
    {
        const int ic = ic0 + li, jc = jc0 + lj;
        cell = ic * g.JK + jc * g.Kd + kc0 + lk;
        if(ic >= 1 && jc >= 1 && ic < g.Id && jc < g.Jd)
            for(int c = 0; c < 8; c++)
                for(int m = 0; m < V; m++)
                {
                    const int kc = kc0 + lk + m;
                    if(kc >= 1 && kc < g.Kd && 2 * (ic - 1) + ((c >> 2) & 1) < g.noy && 2 * (jc - 1) + ((c >> 1) & 1) < g.nox && 2 * (kc - 1) + (c & 1) < g.noz)
                        vmask |= 1u << (V * c + m);
                }
    }
    const int sbase = (li + 1) * SJK + (lj + 1) * SK + lk + 1;
    const int qq = Q - 1 - q;                    // block group: the updating threads (q < 3) get the short groups
    constexpr unsigned VM = (1u << V) - 1u;

    // one block of V rows: nine coefficient vectors from HBM / L2
    auto load_block = [&](float (&k)[9][V], const int pass, const int b)
    {
        const int c = (MODE == 0) ? 7 - pass : pass;
        const unsigned vm = (vmask >> (V * c)) & VM;
        if(!vm) return;                                  // no node of this colour here: nothing to fetch (and no stray addresses)
        const int cm = tab.cm[c][b];
        // first of the two uses of this block inside the tile -> keep it in L2; second (or only) use -> let it go
        const bool first = (b != 0) && ((MODE == 0) ? (c > cm) : (c < cm));
        const unsigned long long pol = first ? pol_keep : pol_drop;
        if(b < 14)
        {   // own block: the pass's 14*9 planes are one contiguous run of the tile-major array
            const float *Kp = Kt + ((size_t)(tile * 8 + c) * 126 + b * 9) * CT + lc;
#pragma unroll
            for(int e = 0; e < 9; e++) ccu_ldkv<V>(k[e], Kp + e * CT, pol);
            return;
        }
        // transposed block: stored with the upper neighbour (colour cm, cell shifted by si, sj, sk), maybe in the next tile
        int ni = li + tab.sh[c][b][0], nj = lj + tab.sh[c][b][1], t2i = ti, t2j = tj;
        const int sk = tab.sh[c][b][2];
        if(ni < 0) { ni += TI; t2i--; } else if(ni >= TI) { ni -= TI; t2i++; }
        if(nj < 0) { nj += TJ; t2j--; } else if(nj >= TJ) { nj -= TJ; t2j++; }
        const int trow = (t2i * tab.ntj + t2j) * tab.ntk, lrow = (ni * TJ + nj) * TK;
        const int p0 = (b - 13) * 9;
        if(sk == 0)
        {
            const float *Kp = Kt + ((size_t)((trow + tk) * 8 + cm) * 126 + p0) * CT + lrow + lk;
#pragma unroll
            for(int e = 0; e < 9; e++) ccu_ldkv<V>(k[e], Kp + e * CT, pol);
            return;
        }
#pragma unroll
        for(int m = 0; m < V; m++)
        {   // z-shifted neighbours: one cell at a time (the end cell may sit in the next tile along z)
            int nk = lk + m + sk, t2k = tk;
            if(nk < 0) { nk += TK; t2k--; } else if(nk >= TK) { nk -= TK; t2k++; }
            const float *Kp = Kt + ((size_t)((trow + t2k) * 8 + cm) * 126 + p0) * CT + lrow + nk;
            const bool on = (vm >> m) & 1u;
#pragma unroll
            for(int e = 0; e < 9; e++) k[e][m] = on ? ccu_ldk(Kp + e * CT, pol) : 0.0f;
        }
    };
    auto mul_block = [&](const float (&k)[9][V], const int pass, const int b, double (&r)[3][V])
    {
        const int c = (MODE == 0) ? 7 - pass : pass;
        if(!((vmask >> (V * c)) & VM)) return;
        const int sm = sbase + tab.soff[c][b];
        if(b < 14)
        {
#pragma unroll
            for(int m = 0; m < V; m++)
            {
                const double x0 = xs[sm + m], x1 = xs[8 * BOX + sm + m], x2 = xs[16 * BOX + sm + m];
                r[0][m] += (double)k[0][m] * x0 + (double)k[1][m] * x1 + (double)k[2][m] * x2;
                r[1][m] += (double)k[3][m] * x0 + (double)k[4][m] * x1 + (double)k[5][m] * x2;
                r[2][m] += (double)k[6][m] * x0 + (double)k[7][m] * x1 + (double)k[8][m] * x2;
            }
        }
        else
        {
#pragma unroll
            for(int m = 0; m < V; m++)
            {
                const double x0 = xs[sm + m], x1 = xs[8 * BOX + sm + m], x2 = xs[16 * BOX + sm + m];
                r[0][m] += (double)k[0][m] * x0 + (double)k[3][m] * x1 + (double)k[6][m] * x2;
                r[1][m] += (double)k[1][m] * x0 + (double)k[4][m] * x1 + (double)k[7][m] * x2;
                r[2][m] += (double)k[2][m] * x0 + (double)k[5][m] * x1 + (double)k[8][m] * x2;
            }
        }
    };

    float kA[9][V], kB[9][V];
    load_block(kA, 0, qq);                       // in flight while the solution box arrives
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

#pragma unroll 1
    for(int pass = 0; pass < 8; pass++)
    {
        const int c = (MODE == 0) ? 7 - pass : pass;
        const int s0 = c * g.NC + cell;
        const unsigned vm = (vmask >> (V * c)) & VM;
        double r[3][V];
#pragma unroll
        for(int d = 0; d < 3; d++)
#pragma unroll
            for(int m = 0; m < V; m++) r[d][m] = 0.0;
        // stage t multiplies block qq + t*Q out of one register buffer while the next block streams into the other;
        // the last stage fetches the first block of the NEXT pass, which then arrives during the two barriers below
        static_assert(NST % 2 == 0, "an odd stage count would leave the next pass's first block in the wrong buffer");
#pragma unroll
        for(int t = 0; t < NST; t++)
        {
            const int b = qq + t * Q, bn = b + Q;
            float (&kc_)[9][V] = (t & 1) ? kB : kA;
            float (&kn_)[9][V] = (t & 1) ? kA : kB;
            if(t + 1 < NST) { if(bn < 27) load_block(kn_, pass, bn); }
            else if(pass < 7) load_block(kn_, pass + 1, qq);
            if(b < 27) mul_block(kc_, pass, b, r);
        }
#pragma unroll
        for(int d = 0; d < 3; d++)
#pragma unroll
            for(int m = 0; m < V; m++) ps[(q * 3 + d) * CT + lc + m] = r[d][m];
        double fq[V], bq[V];
        if(q < 3 && vm)
        {
#pragma unroll
            for(int m = 0; m < V; m++)
            {
                fq[m] = (MODE != 1) ? F[q * NS + s0 + m] : 0.0;
                bq[m] = (MODE == 0) ? BI[q * NS + s0 + m] : 0.0;
            }
        }
        __syncthreads();
        if(q < 3 && vm)
        {   // thread (q, cells) owns equation q of its V nodes: fold the Q partial rows in a fixed order
#pragma unroll
            for(int m = 0; m < V; m++)
            {
                if(!((vm >> m) & 1u)) continue;
                if(MODE == 0 && fl && (fl[s0 + m] & CCU_B_SHARED)) continue;
                double a = 0.0;
#pragma unroll
                for(int w = 0; w < Q; w++) a += ps[(w * 3 + q) * CT + lc + m];
                if(MODE == 0)
                {   // General_matrix_functions.c:1250-1259: scalar BI per equation, correction rounded to fp32
                    const float t = (float)((fq[m] - a) * bq[m]);
                    const int sx = q * 8 * BOX + c * BOX + sbase + m;
                    const double xn = xs[sx] + (double)t;
                    xs[sx] = xn;
                    x[q * NS + s0 + m] = xn;
                }
                else
                {
                    if(strip && (fl[s0 + m] & (q == 0 ? CCU_F_VBX : (q == 1 ? CCU_F_VBY : CCU_F_VBZ)))) a = 0.0;
                    out[q * NS + s0 + m] = (MODE == 1) ? a : fq[m] - a;
                }
            }
        }
        __syncthreads();
    }
}
