// ccu_col.cuh -- column-resident Gauss-Seidel smoother and matvec for the large multigrid levels (sm_100a).
//
// Why.  The colour-pass kernels (ccu_k_relax_tab, ccu_kernels.cuh) keep the reference's half-stored stiffness
// (Eqn_k1-3: own 3x3 block + 13 lower-neighbour blocks per node, 504 B/node) and gather every row, so each stored
// block crosses HBM twice per sweep -- once as an "own" block of the node that stores it, once transposed for the lower
// neighbour -- plus uncached neighbour values: 2.5x the 600 B/node a sweep has to move (ncu, profiles/r01_*).
// Here a CTA owns a COLUMN of TI x TJ nodes over all z layers and marches through it layer by layer:
//   * the stiffness of a layer arrives as ONE bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA engine) of a
//     chunk that was laid out in HBM exactly as it is used in shared memory (ccu_col_index.h); chunk k holds the
//     in-plane blocks of layer k and every block between the layers k and k+1, so the rows of layer k read the chunks
//     k-1 and k: a ring of three holds two in use and one in flight;
//   * the solution values of the column plus a one-node rim live in a ring of three layers in shared memory, filled two
//     layers ahead by plain loads (their sectors are shared by the layers of equal parity: L2 hits);
//   * a CTA = two groups of compute warps (in-plane colours 3, 1 and 2, 0) + one service warp that issues the bulk
//     copies and handles the progress words, so that no compute warp ever executes a memory fence;
//   * NINE lanes relax one node (three nodes per warp): lane q owns the in-plane direction (q/3-1, q%3-1) and
//     multiplies the three 3x3 blocks towards the layers below, same and above with the neighbour's three values
//     for all three rows; a block is two 128-bit and one 32-bit shared-memory loads, used as stored or transposed
//     (six selects); the nine partial row triples are folded with four double shuffles, after which lane q < 3 holds
//     row q and applies the reference's update (scalar BI per equation, correction rounded to fp32,
//     General_matrix_functions.c:1250-1259);
//   * per layer, the products with the layers below and above (2/3 of a row, independent of this layer's colour phases)
//     are formed for all four in-plane colours at once; a colour phase then only adds the same-layer blocks;
//   * every stiffness byte is read from HBM once per sweep (+ the duplicated halo blocks, 19 % for 6 x 8 columns);
//   * the kernel's limit is the shared-memory pipe (ncu r02: l1tex data pipe 84 % busy, 44 % of it bank-conflict replays
//     with the blocks in natural order), so the blocks of a chunk sit at permuted positions (ccu_col_perm.h, generated
//     by scripts/gen_col_perm.py from the kernel's own access sets) and the rows of the solution window are padded.
// Gauss-Seidel order: columns are 4-coloured by the parity of their column indices, colours 3..0; inside a column z
// ascends; inside a layer the four (y, x)-parity colours 3..0.  All nodes relaxed concurrently share no stencil
// neighbour, so this is a Gauss-Seidel ordering of the same point-block smoother (General_matrix_functions.c:1231-1260);
// oracle/restate.c `ccu_r_ordered_gs` mode 10 states it on the CPU and contracts like the lexicographic order inside
// the multigrid cycle (tests/test_oracle_restate.py).
// The four column colours run in ONE launch (WF = 1): CTAs draw columns from a ticket counter in colour order and a
// column waits, layer by layer, until the columns of earlier colours around it are three layers ahead (a progress word
// per column, release/acquire at gpu scope).  A ticket only ever waits for lower tickets, which are running or done,
// so the launch cannot deadlock however the CTAs are scheduled; results are bitwise those of four launches (WF = 0).
// MODE 1 / 2 run the same march without colours: Au = K u, or rhs - K u with boundary rows stripped
// (n_assemble_del2_u, Element_calculations.c:552).
#pragma once
#include "ccu_kernels.cuh"
#include "ccu_col_index.h"

template <int TI_, int TJ_, int CTAS_>
struct CcuColShape
{
    static constexpr int TI = TI_, TJ = TJ_, S = 3, CTAS = CTAS_;
    static constexpr int NT = TI * TJ;                    // nodes of one layer of a full column
    static constexpr int NQ = NT / 4;                     // nodes of one in-plane colour
    static constexpr int NW = (NQ + 2) / 3;               // warps per colour: three nodes each
    static constexpr int NTC = 2 * NW * 32;               // compute threads: two warp groups, two colours each
    static constexpr int THREADS = NTC + 32;              // + the service warp (bulk copies, progress words)
    static constexpr int BJR = TJ + 2, BOXR = (TI + 2) * BJR;      // solution window of one layer (column + rim)
    static constexpr int BJ = TJ + 3, BOX = (TI + 2) * BJ;         // ... as stored: rows padded so that the nine-point reads of a warp spread over the banks
    static constexpr int XE = (3 * BOXR + NTC - 1) / NTC; // window entries per compute thread
    static constexpr int CHUNK = ccu_col_dims(TI_, TJ_).cb;        // bytes of a full column's chunk
    static constexpr int XLAYER = 3 * BOX * 8;            // bytes of one layer of the solution ring
    static constexpr size_t SMEM = (size_t)S * CHUNK + (size_t)S * XLAYER + (size_t)S * 8;
    static_assert(TI_ % 2 == 0 && TJ_ % 2 == 0, "column extents must be even (in-plane colours)");
};

struct CcuColArgs
{
    CcuGeom g;
    const unsigned char *Kc;      // chunks: column-major, per column layers 0 .. noz-1
    const size_t *colofs;         // [nI * nJ] byte offset of a column's first chunk
    const double *F;              // MODE 0: right-hand side; MODE 2: rhs of the residual
    double *x;                    // MODE 0: solution (in/out); MODE 1, 2: the vector to multiply
    double *out;                  // MODE 1, 2: result
    int nI, nJ;                   // columns along y and x
    int cc;                       // MODE 0, WF 0: column colour of this launch
    int strip;                    // MODE 1: zero the boundary rows of the product
    // MODE 0, WF 1: one launch for the four column colours
    int cstart[5];                // first ticket of the colours 3, 2, 1, 0 and the ticket count
    unsigned *ticket;             // [4] next ticket, finished CTAs (the last one resets both), epoch (the last one bumps it), pad
    unsigned *progress;           // [nI * nJ] epoch * CCU_COL_EPOCH + layers finished; the epoch lives on the device so that a
                                  // launch replayed from a CUDA graph still gets a fresh one and the words never need a reset
};
#define CCU_COL_EPOCH 4096u       // > noz of any level

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
__device__ __forceinline__ void ccu_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do
    {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while(!ok);
}
__device__ __forceinline__ unsigned ccu_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ccu_st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// r[0..2] += B x  (tr = false)  or  B^T x  (tr = true), B = the 3x3 block at `A4`, `B4`, `C1` (ccu_col_coef_ofs)
__device__ __forceinline__ void ccu_col_block(double (&r)[3], const float4 a, const float4 b, const float c, const bool tr,
                                              const double x0, const double x1, const double x2)
{
    // e = 3 row + col: a = (e0 e1 e2 e3), b = (e4 e5 e6 e7), c = e8; transposing swaps e1/e3, e2/e6, e5/e7
    const float e1 = tr ? a.w : a.y, e3 = tr ? a.y : a.w, e2 = tr ? b.z : a.z, e6 = tr ? a.z : b.z, e5 = tr ? b.w : b.y, e7 = tr ? b.y : b.w;
    r[0] += (double)a.x * x0; r[1] += (double)e3 * x0; r[2] += (double)e6 * x0;
    r[0] += (double)e1 * x1;  r[1] += (double)b.x * x1; r[2] += (double)e7 * x1;
    r[0] += (double)e2 * x2;  r[1] += (double)e5 * x2;  r[2] += (double)c * x2;
}

// Row sums over the nine lanes q = 3 tri + mm of a node: returns, on every lane, the total of row mm (taken on q < 3).
// Within a triple each lane hands the two rows it does not keep to the lanes that keep them, then the three triples add up.
__device__ __forceinline__ double ccu_col_fold9(const double (&r)[3], const int mm, const int srcA, const int srcB, const int src3, const int src6)
{
    const double own = mm == 0 ? r[0] : (mm == 1 ? r[1] : r[2]);
    const double toA = mm == 0 ? r[2] : (mm == 1 ? r[0] : r[1]);     // row (mm+2)%3: wanted by the lane that reads me in round A
    const double toB = mm == 0 ? r[1] : (mm == 1 ? r[2] : r[0]);     // row (mm+1)%3: wanted by the lane that reads me in round B
    const double a = __shfl_sync(0xffffffffu, toA, srcA), b = __shfl_sync(0xffffffffu, toB, srcB);
    const double s = (own + a) + b;
    const double s3 = __shfl_sync(0xffffffffu, s, src3), s6 = __shfl_sync(0xffffffffu, s, src6);
    return (s + s3) + s6;
}

__device__ __forceinline__ void ccu_bar_sync(const int id, const int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <class SH, int MODE, int WF>
__global__ void __launch_bounds__(SH::THREADS, SH::CTAS) ccu_k_col(const __grid_constant__ CcuColArgs A)
{
    constexpr int TI = SH::TI, TJ = SH::TJ, S = SH::S, BJ = SH::BJ, BOX = SH::BOX, BJR = SH::BJR, BOXR = SH::BOXR, CH = SH::CHUNK, XL = SH::XLAYER;
    constexpr int NQ = SH::NQ, NW = SH::NW, HJ = TJ / 2, XE = SH::XE, NTC = SH::NTC, NTH = SH::THREADS;
    extern __shared__ __align__(128) unsigned char ccu_col_smem[];
    __shared__ int s_ticket;
    __shared__ unsigned s_epoch;
    unsigned char *stg = ccu_col_smem;                                   // [S][CH] stiffness chunks
    unsigned char *xrb = stg + (size_t)S * CH;                           // [S][3][BOX] doubles, solution ring
    const unsigned bar0 = ccu_smem_u32(xrb + (size_t)S * XL);            // [S] mbarriers
    const CcuGeom &g = A.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int I, J, mycc = 0;
    unsigned pbase = 0u;
    if(MODE == 0 && WF)
    {
        if(tid == 0) { s_ticket = (int)atomicAdd(A.ticket, 1u); s_epoch = *(volatile unsigned *)(A.ticket + 2); }
        __syncthreads();
        const int t = s_ticket;
        pbase = s_epoch * CCU_COL_EPOCH;
        int grp = 0;
        while(grp < 3 && t >= A.cstart[grp + 1]) grp++;
        mycc = 3 - grp;
        const int ci = mycc >> 1, cj = mycc & 1, nJc = (A.nJ - cj + 1) / 2, r = t - A.cstart[grp];
        I = 2 * (r / nJc) + ci; J = 2 * (r % nJc) + cj;
    }
    else if(MODE == 0)
    {
        const int ci = A.cc >> 1, cj = A.cc & 1, nJc = (A.nJ - cj + 1) / 2;
        I = 2 * ((int)blockIdx.x / nJc) + ci; J = 2 * ((int)blockIdx.x % nJc) + cj;
    }
    else { I = (int)blockIdx.x / A.nJ; J = (int)blockIdx.x % A.nJ; }
    const int i0 = I * TI, j0 = J * TJ, noz = g.noz;
    const CcuColDims cd = ccu_col_dims(min(TI, g.noy - i0), min(TJ, g.nox - j0));
    const unsigned char *chunks = A.Kc + A.colofs[I * A.nJ + J];

    // ---- prologue: barriers, zeros where chunk -1 would be
    if(tid == 0)
    {
        for(int s = 0; s < S; s++) ccu_mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for(int w = tid; w < CH / 16; w += NTH) ((float4 *)(stg + (size_t)(S - 1) * CH))[w] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if(warp == 2 * NW)
    {   // ================= service warp: bulk copies of the chunks, progress words of the one-launch sweep
        auto issue = [&](int layer)        // lane 0: chunk `layer` (0 .. noz-1) into its ring stage
        {
            const int s = layer % S;
            ccu_mbar_expect_tx(bar0 + 8 * s, (unsigned)cd.cb);
            ccu_bulk_g2s(ccu_smem_u32(stg + (size_t)s * CH), chunks + (size_t)layer * cd.cb, (unsigned)cd.cb, bar0 + 8 * s);
        };
        if(lane == 0)
            for(int layer = 0; layer <= S - 2 && layer < noz; layer++) issue(layer);
        // the columns of earlier colours around this one (a column of colour (ci, cj) has the neighbour colours
        // (ci ^ |a|, cj ^ |b|)); lane j < 8 watches neighbour j
        const unsigned *watch = nullptr;
        unsigned seen = pbase;
        if(MODE == 0 && WF && lane < 8)
        {
            const int a = (lane < 3) ? -1 : (lane < 5 ? 0 : 1), b = (lane < 3) ? lane - 1 : (lane < 5 ? (lane == 3 ? -1 : 1) : lane - 6);
            const int nc = 2 * ((mycc >> 1) ^ (a & 1)) + ((mycc & 1) ^ (b & 1));
            const int In = I + a, Jn = J + b;
            if(nc > mycc && In >= 0 && In < A.nI && Jn >= 0 && Jn < A.nJ) watch = A.progress + (In * A.nJ + Jn);
        }
        auto wait_for = [&](int layers)    // until every watched column has finished `layers` layers of this sweep
        {
            const unsigned need = pbase + (unsigned)min(layers, noz);
            if(watch)
            {   // bounded: a wait that cannot end (it never should: lower tickets run or are done) flags the launch instead of hanging the device
                unsigned spins = 0;
                while((int)(seen - need) < 0)
                {
                    seen = ccu_ld_acquire(watch);
                    if((int)(seen - need) < 0) { __nanosleep(64); if(++spins > (1u << 24)) { A.ticket[3] = 1u; break; } }
                }
            }
            __syncwarp();
        };
        // the compute warps load the rim of the layers -1 .. 2 before their first end-of-layer barrier, and of layer k + 3
        // right after the end-of-layer barrier of layer k: the neighbours must have finished it by then
        if(MODE == 0 && WF) wait_for(3);
        __syncthreads();
        for(int k = 0; k < noz; k++)
        {
            if(MODE == 0 && WF) wait_for(k + 4);
            __syncthreads();                                         // end of layer k
            if(lane == 0)
            {
                if(k + S - 1 < noz) issue(k + S - 1);                // the stage of chunk k - 1 is free now
                if(MODE == 0 && WF) ccu_st_release(A.progress + (I * A.nJ + J), pbase + (unsigned)(k + 1));
            }
        }
        if(MODE == 0 && WF && lane == 0)
        {   // the last CTA of the launch resets the ticket counter and bumps the epoch for the next one
            if(atomicAdd(A.ticket + 1, 1u) + 1u == (unsigned)A.cstart[4]) { A.ticket[1] = 0u; A.ticket[0] = 0u; A.ticket[2] = s_epoch + 1u; }
        }
        return;
    }

    // ================= compute warps
    const size_t NS = (size_t)g.NS;
    const int grp = warp / NW, wl = warp % NW;           // group 0: in-plane colours 3 and 1, group 1: colours 2 and 0
    const int n3 = lane / 9, q = lane % 9, tri = q / 3, mm = q % 3;
    const bool act = lane < 27;
    const int m = 3 * wl + n3;                           // this lane's node among the NQ nodes of an in-plane colour
    const int wa = m / HJ, wb = m % HJ;
    const int lb = 9 * n3;
    const int srcA = act ? lb + 3 * tri + (mm + 1) % 3 : lane, srcB = act ? lb + 3 * tri + (mm + 2) % 3 : lane;
    const int src3 = act ? lb + 3 * ((tri + 1) % 3) + mm : lane, src6 = act ? lb + 3 * ((tri + 2) % 3) + mm : lane;
    const bool upd = act && q < 3;                       // lane q < 3 ends up with row d = q
    int kA[2][3], xof[2], nodeA[2], xself[2], pidx[2];
    bool nv[2], tr[3];
#pragma unroll
    for(int t = 0; t < 3; t++)
    {
        const int di = q / 3 - 1, dj = q % 3 - 1, dk = t - 1;
        tr[t] = !((di == 0 && dj == 0 && dk == 0) || ccu_lo_index(di, dj, dk) >= 0);
    }
#pragma unroll
    for(int ci = 0; ci < 2; ci++)
    {
        const int c2 = (ci == 0 ? 3 : 1) - grp;
        const int li = 2 * wa + (c2 >> 1), lj = 2 * wb + (c2 & 1);
        nv[ci] = act && m < NQ && li < cd.ti && lj < cd.tj;
        xof[ci] = 0;
#pragma unroll
        for(int t = 0; t < 3; t++)
        {
            CcuColDesc ds = { 0, 0, 0 };
            if(nv[ci]) ds = ccu_col_desc(cd, BJ, li, lj, q, t);
            kA[ci][t] = cd.kofs + 16 * ccu_col_pos(cd, ds.id); xof[ci] = ds.xof;
        }
        const int gi = i0 + li, gj = j0 + lj;
        nodeA[ci] = (4 * (gi & 1) + 2 * (gj & 1)) * g.NC + ((gi >> 1) + 1) * g.JK + ((gj >> 1) + 1) * g.Kd + 1;
        xself[ci] = (li + 1) * BJ + (lj + 1);
        pidx[ci] = nv[ci] ? li * cd.tj + lj : 0;
    }
    const int oB = 16 * cd.nbp, oC = cd.kofs + 32 * cd.nbp; // from a block's A entry to its B entry; C entry = oC + 4 position
    // solution-window loader: compute thread `tid` owns entries tid, tid + NTC, ... (dof plane dx, window position bn) of every layer
    size_t xA[XE];
    bool xin[XE];
    int xdst[XE];
#pragma unroll
    for(int e = 0; e < XE; e++)
    {
        const int w = tid + e * NTC;
        const bool xl = w < 3 * BOXR;
        const int dx = xl ? w / BOXR : 0, bn = xl ? w % BOXR : 0;
        const int xgi = i0 + bn / BJR - 1, xgj = j0 + bn % BJR - 1;
        xdst[e] = dx * BOX + (bn / BJR) * BJ + bn % BJR;
        xin[e] = xl && xgi >= 0 && xgi < g.noy && xgj >= 0 && xgj < g.nox;
        xA[e] = xin[e] ? (size_t)dx * NS + (size_t)((4 * (xgi & 1) + 2 * (xgj & 1)) * g.NC + ((xgi >> 1) + 1) * g.JK + ((xgj >> 1) + 1) * g.Kd + 1) : 0;
    }
    auto xload = [&](int e, int k) -> double { return (xin[e] && k >= 0 && k < noz) ? __ldcg(A.x + xA[e] + (size_t)((k & 1) * g.NC + (k >> 1))) : 0.0; };
    auto xstore = [&](int slot, const double (&v)[XE])
    {
#pragma unroll
        for(int e = 0; e < XE; e++)
            if(tid + e * NTC < 3 * BOXR) ((double *)(xrb + (size_t)slot * XL))[xdst[e]] = v[e];
    };
    __syncthreads();                                     // the service warp has seen the neighbours past layer 2 (one-launch sweep)
    {   // layers -1, 0, 1 of the solution ring
        double v[XE];
#pragma unroll
        for(int e = 0; e < XE; e++) v[e] = xload(e, -1);
        xstore(S - 1, v);
#pragma unroll
        for(int e = 0; e < XE; e++) v[e] = xload(e, 0);
        xstore(0, v);
#pragma unroll
        for(int e = 0; e < XE; e++) v[e] = xload(e, 1);
        xstore(1, v);
    }
    double fnx[2] = { 0.0, 0.0 };                          // right-hand side of this lane's rows, one layer ahead
    if(MODE != 1 && upd)
#pragma unroll
        for(int ci = 0; ci < 2; ci++)
            if(nv[ci]) fnx[ci] = A.F[(size_t)q * NS + (size_t)nodeA[ci]];
    ccu_bar_sync(1, NTC);

    for(int k0 = 0; k0 < noz; k0 += S)
    {
#pragma unroll
        for(int JJ = 0; JJ < S; JJ++)
        {
            const int k = k0 + JJ;               // k % S == JJ: ring positions are compile-time constants below
            if(k >= noz) break;
            const int PJ = (JJ + S - 1) % S, NJ = (JJ + 1) % S;          // ring stages of the layers k - 1 and k + 1
            double xpre[XE], fcur[2];
#pragma unroll
            for(int e = 0; e < XE; e++) xpre[e] = xload(e, k + 2);       // lands in the ring before this layer's last barrier
#pragma unroll
            for(int ci = 0; ci < 2; ci++)
            {
                fcur[ci] = fnx[ci];
                if(MODE != 1 && upd && nv[ci] && k + 1 < noz) fnx[ci] = A.F[(size_t)q * NS + (size_t)(nodeA[ci] + ((k + 1) & 1) * g.NC + ((k + 1) >> 1))];
            }
            ccu_mbar_wait(bar0 + 8 * JJ, (unsigned)((k / S) & 1));       // chunk k (chunk k - 1 arrived a layer ago)
            const int zoff = (k & 1) * g.NC + (k >> 1);
            const unsigned char *cur = stg + (size_t)JJ * CH, *prv = stg + (size_t)PJ * CH;
            // r += the block of direction layer t - 1 times the neighbour's values, for this lane's node of its colour ci
            auto prod = [&](double (&r)[3], const int ci, const int t)
            {
                const unsigned char *ck = t == 0 ? prv : cur;
                const unsigned char *kb = ck + kA[ci][t];
                const float4 a = *(const float4 *)kb, b = *(const float4 *)(kb + oB);
                const float c = *(const float *)(ck + oC + ((kA[ci][t] - cd.kofs) >> 2));
                const double *xp = (const double *)(xrb + ((t == 0 ? PJ : (t == 1 ? JJ : NJ)) * XL + xof[ci]));
                ccu_col_block(r, a, b, c, tr[t], xp[0], xp[BOX], xp[2 * BOX]);
            };
            double acc[3];
            // what a row needs from the layers below and above is independent of this layer's colour phases
            auto outer = [&](const int ci) { acc[0] = acc[1] = acc[2] = 0.0; prod(acc, ci, 0); prod(acc, ci, 2); };
            auto phase = [&](const int ci)
            {
                prod(acc, ci, 1);
                const double r = ccu_col_fold9(acc, mm, srcA, srcB, src3, src6);
                if(upd && nv[ci])
                {   // General_matrix_functions.c:1250-1259: scalar BI per equation, correction rounded to fp32
                    const double bi = ((const double *)cur)[q * cd.nt + pidx[ci]];
                    double *xs = (double *)(xrb + (size_t)JJ * XL) + q * BOX + xself[ci];
                    const double xn = *xs + (double)(float)((fcur[ci] - r) * bi);
                    *xs = xn;
                    A.x[(size_t)q * NS + (size_t)(nodeA[ci] + zoff)] = xn;
                }
            };
            if(MODE == 0)
            {   // in-plane colours 3, 2, 1, 0: the two warp groups alternate, the idle one prepares its next colour
                outer(0);
                if(grp == 0) phase(0);
                ccu_bar_sync(1, NTC);
                if(grp == 1) phase(0); else outer(1);
                ccu_bar_sync(1, NTC);
                if(grp == 0) phase(1); else outer(1);
                ccu_bar_sync(1, NTC);
                if(grp == 1) phase(1);
                xstore(PJ, xpre);                                    // stage of layer k - 1: its last readers were the outer products
            }
            else
            {
#pragma unroll
                for(int ci = 0; ci < 2; ci++)
                {
                    outer(ci);
                    prod(acc, ci, 1);
                    double r = ccu_col_fold9(acc, mm, srcA, srcB, src3, src6);
                    if(upd && nv[ci])
                    {
                        const unsigned char flg = cur[cd.flofs + pidx[ci]];
                        if((MODE == 2 || A.strip) && ((flg >> q) & 1)) r = 0.0;
                        if(MODE == 2) r = fcur[ci] - r;
                        A.out[(size_t)q * NS + (size_t)(nodeA[ci] + zoff)] = r;
                    }
                }
                ccu_bar_sync(1, NTC);                                // every product of this layer has read the stage of layer k - 1
                xstore(PJ, xpre);
            }
            __syncthreads();                                         // end of layer k (with the service warp)
        }
    }
}

// K (coefficient-major, colour-blocked), BI, flags -> column chunks: the same bytes ccu_col_fill_chunk (ccu_col_index.h, the
// statement the host emulation uses) produces.  A CTA takes one column and 32 consecutive layers.  Reads run along z (a lane
// per layer: the two parities of a warp read two 64-byte runs); the blocks go through a shared-memory tile in batches of 32
// consecutive chunk POSITIONS (`inv` = position -> block id for the full column shape, identity for clipped columns), so
// that every store to HBM is a full 512-byte (A, B halves) or 128-byte (ninth coefficients) run of one chunk.
template <int TI, int TJ>
__global__ void __launch_bounds__(256) ccu_k_col_relayout(const CcuGeom g, const int nJ, const size_t *__restrict__ colofs,
                                                           const float *__restrict__ K, const double *__restrict__ BI,
                                                           const unsigned char *__restrict__ flags, const unsigned char *__restrict__ bits,
                                                           const short *__restrict__ inv, unsigned char *Kc)
{
    __shared__ float tile[32][32 * 9 + 1];                      // [layer][position in batch][coefficient]
    const int col = blockIdx.x, I = col / nJ, J = col % nJ;
    const int k0 = (int)blockIdx.y * 32, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = I * TI, j0 = J * TJ;
    const CcuColDims cd = ccu_col_dims(min(TI, g.noy - i0), min(TJ, g.nox - j0));
    const bool full = cd.ti == TI && cd.tj == TJ;
    unsigned char *chunk0 = Kc + colofs[col] + (size_t)k0 * cd.cb;
    const size_t NS = (size_t)g.NS;
    const int nlay = min(32, g.noz - k0);
    // inverse diagonal and flags: thread per (layer, entry)
    for(int w = threadIdx.x; w < nlay * (cd.kofs / 8); w += 256)
    {
        const int l = w / (cd.kofs / 8), e = w % (cd.kofs / 8);
        double v = 0.0;
        if(e < 3 * cd.nt)
        {
            const int dd = e / cd.nt, p = e % cd.nt;
            const int s = ccu_sidx(g, i0 + p / cd.tj, j0 + p % cd.tj, k0 + l);
            v = BI[dd * NS + s];
            if(bits && (bits[s] & 2)) v = 0.0;
        }
        ((double *)(chunk0 + (size_t)l * cd.cb))[e] = v;
    }
    for(int w = threadIdx.x; w < nlay * (cd.cb - cd.flofs); w += 256)
    {
        const int l = w / (cd.cb - cd.flofs), p = w % (cd.cb - cd.flofs);
        chunk0[(size_t)l * cd.cb + cd.flofs + p] = p < cd.nt ? flags[ccu_sidx(g, i0 + p / cd.tj, j0 + p % cd.tj, k0 + l)] : (unsigned char)0;
    }
    for(int b0 = 0; b0 < cd.nbp; b0 += 32)
    {
        // gather: warp w takes the positions b0 + 4 w .. + 3, lane = layer
#pragma unroll
        for(int pp = 0; pp < 4; pp++)
        {
            const int pl = 4 * warp + pp, pos = b0 + pl;
            const int id = full ? (int)inv[pos] : (pos < cd.nb ? pos : -1);
            float v[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
            if(id >= 0 && lane < nlay)
            {
                int sli, slj, slot;
                if(id < 14 * cd.nt) { const int p = id / 14; slot = id % 14; sli = p / cd.tj; slj = p % cd.tj; }
                else { int b; ccu_col_halo_decode(cd.ti, cd.tj, id - 14 * cd.nt, sli, slj, b); slot = b + 1; }
                int di = 0, dj = 0, dk = 0;
                if(slot) ccu_lo_offset(slot - 1, di, dj, dk);
                const int ks = k0 + lane + (dk < 0 ? 1 : 0);           // blocks that reach down belong to the layer above
                const int gi = i0 + sli, gj = j0 + slj;
                if(ks < g.noz && gi >= 0 && gi < g.noy && gj >= 0 && gj < g.nox)
                {
                    const int s = ccu_sidx(g, gi, gj, ks);
#pragma unroll
                    for(int e = 0; e < 9; e++) v[e] = K[(size_t)(slot * 9 + e) * NS + s];
                }
            }
#pragma unroll
            for(int e = 0; e < 9; e++) tile[lane][pl * 9 + e] = v[e];
        }
        __syncthreads();
        // scatter: 32 layers x 32 positions; a warp writes the 512 bytes of one layer's A (then B) halves, 128 bytes of its C entries
        for(int l = warp; l < nlay; l += 8)
        {
            unsigned char *ch = chunk0 + (size_t)l * cd.cb + cd.kofs;
            const float *t = &tile[l][lane * 9];
            ((float4 *)(ch + 16 * b0))[lane] = make_float4(t[0], t[1], t[2], t[3]);
            ((float4 *)(ch + 16 * cd.nbp + 16 * b0))[lane] = make_float4(t[4], t[5], t[6], t[7]);
            ((float *)(ch + 32 * cd.nbp + 4 * b0))[lane] = t[8];
        }
        __syncthreads();
    }
}
