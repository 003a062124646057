// col_emul.cpp -- HOST emulation of the column-resident smoother / matvec kernel (csrc/ccu_col.cuh), test infrastructure.
//
// It re-states the kernel's control flow (ring of S = 3 stiffness chunks filled at the kernel's own issue points, ring of S
// solution layers, nine lanes q per node x 3 layer directions x all three rows, the fold order, the update) on
// the CPU, using the SAME index functions (csrc/ccu_col_index.h: chunk layout, halo block ids, lane descriptors,
// chunk fill) the device code uses.  tests/test_col_emul.py compares it with the oracle's column-ordered Gauss-Seidel
// (oracle/restate.c mode 10) and the reference's matvec known answers, so that the index logic and the ring schedule
// are verified without a GPU; the GPU tests then only have to prove the CUDA-specific parts.
#include "../citcomcu_b200/csrc/ccu_col_index.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace
{
struct Emul
{
    CcuGeom g;
    int TI, TJ, S, BJ, BOX, BJR, BOXR, CH, nI, nJ;
    std::vector<float> K;
    std::vector<double> BI;
    std::vector<unsigned char> flags, Kc;
    std::vector<size_t> colofs;
};

void run_column(const Emul &E, int mode, int I, int J, const double *F, double *x, double *out, int strip)
{
    const CcuGeom &g = E.g;
    const int TI = E.TI, TJ = E.TJ, S = E.S, BJ = E.BJ, BOX = E.BOX, BJR = E.BJR, BOXR = E.BOXR, CH = E.CH, noz = g.noz;
    const int i0 = I * TI, j0 = J * TJ;
    const CcuColDims cd = ccu_col_dims(std::min(TI, g.noy - i0), std::min(TJ, g.nox - j0));
    const unsigned char *chunks = E.Kc.data() + E.colofs[I * E.nJ + J];
    const size_t NS = (size_t)g.NS;
    std::vector<unsigned char> stg((size_t)S * CH, 0xff);          // poison: reads of an unfilled stage show up as NaN
    std::vector<double> xr((size_t)S * 3 * BOX, 1e300);
    const int NQ = TI * TJ / 4, HJ = TJ / 2;
    auto issue = [&](int layer) { memcpy(stg.data() + (size_t)(layer % S) * CH, chunks + (size_t)layer * cd.cb, cd.cb); };
    auto xdst = [&](int tid) { const int dx = tid / BOXR, bn = tid % BOXR; return dx * BOX + (bn / BJR) * BJ + bn % BJR; };
    auto xload = [&](int tid, int k) -> double
    {
        const int dx = tid / BOXR, bn = tid % BOXR, gi = i0 + bn / BJR - 1, gj = j0 + bn % BJR - 1;
        if(!(gi >= 0 && gi < g.noy && gj >= 0 && gj < g.nox && k >= 0 && k < noz)) return 0.0;
        const size_t xA = (size_t)dx * NS + (size_t)((4 * (gi & 1) + 2 * (gj & 1)) * g.NC + ((gi >> 1) + 1) * g.JK + ((gj >> 1) + 1) * g.Kd + 1);
        return x[xA + (size_t)((k & 1) * g.NC + (k >> 1))];
    };
    memset(stg.data() + (size_t)(S - 1) * CH, 0, CH);              // where chunk -1 would be
    for(int layer = 0; layer <= S - 2 && layer < noz; layer++) issue(layer);
    for(int tid = 0; tid < 3 * BOXR; tid++)
    {
        xr[(size_t)(S - 1) * 3 * BOX + xdst(tid)] = xload(tid, -1);
        xr[(size_t)0 * 3 * BOX + xdst(tid)] = xload(tid, 0);
        xr[(size_t)1 * 3 * BOX + xdst(tid)] = xload(tid, 1);
    }
    std::vector<double> xpre(3 * BOXR);
    for(int k = 0; k < noz; k++)
    {
        const int JJ = k % S, PJ = (JJ + S - 1) % S, NJ = (JJ + 1) % S;
        for(int tid = 0; tid < 3 * BOXR; tid++) xpre[tid] = xload(tid, k + 2);
        const int zoff = (k & 1) * g.NC + (k >> 1);
        const unsigned char *cur = stg.data() + (size_t)JJ * CH, *prv = stg.data() + (size_t)PJ * CH;
        for(int ph = 0; ph < 4; ph++)
        {
            const int c2 = mode == 0 ? 3 - ph : ph;
            for(int m = 0; m < NQ; m++)
            {
                const int wa = m / HJ, wb = m % HJ;
                const int li = 2 * wa + (c2 >> 1), lj = 2 * wb + (c2 & 1);
                if(!(li < cd.ti && lj < cd.tj)) continue;
                double r[9][3];
                for(int q = 0; q < 9; q++)
                {
                    double acc[3] = { 0, 0, 0 };
                    const int order[3] = { 0, 2, 1 };                // layers below and above first, then the same layer
                    for(int oi = 0; oi < 3; oi++)
                    {
                        const int t = order[oi];
                        const CcuColDesc ds = ccu_col_desc(cd, BJ, li, lj, q, t);
                        const unsigned char *ck = t == 0 ? prv : cur;
                        float e[9];
                        for(int ee = 0; ee < 9; ee++) memcpy(&e[ee], ck + ccu_col_coef_ofs(cd, ds.id, ee), 4);
                        const int ring = t == 0 ? PJ : (t == 1 ? JJ : NJ);
                        const double *xp = (const double *)((const unsigned char *)xr.data() + (size_t)ring * 3 * BOX * 8 + ds.xof);
                        const double xv[3] = { xp[0], xp[BOX], xp[2 * BOX] };
                        for(int bb = 0; bb < 3; bb++)
                            for(int a = 0; a < 3; a++) acc[a] += (double)(ds.tr ? e[3 * bb + a] : e[3 * a + bb]) * xv[bb];
                    }
                    for(int a = 0; a < 3; a++) r[q][a] = acc[a];
                }
                // the kernel's fold (ccu_col_fold9): inside a triple of lanes, then over the three triples
                double row[3];
                for(int d = 0; d < 3; d++)
                {
                    double s[3];
                    for(int tri = 0; tri < 3; tri++) s[tri] = (r[3 * tri + d][d] + r[3 * tri + (d + 1) % 3][d]) + r[3 * tri + (d + 2) % 3][d];
                    row[d] = (s[0] + s[1]) + s[2];
                }
                const int gi = i0 + li, gj = j0 + lj;
                const int nodeA = (4 * (gi & 1) + 2 * (gj & 1)) * g.NC + ((gi >> 1) + 1) * g.JK + ((gj >> 1) + 1) * g.Kd + 1;
                for(int d = 0; d < 3; d++)
                {
                    const size_t sn = (size_t)d * NS + (size_t)(nodeA + zoff);
                    const double rr = row[d];
                    const int p = li * cd.tj + lj;
                    if(mode == 0)
                    {
                        const double bi = ((const double *)cur)[d * cd.nt + p];
                        double *xs = xr.data() + (size_t)JJ * 3 * BOX + d * BOX + (li + 1) * BJ + (lj + 1);
                        const double xn = *xs + (double)(float)((F[sn] - rr) * bi);
                        *xs = xn; x[sn] = xn;
                    }
                    else
                    {
                        const unsigned char fl = cur[cd.flofs + p];
                        double a = rr;
                        if((mode == 2 || strip) && ((fl >> d) & 1)) a = 0.0;
                        out[sn] = mode == 1 ? a : F[sn] - a;
                    }
                }
            }
            if(ph == 3) for(int tid = 0; tid < 3 * BOXR; tid++) xr[(size_t)((JJ + 2) % S) * 3 * BOX + xdst(tid)] = xpre[tid];
        }
        if(k + S - 1 < noz) issue(k + S - 1);
    }
}
}

// mode 0: `cycles` column-ordered sweeps on x (in/out) with right-hand side F; mode 1: out = K x (boundary rows zeroed if
// strip); mode 2: out = F - K x with boundary rows of K x zeroed.  Arrays in the REFERENCE's layouts (Eqn_k1-3: 42 per
// node; vectors 3n+d; NODE flags), exactly what ccu_set_stiffness / ccu_set_node_flags take.
extern "C" int ccu_col_emul(int nox, int noy, int noz, int TI, int TJ, int S, int mode, const float *k1, const float *k2,
                            const float *k3, const double *BIh, const unsigned *node, double *xh, const double *Fh, double *outh,
                            int cycles, int strip)
{
    const int LO[13][3] = CCU_LO_INIT;
    Emul E;
    E.g = ccu_make_geom(nox, noy, noz);
    const CcuGeom &g = E.g;
    E.TI = TI; E.TJ = TJ; E.S = S; E.BJR = TJ + 2; E.BOXR = (TI + 2) * (TJ + 2); E.BJ = TJ + 3; E.BOX = (TI + 2) * (TJ + 3);
    E.CH = ccu_col_dims(TI, TJ).cb;
    E.nI = (noy + TI - 1) / TI; E.nJ = (nox + TJ - 1) / TJ;
    const size_t NS = (size_t)g.NS;
    E.K.assign(126 * NS, 0.0f); E.BI.assign(3 * NS, 0.0); E.flags.assign(NS, 0);
    std::vector<double> x(3 * NS, 0.0), F(3 * NS, 0.0), out(3 * NS, 0.0);
    for(int n = 0; n < g.nno; n++)
    {   // ccu_k_stiffness_to_dev / ccu_k_vec_to_dev / ccu_k_flags_to_dev
        const int k = n % noz, j = (n / noz) % nox, i = n / (noz * nox);
        const int s = ccu_sidx(g, i, j, k);
        const size_t base = (size_t)n * 42;
        const float *kk[3] = { k1 + base, k2 + base, k3 + base };
        for(int a = 0; a < 3; a++) for(int b = 0; b < 3; b++) E.K[(size_t)(a * 3 + b) * NS + s] = kk[a][b];
        int rs = 0;
        for(int q = 0; q < 13; q++)
        {
            const int ii = i + LO[q][0], jj = j + LO[q][1], kz = k + LO[q][2];
            const bool in = ii >= 0 && jj >= 0 && jj < nox && kz >= 0 && kz < noz;
            if(in) rs++;
            for(int a = 0; a < 3; a++) for(int b = 0; b < 3; b++) E.K[(size_t)((q + 1) * 9 + a * 3 + b) * NS + s] = in ? kk[a][3 * rs + b] : 0.0f;
        }
        for(int d = 0; d < 3; d++)
        {
            E.BI[d * NS + s] = BIh[3 * (size_t)n + d];
            x[d * NS + s] = xh[3 * (size_t)n + d];
            if(Fh) F[d * NS + s] = Fh[3 * (size_t)n + d];
        }
        const unsigned f = node[n];
        E.flags[s] = (unsigned char)(128 | ((f & 0x2u) ? 1 : 0) | ((f & 0x8u) ? 2 : 0) | ((f & 0x4u) ? 4 : 0));
    }
    E.colofs.resize((size_t)E.nI * E.nJ);
    size_t total = 0;
    for(int I = 0; I < E.nI; I++)
        for(int J = 0; J < E.nJ; J++)
        {
            E.colofs[I * E.nJ + J] = total;
            total += (size_t)noz * ccu_col_dims(std::min(TI, noy - I * TI), std::min(TJ, nox - J * TJ)).cb;
        }
    E.Kc.assign(total, 0xee);
    for(int I = 0; I < E.nI; I++)
        for(int J = 0; J < E.nJ; J++)
        {
            const CcuColDims cd = ccu_col_dims(std::min(TI, noy - I * TI), std::min(TJ, nox - J * TJ));
            for(int kk = 0; kk < noz; kk++)
                ccu_col_fill_chunk(g, cd, I * TI, J * TJ, kk, E.K.data(), E.BI.data(), E.flags.data(), nullptr,
                                   E.Kc.data() + E.colofs[I * E.nJ + J] + (size_t)kk * cd.cb, 0, 1);
        }
    if(mode == 0)
    {
        for(int s = 0; s < cycles; s++)
            for(int cc = 3; cc >= 0; cc--)
                for(int I = cc >> 1; I < E.nI; I += 2)
                    for(int J = cc & 1; J < E.nJ; J += 2) run_column(E, 0, I, J, F.data(), x.data(), nullptr, 0);
    }
    else
        for(int I = 0; I < E.nI; I++)
            for(int J = 0; J < E.nJ; J++) run_column(E, mode, I, J, F.data(), x.data(), out.data(), strip);
    for(int n = 0; n < g.nno; n++)
    {
        const int k = n % noz, j = (n / noz) % nox, i = n / (noz * nox);
        const int s = ccu_sidx(g, i, j, k);
        for(int d = 0; d < 3; d++)
        {
            if(mode == 0) xh[3 * (size_t)n + d] = x[d * NS + s];
            else outh[3 * (size_t)n + d] = out[d * NS + s];
        }
    }
    return 0;
}
