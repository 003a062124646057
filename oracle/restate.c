/* restate.c -- CPU oracle: plain-C restatement of CitcomCU's Stokes hot path.
 * TEST INFRASTRUCTURE ONLY (see restate.h).  Single-rank semantics (no OFFSIDE
 * nodes, exchange_id_d20 is the identity).  Loop orders follow the reference so
 * results are bit-identical to it when built without FMA contraction.
 *
 * Two families:
 *   ccu_r_*      reference algorithms (file:line cited per function)
 *   ccu_r_mc_*   the 8-colour Gauss-Seidel (colours 7..0 each sweep) the CUDA path implements,
 *                stated on the reference's own arrays; it is the exact checker
 *                for the kernels, and is itself checked against the reference at
 *                converged tolerance (colouring changes iterates, not solutions).
 */
#include "restate.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define MAXEQ 42

/* local node a (1..8) -> offsets {dz, dx, dy}: element_definitions.h:211-222 + Construct_arrays.c:74-82 */
static const int OFFS[9][3] = { {0,0,0}, {0,0,0}, {0,1,0}, {0,1,1}, {0,0,1}, {1,0,0}, {1,1,0}, {1,1,1}, {1,0,1} };

static inline int nid(const ccu_r_level *L, int i, int j, int k) { return k + L->noz * (j + L->nox * i); }
static inline void nijk(const ccu_r_level *L, int n, int *i, int *j, int *k)
{
    *k = n % L->noz; *j = (n / L->noz) % L->nox; *i = n / (L->noz * L->nox);
}
static inline int elz_(const ccu_r_level *L) { return L->noz - 1; }
static inline int elx_(const ccu_r_level *L) { return L->nox - 1; }
static inline int eid(const ccu_r_level *L, int ey, int ex, int ez) { return ez + elz_(L) * (ex + elx_(L) * ey); }
static inline int enode(const ccu_r_level *L, int e, int a)
{
    int ez = e % elz_(L), ex = (e / elz_(L)) % elx_(L), ey = e / (elz_(L) * elx_(L));
    return nid(L, ey + OFFS[a][2], ex + OFFS[a][1], ez + OFFS[a][0]);
}

/* lower-numbered neighbours of node (i,j,k) in the reference's slot order
 * (Construct_arrays.c:320-341); returns count, fills m[1..]. */
static int lower_nbrs(const ccu_r_level *L, int i, int j, int k, int *m)
{
    int ia = 0, di, dj, dk;
    const int nn = nid(L, i, j, k);
    for(di = (i == 0) ? 0 : -1; di <= 0; di++)
        for(dj = (j == 0) ? 0 : -1; dj <= ((j == L->nox - 1) ? 0 : 1); dj++)
            for(dk = (k == 0) ? 0 : -1; dk <= ((k == L->noz - 1) ? 0 : 1); dk++)
            {
                int ja = nid(L, i + di, j + dj, k + dk);
                if(ja < nn) m[++ia] = ja;
            }
    return ia;
}

/* Boundary_conditions.c:926-947 */
void ccu_r_strip_bcs(const ccu_r_level *L, double *v)
{
    int n;
    for(n = 0; n < L->nno; n++)
    {
        if(L->node[n] & CCU_R_VBX) v[3 * n] = 0.0;
        if(L->node[n] & CCU_R_VBY) v[3 * n + 1] = 0.0;
        if(L->node[n] & CCU_R_VBZ) v[3 * n + 2] = 0.0;
    }
}

/* Element_calculations.c:552-621 (n_assemble_del2_u) */
void ccu_r_matvec(const ccu_r_level *L, const double *u, double *Au, int strip)
{
    int n, s, d, i, j, k, m[14];
    for(n = 0; n < L->neq; n++) Au[n] = 0.0;
    for(n = 0; n < L->nno; n++)
    {
        const float *B1 = L->k1 + (size_t)n * MAXEQ, *B2 = L->k2 + (size_t)n * MAXEQ, *B3 = L->k3 + (size_t)n * MAXEQ;
        const double U1 = u[3 * n], U2 = u[3 * n + 1], U3 = u[3 * n + 2];
        int ns;
        nijk(L, n, &i, &j, &k);
        ns = lower_nbrs(L, i, j, k, m);
        for(s = 1; s <= ns; s++)
            for(d = 0; d < 3; d++)
            {
                const double UU = u[3 * m[s] + d];
                Au[3 * n] += B1[3 * s + d] * UU;
                Au[3 * n + 1] += B2[3 * s + d] * UU;
                Au[3 * n + 2] += B3[3 * s + d] * UU;
            }
        for(d = 0; d < 3; d++)
            Au[3 * n + d] += B1[d] * U1 + B2[d] * U2 + B3[d] * U3;
        for(s = 1; s <= ns; s++)
            for(d = 0; d < 3; d++)
                Au[3 * m[s] + d] += B1[3 * s + d] * U1 + B2[3 * s + d] * U2 + B3[3 * s + d] * U3;
    }
    if(strip) ccu_r_strip_bcs(L, Au);
}

/* General_matrix_functions.c:1160-1361 (gauss_seidel, 3-D branch, no OFFSIDE nodes) */
void ccu_r_gauss_seidel(const ccu_r_level *L, double *d0, const double *F, double *Ad, int cycles, int guess)
{
    int n, s, d, i, j, k, m[14], count;
    float *temp = (float *)malloc(sizeof(float) * (L->neq + 2));
    if(guess) ccu_r_matvec(L, d0, Ad, 1);
    else for(n = 0; n < L->neq; n++) d0[n] = Ad[n] = 0.0;
    for(count = 0; count < cycles; count++)
    {
        for(n = 0; n < L->neq + 2; n++) temp[n] = 0.0f;
        for(n = 0; n < L->nno; n++)
        {
            const float *B1 = L->k1 + (size_t)n * MAXEQ, *B2 = L->k2 + (size_t)n * MAXEQ, *B3 = L->k3 + (size_t)n * MAXEQ;
            const int e1 = 3 * n, e2 = 3 * n + 1, e3 = 3 * n + 2;
            int ns;
            nijk(L, n, &i, &j, &k);
            ns = lower_nbrs(L, i, j, k, m);
            for(s = 1; s <= ns; s++)
                for(d = 0; d < 3; d++)
                {
                    const double UU = temp[3 * m[s] + d];
                    Ad[e1] += B1[3 * s + d] * UU;
                    Ad[e2] += B2[3 * s + d] * UU;
                    Ad[e3] += B3[3 * s + d] * UU;
                }
            temp[e1] = (F[e1] - Ad[e1]) * L->BI[e1];
            temp[e2] = (F[e2] - Ad[e2]) * L->BI[e2];
            temp[e3] = (F[e3] - Ad[e3]) * L->BI[e3];
            for(d = 0; d < 3; d++)
                Ad[3 * n + d] += B1[d] * temp[e1] + B2[d] * temp[e2] + B3[d] * temp[e3];
            for(s = 1; s <= ns; s++)
                for(d = 0; d < 3; d++)
                    Ad[3 * m[s] + d] += B1[3 * s + d] * temp[e1] + B2[3 * s + d] * temp[e2] + B3[3 * s + d] * temp[e3];
            d0[e1] += temp[e1];
            d0[e2] += temp[e2];
            d0[e3] += temp[e3];
        }
    }
    free(temp);
}

/* ------------------------------------------------------------------------------------------
 * 8-colour Gauss-Seidel model (what the CUDA path computes).
 * colour(n) = 4*(i&1) + 2*(j&1) + (k&1).  Nodes of one colour share no stencil neighbour, so a
 * colour pass is order-independent; it is stated here on the reference's half-stored arrays:
 * row of n = self block + blocks n owns (lower neighbours) + transposed blocks its upper
 * neighbours own.
 * ------------------------------------------------------------------------------------------ */
static inline int colour_of(int i, int j, int k) { return ((i & 1) << 2) | ((j & 1) << 1) | (k & 1); }

static int slot_in(const ccu_r_level *L, int n, int m)       /* slot of lower neighbour m in n's row */
{
    int i, j, k, lst[14], ns, s;
    nijk(L, n, &i, &j, &k);
    ns = lower_nbrs(L, i, j, k, lst);
    for(s = 1; s <= ns; s++) if(lst[s] == m) return s;
    fprintf(stderr, "restate: slot_in failed\n"); abort();
}

/* B[a][b] = K(row dof a of n, col dof b of m) from the half-stored reference arrays */
static void get_block(const ccu_r_level *L, int n, int m, double B[3][3])
{
    int a, b;
    const float *kk[3];
    if(m <= n)
    {
        const int s = (m == n) ? 0 : slot_in(L, n, m);
        kk[0] = L->k1 + (size_t)n * MAXEQ; kk[1] = L->k2 + (size_t)n * MAXEQ; kk[2] = L->k3 + (size_t)n * MAXEQ;
        for(a = 0; a < 3; a++) for(b = 0; b < 3; b++) B[a][b] = kk[a][3 * s + b];
    }
    else
    {
        const int s = slot_in(L, m, n);
        kk[0] = L->k1 + (size_t)m * MAXEQ; kk[1] = L->k2 + (size_t)m * MAXEQ; kk[2] = L->k3 + (size_t)m * MAXEQ;
        for(a = 0; a < 3; a++) for(b = 0; b < 3; b++) B[a][b] = kk[b][3 * s + a];
    }
}

/* sum over the lower-numbered (which<0) or higher-numbered (which>0) neighbours of n */
static void tri_product(const ccu_r_level *L, int n, const double *x, int which, double out[3])
{
    int i, j, k, o, a;
    double B[3][3];
    nijk(L, n, &i, &j, &k);
    const int c = colour_of(i, j, k);
    out[0] = out[1] = out[2] = 0.0;
    for(o = 0; o < 27; o++)
    {
        const int di = o / 9 - 1, dj = (o / 3) % 3 - 1, dk = o % 3 - 1;
        const int ii = i + di, jj = j + dj, kk = k + dk;
        int cm, m;
        if(o == 13) continue;
        if(ii < 0 || ii >= L->noy || jj < 0 || jj >= L->nox || kk < 0 || kk >= L->noz) continue;
        (void)cm; (void)c;
        m = nid(L, ii, jj, kk);
        if((which < 0 && m > n) || (which > 0 && m < n)) continue;      /* lower / upper neighbours (natural numbering) */
        get_block(L, n, m, B);
        for(a = 0; a < 3; a++)
            out[a] += B[a][0] * x[3 * m] + B[a][1] * x[3 * m + 1] + B[a][2] * x[3 * m + 2];
    }
}

static void self_product(const ccu_r_level *L, int n, const double *x, double out[3])
{
    double B[3][3];
    int a;
    get_block(L, n, n, B);
    for(a = 0; a < 3; a++)
        out[a] = B[a][0] * x[3 * n] + B[a][1] * x[3 * n + 1] + B[a][2] * x[3 * n + 2];
}

void ccu_r_mc_matvec(const ccu_r_level *L, const double *u, double *Au, int strip)
{
    int n, a;
    for(n = 0; n < L->nno; n++)
    {
        double lo[3], up[3], s[3];
        tri_product(L, n, u, -1, lo);
        tri_product(L, n, u, +1, up);
        self_product(L, n, u, s);
        for(a = 0; a < 3; a++) Au[3 * n + a] = (s[a] + lo[a]) + up[a];
    }
    if(strip) ccu_r_strip_bcs(L, Au);
}

/* One colour pass = every node of that colour relaxed from the full row with the values its
 * neighbours hold now (neighbours never share its colour).  Colours go 7,6,...,0 on every sweep:
 * the design study below (ccu_r_ordered_gs) shows this fixed order matches the lexicographic
 * smoother's multigrid convergence, while alternating directions does not. */
void ccu_r_mc_gauss_seidel(const ccu_r_level *L, double *d0, const double *F, double *Ad, int cycles, int guess)
{
    int n, c, sweep, i, j, k, a;
    if(!guess) for(n = 0; n < L->neq; n++) d0[n] = 0.0;
    for(sweep = 0; sweep < cycles; sweep++)
        for(c = 7; c >= 0; c--)
            for(n = 0; n < L->nno; n++)
            {
                double lo[3], up[3], s[3];
                nijk(L, n, &i, &j, &k);
                if(colour_of(i, j, k) != c) continue;
                tri_product(L, n, d0, -1, lo); tri_product(L, n, d0, +1, up); self_product(L, n, d0, s);
                for(a = 0; a < 3; a++)
                {
                    const double r = F[3 * n + a] - ((s[a] + lo[a]) + up[a]);
                    const float t = (float)(r * L->BI[3 * n + a]);   /* fp32 correction, General_matrix_functions.c:1172,1250 */
                    d0[3 * n + a] += t;
                }
            }
    ccu_r_mc_matvec(L, d0, Ad, 1);
}

/* Solver_multigrid.c:72-159 (project_vector, 3-D) */
void ccu_r_project_vector(const ccu_r_level *fine, const ccu_r_level *coarse, const double *AU, double *AD)
{
    int el, i, j, n;
    for(n = 0; n < coarse->neq; n++) AD[n] = 0.0;
    for(el = 0; el < coarse->nel; el++)
    {
        const int ez = el % elz_(coarse), ex = (el / elz_(coarse)) % elx_(coarse), ey = el / (elz_(coarse) * elx_(coarse));
        for(i = 1; i <= 8; i++)
        {
            const int e1 = eid(fine, 2 * ey + OFFS[i][2], 2 * ex + OFFS[i][1], 2 * ez + OFFS[i][0]);
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for(j = 1; j <= 8; j++)
            {
                const int nf = enode(fine, e1, j);
                a1 += AU[3 * nf]; a2 += AU[3 * nf + 1]; a3 += AU[3 * nf + 2];
            }
            n = enode(coarse, el, i);
            AD[3 * n] += coarse->TWW[(size_t)el * 8 + i - 1] * a1;
            AD[3 * n + 1] += coarse->TWW[(size_t)el * 8 + i - 1] * a2;
            AD[3 * n + 2] += coarse->TWW[(size_t)el * 8 + i - 1] * a3;
        }
    }
    for(n = 0; n < coarse->nno; n++)
    {
        AD[3 * n] = AD[3 * n] * coarse->MASS[n];
        AD[3 * n + 1] = AD[3 * n + 1] * coarse->MASS[n];
        AD[3 * n + 2] = AD[3 * n + 2] * coarse->MASS[n];
    }
}

/* first (lowest-numbered) element containing a node: NEI[level].element[8*(node-1)] (Construct_arrays.c:88-99) */
static inline int first_elt(const ccu_r_level *L, int i, int j, int k)
{
    return eid(L, i > 0 ? i - 1 : 0, j > 0 ? j - 1 : 0, k > 0 ? k - 1 : 0);
}

/* Solver_multigrid.c:173-298 (interp_vector) + :581-634 (un_inject_vector) */
void ccu_r_interp_vector(const ccu_r_level *coarse, const ccu_r_level *fine, const double *AD, double *AU)
{
    int i, j, k, d, n;
    for(n = 0; n < fine->neq; n++) AU[n] = 0.0;
    for(i = 0; i < coarse->noy; i++)
        for(j = 0; j < coarse->nox; j++)
            for(k = 0; k < coarse->noz; k++)
            {
                const int nc = nid(coarse, i, j, k), nf = nid(fine, 2 * i, 2 * j, 2 * k);
                for(d = 0; d < 3; d++) AU[3 * nf + d] = AD[3 * nc + d];
            }
    /* x direction */
    for(i = 0; i < fine->noy; i += 2)
        for(k = 0; k < fine->noz; k += 2)
            for(j = 1; j < fine->nox - 1; j += 2)
            {
                const int n0 = nid(fine, i, j, k), n1 = nid(fine, i, j - 1, k), n2 = nid(fine, i, j + 1, k);
                const float x1 = fine->eco_size[(size_t)first_elt(fine, i, j - 1, k) * 3 + 0];
                const float x2 = fine->eco_size[(size_t)first_elt(fine, i, j + 1, k) * 3 + 0];
                const float w1 = x2 / (x1 + x2), w2 = x1 / (x1 + x2);
                for(d = 0; d < 3; d++) AU[3 * n0 + d] = w1 * AU[3 * n1 + d] + w2 * AU[3 * n2 + d];
            }
    /* z direction */
    for(i = 0; i < fine->noy; i += 2)
        for(j = 0; j < fine->nox; j++)
            for(k = 1; k < fine->noz - 1; k += 2)
            {
                const int n0 = nid(fine, i, j, k), n1 = n0 - 1, n2 = n0 + 1;
                const float x1 = fine->eco_size[(size_t)first_elt(fine, i, j, k - 1) * 3 + 2];
                const float x2 = fine->eco_size[(size_t)first_elt(fine, i, j, k + 1) * 3 + 2];
                const float w1 = x2 / (x1 + x2), w2 = x1 / (x1 + x2);
                for(d = 0; d < 3; d++) AU[3 * n0 + d] = w1 * AU[3 * n1 + d] + w2 * AU[3 * n2 + d];
            }
    /* y direction */
    for(j = 0; j < fine->nox; j++)
        for(k = 0; k < fine->noz; k++)
            for(i = 1; i < fine->noy - 1; i += 2)
            {
                const int n0 = nid(fine, i, j, k), n1 = nid(fine, i - 1, j, k), n2 = nid(fine, i + 1, j, k);
                const float x1 = fine->eco_size[(size_t)first_elt(fine, i - 1, j, k) * 3 + 1];
                const float x2 = fine->eco_size[(size_t)first_elt(fine, i + 1, j, k) * 3 + 1];
                const float w1 = x2 / (x1 + x2), w2 = x1 / (x1 + x2);
                for(d = 0; d < 3; d++) AU[3 * n0 + d] = w1 * AU[3 * n1 + d] + w2 * AU[3 * n2 + d];
            }
}

/* Element_calculations.c:691-720 */
void ccu_r_div_u(const ccu_r_level *L, const double *U, double *divU)
{
    int e, a;
    for(e = 0; e < L->npno; e++) divU[e] = 0.0;
    for(a = 1; a <= 8; a++)
        for(e = 0; e < L->nel; e++)
        {
            const int n = enode(L, e, a);
            const float *g = L->elt_del + (size_t)e * 24 + 3 * (a - 1);
            divU[e] += g[0] * U[3 * n] + g[1] * U[3 * n + 1] + g[2] * U[3 * n + 2];
        }
}

/* Element_calculations.c:727-769 */
void ccu_r_grad_p(const ccu_r_level *L, const double *P, double *gradP)
{
    int e, a, n;
    for(n = 0; n < L->neq; n++) gradP[n] = 0.0;
    for(e = 0; e < L->nel; e++)
    {
        if(P[e] == 0.0) continue;
        for(a = 1; a <= 8; a++)
        {
            const float *g = L->elt_del + (size_t)e * 24 + 3 * (a - 1);
            n = enode(L, e, a);
            gradP[3 * n] += g[0] * P[e];
            gradP[3 * n + 1] += g[1] * P[e];
            gradP[3 * n + 2] += g[2] * P[e];
        }
    }
    ccu_r_strip_bcs(L, gradP);
}

/* Global_operations.c:339-357, 359-375 (single rank: IDD is all ones) */
double ccu_r_vdot(const ccu_r_level *L, const double *a, const double *b)
{
    double t = 0.0; int i;
    for(i = 0; i < L->neq; i++) t += a[i] * b[i];
    return t;
}
double ccu_r_pdot(const ccu_r_level *L, const double *a, const double *b)
{
    double t = 0.0; int i;
    for(i = 0; i < L->npno; i++) t += a[i] * b[i];
    return t;
}

void ccu_r_ordered_gs(const ccu_r_level *L, double *d0, const double *F, double *Ad, int cycles, int guess, int mode);
static void smooth(const ccu_r_mg *M, int lev, double *d0, const double *F, double *Ad, int cycles, int guess)
{
    if(M->smoother >= 10) ccu_r_ordered_gs(&M->lev[lev], d0, F, Ad, cycles, guess, M->smoother - 10);
    else if(M->smoother == 0) ccu_r_gauss_seidel(&M->lev[lev], d0, F, Ad, cycles, guess);
    else ccu_r_mc_gauss_seidel(&M->lev[lev], d0, F, Ad, cycles, guess);
}

/* General_matrix_functions.c:525-653 */
double ccu_r_multi_grid(const ccu_r_mg *M, double *d1, double *F, double acc)
{
    const int levmin = M->levmin, levmax = M->levmax;
    double *res[12], *rhs[12], *AU[12], *vel[12], *fl[12], *del_vel[12];
    int lev, dlev, ulev, i, Vn;
    double residual;
    (void)acc;
    for(i = levmin; i <= levmax; i++)
    {
        const size_t nb = sizeof(double) * (M->lev[i].neq + 2);
        res[i] = calloc(1, nb); rhs[i] = calloc(1, nb); AU[i] = calloc(1, nb);
        vel[i] = calloc(1, nb); fl[i] = calloc(1, nb); del_vel[i] = calloc(1, nb);
    }
    memcpy(fl[levmax], F, sizeof(double) * M->lev[levmax].neq);
    for(lev = levmax; lev > levmin; lev--)
    {
        ccu_r_project_vector(&M->lev[lev], &M->lev[lev - 1], fl[lev], fl[lev - 1]);
        ccu_r_strip_bcs(&M->lev[lev - 1], fl[lev - 1]);
    }
    smooth(M, levmin, vel[levmin], fl[levmin], AU[levmin], M->v_steps_low, 0);
    for(lev = levmin + 1; lev <= levmax; lev++)
    {
        ccu_r_interp_vector(&M->lev[lev - 1], &M->lev[lev], vel[lev - 1], vel[lev]);
        ccu_r_strip_bcs(&M->lev[lev], vel[lev]);
        memcpy(rhs[lev], fl[lev], sizeof(double) * M->lev[lev].neq);
        for(Vn = 1; Vn <= M->mg_cycle; Vn++)
        {
            for(dlev = lev; dlev >= levmin + 1; dlev--)
            {
                const int cycles = (dlev == levmax) ? M->v_steps_high : M->down_heavy;
                smooth(M, dlev, vel[dlev], rhs[dlev], AU[dlev], cycles, dlev == lev);
                for(i = 0; i < M->lev[dlev].neq; i++) res[dlev][i] = rhs[dlev][i] - AU[dlev][i];
                ccu_r_project_vector(&M->lev[dlev], &M->lev[dlev - 1], res[dlev], rhs[dlev - 1]);
                ccu_r_strip_bcs(&M->lev[dlev - 1], rhs[dlev - 1]);
            }
            smooth(M, levmin, vel[levmin], rhs[levmin], AU[levmin], M->v_steps_low, 0);
            for(ulev = levmin + 1; ulev <= lev; ulev++)
            {
                const int cycles = (ulev == levmax) ? M->v_steps_high : M->up_heavy;
                double AudotAu, alpha;
                ccu_r_interp_vector(&M->lev[ulev - 1], &M->lev[ulev], vel[ulev - 1], del_vel[ulev]);
                ccu_r_strip_bcs(&M->lev[ulev], del_vel[ulev]);
                smooth(M, ulev, del_vel[ulev], res[ulev], AU[ulev], cycles, 1);
                AudotAu = ccu_r_vdot(&M->lev[ulev], AU[ulev], AU[ulev]);
                alpha = ccu_r_vdot(&M->lev[ulev], AU[ulev], res[ulev]) / AudotAu;
                for(i = 0; i < M->lev[ulev].neq; i++) vel[ulev][i] += alpha * del_vel[ulev][i];
                if(ulev == levmax)
                    for(i = 0; i < M->lev[ulev].neq; i++) res[ulev][i] -= alpha * AU[ulev][i];
            }
        }
    }
    for(i = 0; i < M->lev[levmax].neq; i++) { F[i] = res[levmax][i]; d1[i] = vel[levmax][i]; }
    residual = sqrt(ccu_r_vdot(&M->lev[levmax], F, F) / M->lev[levmax].neq);
    for(i = levmin; i <= levmax; i++) { free(res[i]); free(rhs[i]); free(AU[i]); free(vel[i]); free(fl[i]); free(del_vel[i]); }
    return residual;
}

/* General_matrix_functions.c:368-520 (multigrid branch) */
int ccu_r_solve_del2_u(const ccu_r_mg *M, double *d0, const double *F, double acc, int *mg_cycles_out)
{
    const int neq = M->lev[M->levmax].neq;
    double *r = malloc(sizeof(double) * (neq + 2)), *D1 = calloc(neq + 2, sizeof(double));
    double residual, r0;
    int i, valid, count = 0;
    for(i = 0; i < neq; i++) { r[i] = F[i]; d0[i] = 0.0; }
    r0 = residual = sqrt(ccu_r_vdot(&M->lev[M->levmax], r, r) / neq);
    acc = (acc > r0 * M->accuracy) ? acc : r0 * M->accuracy;
    valid = (residual < acc) ? 0 : 1;
    while(residual > acc)
    {
        residual = ccu_r_multi_grid(M, D1, r, acc);
        for(i = 0; i < neq; i++) { d0[i] += D1[i]; D1[i] = 0.0; }
        count++;
        if(count > 200) break;      /* guard for the checker only; the reference loops unbounded */
    }
    if(mg_cycles_out) *mg_cycles_out = count;
    free(r); free(D1);
    return valid;
}

/* Stokes_flow_Incomp.c:295-497 */
float ccu_r_solve_Ahat_p_fhat(const ccu_r_mg *M, double *V, double *P, const double *F, double imp, int *steps_max, double *hist)
{
    const ccu_r_level *L = &M->lev[M->levmax];
    const int neq = L->neq, npno = L->npno;
    double *r0 = calloc(npno, 8), *r1 = calloc(npno, 8), *r2 = calloc(npno, 8), *z0 = calloc(npno, 8), *z1 = calloc(npno, 8);
    double *s1 = calloc(npno, 8), *s2 = calloc(npno, 8), *Ah = calloc(neq + 2, 8), *u1 = calloc(neq + 2, 8), *shuffle;
    double alpha, delta, s2dotAhat, r0dotr0, r1dotz1, residual, v_res;
    float dpressure, dvelocity, vdotv, pdotp;
    int i, j, count, valid;

    v_res = sqrt(ccu_r_vdot(L, F, F) / neq);
    ccu_r_grad_p(L, P, Ah);
    if(M->smoother == 0) ccu_r_matvec(L, V, u1, 1); else ccu_r_mc_matvec(L, V, u1, 1);
    for(i = 0; i < neq; i++) Ah[i] = F[i] - Ah[i] - u1[i];
    ccu_r_strip_bcs(L, Ah);
    valid = ccu_r_solve_del2_u(M, u1, Ah, imp * v_res, NULL);
    ccu_r_strip_bcs(L, u1);
    for(i = 0; i < neq; i++) V[i] += u1[i];
    ccu_r_div_u(L, V, r1);
    residual = sqrt(ccu_r_pdot(L, r1, r1) / npno);
    count = 0;
    dpressure = 1.0; dvelocity = 1.0;
    while((count < *steps_max) && (dpressure >= imp || dvelocity >= imp))
    {
        for(j = 0; j < npno; j++) z1[j] = L->BPI[j] * r1[j];
        r1dotz1 = ccu_r_pdot(L, r1, z1);
        if(count == 0) for(j = 0; j < npno; j++) s2[j] = z1[j];
        else
        {
            r0dotr0 = ccu_r_pdot(L, r0, z0);
            delta = r1dotz1 / r0dotr0;
            for(j = 0; j < npno; j++) s2[j] = z1[j] + delta * s1[j];
        }
        ccu_r_grad_p(L, s2, Ah);
        valid = ccu_r_solve_del2_u(M, u1, Ah, imp * v_res, NULL);
        ccu_r_strip_bcs(L, u1);
        ccu_r_div_u(L, u1, Ah);
        s2dotAhat = ccu_r_pdot(L, s2, Ah);
        alpha = valid ? r1dotz1 / s2dotAhat : 0.0;
        for(j = 0; j < npno; j++) { r2[j] = r1[j] - alpha * Ah[j]; P[j] += alpha * s2[j]; }
        for(j = 0; j < neq; j++) V[j] -= alpha * u1[j];
        ccu_r_div_u(L, V, Ah);
        vdotv = ccu_r_vdot(L, V, V);
        pdotp = ccu_r_pdot(L, P, P);
        dpressure = alpha * sqrt(ccu_r_pdot(L, s2, s2) / (1.0e-32 + pdotp));
        dvelocity = alpha * sqrt(ccu_r_vdot(L, u1, u1) / (1.0e-32 + vdotv));
        if(hist) { hist[4 * count] = sqrt(vdotv / neq); hist[4 * count + 1] = dvelocity; hist[4 * count + 2] = sqrt(pdotp / npno); hist[4 * count + 3] = dpressure; }
        count++;
        shuffle = s1; s1 = s2; s2 = shuffle;
        shuffle = r0; r0 = r1; r1 = r2; r2 = shuffle;
        shuffle = z0; z0 = z1; z1 = shuffle;
    }
    *steps_max = count;
    free(r0); free(r1); free(r2); free(z0); free(z1); free(s1); free(s2); free(Ah); free(u1);
    return (float)residual;
}

/* General_matrix_functions.c:661-770: Jacobi-preconditioned conjugate gradients on K (Solver=cgrad).
 * d0 out (zeroed first), returns the residual norm sqrt(r.r/neq); *cycles in = max iterations, out = done. */
double ccu_r_conj_grad(const ccu_r_level *L, double *d0, const double *F, double acc, int *cycles)
{
    const int n = L->neq, steps = *cycles;
    double *r1 = calloc(n + 2, 8), *r2 = calloc(n + 2, 8), *z1 = calloc(n + 2, 8), *p1 = calloc(n + 2, 8), *p2 = calloc(n + 2, 8),
           *Ap = calloc(n + 2, 8), *sh;
    double residual, alpha, beta, dotprod, dotr1z1, dotr0z0 = 0.0;
    int i, count = 0;
    for(i = 0; i < n; i++) { r1[i] = F[i]; d0[i] = 0.0; }
    residual = sqrt(ccu_r_vdot(L, r1, r1) / n);
    while(((residual > acc) && (count < steps)) || count == 0)
    {
        for(i = 0; i < n; i++) z1[i] = L->BI[i] * r1[i];
        dotr1z1 = ccu_r_vdot(L, r1, z1);
        if(count == 0) for(i = 0; i < n; i++) p2[i] = z1[i];
        else { beta = dotr1z1 / dotr0z0; for(i = 0; i < n; i++) p2[i] = z1[i] + beta * p1[i]; }
        dotr0z0 = dotr1z1;
        ccu_r_matvec(L, p2, Ap, 1);
        dotprod = ccu_r_vdot(L, p2, Ap);
        alpha = (dotprod == 0.0) ? 1.0e-3 : dotr1z1 / dotprod;
        for(i = 0; i < n; i++) { d0[i] += alpha * p2[i]; r2[i] = r1[i] - alpha * Ap[i]; }
        residual = sqrt(ccu_r_vdot(L, r2, r2) / n);
        sh = r1; r1 = r2; r2 = sh;
        sh = p1; p1 = p2; p2 = sh;
        count++;
    }
    *cycles = count;
    ccu_r_strip_bcs(L, d0);
    free(r1); free(r2); free(z1); free(p1); free(p2); free(Ap);
    return residual;
}

/* ------------------------------------------------------------------------------------------
 * Experimental orderings (design studies only; not a checker): point-block GS with the
 * reference's scalar-BI update, nodes visited in a chosen order, residual from full rows.
 *   mode 0 lexicographic forward          mode 1 8-colour forward-only
 *   mode 2 8-colour symmetric             mode 3 lexicographic symmetric
 *   mode 4 4-colour (x,y parity) lines, z ascending within a line, forward
 *   mode 5 as 4 but symmetric (z descending on backward sweeps)
 *   mode 9 tile-ordered 8-colour: tiles of g_ccu_r_tile[] = TI x TJ x TK colour cells (cell = (index>>1)+1, the
 *          device layout's halo offset), tile colours 7..0 outside, node colours 7..0 inside each tile -- the order
 *          of the round-1 tile kernels (removed in round 2; kept as an ordering study)
 *   mode 10 column-ordered: columns of g_ccu_r_col[] = TI (y) x TJ (x) NODES spanning all of z, 4-coloured by the parity
 *          of their column indices (colours 3..0); inside a column the z layers ascending, and inside a layer the four
 *          (y,x)-parity colours 3..0 -- the order of ccu_k_col (csrc/ccu_col.cuh); THIS mode is the checker for that kernel
 * ------------------------------------------------------------------------------------------ */
static void full_row(const ccu_r_level *L, int n, const double *x, double out[3])
{
    double lo[3], up[3], s[3]; int a;
    tri_product(L, n, x, -1, lo); tri_product(L, n, x, +1, up); self_product(L, n, x, s);
    for(a = 0; a < 3; a++) out[a] = lo[a] + s[a] + up[a];
}
int g_ccu_r_tile[3] = { 2, 4, 16 };      /* TI (y), TJ (x), TK (z) colour cells per tile */
int g_ccu_r_col[2] = { 8, 4 };           /* TI (y), TJ (x) nodes per column (mode 10) */
static int ord_key(const ccu_r_level *L, int n, int mode, int back)
{
    int i, j, k; nijk(L, n, &i, &j, &k);
    if(mode == 9)
    {
        const int ti = ((i >> 1) + 1) / g_ccu_r_tile[0], tj = ((j >> 1) + 1) / g_ccu_r_tile[1], tk = ((k >> 1) + 1) / g_ccu_r_tile[2];
        const int ntj = ((L->nox + 1) / 2 + 2) / g_ccu_r_tile[1] + 1, ntk = ((L->noz + 1) / 2 + 2) / g_ccu_r_tile[2] + 1;
        const int tc = ((ti & 1) << 2) | ((tj & 1) << 1) | (tk & 1);
        const int tile = tk + ntk * (tj + ntj * ti);
        return ((7 - tc) * (1 << 20) + tile) * 8 + (7 - colour_of(i, j, k));
    }
    if(mode == 10)
    {
        const int I = i / g_ccu_r_col[0], J = j / g_ccu_r_col[1];
        const int nJ = (L->nox + g_ccu_r_col[1] - 1) / g_ccu_r_col[1], nI = (L->noy + g_ccu_r_col[0] - 1) / g_ccu_r_col[0];
        const int cc = ((I & 1) << 1) | (J & 1), c2 = ((i & 1) << 1) | (j & 1);
        return ((((3 - cc) * nI * nJ + I * nJ + J) * L->noz) + k) * 4 + (3 - c2);
    }
    if(mode == 0 || mode == 3) return back ? (L->nno - 1 - n) : n;
    if(mode == 1 || mode == 2) { int c = colour_of(i, j, k); return back ? 7 - c : c; }
    if(mode == 8) { int c = colour_of(i, j, k); return back ? c : 7 - c; }            /* sym starting 7..0 */
    if(mode == 6) { int c = colour_of(i, j, k); return 7 - c; }                       /* 8-colour, always 7..0 */
    if(mode == 7) { int c = colour_of(i, j, k); return back ? (c == 0 ? 0 : 8 - c) : c; } /* sym but colour 0 first both ways */
    { int c = ((i & 1) << 1) | (j & 1); int kk = back ? (L->noz - 1 - k) : k; if(back) c = 3 - c; return c * L->noz + kk; }
}
static const ccu_r_level *g_sortL; static int g_sortmode, g_sortback;
static int cmp_ord(const void *a, const void *b)
{
    int na = *(const int *)a, nb = *(const int *)b;
    int ka = ord_key(g_sortL, na, g_sortmode, g_sortback), kb = ord_key(g_sortL, nb, g_sortmode, g_sortback);
    if(ka != kb) return ka < kb ? -1 : 1;
    return na < nb ? -1 : (na > nb);
}
void ccu_r_ordered_gs(const ccu_r_level *L, double *d0, const double *F, double *Ad, int cycles, int guess, int mode)
{
    int n, s, a, q;
    int *ord = malloc(sizeof(int) * L->nno);
    if(!guess) for(n = 0; n < L->neq; n++) d0[n] = 0.0;
    for(s = 0; s < cycles; s++)
    {
        const int back = ((mode == 2 || mode == 3 || mode == 5 || mode == 7 || mode == 8) && (s & 1));
        for(n = 0; n < L->nno; n++) ord[n] = n;
        g_sortL = L; g_sortmode = mode; g_sortback = back;
        qsort(ord, L->nno, sizeof(int), cmp_ord);
        for(q = 0; q < L->nno; q++)
        {
            double r[3];
            n = ord[q];
            full_row(L, n, d0, r);
            for(a = 0; a < 3; a++) { const float t = (float)((F[3 * n + a] - r[a]) * L->BI[3 * n + a]); d0[3 * n + a] += t; }
        }
    }
    for(n = 0; n < L->nno; n++) full_row(L, n, d0, Ad + 3 * n);
    ccu_r_strip_bcs(L, Ad);
    free(ord);
}
int g_ccu_r_exp_mode = -1;
