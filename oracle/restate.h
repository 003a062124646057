/* restate.h -- CPU oracle: plain-C restatement of CitcomCU's Stokes hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under citcomcu_b200/ may include, link or
 * call this; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA path.  Parity status: PINNED -- every
 * function here is validated bit-for-bit (or to the stated tolerance) against
 * outputs of the reference itself (oracle/_ref/ref_harness known-answer dumps,
 * tests/test_oracle_restate.py) and against committed fixtures in tests/golden.
 *
 * Conventions: all vectors 0-based; nodes n = k + noz*(j + nox*i) with k the z
 * index (fastest), j the x index, i the y index (Construct_arrays.c:158-165,
 * 311-317); equation 3n+d; pressure/element e = ez + elz*(ex + elx*ey).
 */
#ifndef CCU_RESTATE_H
#define CCU_RESTATE_H
#ifdef __cplusplus
extern "C" {
#endif

#define CCU_R_VBX 0x2u
#define CCU_R_VBZ 0x4u
#define CCU_R_VBY 0x8u

typedef struct {
    int nox, noy, noz;          /* nodes in x, y, z */
    int nno, nel, neq, npno;
    const unsigned *node;       /* [nno] NODE flags (global_defs.h:65-89) */
    const float *k1, *k2, *k3;  /* [nno*42] Eqn_k1-3, reference layout (Construct_arrays.c:359) */
    const double *BI;           /* [neq] */
    const double *BPI;          /* [npno] */
    const float *elt_del;       /* [nel*24] */
    const float *TWW;           /* [nel*8] */
    const float *MASS;          /* [nno] */
    const float *eco_size;      /* [nel*3] element sizes (x, y, z) = ECO.size[1..3] */
} ccu_r_level;

typedef struct {
    int levmin, levmax;
    int v_steps_low, v_steps_high, down_heavy, up_heavy, mg_cycle;
    int smoother;               /* 0 = reference lexicographic GS, 1 = 8-colour GS model (colours 7..0) */
    double accuracy;            /* E->control.accuracy */
    ccu_r_level lev[12];
} ccu_r_mg;

void ccu_r_strip_bcs(const ccu_r_level *L, double *v);
void ccu_r_matvec(const ccu_r_level *L, const double *u, double *Au, int strip);
void ccu_r_gauss_seidel(const ccu_r_level *L, double *d0, const double *F, double *Ad, int cycles, int guess);
void ccu_r_mc_matvec(const ccu_r_level *L, const double *u, double *Au, int strip);
void ccu_r_mc_gauss_seidel(const ccu_r_level *L, double *d0, const double *F, double *Ad, int cycles, int guess);
void ccu_r_project_vector(const ccu_r_level *fine, const ccu_r_level *coarse, const double *AU, double *AD);
void ccu_r_interp_vector(const ccu_r_level *coarse, const ccu_r_level *fine, const double *AD, double *AU);
void ccu_r_div_u(const ccu_r_level *L, const double *U, double *divU);
void ccu_r_grad_p(const ccu_r_level *L, const double *P, double *gradP);
double ccu_r_vdot(const ccu_r_level *L, const double *a, const double *b);
double ccu_r_pdot(const ccu_r_level *L, const double *a, const double *b);
double ccu_r_conj_grad(const ccu_r_level *L, double *d0, const double *F, double acc, int *cycles);
double ccu_r_multi_grid(const ccu_r_mg *M, double *d1, double *F, double acc);
int ccu_r_solve_del2_u(const ccu_r_mg *M, double *d0, const double *F, double acc, int *mg_cycles_out);
float ccu_r_solve_Ahat_p_fhat(const ccu_r_mg *M, double *V, double *P, const double *F, double imp,
                              int *steps_max, double *hist /* [steps][4]: v, dv, p, dp; may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
