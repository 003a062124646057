/* citcom_dropin_funcs.c -- function-level bindings: the inner functions of the hot path with the
 * reference's own signatures (src/prototypes.h) on top of the C ABI, so that a maintainer can move the path to the
 * device one function at a time.
 *
 *     CCU_DROPIN_STOKES=0 CCU_DROPIN_FUNCS=n_assemble_del2_u,assemble_div_u,assemble_grad_p \
 *         LD_PRELOAD=dropin/libcitcomcu_dropin.so  citcom.mpi  input_file
 *
 * Every function named in CCU_DROPIN_FUNCS (comma separated, or "all") runs on the device; the others forward to the
 * reference's own definition (dlsym RTLD_NEXT), so the library can be preloaded whole.  The operator arrays are the
 * REFERENCE's: after each construct_stiffness_B_matrix (Construct_arrays.c:834, wrapped here: the reference's own runs,
 * then the arrays are marked stale) E->Eqn_k1-3, BI, BPI, elt_del, TWW, MASS, ECO.size go to the device in the layouts
 * the reference keeps them in; vectors cross the boundary on every call (host pointers in, host pointers out), exactly the
 * semantics of the functions they replace.  tests/test_gpu_dropin.py runs the unmodified reference with several such
 * sets and compares with the pure-CPU run.
 *
 * Bound here: n_assemble_del2_u / assemble_del2_u / e_assemble_del2_u (Element_calculations.c:552,480,494), gauss_seidel
 * (General_matrix_functions.c:1160), multi_grid (:525), solve_del2_u (:368), conj_grad (:661), project_vector /
 * interp_vector (Solver_multigrid.c:72,173), strip_bcs_from_residual (Boundary_conditions.c:926), assemble_div_u /
 * assemble_grad_p (Element_calculations.c:691,727), global_vdot / global_pdot (Global_operations.c:339,359),
 * solve_Ahat_p_fhat (Stokes_flow_Incomp.c:295), PG_timestep_particle (Advection_diffusion.c:128, with its static on_off),
 * Euler / Runge_Kutta (Composition_adv.c:108,61).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include "global_defs.h"
#include "prototypes.h"
#include "citcomcu_b200.h"

extern ccu_ctx *g_ctx;
extern int g_ccu_device_geometry, g_ccu_cartesian;
void ccu_dropin_init(struct All_variables *E);
void ccu_dropin_die(const char *msg);
#define CCU(call) do { if((call) != 0) { fprintf(stderr, "citcomcu_b200 drop-in: %s failed: %s\n", #call, ccu_last_error()); exit(9); } } while(0)

static int bound(const char *name)
{
    const char *list = getenv("CCU_DROPIN_FUNCS");
    size_t n = strlen(name);
    const char *p;
    if(!list || !*list) return 0;
    if(strcmp(list, "all") == 0) return 1;
    for(p = list; (p = strstr(p, name)) != NULL; p += n)
        if((p == list || p[-1] == ',') && (p[n] == 0 || p[n] == ',')) return 1;
    return 0;
}
#define NEXT(var, name) do { if(!(var)) { *(void **)(&(var)) = dlsym(RTLD_NEXT, name); if(!(var)) ccu_dropin_die("no reference definition of " name " to forward to"); } } while(0)

/* ---- the reference's operator arrays on the device ---- */
static int g_op_stale = 1;
static void sync_operator(struct All_variables *E)
{
    int lev, e, a;
    if(!g_ctx) ccu_dropin_init(E);
    if(!g_op_stale) return;
    for(lev = E->mesh.levmin; lev <= E->mesh.levmax; lev++)
    {
        const int nel = E->lmesh.NEL[lev];
        float *g = (float *)malloc((size_t)nel * 24 * sizeof(float));
        float *tw = (float *)malloc((size_t)nel * 8 * sizeof(float));
        float *sz = (float *)malloc((size_t)nel * 3 * sizeof(float));
        if(!E->Eqn_k1[lev]) ccu_dropin_die("function-level bindings need the reference's node-assembled operator (node_assemble=1)");
        for(e = 1; e <= nel; e++)
        {   /* struct strides: FNODE and SIZE carry an unused entry 0 */
            for(a = 0; a < 24; a++) g[(size_t)(e - 1) * 24 + a] = E->elt_del[lev][e].g[a][0];
            for(a = 1; a <= 8; a++) tw[(size_t)(e - 1) * 8 + a - 1] = E->TWW[lev][e].node[a];
            for(a = 1; a <= 3; a++) sz[(size_t)(e - 1) * 3 + a - 1] = E->ECO[lev][e].size[a];
        }
        CCU(ccu_set_stiffness(g_ctx, lev, E->Eqn_k1[lev], E->Eqn_k2[lev], E->Eqn_k3[lev], E->BI[lev]));
        CCU(ccu_set_pressure_ops(g_ctx, lev, g, E->BPI[lev] + 1));
        CCU(ccu_set_transfer_weights(g_ctx, lev, tw, E->MASS[lev] + 1, sz));
        free(g); free(tw); free(sz);
    }
    g_op_stale = 0;
}

void construct_stiffness_B_matrix(struct All_variables *E)
{   /* always the reference's own (the whole-step binding builds on the device and never comes here) */
    static void (*next)(struct All_variables *) = NULL;
    NEXT(next, "construct_stiffness_B_matrix");
    next(E);
    g_op_stale = 1;
}

void n_assemble_del2_u(struct All_variables *E, double *u, double *Au, int level, int strip_bcs)
{   /* Element_calculations.c:552 */
    static void (*next)(struct All_variables *, double *, double *, int, int) = NULL;
    if(!bound("n_assemble_del2_u")) { NEXT(next, "n_assemble_del2_u"); next(E, u, Au, level, strip_bcs); return; }
    sync_operator(E);
    CCU(ccu_n_assemble_del2_u(g_ctx, level, u, Au, strip_bcs));
}
void e_assemble_del2_u(struct All_variables *E, double *u, double *Au, int level, int strip_bcs)
{   /* Element_calculations.c:494: the same product from the node-stored operator */
    static void (*next)(struct All_variables *, double *, double *, int, int) = NULL;
    if(!bound("e_assemble_del2_u")) { NEXT(next, "e_assemble_del2_u"); next(E, u, Au, level, strip_bcs); return; }
    sync_operator(E);
    CCU(ccu_e_assemble_del2_u(g_ctx, level, u, Au, strip_bcs));
}
void assemble_del2_u(struct All_variables *E, double *u, double *Au, int level, int strip_bcs)
{   /* Element_calculations.c:480: dispatch */
    static void (*next)(struct All_variables *, double *, double *, int, int) = NULL;
    if(!bound("assemble_del2_u")) { NEXT(next, "assemble_del2_u"); next(E, u, Au, level, strip_bcs); return; }
    sync_operator(E);
    CCU(ccu_n_assemble_del2_u(g_ctx, level, u, Au, strip_bcs));
}
void gauss_seidel(struct All_variables *E, double *d0, double *F, double *Ad, double acc, int *cycles, int level, int guess)
{   /* General_matrix_functions.c:1160: `cycles` sweeps (8-colour on the device), d0 and Ad = K d0 out */
    static void (*next)(struct All_variables *, double *, double *, double *, double, int *, int, int) = NULL;
    if(!bound("gauss_seidel")) { NEXT(next, "gauss_seidel"); next(E, d0, F, Ad, acc, cycles, level, guess); return; }
    sync_operator(E);
    CCU(ccu_gauss_seidel(g_ctx, level, d0, F, Ad, *cycles, guess));
}
double multi_grid(struct All_variables *E, double *d1, double *F, double *Au, double acc, int hl)
{   /* General_matrix_functions.c:525: one full-multigrid cycle from the finest level; d1 = correction, F = residual in / out.
       Au is an output of the reference's (K d1), which the residual update already contains: recomputed for the caller. */
    static double (*next)(struct All_variables *, double *, double *, double *, double, int) = NULL;
    double res = 0.0;
    if(!bound("multi_grid")) { NEXT(next, "multi_grid"); return next(E, d1, F, Au, acc, hl); }
    if(hl != E->mesh.levmax) ccu_dropin_die("multi_grid: the device cycle starts from the finest level");
    sync_operator(E);
    CCU(ccu_multi_grid(g_ctx, d1, F, &res));
    if(Au) CCU(ccu_n_assemble_del2_u(g_ctx, hl, d1, Au, 1));
    return res;
}
int solve_del2_u(struct All_variables *E, double *d0, double *F, double acc, int high_lev)
{   /* General_matrix_functions.c:368 */
    static int (*next)(struct All_variables *, double *, double *, double, int) = NULL;
    int valid = 0, cycles = 0;
    if(!bound("solve_del2_u")) { NEXT(next, "solve_del2_u"); return next(E, d0, F, acc, high_lev); }
    if(high_lev != E->mesh.levmax) ccu_dropin_die("solve_del2_u: the device solve runs on the finest level");
    sync_operator(E);
    CCU(ccu_solve_del2_u(g_ctx, d0, F, acc, &valid, &cycles));
    return valid;
}
double conj_grad(struct All_variables *E, double *d0, double *F, double *Au, double acc, int *cycles, int level)
{   /* General_matrix_functions.c:661 */
    static double (*next)(struct All_variables *, double *, double *, double *, double, int *, int) = NULL;
    double res = 0.0;
    if(!bound("conj_grad")) { NEXT(next, "conj_grad"); return next(E, d0, F, Au, acc, cycles, level); }
    sync_operator(E);
    CCU(ccu_conj_grad(g_ctx, level, d0, F, acc, cycles, &res));
    if(Au) CCU(ccu_n_assemble_del2_u(g_ctx, level, d0, Au, 1));
    return res;
}
void strip_bcs_from_residual(struct All_variables *E, double *Res, int level)
{   /* Boundary_conditions.c:926 */
    static void (*next)(struct All_variables *, double *, int) = NULL;
    if(!bound("strip_bcs_from_residual")) { NEXT(next, "strip_bcs_from_residual"); next(E, Res, level); return; }
    if(!g_ctx) ccu_dropin_init(E);
    CCU(ccu_strip_bcs_from_residual(g_ctx, level, Res));
}
void project_vector(struct All_variables *E, int start_lev, double *AU, double *AD, int ic)
{   /* Solver_multigrid.c:72; ic = strip the boundary rows of the result (:155) */
    static void (*next)(struct All_variables *, int, double *, double *, int) = NULL;
    if(!bound("project_vector")) { NEXT(next, "project_vector"); next(E, start_lev, AU, AD, ic); return; }
    sync_operator(E);
    CCU(ccu_project_vector(g_ctx, start_lev, AU, AD));
    if(ic) CCU(ccu_strip_bcs_from_residual(g_ctx, start_lev - 1, AD));
}
void interp_vector(struct All_variables *E, int start_lev, double *AD, double *AU)
{   /* Solver_multigrid.c:173 */
    static void (*next)(struct All_variables *, int, double *, double *) = NULL;
    if(!bound("interp_vector")) { NEXT(next, "interp_vector"); next(E, start_lev, AD, AU); return; }
    sync_operator(E);
    CCU(ccu_interp_vector(g_ctx, start_lev, AD, AU));
}
void assemble_div_u(struct All_variables *E, double *U, double *divU, int level)
{   /* Element_calculations.c:691; divU is 1-based */
    static void (*next)(struct All_variables *, double *, double *, int) = NULL;
    if(!bound("assemble_div_u")) { NEXT(next, "assemble_div_u"); next(E, U, divU, level); return; }
    sync_operator(E);
    CCU(ccu_assemble_div_u(g_ctx, level, U, divU + 1));
}
void assemble_grad_p(struct All_variables *E, double *P, double *gradP, int lev)
{   /* Element_calculations.c:727; P is 1-based */
    static void (*next)(struct All_variables *, double *, double *, int) = NULL;
    if(!bound("assemble_grad_p")) { NEXT(next, "assemble_grad_p"); next(E, P, gradP, lev); return; }
    sync_operator(E);
    CCU(ccu_assemble_grad_p(g_ctx, lev, P + 1, gradP));
}
double global_vdot(struct All_variables *E, double *A, double *B, int lev)
{   /* Global_operations.c:339 */
    static double (*next)(struct All_variables *, double *, double *, int) = NULL;
    double r = 0.0;
    if(!bound("global_vdot")) { NEXT(next, "global_vdot"); return next(E, A, B, lev); }
    if(!g_ctx) ccu_dropin_init(E);
    CCU(ccu_global_vdot(g_ctx, lev, A, B, &r));
    return r;
}
double global_pdot(struct All_variables *E, double *A, double *B, int lev)
{   /* Global_operations.c:359; A, B 1-based */
    static double (*next)(struct All_variables *, double *, double *, int) = NULL;
    double r = 0.0;
    if(!bound("global_pdot")) { NEXT(next, "global_pdot"); return next(E, A, B, lev); }
    if(!g_ctx) ccu_dropin_init(E);
    CCU(ccu_global_pdot(g_ctx, lev, A + 1, B + 1, &r));
    return r;
}
float solve_Ahat_p_fhat(struct All_variables *E, double *V, double *P, double *F, double imp, int *steps_max)
{   /* Stokes_flow_Incomp.c:295: the Uzawa pressure iteration with the multigrid velocity solves, from V, P in */
    static float (*next)(struct All_variables *, double *, double *, double *, double, int *) = NULL;
    float res = 0.0f;
    if(!bound("solve_Ahat_p_fhat")) { NEXT(next, "solve_Ahat_p_fhat"); return next(E, V, P, F, imp, steps_max); }
    sync_operator(E);
    CCU(ccu_solve_Ahat_p_fhat(g_ctx, V, P + 1, F, imp, steps_max, &res, NULL));
    return res;
}

/* ---- markers (Composition_adv.c): the marker state lives in E; every call uploads it, steps, and brings it back ---- */
static int g_markers = 0;
static void markers_to_device(struct All_variables *E)
{
    if(!g_ctx) ccu_dropin_init(E);
    if(!g_ccu_device_geometry) ccu_dropin_die("markers: the context was created without device geometry (CCU_DROPIN_STOKES=0 on a regional-spherical run)");
    if(!g_markers)
    {
        CCU(ccu_markers_setup(g_ctx, E->advection.markers_uplimit, E->advection.markers_per_ele, E->lmesh.rnoz, E->XP[1] + 1, E->XP[2] + 1,
                              E->XP[3] + 1, E->RG[3], E->XG1 + 1, E->XG2 + 1, E->Element + 1, E->control.Acomp));
        g_markers = 1;
    }
    CCU(ccu_markers_upload(g_ctx, E->advection.markers, E->XMC[1] + 1, E->XMC[2] + 1, E->XMC[3] + 1, E->C12 + 1, E->CElement + 1, E->CE + 1));
    CCU(ccu_set_velocity(g_ctx, E->V[1] + 1, E->V[2] + 1, E->V[3] + 1));
}
static void markers_from_device(struct All_variables *E, int corrector)
{
    const int n = ccu_markers_count(g_ctx);
    double *X, *Xp; float *VO, *Vp;
    int d, i;
    if(n < 0 || n > E->advection.markers_uplimit) ccu_dropin_die("markers: count out of range after the step");
    X = (double *)malloc(sizeof(double) * 3 * (size_t)(n + 1)); Xp = (double *)malloc(sizeof(double) * 3 * (size_t)(n + 1));
    VO = (float *)malloc(sizeof(float) * 3 * (size_t)(n + 1)); Vp = (float *)malloc(sizeof(float) * 3 * (size_t)(n + 1));
    CCU(ccu_markers_download(g_ctx, X, Xp, VO, Vp, E->CElement + 1, E->C + 1, E->CE + 1));
    for(d = 0; d < 3; d++)
        for(i = 0; i < n; i++)
        {
            E->XMC[d + 1][i + 1] = X[(size_t)d * n + i]; E->XMCpred[d + 1][i + 1] = Xp[(size_t)d * n + i];
            E->VO[d + 1][i + 1] = VO[(size_t)d * n + i]; E->Vpred[d + 1][i + 1] = Vp[(size_t)d * n + i];
        }
    E->advection.markers = n;
    free(X); free(Xp); free(VO); free(Vp);
    (void)corrector;
}
void Euler(struct All_variables *E, float *C, float *V[4], int on_off)
{   /* Composition_adv.c:108 */
    static void (*next)(struct All_variables *, float *, float *[4], int) = NULL;
    if(!bound("Euler")) { NEXT(next, "Euler"); next(E, C, V, on_off); return; }
    markers_to_device(E);
    CCU(ccu_Euler(g_ctx, E->advection.timestep));
    markers_from_device(E, 0);
}
void Runge_Kutta(struct All_variables *E, float *C, float *V[4], int on_off)
{   /* Composition_adv.c:61 */
    static void (*next)(struct All_variables *, float *, float *[4], int) = NULL;
    if(!bound("Runge_Kutta")) { NEXT(next, "Runge_Kutta"); next(E, C, V, on_off); return; }
    markers_to_device(E);
    CCU(ccu_Runge_Kutta(g_ctx, E->advection.timestep));
    markers_from_device(E, 1);
}

void PG_timestep_particle(struct All_variables *E)
{   /* Advection_diffusion.c:128-240: alternates (static on_off) between std_timestep + thermal step + Euler predictor of the
       markers, and the Runge_Kutta corrector with the new velocity; thermal_buoyancy after either */
    static void (*next)(struct All_variables *) = NULL;
    static int on_off = 0, energy = 0;
    float dt = 0.0f, Tint = 0.0f;
    int n;
    if(!bound("PG_timestep_particle")) { NEXT(next, "PG_timestep_particle"); next(E); return; }
    if(!g_ctx) ccu_dropin_init(E);
    if(!g_ccu_device_geometry) ccu_dropin_die("PG_timestep_particle: the context was created without device geometry (CCU_DROPIN_STOKES=0 on a regional-spherical run)");
    if(on_off == 0)
    {
        if(E->control.composition != 2)
        {
            if(!energy)
            {
                if(!E->advection.ADVECTION) ccu_dropin_die("ADVECTION=off is not on the device path");
                for(n = 1; n <= E->lmesh.nno; n++)
                    if(E->node[n] & FBZ) ccu_dropin_die("heat-flux boundary conditions are not on the device path");
                CCU(ccu_set_energy_params(g_ctx, E->advection.fine_tune_dt, E->advection.fixed_timestep, E->advection.gamma,
                                          E->advection.temp_iterations, E->diffusivity + 1, E->expansivity + 1, E->control.Q0));
                energy = 1;
            }
            E->advection.timesteps++;
            if(E->control.adi_heating || E->control.visc_heating || E->control.Ra_410 != 0.0 || E->control.Ra_670 != 0.0)
                CCU(ccu_set_heating_arrays(g_ctx, E->heating_adi + 1, E->heating_visc + 1, E->heating_latent + 1));
            CCU(ccu_set_velocity(g_ctx, E->V[1] + 1, E->V[2] + 1, E->V[3] + 1));
            CCU(ccu_PG_timestep(g_ctx, E->T + 1, E->Tdot + 1, &dt, &Tint));      /* std_timestep + predictor/corrector + Tmax safeguard */
            E->advection.timestep = dt;
            E->monitor.T_interior = Tint;
            E->advection.dt_reduced = 1.0;
            E->advection.total_timesteps++;
            temperatures_conform_bcs(E);
            E->advection.last_sub_iterations = 1;
        }
        else
        {   /* purely compositional: only the timestep */
            E->advection.timesteps++;
            std_timestep(E);
            E->advection.total_timesteps++;
            E->advection.last_sub_iterations = 0;
        }
        Euler(E, E->C, E->V, on_off);
        E->monitor.elapsed_time += E->advection.timestep;
    }
    else
        Runge_Kutta(E, E->C, E->V, on_off);
    thermal_buoyancy(E);
    E->control.keep_going = (E->monitor.solution_cycles < E->advection.max_timesteps) ? 1 : 0;
    on_off = on_off ? 0 : 1;
}
