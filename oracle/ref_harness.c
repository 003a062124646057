/* ref_harness.c -- driver around the UNMODIFIED reference (TEST INFRASTRUCTURE).
 *
 * Links against oracle/_ref/libcitcom_ref.so (the reference's own objects) and
 * calls the reference's own functions through their public prototypes
 * (src/prototypes.h).  It replays main()'s sequence (Citcom.c:54-175) and, at
 * chosen points, writes raw binary dumps of the arrays the B200 path mirrors
 * plus known-answer vectors produced by the reference's own hot-path functions
 * on seeded inputs.  With LD_PRELOAD=libcitcomcu_dropin.so the very same binary
 * exercises the drop-in (symbol interposition), so both arms share one harness.
 *
 *   ref_harness dump <input> <outdir> <nsteps> [kat]
 *   ref_harness time <input> <nsteps>
 *   ref_harness timezero <input> <repeats>      (step-0 Stokes solve repeated from a zero guess)
 *
 * Dump format: <outdir>/<name>.bin raw little-endian + manifest.txt lines
 * "name dtype count".
 */
#include <mpi.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "element_definitions.h"
#include "global_defs.h"

extern int Emergency_stop;

static char g_outdir[1000];
static FILE *g_manifest = NULL;
static int g_rank = 0;

static void dump(const char *name, const void *ptr, size_t elsize, size_t count, const char *dtype)
{
    char path[1400];
    snprintf(path, sizeof path, "%s/%s.r%d.bin", g_outdir, name, g_rank);
    FILE *f = fopen(path, "wb");
    if(!f) { perror(path); exit(3); }
    if(count) fwrite(ptr, elsize, count, f);
    fclose(f);
    fprintf(g_manifest, "%s %s %zu\n", name, dtype, count);
    fflush(g_manifest);
}
#define DUMP_F64(n,p,c) dump(n,p,8,c,"f64")
#define DUMP_F32(n,p,c) dump(n,p,4,c,"f32")
#define DUMP_I32(n,p,c) dump(n,p,4,c,"i32")
#define DUMP_U32(n,p,c) dump(n,p,4,c,"u32")

static void dump_scalar_d(const char *name, double v) { DUMP_F64(name, &v, 1); }
static void dump_scalar_i(const char *name, int v) { DUMP_I32(name, &v, 1); }

/* xorshift64* : documented so the Python side could regenerate, though inputs are dumped too */
static unsigned long long g_rng = 88172645463325252ULL;
static double rnd(void)
{
    g_rng ^= g_rng >> 12; g_rng ^= g_rng << 25; g_rng ^= g_rng >> 27;
    unsigned long long r = g_rng * 2685821657736338717ULL;
    return ((double)(r >> 11) / 9007199254740992.0) * 2.0 - 1.0;
}

static void dump_level_arrays(struct All_variables *E)
{
    char nm[200];
    int lev, e, a;
    for(lev = E->mesh.levmin; lev <= E->mesh.levmax; lev++)
    {
        const int nno = E->lmesh.NNO[lev], nel = E->lmesh.NEL[lev], neq = E->lmesh.NEQ[lev], npno = E->lmesh.NPNO[lev];
        int dims[10] = { E->lmesh.NOX[lev], E->lmesh.NOY[lev], E->lmesh.NOZ[lev], E->lmesh.ELX[lev], E->lmesh.ELY[lev],
                         E->lmesh.ELZ[lev], nno, nel, neq, npno };
        snprintf(nm, sizeof nm, "L%d_dims", lev); DUMP_I32(nm, dims, 10);
        snprintf(nm, sizeof nm, "L%d_NODE", lev); DUMP_U32(nm, E->NODE[lev] + 1, nno);
        if(lev == E->mesh.levmax)
        {   /* imposed boundary velocities (finest level only, global_defs.h:1038) */
            DUMP_F32("VB1", E->VB[1] + 1, nno); DUMP_F32("VB2", E->VB[2] + 1, nno); DUMP_F32("VB3", E->VB[3] + 1, nno);
        }
        if(!E->Eqn_k1[lev] || !E->Node_map[lev]) continue;   /* operator lives on the device (drop-in preloaded) */
        snprintf(nm, sizeof nm, "L%d_Eqn_k1", lev); DUMP_F32(nm, E->Eqn_k1[lev], (size_t)nno * 42);
        snprintf(nm, sizeof nm, "L%d_Eqn_k2", lev); DUMP_F32(nm, E->Eqn_k2[lev], (size_t)nno * 42);
        snprintf(nm, sizeof nm, "L%d_Eqn_k3", lev); DUMP_F32(nm, E->Eqn_k3[lev], (size_t)nno * 42);
        snprintf(nm, sizeof nm, "L%d_Node_map", lev); DUMP_I32(nm, E->Node_map[lev], (size_t)nno * 42);
        snprintf(nm, sizeof nm, "L%d_BI", lev); DUMP_F64(nm, E->BI[lev], neq);
        snprintf(nm, sizeof nm, "L%d_BPI", lev); DUMP_F64(nm, E->BPI[lev] + 1, npno);
        snprintf(nm, sizeof nm, "L%d_IDD", lev); DUMP_I32(nm, E->parallel.IDD[lev], neq);
        snprintf(nm, sizeof nm, "L%d_MASS", lev); DUMP_F32(nm, E->MASS[lev] + 1, nno);
        snprintf(nm, sizeof nm, "L%d_EVI", lev); DUMP_F32(nm, E->EVI[lev] + 1, (size_t)nel * 8);
        {
            float *g = (float *)malloc((size_t)nel * 24 * sizeof(float));
            float *tw = (float *)malloc((size_t)nel * 8 * sizeof(float));
            float *sz = (float *)malloc((size_t)nel * 3 * sizeof(float));
            for(e = 1; e <= nel; e++)
            {
                for(a = 0; a < 24; a++) g[(size_t)(e - 1) * 24 + a] = E->elt_del[lev][e].g[a][0];
                for(a = 1; a <= 8; a++) tw[(size_t)(e - 1) * 8 + a - 1] = E->TWW[lev][e].node[a];
                for(a = 1; a <= 3; a++) sz[(size_t)(e - 1) * 3 + a - 1] = E->ECO[lev][e].size[a];
            }
            snprintf(nm, sizeof nm, "L%d_elt_del", lev); DUMP_F32(nm, g, (size_t)nel * 24);
            snprintf(nm, sizeof nm, "L%d_TWW", lev); DUMP_F32(nm, tw, (size_t)nel * 8);
            snprintf(nm, sizeof nm, "L%d_eco_size", lev); DUMP_F32(nm, sz, (size_t)nel * 3);
            free(g); free(tw); free(sz);
        }
        snprintf(nm, sizeof nm, "L%d_GNX", lev); DUMP_F32(nm, &E->GNX[lev][1], (size_t)nel * (sizeof(struct Shape_function_dx) / 4));
        snprintf(nm, sizeof nm, "L%d_GDA", lev); DUMP_F32(nm, &E->GDA[lev][1], (size_t)nel * (sizeof(struct Shape_function_dA) / 4));
        for(a = 1; a <= 3; a++)
        {
            snprintf(nm, sizeof nm, "L%d_XX%d", lev, a); DUMP_F32(nm, E->XX[lev][a] + 1, nno);
            if(E->control.Rsphere) { snprintf(nm, sizeof nm, "L%d_SXX%d", lev, a); DUMP_F32(nm, E->SXX[lev][a] + 1, nno); }
        }
    }
}

static void dump_fields(struct All_variables *E, const char *tag)
{
    char nm[200], nm_[200];
    const int nno = E->lmesh.nno, neq = E->lmesh.neq, npno = E->lmesh.npno, nel = E->lmesh.nel;
    int d;
    snprintf(nm, sizeof nm, "%s_mat", tag); DUMP_I32(nm, E->mat + 1, nel);
    snprintf(nm, sizeof nm, "%s_F", tag); DUMP_F64(nm, E->F, neq);
    snprintf(nm, sizeof nm, "%s_U", tag); DUMP_F64(nm, E->U, neq);
    snprintf(nm, sizeof nm, "%s_P", tag); DUMP_F64(nm, E->P + 1, npno);
    snprintf(nm, sizeof nm, "%s_T", tag); DUMP_F32(nm, E->T + 1, nno);
    snprintf(nm, sizeof nm, "%s_Tdot", tag); DUMP_F32(nm, E->Tdot + 1, nno);
    snprintf(nm, sizeof nm, "%s_buoyancy", tag); DUMP_F32(nm, E->buoyancy + 1, nno);
    snprintf(nm, sizeof nm, "%s_EVI", tag); DUMP_F32(nm, E->EVI[E->mesh.levmax] + 1, (size_t)nel * 8);
    for(d = 1; d <= 3; d++)
    {
        snprintf(nm, sizeof nm, "%s_V%d", tag, d); DUMP_F32(nm, E->V[d] + 1, nno);
    }
    if(E->control.composition && E->advection.markers > 0)
    {   /* marker state (Composition_adv.c): positions, dense/regular flag, element assignment, nodal and elemental C */
        const int nm = E->advection.markers;
        int nmv = nm;
        snprintf(nm_, sizeof nm_, "%s_nmarkers", tag); DUMP_I32(nm_, &nmv, 1);
        for(d = 1; d <= 3; d++)
        {
            snprintf(nm_, sizeof nm_, "%s_XMC%d", tag, d); DUMP_F64(nm_, E->XMC[d] + 1, nm);
            snprintf(nm_, sizeof nm_, "%s_XMCpred%d", tag, d); DUMP_F64(nm_, E->XMCpred[d] + 1, nm);
            snprintf(nm_, sizeof nm_, "%s_VO%d", tag, d); DUMP_F32(nm_, E->VO[d] + 1, nm);
        }
        snprintf(nm_, sizeof nm_, "%s_C12", tag); DUMP_I32(nm_, E->C12 + 1, nm);
        snprintf(nm_, sizeof nm_, "%s_CElement", tag); DUMP_I32(nm_, E->CElement + 1, nm);
        snprintf(nm_, sizeof nm_, "%s_C", tag); DUMP_F32(nm_, E->C + 1, nno);
        snprintf(nm_, sizeof nm_, "%s_CE", tag); DUMP_F32(nm_, E->CE + 1, nel);
    }
    if(E->control.adi_heating || E->control.visc_heating)
    {   /* extended-Boussinesq heating terms as process_heating (Advection_diffusion.c:813) left them for the step just taken */
        double eb[4] = { E->data.disptn_number, E->data.surf_temp, E->control.Atemp, E->control.Q0 };
        snprintf(nm, sizeof nm, "%s_heating_adi", tag); DUMP_F32(nm, E->heating_adi + 1, nel);
        snprintf(nm, sizeof nm, "%s_heating_visc", tag); DUMP_F32(nm, E->heating_visc + 1, nel);
        snprintf(nm, sizeof nm, "%s_heating_latent", tag); DUMP_F32(nm, E->heating_latent + 1, nel);
        snprintf(nm, sizeof nm, "%s_eba", tag); DUMP_F64(nm, eb, 4);
    }
    if(E->control.Ra_670 != 0.0 || E->control.Ra_410 != 0.0)
    {   /* phase functions of the last phase_change call (Phase_change.c:43) and its (rescaled) parameters */
        double ph[10] = { E->viscosity.zlm, E->viscosity.z410, E->control.Ra_670, E->control.clapeyron670, E->control.width670,
                          E->control.transT670, E->control.Ra_410, E->control.clapeyron410, E->control.width410, E->control.transT410 };
        snprintf(nm, sizeof nm, "%s_Fas670", tag); DUMP_F32(nm, E->Fas670 + 1, nno);
        snprintf(nm, sizeof nm, "%s_Fas410", tag); DUMP_F32(nm, E->Fas410 + 1, nno);
        snprintf(nm, sizeof nm, "%s_phase", tag); DUMP_F64(nm, ph, 10);
    }
    {
        double sc[8] = { E->monitor.elapsed_time, E->advection.timestep, E->slice.Nut, E->slice.Nub,
                         E->monitor.T_interior, (double)E->monitor.solution_cycles, E->monitor.vdotv, E->monitor.pdotp };
        snprintf(nm, sizeof nm, "%s_scalars", tag); DUMP_F64(nm, sc, 8);
    }
    /* layer averages of the reference's averages() (Process_velocity.c:179; current when storage_spacing divides the step) */
    snprintf(nm, sizeof nm, "%s_Have_vrms", tag); DUMP_F32(nm, E->Have.vrms + 1, E->lmesh.noz);
    snprintf(nm, sizeof nm, "%s_Have_Vi", tag); DUMP_F32(nm, E->Have.Vi + 1, E->lmesh.noz);
    snprintf(nm, sizeof nm, "%s_XP3", tag); DUMP_F64(nm, E->XP[3] + 1, E->lmesh.noz);
}

static void strip(struct All_variables *E, double *v, int lev) { strip_bcs_from_residual(E, v, lev); }

/* Known-answer vectors: the reference's own hot-path functions on seeded inputs. */
static void dump_kats(struct All_variables *E)
{
    char nm[200];
    int lev, i;
    const int levmax = E->mesh.levmax, levmin = E->mesh.levmin;
    const int neqmax = E->lmesh.NEQ[levmax];
    double *u = (double *)calloc(neqmax + 10, 8), *Au = (double *)calloc(neqmax + 10, 8);
    double *f = (double *)calloc(neqmax + 10, 8), *d0 = (double *)calloc(neqmax + 10, 8);
    double *w = (double *)calloc(neqmax + 10, 8);
    double *p = (double *)calloc(E->lmesh.npno + 10, 8), *q = (double *)calloc(E->lmesh.npno + 10, 8);
    {   /* get_stress / get_STD_topo (Topo_gravity.c:352,307) on the current state (V, P, EVI of the last solve) */
        const int nno = E->lmesh.nno, nsf = E->lmesh.nsf;
        float *S = (float *)calloc((size_t)6 * (nno + 1), sizeof(float));
        float *tp = (float *)calloc((size_t)2 * (nsf + 2), sizeof(float));
        get_stress(S, S + (nno + 1), S + 2 * (nno + 1), S + 3 * (nno + 1), S + 4 * (nno + 1), S + 5 * (nno + 1), E);
        for(i = 0; i < 6; i++) { snprintf(nm, sizeof nm, "kat_stress%d", i); DUMP_F32(nm, S + (size_t)i * (nno + 1) + 1, nno); }
        get_STD_topo(E, tp, tp + nsf + 2, 0);
        DUMP_F32("kat_tpg", tp + 1, nsf); DUMP_F32("kat_tpgb", tp + nsf + 2 + 1, nsf);
        free(S); free(tp);
    }

    for(lev = levmin; lev <= levmax; lev++)
    {
        const int neq = E->lmesh.NEQ[lev];
        int cycles;
        for(i = 0; i < neq; i++) { u[i] = rnd(); f[i] = rnd(); }
        u[neq] = u[neq + 1] = f[neq] = f[neq + 1] = 0.0;
        strip(E, u, lev); strip(E, f, lev);
        snprintf(nm, sizeof nm, "kat_L%d_u", lev); DUMP_F64(nm, u, neq);
        snprintf(nm, sizeof nm, "kat_L%d_f", lev); DUMP_F64(nm, f, neq);

        n_assemble_del2_u(E, u, Au, lev, 1);
        snprintf(nm, sizeof nm, "kat_L%d_Au", lev); DUMP_F64(nm, Au, neq);

        cycles = 2;
        gauss_seidel(E, d0, f, Au, 0.01, &cycles, lev, 0);
        snprintf(nm, sizeof nm, "kat_L%d_gs0_d", lev); DUMP_F64(nm, d0, neq);
        snprintf(nm, sizeof nm, "kat_L%d_gs0_Ad", lev); DUMP_F64(nm, Au, neq);

        for(i = 0; i < neq; i++) d0[i] = u[i];
        cycles = 3;
        gauss_seidel(E, d0, f, Au, 0.01, &cycles, lev, 1);
        snprintf(nm, sizeof nm, "kat_L%d_gs1_d", lev); DUMP_F64(nm, d0, neq);
        snprintf(nm, sizeof nm, "kat_L%d_gs1_Ad", lev); DUMP_F64(nm, Au, neq);

        if(lev > levmin)
        {
            project_vector(E, lev, u, w, 1);
            snprintf(nm, sizeof nm, "kat_L%d_proj", lev); DUMP_F64(nm, w, E->lmesh.NEQ[lev - 1]);
        }
        if(lev < levmax)
        {
            interp_vector(E, lev, u, w);
            snprintf(nm, sizeof nm, "kat_L%d_interp", lev); DUMP_F64(nm, w, E->lmesh.NEQ[lev + 1]);
        }
        snprintf(nm, sizeof nm, "kat_L%d_vdot", lev); dump_scalar_d(nm, global_vdot(E, u, f, lev));
    }

    /* finest level: div / grad / pdot */
    {
        const int neq = E->lmesh.NEQ[levmax], npno = E->lmesh.npno;
        for(i = 0; i < neq; i++) u[i] = rnd();
        strip(E, u, levmax);
        for(i = 1; i <= npno; i++) p[i] = rnd();
        DUMP_F64("kat_div_u", u, neq);
        assemble_div_u(E, u, q, levmax);
        DUMP_F64("kat_div_out", q + 1, npno);
        DUMP_F64("kat_grad_p", p + 1, npno);
        assemble_grad_p(E, p, Au, levmax);
        DUMP_F64("kat_grad_out", Au, neq);
        dump_scalar_d("kat_pdot", global_pdot(E, p, q, levmax));
    }

    /* one multigrid cycle and a tight velocity solve on a seeded right-hand side */
    {
        const int neq = E->lmesh.NEQ[levmax];
        float acc_save = E->control.accuracy;
        double res;
        int valid;
        for(i = 0; i < neq; i++) f[i] = rnd();
        strip(E, f, levmax);
        DUMP_F64("kat_solve_f", f, neq);
        for(i = 0; i < neq; i++) { w[i] = f[i]; d0[i] = 0.0; Au[i] = 0.0; }
        res = multi_grid(E, d0, w, Au, 1e-30, levmax);
        DUMP_F64("kat_mg_d1", d0, neq);
        DUMP_F64("kat_mg_res", w, neq);
        dump_scalar_d("kat_mg_residual", res);
        E->control.accuracy = 1.0e-11;
        for(i = 0; i < neq; i++) w[i] = f[i];
        valid = solve_del2_u(E, d0, w, 1e-30, levmax);
        E->control.accuracy = acc_save;
        DUMP_F64("kat_solve_d0", d0, neq);
        dump_scalar_i("kat_solve_valid", valid);
        {   /* conj_grad (General_matrix_functions.c:661): 25 iterations of the Jacobi-preconditioned CG on the same rhs */
            int cyc = 25;
            double r;
            for(i = 0; i < neq; i++) { w[i] = f[i]; d0[i] = 0.0; }
            r = conj_grad(E, d0, w, Au, 1e-30, &cyc, levmax);
            DUMP_F64("kat_cg_d0", d0, neq);
            dump_scalar_d("kat_cg_residual", r);
            dump_scalar_i("kat_cg_cycles", cyc);
        }
    }
    free(u); free(Au); free(f); free(d0); free(w); free(p); free(q);
    {   /* energy step known answers on the state after the step-0 Stokes solve: std_timestep (Advection_diffusion.c:737)
         * and one pg_solver call (:398) exactly as PG_timestep invokes it (:303) */
        const int nno = E->lmesh.nno;
        float *DTdot = (float *)calloc(nno + 2, sizeof(float));
        double adv[6];
        std_timestep(E);
        pg_solver(E, E->T, E->Tdot, DTdot, E->V, E->convection.heat_sources, 1.0, 1, E->TB, E->node);
        adv[0] = E->advection.fine_tune_dt; adv[1] = E->advection.fixed_timestep; adv[2] = E->advection.gamma;
        adv[3] = E->advection.temp_iterations; adv[4] = E->control.Q0; adv[5] = E->control.Atemp;
        DUMP_F64("kat_adv_params", adv, 6);
        dump_scalar_d("kat_dt", (double)E->advection.timestep);
        DUMP_F32("kat_pg_DTdot", DTdot + 1, nno);
        DUMP_F32("kat_diffusivity", E->diffusivity + 1, E->lmesh.noz);
        DUMP_F32("kat_expansivity", E->expansivity + 1, E->lmesh.noz);
        DUMP_U32("kat_node", E->node + 1, nno);
        free(DTdot);
    }
}

static void setup(struct All_variables *E, int *argc, char ***argv, char *input)
{
    static char *fake[3];
    srand48((long int)-1);
    E->parallel.me = 0;
    E->parallel.nproc = 1;
    E->parallel.me_loc[1] = E->parallel.me_loc[2] = E->parallel.me_loc[3] = 0;
    MPI_Init(argc, argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &(E->parallel.me));
    MPI_Comm_size(MPI_COMM_WORLD, &(E->parallel.nproc));
    gethostname(E->parallel.machinename, 160);
    E->monitor.solution_cycles = 0;
    E->monitor.elapsed_time = 0;
    E->advection.timestep = 0;
    E->advection.timesteps = 0;
    fake[0] = (*argv)[0]; fake[1] = input; fake[2] = NULL;
    read_instructions(E, 2, fake);
    E->control.keep_going = 1;
    g_rank = E->parallel.me;
}

/* one pass of the body of main()'s while loop (Citcom.c:111-161) */
static void one_timestep(struct All_variables *E)
{
    if(E->control.composition != 2)
        process_heating(E);
    E->monitor.solution_cycles++;
    if(E->monitor.solution_cycles > E->control.print_convergence)
        E->control.print_convergence = 1;
    (E->next_buoyancy_field) (E);
    process_temp_field(E, E->monitor.solution_cycles);
    general_stokes_solver(E);
    if(E->control.composition)
        (E->next_buoyancy_field) (E);
    process_new_velocity(E, E->monitor.solution_cycles);
}

int main(int argc, char **argv)
{
    static struct All_variables E;
    int nsteps, step, want_kat = 0;
    char tag[64];
    double t0, t1;

    if(argc < 4) { fprintf(stderr, "usage: ref_harness dump <input> <outdir> <nsteps> [kat] | time <input> <nsteps>\n"); return 2; }

    if(strcmp(argv[1], "timezero") == 0)
    {   /* repeat the step-0 general_stokes_solver from a zero guess (bench.py's step, identical work each time) */
        int rep, i;
        nsteps = atoi(argv[3]);
        setup(&E, &argc, &argv, argv[2]);
        for(rep = 0; rep < nsteps; rep++)
        {
            for(i = 0; i < E.lmesh.neq + 2; i++) E.U[i] = 0.0;
            for(i = 0; i <= E.lmesh.npno; i++) E.P[i] = 0.0;
            MPI_Barrier(MPI_COMM_WORLD);
            t0 = MPI_Wtime();
            general_stokes_solver(&E);
            MPI_Barrier(MPI_COMM_WORLD);
            t1 = MPI_Wtime();
            if(E.parallel.me == 0) printf("CCU_TIME step %d stokes_s %.6f\n", rep, t1 - t0);
        }
        if(getenv("CCU_TZ_DUMP"))
        {   /* bench.py's parity gate: every rank writes its converged U, P and where it sits in the processor grid */
            char path[1400];
            int meta[9] = { E.parallel.nprocx, E.parallel.nprocy, E.parallel.nprocz, E.parallel.me_loc[1], E.parallel.me_loc[2],
                            E.parallel.me_loc[3], E.lmesh.nox, E.lmesh.noy, E.lmesh.noz };
            snprintf(g_outdir, sizeof g_outdir, "%s", getenv("CCU_TZ_DUMP"));
            snprintf(path, sizeof path, "%s/manifest.r%d.txt", g_outdir, g_rank);
            g_manifest = fopen(path, "w");
            if(!g_manifest) { perror(path); return 3; }
            DUMP_I32("tz_meta", meta, 9);
            DUMP_F64("tz_U", E.U, E.lmesh.neq);
            DUMP_F64("tz_P", E.P + 1, E.lmesh.npno);
            fclose(g_manifest);
        }
        fflush(stdout);
        MPI_Finalize();
        return 0;
    }
    if(strcmp(argv[1], "time") == 0)
    {
        nsteps = atoi(argv[3]);
        setup(&E, &argc, &argv, argv[2]);
        t0 = MPI_Wtime();
        general_stokes_solver(&E);
        t1 = MPI_Wtime();
        if(E.parallel.me == 0) printf("CCU_TIME step 0 stokes_s %.6f\n", t1 - t0);
        for(step = 1; step <= nsteps; step++)
        {
            double ta, tb, tc;
            if(E.control.composition != 2) process_heating(&E);
            E.monitor.solution_cycles++;
            ta = MPI_Wtime();
            (E.next_buoyancy_field) (&E);
            tb = MPI_Wtime();
            general_stokes_solver(&E);
            tc = MPI_Wtime();
            if(E.control.composition) (E.next_buoyancy_field) (&E);
            if(E.parallel.me == 0) printf("CCU_TIME step %d energy_s %.6f stokes_s %.6f\n", step, tb - ta, tc - tb);
        }
        fflush(stdout);
        MPI_Finalize();
        return 0;
    }

    nsteps = atoi(argv[4]);
    if(argc > 5 && strcmp(argv[5], "kat") == 0) want_kat = 1;
    snprintf(g_outdir, sizeof g_outdir, "%s", argv[3]);
    setup(&E, &argc, &argv, argv[2]);
    {
        char path[1400];
        snprintf(path, sizeof path, "%s/manifest.r%d.txt", g_outdir, g_rank);
        g_manifest = fopen(path, "w");
        if(!g_manifest) { perror(path); return 3; }
    }
    {
        int meta[16] = { E.mesh.levmin, E.mesh.levmax, E.parallel.nprocx, E.parallel.nprocy, E.parallel.nprocz,
                         E.parallel.me_loc[1], E.parallel.me_loc[2], E.parallel.me_loc[3],
                         E.control.v_steps_low, E.control.v_steps_high, E.control.down_heavy, E.control.up_heavy,
                         E.control.mg_cycle, E.control.p_iterations, E.control.precondition, E.control.augmented_Lagr };
        double ctl[4] = { E.control.accuracy, E.control.augmented, E.control.Atemp, E.data.therm_diff };
        DUMP_I32("meta", meta, 16);
        DUMP_F64("ctl", ctl, 4);
    }

    if(getenv("CCU_SETUP_ONLY"))
    {   /* operator + right-hand side only: general_stokes_solver (Drive_solvers.c:105-126) without the solve */
        velocities_conform_bcs(&E, E.U);
        assemble_forces(&E, 0);
        if(E.viscosity.update_allowed)
            get_system_viscosity(&E, 1, E.EVI[E.mesh.levmax], E.VI[E.mesh.levmax]);
        construct_stiffness_B_matrix(&E);
    }
    else
    {
        general_stokes_solver(&E);
        process_temp_field(&E, E.monitor.solution_cycles);
        process_new_velocity(&E, E.monitor.solution_cycles);
    }
    dump_level_arrays(&E);
    dump_fields(&E, "s0");
    if(want_kat) dump_kats(&E);

    for(step = 1; step <= nsteps && Emergency_stop == 0; step++)
    {
        one_timestep(&E);
        snprintf(tag, sizeof tag, "s%d", step);
        dump_fields(&E, tag);
    }
    if(E.control.composition && E.advection.markers > 0 && (want_kat || getenv("CCU_MARKER_KAT")))
    {   /* marker known answers on the final state: the lookup tables, then Euler and Runge_Kutta (Composition_adv.c:108,61)
         * with the current velocity, exactly as PG_timestep_particle calls them (Advection_diffusion.c:229,162) */
        const int nm = E.advection.markers, nno = E.lmesh.nno, nel = E.lmesh.nel;
        int d, ip[4] = { E.lmesh.rnoz, E.advection.markers_per_ele, E.advection.markers, E.advection.markers_uplimit };
        double dp[8] = { E.XG1[1], E.XG1[2], E.XG1[3], E.XG2[1], E.XG2[2], E.XG2[3], E.advection.timestep, E.control.Acomp };
        char nm_[100];
        DUMP_I32("mk_ints", ip, 4);
        DUMP_F64("mk_doubles", dp, 8);
        DUMP_F64("mk_XP1", E.XP[1] + 1, E.lmesh.nox); DUMP_F64("mk_XP2", E.XP[2] + 1, E.lmesh.noy); DUMP_F64("mk_XP3", E.XP[3] + 1, E.lmesh.noz);
        DUMP_I32("mk_RG3", E.RG[3], E.lmesh.rnoz + 1);
        DUMP_U32("mk_Element", E.Element + 1, nel);
        for(d = 1; d <= 3; d++) { snprintf(nm_, sizeof nm_, "mk_in_XMC%d", d); DUMP_F64(nm_, E.XMC[d] + 1, nm); }
        DUMP_I32("mk_in_C12", E.C12 + 1, nm); DUMP_I32("mk_in_CElement", E.CElement + 1, nm); DUMP_F32("mk_in_CE", E.CE + 1, nel);
        for(d = 1; d <= 3; d++) { snprintf(nm_, sizeof nm_, "mk_in_V%d", d); DUMP_F32(nm_, E.V[d] + 1, nno); }
        Euler(&E, E.C, E.V, 0);
        {   /* several ranks: markers changed owner inside Euler (transfer_markers_processors), the count with them */
            const int nm = E.advection.markers;
            dump_scalar_i("mk_euler_nmarkers", nm);
            for(d = 1; d <= 3; d++) { snprintf(nm_, sizeof nm_, "mk_euler_XMC%d", d); DUMP_F64(nm_, E.XMC[d] + 1, nm); }
            DUMP_I32("mk_euler_C12", E.C12 + 1, nm);
        for(d = 1; d <= 3; d++)
        {
            snprintf(nm_, sizeof nm_, "mk_euler_XMCpred%d", d); DUMP_F64(nm_, E.XMCpred[d] + 1, nm);
            snprintf(nm_, sizeof nm_, "mk_euler_VO%d", d); DUMP_F32(nm_, E.VO[d] + 1, nm);
        }
        DUMP_I32("mk_euler_CElement", E.CElement + 1, nm);
        } DUMP_F32("mk_euler_C", E.C + 1, nno); DUMP_F32("mk_euler_CE", E.CE + 1, nel);
        Runge_Kutta(&E, E.C, E.V, 1);
        {
            const int nm = E.advection.markers;
        for(d = 1; d <= 3; d++)
        {
            snprintf(nm_, sizeof nm_, "mk_rk_XMC%d", d); DUMP_F64(nm_, E.XMC[d] + 1, nm);
            snprintf(nm_, sizeof nm_, "mk_rk_Vpred%d", d); DUMP_F32(nm_, E.Vpred[d] + 1, nm);
        }
        DUMP_I32("mk_rk_CElement", E.CElement + 1, nm);
        } DUMP_F32("mk_rk_C", E.C + 1, nno); DUMP_F32("mk_rk_CE", E.CE + 1, nel);
        dump_scalar_i("mk_rk_nmarkers", E.advection.markers);
    }
    fclose(g_manifest);
    fflush(NULL);
    MPI_Finalize();
    return 0;
}
