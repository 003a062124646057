/* citcom_dropin.c -- reference-side binding of libcitcomcu_b200.so.
 *
 * Defines `general_stokes_solver` (Drive_solvers.c:45) and `PG_timestep` (Advection_diffusion.c:251) with the
 * reference's own signatures (src/prototypes.h) on top of the C ABI in include/citcomcu_b200.h.  A maintainer either adds this
 * file to src/Makefile's CFILES in place of the body of Drive_solvers.c:general_stokes_solver, or --
 * without touching the reference at all -- preloads it:
 *
 *     LD_PRELOAD=dropin/libcitcomcu_dropin.so  citcom.mpi  input_file
 *
 * (the reference's objects are position independent, so the call sites in Citcom.c:97,136 resolve to
 * this definition).  It is compiled against the reference's own headers where they lie
 * (-I$(REF)/src); nothing of the reference is copied into this repository.
 *
 * Per call: E->T and E->buoyancy go to the device, viscosity / stiffness / BI / BPI are rebuilt there
 * when the reference would rebuild them (Construct_arrays.c:849), the Uzawa + full-multigrid solve runs
 * on the device from the previous E->U / E->P, and E->U, E->P, E->V and E->EVI[levmax] come back for
 * the reference's diagnostics and its energy step.
 *
 * PG_timestep: E->V, E->T, E->Tdot go to the device, std_timestep + predictor + (pg_solver, corrector) x temp_iterations
 * with the Tmax safeguard run there, T / Tdot come back; temperatures_conform_bcs and thermal_buoyancy stay the
 * reference's (they are O(nno) host loops whose results the host needs anyway), and so does process_heating: its element
 * heating arrays (adiabatic, viscous, phase-change latent) are uploaded before the step.  CCU_DROPIN_ENERGY=0 keeps the
 * reference's own energy step.
 *
 * Several ranks: one rank per GPU (CCU_DEVICE, or rank modulo CCU_DEVICES); the reference's own MPI_Bcast carries the NCCL
 * id once, after that the library does the halo sums and reductions over NCCL (ccu_comm_init).
 *
 * Regional-spherical runs (Geometry=Rsphere): the node positions E->XX and E->SXX go up and the device uses the Rsphere branches of
 * the element routines; process_heating / thermal_buoyancy / heat_flux stay the reference's host code.
 *
 * Unsupported configurations stop the run loudly (there is no CPU fallback): the selective / strain-weakening variants of plastic yielding,
 * anisotropic viscosity, periodic side walls, heat-flux boundary conditions.
 *
 * citcom_dropin_funcs.c (same library) binds the INNER functions of the path one by one (CCU_DROPIN_FUNCS).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <mpi.h>
#include "global_defs.h"
#include "prototypes.h"
#include "citcomcu_b200.h"

ccu_ctx *g_ctx = NULL;              /* shared with citcom_dropin_funcs.c */
int g_ccu_device_geometry = 1;      /* 0: spherical mesh with CCU_DROPIN_STOKES=0, only the function-level bindings (operator assembled by the reference) */
int g_ccu_cartesian = 1;            /* 0: regional-spherical mesh: energy step, heating and markers stay the reference's host code */
static int g_calls = 0;

void ccu_dropin_die(const char *msg);
static void die(const char *msg) { ccu_dropin_die(msg); }
void ccu_dropin_die(const char *msg)
{
    fprintf(stderr, "citcomcu_b200 drop-in: %s\n", msg);
    exit(9);
}
#define CCU(call) do { if((call) != 0) { fprintf(stderr, "citcomcu_b200 drop-in: %s failed: %s\n", #call, ccu_last_error()); exit(9); } } while(0)

void ccu_dropin_init(struct All_variables *E);
static void dropin_init(struct All_variables *E) { ccu_dropin_init(E); }
void ccu_dropin_init(struct All_variables *E)
{
    ccu_config cfg;
    int lev;
    const char *dev = getenv("CCU_DEVICE");
    /* Regional-spherical runs: the device has no Rsphere operator construction / energy step yet.  What it can run is the SOLVER on
       the operator the reference assembled (function-level bindings, citcom_dropin_funcs.c: the node-stored stiffness, the pressure
       operators and the transfer weights are geometry-free once assembled), so the context is created without device geometry. */
    g_ccu_cartesian = E->control.CART3D ? 1 : 0;
    {   /* Rsphere: the Stokes operator is assembled on the device (Rsphere branches of get_elt_k / get_elt_g / get_elt_f) unless the run
           keeps the reference's driver (CCU_DROPIN_STOKES=0) and binds inner functions only */
        const char *sw = getenv("CCU_DROPIN_STOKES");
        g_ccu_device_geometry = (g_ccu_cartesian || !(sw && atoi(sw) == 0)) ? 1 : 0;
    }
    if(E->viscosity.BDEPV && (E->viscosity.psrw || E->viscosity.pdepv_for_flavor || E->viscosity.pdepv_for_zero_comp ||
                              E->viscosity.pdepv_for_unity_comp || E->viscosity.strain_dep_plasticity))
        die("plastic yielding (BDEPV): the strain-rate-weakening, flavour / composition-selective and strain-dependent variants of visc_from_B are not on the device path");
    if(E->viscosity.BDEPV && E->control.restart) die("plastic yielding with restart (strain rate of the restart velocity) is not on the device path");
    if(E->viscosity.CDEPV && (E->viscosity.cdepv_for_flavor || E->viscosity.const_lith_visc || E->viscosity.crust_option == 2))
        die("composition-dependent viscosity: the flavour / crust / constant-lithosphere variants of visc_from_C are not on the device path");
    if(E->viscosity.SDEPV && E->viscosity.sdepv_rheology != 1 && E->viscosity.sdepv_rheology != 2)
        die("stress-dependent viscosity: sdepv_rheology 1 and 2 are on the device path, 3 (dimensional Arrhenius law) is not");
    if(E->viscosity.SDEPV && E->control.restart) die("stress-dependent viscosity with restart (strain rate of the restart velocity) is not on the device path");
    if(E->control.force_initial_stokes_iteration) die("force_initial_stokes_iteration is not on the device path");
    /* options that change the operator and that the device build does not implement: stop, never differ silently */
    /* VISC_SMOOTH: apply_viscosity_smoother (Viscosity_structures.c:446) smooths the NODAL array E->VI only; the Gauss-point viscosity the
       operator is built from is untouched (the reference's U, T, EVI are bitwise the same with it on and off), so nothing to do here */
    if(E->viscosity.allow_anisotropic_viscosity) die("anisotropic viscosity is not on the device path");
#ifdef USE_GGRD
    if(E->control.ggrd.mat_control) die("ggrd material control (viscosity prefactors from grids) is not on the device path");
#endif
    if(E->mesh.periodic_x || E->mesh.periodic_y) die("periodic side walls are not on the device path");
    if(!E->control.NMULTIGRID) die("Solver=multigrid is required");
    memset(&cfg, 0, sizeof cfg);
    cfg.levmin = E->mesh.levmin; cfg.levmax = E->mesh.levmax;
    for(lev = cfg.levmin; lev <= cfg.levmax; lev++)
    {
        cfg.nox[lev] = E->lmesh.NOX[lev]; cfg.noy[lev] = E->lmesh.NOY[lev]; cfg.noz[lev] = E->lmesh.NOZ[lev];
    }
    cfg.v_steps_low = E->control.v_steps_low; cfg.v_steps_high = E->control.v_steps_high;
    cfg.down_heavy = E->control.down_heavy; cfg.up_heavy = E->control.up_heavy; cfg.mg_cycle = E->control.mg_cycle;
    cfg.p_iterations = E->control.p_iterations; cfg.accuracy = E->control.accuracy;
    /* one rank per GPU: CCU_DEVICE pins the device, else rank modulo CCU_DEVICES (default: one device per rank up to 8) */
    {
        const char *nd = getenv("CCU_DEVICES");
        const int ndev = nd ? atoi(nd) : 8;
        cfg.device = dev ? atoi(dev) : (E->parallel.nproc > 1 ? E->parallel.me % (ndev > 0 ? ndev : 1) : 0);
    }
    CCU(ccu_create(&cfg, &g_ctx));
    if(E->parallel.nproc > 1)
    {   /* the reference's own MPI carries the 128-byte NCCL id, nothing else (Parallel_related.c:80-173 decomposition) */
        char id[128];
        memset(id, 0, sizeof id);
        if(E->parallel.me == 0) CCU(ccu_comm_unique_id(id));
        MPI_Bcast(id, 32, MPI_INT, 0, MPI_COMM_WORLD);      /* 128 bytes */
        CCU(ccu_comm_init(g_ctx, E->parallel.nprocx, E->parallel.nprocy, E->parallel.nprocz,
                          E->parallel.me_loc[1], E->parallel.me_loc[2], E->parallel.me_loc[3], id));
    }
    for(lev = cfg.levmin; lev <= cfg.levmax; lev++)
    {
        CCU(ccu_set_node_flags(g_ctx, lev, E->NODE[lev] + 1));
        if(g_ccu_device_geometry) CCU(ccu_set_coordinates(g_ctx, lev, E->XX[lev][1] + 1, E->XX[lev][2] + 1, E->XX[lev][3] + 1));
        if(g_ccu_device_geometry && !g_ccu_cartesian) CCU(ccu_set_spherical_coordinates(g_ctx, lev, E->SXX[lev][1] + 1, E->SXX[lev][2] + 1, E->SXX[lev][3] + 1));
    }
    if(!g_ccu_device_geometry)
    {
        if(E->parallel.me == 0) fprintf(stderr, "citcomcu_b200 drop-in: Rsphere geometry: solver functions on CUDA device %d, operator assembly by the reference\n", cfg.device);
        return;
    }
    CCU(ccu_build_geometry(g_ctx));
    CCU(ccu_set_viscosity_law(g_ctx, E->viscosity.TDEPV, E->viscosity.RHEOL, E->viscosity.num_mat, E->viscosity.N0, E->viscosity.E,
                              E->viscosity.T, E->viscosity.Z, E->viscosity.MIN, E->viscosity.min_value, E->viscosity.MAX,
                              E->viscosity.max_value, E->viscosity.smooth_cycles));
    CCU(ccu_set_material(g_ctx, E->mat + 1));
    {   /* imposed boundary velocities: uploaded when any subdomain has a non-zero one (E->VB, Boundary_conditions.c) */
        int n, d, any = 0, all = 0;
        for(d = 1; d <= 3 && !any; d++)
            for(n = 1; n <= E->lmesh.nno && !any; n++) any = E->VB[d][n] != 0.0;
        all = any;
        if(E->parallel.nproc > 1) MPI_Allreduce(&any, &all, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
        if(all) CCU(ccu_set_velocity_bcs(g_ctx, E->VB[1] + 1, E->VB[2] + 1, E->VB[3] + 1));
    }
    if(E->viscosity.CDEPV)
        CCU(ccu_set_cdepv(g_ctx, 1, E->viscosity.layer_pre_comp, E->viscosity.pre_comp, E->viscosity.cdepv_absolute, E->control.check_c_irange));
    if(E->viscosity.BDEPV)
    {
        CCU(ccu_set_bdepv(g_ctx, 1, E->viscosity.abyerlee, E->viscosity.bbyerlee, E->viscosity.lbyerlee, E->viscosity.plasticity_dimensional,
                          E->monitor.length_scale, E->monitor.tau_scale, E->viscosity.plasticity_trans, E->viscosity.plasticity_viscosity_offset));
        if(!E->viscosity.SDEPV)     /* the iteration controls are shared with SDEPV (Drive_solvers.c:157-159) */
            CCU(ccu_set_sdepv(g_ctx, 0, 1, NULL, NULL, E->viscosity.sdepv_misfit, E->viscosity.sdepv_iter_damp, E->monitor.max_sdep_visc_iter, 0, 0.0f, 0.0f));
    }
    if(E->viscosity.SDEPV)
        CCU(ccu_set_sdepv(g_ctx, 1, E->viscosity.sdepv_rheology, E->viscosity.sdepv_expt, E->viscosity.sdepv_trns, E->viscosity.sdepv_misfit,
                          E->viscosity.sdepv_iter_damp, E->monitor.max_sdep_visc_iter, E->viscosity.sdepv_start_from_newtonian,
                          E->viscosity.sdepv_trns_T, E->viscosity.sdepv_trns_c));
    if(E->parallel.me == 0) fprintf(stderr, "citcomcu_b200 drop-in: Stokes solve on CUDA device %d%s\n", cfg.device,
                                    g_ccu_cartesian ? "" : " (regional-spherical element routines)");
}

typedef void (*pg_fn)(struct All_variables *);

void general_stokes_solver(struct All_variables *E)
{
    int rebuild, its = 0, i;
    float res = 0.0f;
    double t0 = CPU_time0();
    const int lm = E->mesh.levmax;
    {   /* CCU_DROPIN_STOKES=0: keep the reference's own driver (its inner functions may still be bound one by one,
           citcom_dropin_funcs.c) */
        const char *sw = getenv("CCU_DROPIN_STOKES");
        if(sw && atoi(sw) == 0)
        {
            static pg_fn next = NULL;
            if(!next) next = (pg_fn)dlsym(RTLD_NEXT, "general_stokes_solver");
            if(!next) die("CCU_DROPIN_STOKES=0 but no other general_stokes_solver is linked");
            next(E);
            return;
        }
    }
    if(!g_ctx) dropin_init(E);
    if(!g_ccu_device_geometry) die("general_stokes_solver: the context was created without device geometry (CCU_DROPIN_STOKES=0 at the first call)");
    E->monitor.elapsed_time_vsoln1 = E->monitor.elapsed_time_vsoln;
    E->monitor.elapsed_time_vsoln = E->monitor.elapsed_time;
    /* Construct_arrays.c:849: first call, or viscosity updates allowed and step % update_every_steps == 0 */
    rebuild = (g_calls == 0) || (E->viscosity.update_allowed && E->monitor.solution_cycles % E->control.KERNEL == 0)
              || (E->monitor.solution_cycles == E->control.freeze_surface_at_step);
    velocities_conform_bcs(E, E->U);
    if(E->viscosity.CDEPV) CCU(ccu_set_composition(g_ctx, E->C + 1));      /* the markers' nodal composition as the host holds it */
    CCU(ccu_general_stokes_solver(g_ctx, E->T + 1, E->buoyancy + 1, rebuild, E->control.augmented_Lagr, E->control.augmented,
                                  E->control.precondition, 1, E->U, E->P + 1, &its, &res));
    if(rebuild) CCU(ccu_get_level_array(g_ctx, lm, CCU_ARR_EVI, E->EVI[lm] + 1));
    v_from_vector(E, E->V, E->U);
    E->monitor.visc_iter_count = 1;
    if(E->viscosity.SDEPV || E->viscosity.BDEPV) { double mis; CCU(ccu_get_sdepv_iterations(g_ctx, &E->monitor.visc_iter_count, &mis)); }
    g_calls++;
    if(E->control.print_convergence && E->parallel.me == 0)
    {
        fprintf(stderr, "citcomcu_b200: after (%03d) pressure loops and %g sec for step %d\n", its, CPU_time0() - t0, E->monitor.solution_cycles);
        fprintf(E->fp, "citcomcu_b200: after (%03d) pressure loops and %g sec for step %d\n", its, CPU_time0() - t0, E->monitor.solution_cycles);
    }
    (void)i;
}

/* ---- energy step (Advection_diffusion.c:251-349) ---- */
static int g_energy = 0;

void PG_timestep(struct All_variables *E)
{
    float dt = 0.0f, Tint = 0.0f;
    int n;
    {   /* CCU_DROPIN_ENERGY=0: leave the energy step to the reference's own definition (preload builds only) */
        const char *sw = getenv("CCU_DROPIN_ENERGY");
        if(sw && atoi(sw) == 0)
        {
            static pg_fn next = NULL;
            if(!next) next = (pg_fn)dlsym(RTLD_NEXT, "PG_timestep");
            if(!next) die("CCU_DROPIN_ENERGY=0 but no other PG_timestep is linked");
            next(E);
            return;
        }
    }
    if(!g_ctx) dropin_init(E);
    if(!g_ccu_device_geometry) die("PG_timestep: the context was created without device geometry (CCU_DROPIN_STOKES=0 on a regional-spherical run)");
    if(!g_energy)
    {
        if(!E->advection.ADVECTION) die("ADVECTION=off is not on the device path");
        for(n = 1; n <= E->lmesh.nno; n++)
            if(E->node[n] & FBZ) die("heat-flux boundary conditions are not on the device path");
        CCU(ccu_set_energy_params(g_ctx, E->advection.fine_tune_dt, E->advection.fixed_timestep, E->advection.gamma,
                                  E->advection.temp_iterations, E->diffusivity + 1, E->expansivity + 1, E->control.Q0));
        g_energy = 1;
        if(E->parallel.me == 0) fprintf(stderr, "citcomcu_b200 drop-in: energy step on the CUDA device\n");
    }
    E->advection.timesteps++;
    /* extended-Boussinesq / phase-change heating terms: the reference's process_heating (Citcom.c:116) has just filled them */
    if(E->control.adi_heating || E->control.visc_heating || E->control.Ra_410 != 0.0 || E->control.Ra_670 != 0.0)
        CCU(ccu_set_heating_arrays(g_ctx, E->heating_adi + 1, E->heating_visc + 1, E->heating_latent + 1));
    CCU(ccu_set_velocity(g_ctx, E->V[1] + 1, E->V[2] + 1, E->V[3] + 1));
    CCU(ccu_PG_timestep(g_ctx, E->T + 1, E->Tdot + 1, &dt, &Tint));
    E->advection.timestep = dt;
    E->monitor.T_interior = Tint;
    E->advection.dt_reduced = 1.0;
    E->advection.total_timesteps++;
    E->monitor.elapsed_time += E->advection.timestep;
    temperatures_conform_bcs(E);
    thermal_buoyancy(E);
    E->advection.last_sub_iterations = 1;
    E->control.keep_going = (E->monitor.solution_cycles < E->advection.max_timesteps) ? 1 : 0;
}
