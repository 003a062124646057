/* citcomcu_b200.h -- C ABI of libcitcomcu_b200.so: CitcomCU's per-timestep hot path
 * (Stokes solve; energy step to follow) as hand-written sm_100a CUDA kernels.
 *
 * Plain pointers and sizes only.  Every entry point names the reference function it
 * replaces (paths relative to the reference tree, src/...).  Host arrays use the
 * REFERENCE's conventions so a caller holding `struct All_variables *E` can pass its
 * members straight through (see INTEGRATION.md for the exact bindings):
 *   - nodes      n = k + noz*(j + nox*i), k = z index (fastest), j = x, i = y   (Construct_arrays.c:158-165)
 *   - equations  3*n + d, vectors are double[neq] (slack slots of the reference are not touched)
 *   - elements / pressure dofs  e = ez + elz*(ex + elx*ey); arrays passed here are 0-based,
 *     i.e. the caller passes `E->P + 1`, `E->BPI[lev] + 1`, `E->NODE[lev] + 1`, ...
 *   - levels are the reference's level numbers (E->mesh.levmin .. levmax)
 * Inside the library every level lives in HBM in an 8-colour blocked layout.
 *
 * All functions return 0 on success, non-zero on failure (ccu_last_error() has the text);
 * there is no CPU fallback: without a CUDA device ccu_create fails.
 */
#ifndef CITCOMCU_B200_H
#define CITCOMCU_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define CCU_MAX_LEVELS 12

typedef struct ccu_ctx ccu_ctx;

/* Mesh + solver controls the hot path reads from E (SURVEY.md 8b "State the callee reads") */
typedef struct {
    int levmin, levmax;                 /* E->mesh.levmin / levmax */
    int nox[CCU_MAX_LEVELS];            /* E->lmesh.NOX[lev] (nodes in x), indexed by level */
    int noy[CCU_MAX_LEVELS];            /* E->lmesh.NOY[lev] */
    int noz[CCU_MAX_LEVELS];            /* E->lmesh.NOZ[lev] */
    int v_steps_low, v_steps_high;      /* E->control.v_steps_low / v_steps_high */
    int down_heavy, up_heavy, mg_cycle; /* E->control.down_heavy / up_heavy / mg_cycle */
    int p_iterations;                   /* E->control.p_iterations */
    double accuracy;                    /* E->control.accuracy */
    int device;                         /* CUDA device ordinal */
} ccu_config;

const char *ccu_last_error(void);
int ccu_create(const ccu_config *cfg, ccu_ctx **out);
void ccu_destroy(ccu_ctx *ctx);
/* run all kernels on this CUDA stream (cudaStream_t cast to void*, not the legacy stream 0: it cannot be captured into
 * CUDA graphs); NULL = the context's own non-blocking stream (the default) */
int ccu_set_stream(ccu_ctx *ctx, void *cuda_stream);
int ccu_synchronize(ccu_ctx *ctx);
/* CCU_OPT_GRAPHS (default 1): replay the coarse-level part of the multigrid cycle as CUDA graphs */
/* kernel selection by level size: levels with nno <= SMALL_NODES run all sweeps in one single-CTA launch, <= WARP_NODES
 * use a warp per node, <= QUAD_NODES four lanes per node, larger levels LANES_LARGE (1 or 4) lanes per node */
/* MATVEC_TAB / RELAX_TAB: table-driven (compact code) row kernels on levels above QUAD_NODES */
enum { CCU_OPT_GRAPHS = 0, CCU_OPT_SMALL_NODES = 1, CCU_OPT_WARP_NODES = 2, CCU_OPT_QUAD_NODES = 3, CCU_OPT_LANES_LARGE = 4,
       CCU_OPT_MATVEC_TAB = 5, CCU_OPT_RELAX_TAB = 6,
       CCU_OPT_SMEM_NODES = 7 /* levels with nno <= this (max 434) run all sweeps out of one SM's shared memory */,
       /* column-resident kernels (csrc/ccu_col.cuh) on levels with nno > COL_NODES (default 2000000: measured slower than the colour passes on the 1.08e6-node level): RELAX_COL / MATVEC_COL
        * (default 1) switch them on or off; COL_SHAPE 0 = columns of 6 (y) x 8 (x) nodes, two CTAs per SM, 1 = 12 x 8, one CTA
        * per SM, 2 = 4 x 8, three CTAs per SM.  A CTA streams the stiffness of its column through shared memory with one bulk
        * asynchronous copy per z layer, so every coefficient crosses HBM once per sweep.  The column smoother is a Gauss-Seidel
        * sweep in column order (column colours 3..0, z ascending, (y,x)-parity colours 3..0 inside a layer) instead of the plain
        * 8-colour order; oracle/restate.c ccu_r_ordered_gs mode 10 is its CPU statement.  COL_WF 1 (default): the four column
        * colours of a sweep run in one launch, columns waiting on progress words of their neighbours; 0: one launch per colour
        * (bitwise the same result). */
       CCU_OPT_COL_NODES = 9, CCU_OPT_RELAX_COL = 10, CCU_OPT_MATVEC_COL = 11, CCU_OPT_COL_SHAPE = 13, CCU_OPT_COL_WF = 14,
       /* levels with nno > FULL_NODES (default 500000) also keep the thirteen upper-neighbour blocks of every node, transposed, at the
        * node's own slot (+468 B/node) when RELAX_FULL or MATVEC_FULL is set (both default 0): the colour-pass smoother / matvec then stream
        * every coefficient of a row once per pass instead of gathering each stored block twice.  Measured on B200 (r02): 1.27x fewer DRAM
        * bytes but only 3 % faster (the pass stops being DRAM-bound), the matvec slower -- kept as an option, see DESIGN.md */
       CCU_OPT_FULL_NODES = 18, CCU_OPT_RELAX_FULL = 19, CCU_OPT_MATVEC_FULL = 20,
       CCU_OPT_P2P_HALO = 21 /* subdomain-per-GPU runs: 1 = halo sums through peer memory (CUDA IPC landing buffers over NVLink: a push kernel,
        * a flag per sender, a wait kernel; no NCCL call per exchange), 0 (default: measured faster on 8 x B200) = grouped ncclSend/ncclRecv.
        * Set it identically on all ranks. */,
       CCU_OPT_HALO_OVERLAP = 22 /* subdomain-per-GPU runs: 1 = the duplicated-node exchange of a smoother sweep runs on a second (high-priority)
        * stream while the colour passes relax the nodes far enough from the subdomain faces; the passes are split by distance so that the
        * result is bitwise that of the serial order (levels with more than 4e5 nodes per subdomain; 2 = every level).  0 (default: the
        * split passes and the cross-stream joins cost more than the hidden exchange, measured on B200) = exchange first, then the passes. */,
       CCU_OPT_BOTTOM_CLUSTER = 15 /* 1 (default): the shared-memory bottom smoother runs on an 8-CTA cluster with fp64 rows in
        * distributed shared memory (ccu_k_relax_bottom); 0: on one SM (ccu_k_relax_smem) */ };
int ccu_set_option(ccu_ctx *ctx, int option, int value);
/* current value of an option; for the column options, *value is what is in effect on level `lev` (0 when the level is too
 * small for the column kernels), so that a benchmark can name the kernel it timed */
int ccu_get_option(ccu_ctx *ctx, int option, int lev, int *value);
/* number of kernels this library has launched since creation (bench.py's gpu_launches) */
long long ccu_launch_count(ccu_ctx *ctx);

/* ---- subdomain-per-GPU runs: the reference's nprocx x nprocy x nprocz block decomposition
 * (Parallel_related.c:80-173), one process and one ccu_ctx per GPU, rank = z + nprocz*x + nprocz*nprocx*y.
 * Every level's mesh sizes in ccu_config are the LOCAL ones (E->lmesh.*), face nodes duplicated as in the reference.
 * ccu_comm_unique_id: rank 0 obtains a 128-byte NCCL id and hands it to all ranks by any means (the reference's
 * MPI_Bcast, torch.distributed, a file); ccu_comm_init: collective over all ranks, replaces parallel_domain_decomp1 /
 * parallel_communication_routs1 (Parallel_related.c:80,477): builds the duplicated-node tables of every level and the
 * NCCL communicator.  After it, every entry point below performs the halo sums (exchange_id_d20 / exchange_node_f20,
 * Parallel_related.c:1181,1270) and global reductions (global_vdot / global_pdot) of the reference internally.
 * nproc = 1x1x1 is allowed (no NCCL needed) and is the default when ccu_comm_init is never called. */
int ccu_comm_unique_id(char *out128);
int ccu_comm_init(ccu_ctx *ctx, int nprocx, int nprocy, int nprocz, int me_x, int me_y, int me_z, const char *unique_id128);
/* Replicated coarse levels: below a few ten thousand nodes a multigrid level is pure latency, and a halo exchange per
 * smoother sweep makes it worse.  After ccu_agglomerate(ctx, lev) the levels <= lev are solved REPLICATED: every rank
 * holds the GLOBAL mesh of those levels in the returned context, one all-gather per visit hands over the restricted
 * right-hand side, all ranks run the identical coarse part of the cycle with no further communication, and each takes
 * its piece of the correction back.  The caller gives the returned context the global node flags and coordinates of
 * levels levmin..lev (ccu_set_node_flags, ccu_set_coordinates) and calls ccu_build_geometry on it; the operators are
 * then (re)built together with ctx's.  The returned context is owned by ctx (do not destroy it). */
int ccu_agglomerate(ccu_ctx *ctx, int agg_lev, ccu_ctx **coarse_out);
/* host-only table builders behind ccu_comm_init (no GPU needed; tests/test_decomp.py): sizes = {neighbours, packed nodes,
 * duplicated nodes, contributions}; see csrc/ccu_comm.cuh for the meaning of the arrays */
int ccu_halo_sizes(const int nproc[3], const int me[3], int nox, int noy, int noz, int sizes[4]);
int ccu_halo_tables(const int nproc[3], const int me[3], int nox, int noy, int noz, int *nb_rank, int *nb_off, int *nb_cnt,
                    int *send_n, int *sh_n, int *sh_ptr, int *sh_src, unsigned char *owned);

/* ---- operator upload (what construct_stiffness_B_matrix, Construct_arrays.c:834, leaves in E) ---- */
/* E->NODE[lev]+1 : flag bits VBX 0x2, VBZ 0x4, VBY 0x8 (global_defs.h:65-89) */
int ccu_set_node_flags(ccu_ctx *ctx, int lev, const unsigned *node /*[nno]*/);
/* E->Eqn_k1/2/3[lev] in the reference's half-matrix layout (Construct_arrays.c:359-535), E->BI[lev] */
int ccu_set_stiffness(ccu_ctx *ctx, int lev, const float *eqn_k1, const float *eqn_k2, const float *eqn_k3 /*[nno*42]*/,
                      const double *BI /*[neq]*/);
/* E->elt_del[lev][1..nel].g (get_elt_g, Element_calculations.c:831) and E->BPI[lev]+1 (build_diagonal_of_Ahat, :654) */
int ccu_set_pressure_ops(ccu_ctx *ctx, int lev, const float *elt_del /*[nel*24]*/, const double *BPI /*[npno]*/);
/* E->TWW[lev][1..nel].node[1..8], E->MASS[lev]+1 (Size_does_matter.c:618), E->ECO[lev][1..nel].size[1..3] */
int ccu_set_transfer_weights(ccu_ctx *ctx, int lev, const float *TWW /*[nel*8]*/, const float *MASS /*[nno]*/,
                             const float *eco_size /*[nel*3]*/);

/* ---- operator construction on the device (rebuilt every `update_every_steps` when TDEPV) ---- */
/* node coordinates E->XX[lev][1..3]+1 (x, y, z), natural node order (Nodal_mesh.c:53 node_locations) */
int ccu_set_coordinates(ccu_ctx *ctx, int lev, const float *X1, const float *X2, const float *X3 /*[nno] each*/);
/* Regional-spherical geometry (E->control.Rsphere): E->SXX[lev][1..3]+1 = (theta, phi, r) of the nodes; ccu_set_coordinates then carries
 * the CARTESIAN node coordinates E->XX[lev], as the reference keeps them.  Switches the context's element routines to the Rsphere branches
 * of get_elt_k (Element_calculations.c:249-260), get_elt_g (:896-920), get_elt_f (:1072-1086) and mass_matrix's ECO.size
 * (Size_does_matter.c:661-680), of the SUPG residual, of strain_rate_2_inv in ccu_process_heating and of the marker position update.
 * ccu_thermal_buoyancy and ccu_averages integrate their layer averages over the spherical shells, ccu_heat_flux takes the radial
 * derivative, ccu_phase_change reads r as the depth, ccu_get_stress_topo uses the Rsphere strain rates of get_stress. */
int ccu_set_spherical_coordinates(ccu_ctx *ctx, int lev, const float *theta, const float *phi, const float *r);
/* mass_matrix (Size_does_matter.c:618): TWW, MASS, ECO.size; construct_elt_gs / get_elt_g (Element_calculations.c:831): elt_del; all levels */
int ccu_build_geometry(ccu_ctx *ctx);
/* E->viscosity.{TDEPV,RHEOL,num_mat,N0,E,T,Z,MIN,min_value,MAX,max_value,smooth_cycles} (Viscosity_structures.c:57-326) */
int ccu_set_viscosity_law(ccu_ctx *ctx, int tdepv, int rheol, int num_mat, const float *N0, const float *E, const float *T,
                          const float *Z, int vmin, float min_value, int vmax, float max_value, int smooth_cycles);
int ccu_set_material(ccu_ctx *ctx, const int *mat /*[nel] = E->mat+1*/);
/* stress-dependent viscosity (visc_from_S, Viscosity_structures.c:744, sdepv_rheology 1 and 2) and the viscosity <-> velocity iteration of
 * general_stokes_solver (Drive_solvers.c:120-159): E->viscosity.{SDEPV, sdepv_rheology, sdepv_expt, sdepv_trns, sdepv_misfit, sdepv_iter_damp,
 * sdepv_start_from_newtonian, sdepv_trns_T, sdepv_trns_c}, E->monitor.max_sdep_visc_iter; call after ccu_set_viscosity_law */
int ccu_set_sdepv(ccu_ctx *ctx, int on, int rheology, const float *expt, const float *trns, float misfit, float iter_damp, int max_iter,
                  int start_from_newtonian, float trns_T, float trns_c);
/* Composition-dependent viscosity, visc_from_C (Viscosity_structures.c:1784-1935) in its prefactor mode (and cdepv_absolute):
 * E->viscosity.{CDEPV, layer_pre_comp, pre_comp[2 or 2*num_mat], cdepv_absolute}, E->control.check_c_irange.  The composition is the
 * device marker set's nodal C when one lives on the device, else what ccu_set_composition handed in (E->C+1). */
int ccu_set_cdepv(ccu_ctx *ctx, int on, int layer_pre_comp, const float *pre_comp, int absolute, int check_c_irange);
/* Byerlee-type plastic yielding, visc_from_B (Viscosity_structures.c:1470-1755), regular branch: E->viscosity.{abyerlee, bbyerlee, lbyerlee}
 * [num_mat], plasticity_dimensional with E->monitor.{length_scale, tau_scale}, plasticity_trans, plasticity_viscosity_offset.  The
 * viscosity <-> velocity iteration of general_stokes_solver runs for it as for SDEPV; misfit, damping and iteration cap come from ccu_set_sdepv
 * (call it with on = 0 when only BDEPV is active). */
int ccu_set_bdepv(ccu_ctx *ctx, int on, const float *abyerlee, const float *bbyerlee, const float *lbyerlee, int dimensional, float length_scale,
                  float tau_scale, int plasticity_trans, float viscosity_offset);
int ccu_set_composition(ccu_ctx *ctx, const float *C /*[nno]*/);
/* iterations and relative velocity change of the last stress-dependent-viscosity loop (E->monitor.visc_iter_count) */
int ccu_get_sdepv_iterations(ccu_ctx *ctx, int *count_out, double *misfit_out);
int ccu_set_temperature(ccu_ctx *ctx, const float *T /*[nno] = E->T+1*/);
int ccu_set_element_viscosity(ccu_ctx *ctx, int lev, const float *EVI /*[nel*8] = E->EVI[lev]+1*/);
/* get_system_viscosity (Viscosity_structures.c:369): EVI[levmax] from the resident temperature / material groups */
int ccu_get_system_viscosity(ccu_ctx *ctx);
/* construct_stiffness_B_matrix (Construct_arrays.c:834): project_viscosity, get_elt_k + get_aug_k -> node-stored K,
 * BI (build_diagonal_of_K), BPI (build_diagonal_of_Ahat) on every level, from the resident EVI[levmax] */
int ccu_construct_stiffness_B_matrix(ccu_ctx *ctx, int augmented_Lagr, double augmented, int precondition);
/* assemble_forces (Element_calculations.c:74): buoyancy[nno] (NULL = resident) -> resident F (CCU_VEC_F); F_out optional */
int ccu_assemble_forces(ccu_ctx *ctx, const float *buoyancy, double *F_out /*[neq] or NULL*/);
/* E->VB (global_defs.h:1038): imposed boundary velocities, [nno] per direction in the reference's node order (all NULL clears them).
 * Non-zero values at flagged nodes enter the force vector -- the K.VB term of get_elt_f, Element_calculations.c:1038-1063, with the
 * viscosity as it stands when assemble_forces runs -- and the velocity vector (velocities_conform_bcs, Boundary_conditions.c:993). */
int ccu_set_velocity_bcs(ccu_ctx *ctx, const float *VB1, const float *VB2, const float *VB3);
int ccu_conform_velocity_bcs(ccu_ctx *ctx);      /* velocities_conform_bcs on the resident U */
/* read-back in the reference's layouts (tests / drop-in diagnostics) */
int ccu_get_stiffness(ccu_ctx *ctx, int lev, float *eqn_k1, float *eqn_k2, float *eqn_k3 /*[nno*42]*/, double *BI /*[neq] or NULL*/);
enum { CCU_ARR_TWW = 0, CCU_ARR_MASS = 1, CCU_ARR_ECO_SIZE = 2, CCU_ARR_ELT_DEL = 3, CCU_ARR_BPI = 4, CCU_ARR_EVI = 5 };
int ccu_get_level_array(ccu_ctx *ctx, int lev, int which, void *out);

/* ---- hot-path operators, host vectors in / out (drop-in granularity of SURVEY.md 8b) ---- */
/* n_assemble_del2_u / assemble_del2_u (Element_calculations.c:552, :480) */
int ccu_n_assemble_del2_u(ccu_ctx *ctx, int lev, const double *u, double *Au, int strip_bcs);
/* e_assemble_del2_u (Element_calculations.c:494): the element-by-element form of the same product (Solver=multigrid-el or
 * cgrad without node_assemble); evaluated from the node-stored matrix, which holds the same assembled coefficients */
int ccu_e_assemble_del2_u(ccu_ctx *ctx, int lev, const double *u, double *Au, int strip_bcs);
/* conj_grad (General_matrix_functions.c:661): Jacobi-preconditioned CG on K at level lev (Solver=cgrad); d0 out (zero
 * start), *cycles in = max iterations / out = iterations done, residual_out = sqrt(r.r/neq) as the reference returns */
int ccu_conj_grad(ccu_ctx *ctx, int lev, double *d0, const double *F, double acc, int *cycles, double *residual_out);
/* gauss_seidel (General_matrix_functions.c:1160): `cycles` 8-colour sweeps; d0 in/out (in only if guess), Ad = K*d0 out */
int ccu_gauss_seidel(ccu_ctx *ctx, int lev, double *d0, const double *F, double *Ad, int cycles, int guess);
/* project_vector (Solver_multigrid.c:72): fine level `lev` -> lev-1 */
int ccu_project_vector(ccu_ctx *ctx, int lev, const double *AU, double *AD);
/* interp_vector (Solver_multigrid.c:173): coarse level `lev` -> lev+1 */
int ccu_interp_vector(ccu_ctx *ctx, int lev, const double *AD, double *AU);
/* strip_bcs_from_residual (Boundary_conditions.c:926) */
int ccu_strip_bcs_from_residual(ccu_ctx *ctx, int lev, double *res);
/* assemble_div_u / assemble_grad_p (Element_calculations.c:691, :727), finest level */
int ccu_assemble_div_u(ccu_ctx *ctx, int lev, const double *U, double *divU /*[npno]*/);
int ccu_assemble_grad_p(ccu_ctx *ctx, int lev, const double *P /*[npno]*/, double *gradP);
/* global_vdot / global_pdot (Global_operations.c:339, :359) */
int ccu_global_vdot(ccu_ctx *ctx, int lev, const double *A, const double *B, double *out);
int ccu_global_pdot(ccu_ctx *ctx, int lev, const double *A, const double *B, double *out);
/* multi_grid (General_matrix_functions.c:525): one FMG cycle at levmax; F in = rhs, out = residual; d1 out */
int ccu_multi_grid(ccu_ctx *ctx, double *d1, double *F, double *residual_out);
/* solve_del2_u (General_matrix_functions.c:368): returns `valid` through *valid_out, MG cycle count through *cycles_out */
int ccu_solve_del2_u(ccu_ctx *ctx, double *d0, const double *F, double acc, int *valid_out, int *cycles_out);
/* solve_Ahat_p_fhat (Stokes_flow_Incomp.c:295): V, P in/out, F in; *steps_max in = p_iterations, out = iterations done.
 * hist (may be NULL) receives per iteration {v, dv/v, div/v, p, dp/p} as printed by generate_log_message (:501). */
int ccu_solve_Ahat_p_fhat(ccu_ctx *ctx, double *V, double *P, const double *F, double imp, int *steps_max,
                          float *residual_out, double *hist /*[steps_max*5]*/);

/* ---- device-resident forms (inputs already in HBM; used by bench.py for `value` and the roofline) ---- */
enum { CCU_VEC_VEL = 0, CCU_VEC_RES = 1, CCU_VEC_RHS = 2, CCU_VEC_FL = 3, CCU_VEC_DEL_VEL = 4, CCU_VEC_AU = 5,
       CCU_VEC_U = 6, CCU_VEC_F = 7, CCU_VEC_T0 = 8, CCU_VEC_T1 = 9, CCU_VEC_T2 = 10, CCU_VEC_COUNT = 11 };
int ccu_vec_upload(ccu_ctx *ctx, int lev, int vec, const double *host /*[neq]*/);
int ccu_vec_download(ccu_ctx *ctx, int lev, int vec, double *host /*[neq]*/);
int ccu_pvec_upload(ccu_ctx *ctx, const double *host /*[npno]*/);    /* pressure P at levmax */
int ccu_pvec_download(ccu_ctx *ctx, double *host /*[npno]*/);
int ccu_dev_matvec(ccu_ctx *ctx, int lev, int vec_u, int vec_Au, int strip_bcs);
int ccu_dev_gauss_seidel(ccu_ctx *ctx, int lev, int vec_d0, int vec_F, int vec_Ad, int cycles, int guess);
int ccu_dev_relax_sweeps(ccu_ctx *ctx, int lev, int vec_d0, int vec_F, int cycles);   /* sweeps only, no Ad */
int ccu_dev_multi_grid(ccu_ctx *ctx, int vec_d1, int vec_F, double *residual_out);
/* whole Stokes solve on resident CCU_VEC_U (in/out), CCU_VEC_F (in) and the resident pressure */
int ccu_dev_solve_Ahat_p_fhat(ccu_ctx *ctx, double imp, int *steps_max, float *residual_out, double *hist);


/* ---- one timestep's Stokes work: general_stokes_solver (Drive_solvers.c:45-162), Newtonian rheologies ----
 * velocities_conform_bcs + assemble_forces + [get_system_viscosity + construct_stiffness_B_matrix] + solve_Ahat_p_fhat
 * + (v_from_vector is the caller's: U is returned in the reference's equation numbering).
 *   T, buoyancy : host float[nno] (= E->T+1, E->buoyancy+1) or NULL to use the device-resident fields
 *   rebuild     : 1 = viscosity + stiffness are rebuilt (the reference does so on the first call and then every
 *                 `update_every_steps` when VISC_UPDATE is on, Construct_arrays.c:849), 0 = keep the resident operator
 *   guess       : 0 = U = P = 0; 1 = U, P hold the initial guess on entry (E->U, E->P + 1); 2 = resident U, P of the last solve
 *   U, P        : host double[neq] / double[npno]; out (and in when guess == 1); NULL leaves the result on the device only */
int ccu_general_stokes_solver(ccu_ctx *ctx, const float *T, const float *buoyancy, int rebuild, int augmented_Lagr,
                              double augmented, int precondition, int guess, double *U, double *P,
                              int *iterations_out, float *residual_out);

/* ---- energy step: PG_timestep (Advection_diffusion.c:251-349), Cartesian, fixed-temperature or zero-flux walls ----
 * All fields are float[nno] in the reference's node order (E->T+1, E->Tdot+1, E->V[d]+1, E->buoyancy+1). */
/* advection_diffusion_parameters (Advection_diffusion.c:63-110): E->advection.{fine_tune_dt, fixed_timestep, gamma,
 * temp_iterations}; E->diffusivity+1, E->expansivity+1 (float[noz], Instructions.c:1096-1120); E->control.Q0 */
int ccu_set_energy_params(ccu_ctx *ctx, float fine_tune_dt, float fixed_timestep, float gamma, int temp_iterations,
                          const float *diffusivity /*[noz]*/, const float *expansivity /*[noz]*/, float Q0);
int ccu_set_tdot(ccu_ctx *ctx, const float *Tdot /*[nno] or NULL = zero*/);
/* extended-Boussinesq heating: E->control.{adi_heating, visc_heating, Atemp}, E->data.{disptn_number, surf_temp} */
int ccu_set_heating_params(ccu_ctx *ctx, int adi_heating, int visc_heating, float disptn_number, float surf_temp, float Atemp);
/* process_heating (Advection_diffusion.c:813; strain_rate_2_inv Viscosity_structures.c:979), Cartesian, no phase changes:
 * element heating terms from the resident T, velocity and EVI[levmax]; they enter pg_solver's element residuals
 * (Advection_diffusion.c:643-647).  heating_*_out: float[nel] or NULL */
int ccu_process_heating(ccu_ctx *ctx, float *heating_adi_out, float *heating_visc_out);
/* phase changes (Phase_change.c:43): E->viscosity.{zlm, z410}, E->control.{Ra_670, clapeyron670, width670, Ra_410, clapeyron410,
 * width410} as the reference holds them AFTER its first phase_change call (it rescales them once, in place, :51-67).
 * ccu_set_step = E->monitor.solution_cycles (transition temperatures are re-read every 10th step, :82).  With phase changes
 * configured, ccu_thermal_buoyancy evaluates the phase functions and subtracts Ra*Fas (Pan_problem_misc_functions.c:129), and
 * ccu_process_heating adds the latent-heating terms (Advection_diffusion.c:889-946).  Cartesian, one subdomain. */
int ccu_set_phase_params(ccu_ctx *ctx, float zlm, float z410, float Ra_670, float clapeyron670, float width670,
                         float Ra_410, float clapeyron410, float width410);
int ccu_set_step(ccu_ctx *ctx, int solution_cycles);
int ccu_phase_change(ccu_ctx *ctx, int update_transT, float *Fas670_out /*[nno] or NULL*/, float *Fas410_out, float *transT_out /*[2] or NULL*/);
int ccu_get_heating_latent(ccu_ctx *ctx, float *heating_latent_out /*[nel]*/);
/* upload E->heating_adi+1, E->heating_visc+1, E->heating_latent+1 (float[nel], any may be NULL) when the reference's own
 * process_heating computed them on the host (phase-change latent heating is not evaluated on the device) */
int ccu_set_heating_arrays(ccu_ctx *ctx, const float *heating_adi, const float *heating_visc, const float *heating_latent);
int ccu_set_velocity(ccu_ctx *ctx, const float *V1, const float *V2, const float *V3 /*[nno] each*/);
/* v_from_vector (Stokes_flow_Incomp.c:530): fp32 nodal velocity from the resident solution U; V_out = float[3*nno] or NULL */
int ccu_v_from_vector(ccu_ctx *ctx, float *V_out);
/* std_timestep (Advection_diffusion.c:737-810) from the resident velocity */
int ccu_std_timestep(ccu_ctx *ctx, float *dt_out);
/* pg_solver (Advection_diffusion.c:398-441; pg_shape_fn :448, element_residual :564) on the resident T, Tdot, V */
int ccu_pg_solver(ccu_ctx *ctx, float *DTdot_out /*[nno]*/);
/* PG_timestep: std_timestep, predictor, temp_iterations x (pg_solver, corrector), the Tmax safeguard (restore, halve dt,
 * up to 5 times).  T, Tdot: host in/out, or NULL to work on the resident fields only. */
int ccu_PG_timestep(ccu_ctx *ctx, float *T, float *Tdot, float *dt_out, float *T_interior_out);
/* thermal_buoyancy (Pan_problem_misc_functions.c:83), thermal only: Atemp * T * expansivity[z] minus its layer average
 * (remove_horiz_ave / return_horiz_ave, Global_operations.c:55,133); updates the resident buoyancy; host copy optional */
int ccu_thermal_buoyancy(ccu_ctx *ctx, float Atemp, float *buoyancy_out /*[nno] or NULL*/);
int ccu_get_temperature(ccu_ctx *ctx, float *T /*[nno]*/, float *Tdot /*[nno] or NULL*/);
/* heat_flux (Process_buoyancy.c:63-203): the Nusselt numbers E->slice.Nut / Nub from the resident temperature and velocity
 * (element heat transport, nodal projection through TWW / Mass, linear extrapolation to the top and bottom surfaces,
 * area-weighted means); across subdomains the nodal sums and the four surface sums are reduced over NCCL */
int ccu_heat_flux(ccu_ctx *ctx, float *Nut_out, float *Nub_out);
/* get_stress (Topo_gravity.c:352) and get_STD_topo (:307) from the resident velocity, viscosity and pressure: nodal stresses S_out =
 * float[6][nno] in the order of get_stress's arguments SXX, SXY, SXZ, SYY, SZY, SZZ; dynamic topography at the top (tpg) and the bottom
 * (tpgb), float[nox*noy] in surf_node order.  Any output may be NULL.  Cartesian. */
int ccu_get_stress_topo(ccu_ctx *ctx, float *S_out, float *tpg_out, float *tpgb_out);
/* ---- output staging (Output.c:54-170): ccu_output_stage enqueues the device->host copies of T, V (C with markers, the nodal heat flux of the
 * last ccu_heat_flux) into pinned buffers and returns at once; ccu_output_write hands them to a background thread that waits for the copies and
 * writes <prefix>.velo.<me>.<n> and <prefix>.temp.<me>.<n> in the reference's ASCII formats while the solver continues; ccu_output_wait joins it. */
int ccu_output_stage(ccu_ctx *ctx);
int ccu_output_write(ccu_ctx *ctx, const char *prefix, int me, int file_number, int timesteps, double elapsed_time, int composition);
int ccu_output_wait(ccu_ctx *ctx);
/* averages (Process_velocity.c:179): horizontal averages per z layer -- E->Have.vrms (sqrt of the layer mean of |V|^2), E->Have.Vi (nodal
 * viscosity), E->Have.C (composition) -- float[noz] each, any may be NULL; summed over the ranks of a horizontal plane */
int ccu_averages(ccu_ctx *ctx, float *vrms_out, float *visc_out, float *C_out);

/* ---- markers of the compositional field (Composition_adv.c), Cartesian, one subdomain ----
 * E->advection.markers / markers_uplimit / markers_per_ele, E->lmesh.rnoz, E->XP[d]+1 (double[nox|noy|noz]), E->RG[3]
 * (int[rnoz+1], Nodal_mesh.c:284 pre_interpolation), E->XG1[1..3], E->XG2[1..3], E->Element+1 (SIDEE = 0x800000),
 * E->control.Acomp */
int ccu_markers_setup(ccu_ctx *ctx, int capacity, int markers_per_ele, int rnoz, const double *XP1, const double *XP2,
                      const double *XP3, const int *RG3, const double *XG1, const double *XG2, const unsigned *Element, float Acomp);
/* E->XMC[d]+1, E->C12+1, E->CElement+1 (1-based element numbers as in the reference), E->CE+1 */
int ccu_markers_upload(ccu_ctx *ctx, int n, const double *XMC1, const double *XMC2, const double *XMC3, const int *C12,
                       const int *CElement, const float *CE);
/* any pointer may be NULL; X / Xpred are double[3*n] (component-major), VO / Vpred float[3*n], C float[nno], CE float[nel] */
int ccu_markers_download(ccu_ctx *ctx, double *X, double *Xpred, float *VO, float *Vpred, int *CElement, float *C, float *CE);
/* Euler (Composition_adv.c:108): velocity_markers (:990, get_element :1086) at XMC, XMCpred = XMC + dt*VO,
 * transfer_marker_properties (:682): boundary clamp, element_markers (:960), get_C_from_markers (:709, ratio method) */
int ccu_Euler(ccu_ctx *ctx, float timestep);
/* Runge_Kutta (Composition_adv.c:61): velocity at XMCpred, XMC += dt/2 (VO + Vpred), transfer_marker_properties */
int ccu_Runge_Kutta(ccu_ctx *ctx, float timestep);
/* Subdomain-per-GPU runs (after ccu_comm_init): ccu_Euler / ccu_Runge_Kutta also move the markers that left the subdomain
 * to the neighbour that now holds them (transfer_markers_processors, Composition_adv.c:148: locate_processor :628 on the
 * markers of side elements, counts exchanged with one all-gather, records {XMC, XMCpred, VO, C12} in one grouped NCCL
 * send/recv round) and sum the nodal composition across subdomains (exchange_node_f20, :797).
 * The same step in two halves, with the migrating records handed over on the host -- two subdomains in ONE process for
 * tests and for bindings that keep the reference's MPI for the markers: records are 8 doubles per marker
 * {XMC[3], XMCpred[3], (float VO0, VO1), (float VO2, int C12)}, grouped by neighbour code (ox+1) + 3 (oy+1) + 9 (oz+1),
 * sendcnt[code] of them each. */
int ccu_markers_set_decomp(ccu_ctx *ctx, const int nproc[3], const int me[3]);
int ccu_markers_step_export(ccu_ctx *ctx, float timestep, int corrector, int sendcnt[27], double *records_out, int max_records);
int ccu_markers_import_finish(ccu_ctx *ctx, int corrector, int nrecv, const double *records);
int ccu_markers_count(ccu_ctx *ctx);
/* host-only routing table of the marker exchange (no GPU needed): neighbour rank and arriving record count per direction
 * code for subdomain `me`, from the gathered send counts all_counts[rank][27] of all ranks */
int ccu_marker_routes(const int nproc[3], const int me[3], const int *all_counts, int nb_rank[27], int recvcnt[27]);

/* ---- CUDA-event timing of the finest-level kernels inside a solve (bench.py's live roofline) ---- */
/* classes: finest-level smoother (units = colour-pass launches), finest-level matvec / residual (units = products),
 * operator rebuild, everything below the finest level (units = graph segments), finest-level transfers (project / interp) */
enum { CCU_PROF_RELAX_FINE = 0, CCU_PROF_MATVEC_FINE = 1, CCU_PROF_BUILD = 2, CCU_PROF_COARSE = 3, CCU_PROF_TRANSFER_FINE = 4,
       CCU_PROF_LEVEL0 = 5 /* + lev: smoother + matvec + transfers of multigrid level lev; only recorded with CCU_OPT_GRAPHS = 0 */,
       CCU_PROF_FACES_FINE = 5 + CCU_MAX_LEVELS /* subdomain-per-GPU runs: the duplicated-node part of the finest-level sweeps (partial rows,
        * exchange, Jacobi update), a subset of RELAX_FINE */,
       CCU_PROF_COUNT = 6 + CCU_MAX_LEVELS };
int ccu_profile_enable(ccu_ctx *ctx, int on);
/* synchronises; total milliseconds and units (see above) recorded for class `cls` since the last reset */
int ccu_profile_read(ccu_ctx *ctx, int cls, double *ms_total, long long *launches);
int ccu_profile_reset(ccu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
