// ccu_ctx.cuh -- device context shared by the translation units of libcitcomcu_b200.so
#pragma once
#include "../../include/citcomcu_b200.h"
#include "ccu_layout.cuh"
#include <cuda_runtime.h>
#include <string>
#include <vector>

extern thread_local std::string g_ccu_err;

#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { \
    g_ccu_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + std::to_string(__LINE__); return 1; } } while(0)
#define FAIL(msg) do { g_ccu_err = (msg); return 2; } while(0)

enum { S_DOT0 = 0, S_DOT1, S_DOT2, S_R1Z1, S_R0Z0, S_S2AH, S_VDOTV, S_PDOTP, S_AHAH, S_S2S2, S_U1U1, S_TMP, S_ONE, S_COUNT = 32 };

// viscosity law parameters (E->viscosity.*, Viscosity_structures.c:57-326)
struct CcuViscParams
{
    int rheol = 0, tdepv = 0, num_mat = 1, vmin = 0, vmax = 0, smooth_cycles = 1;
    float N0[40] = { 1.0f }, E[40] = { 0 }, T[40] = { 0 }, Z[40] = { 0 };
    float min_value = 0, max_value = 0;
    // stress-dependent viscosity (visc_from_S, Viscosity_structures.c:744; the outer loop of Drive_solvers.c:120-159)
    int sdepv = 0, sdepv_rheology = 1, sdepv_start_from_newtonian = 0, sdepv_max_iter = 50, sdepv_visits = 0;
    float sdepv_expt[40] = { 1.0f }, sdepv_trns[40] = { 1.0f }, sdepv_misfit = 0.001f, sdepv_iter_damp = 1.0f, sdepv_trns_T = 0, sdepv_trns_c = 0;
    // composition-dependent viscosity (visc_from_C, Viscosity_structures.c:1784-1935; the plain prefactor mode and cdepv_absolute)
    int cdepv = 0, cdepv_layer = 0, cdepv_absolute = 0, cdepv_check_range = 0;
    double cdepv_logv[80] = { 0 };     // log(pre_comp[2 l]), log(pre_comp[2 l + 1]) per material layer (or one pair)
    // Byerlee-type plastic yielding (visc_from_B, Viscosity_structures.c:1470-1755, the regular branch without flavours / strain weakening)
    int bdepv = 0, bdepv_dimensional = 0, bdepv_trans = 0, bdepv_visits = 0;
    float abyerlee[40] = { 0 }, bbyerlee[40] = { 0 }, lbyerlee[40] = { 0 }, bdepv_offset = 0, bdepv_ndz_to_m = 1, bdepv_tau_scale = 1;
};

struct Level
{
    CcuGeom g;
    float *K = nullptr;
    float *KT = nullptr;          // [117][NS] the thirteen upper-neighbour blocks transposed to the node's own slot (full rows: ccu_k_relax_full); null: none
    bool have_KT = false;
    unsigned char *Kc = nullptr;  // column-major copy of K, BI, flags for the column kernels (ccu_col.cuh); col_shape < 0: none
    size_t *colofs = nullptr;     // [col_nI * col_nJ] byte offset of each column's first chunk
    size_t Kc_bytes = 0; int col_shape = -1, col_nI = 0, col_nJ = 0;
    int col_inv_key = 0;
    short *col_inv = nullptr;     // [nbp of the full column shape] chunk position -> block id (re-layout kernel)
    unsigned *col_sync = nullptr; // [4 + col_nI * col_nJ] ticket counter, finished CTAs, epoch, pad, progress word per column (one-launch sweeps)
    double *BI = nullptr;
    unsigned char *flags = nullptr;
    float *MASS = nullptr, *TWW = nullptr, *eco = nullptr, *elt_del = nullptr;
    float *ecoT = nullptr;        // [3][nel] the element sizes direction-major (interp_vector weights); refreshed by ccu_eco_changed
    float *elt_delT = nullptr;    // [24][nel] coefficient-major copy of elt_del (div_u / grad_p read it)
    double *BPI = nullptr;
    double *vec[CCU_VEC_COUNT] = { nullptr };
    // operator construction (ccu_build.cu)
    float *XX = nullptr;          // [3][nno] node coordinates, natural order (E->XX[lev][1..3]); Cartesian also for Rsphere
    float *SXX = nullptr;         // [3][nno] (theta, phi, r) of the nodes, Rsphere only (E->SXX[lev][1..3])
    bool have_sxx = false;
    float *EVI = nullptr;         // [nel*8] viscosity at Gauss points
    unsigned *node = nullptr;     // [nno] raw NODE flags, natural order
    // shared-memory resident bottom smoother (ccu_k_relax_smem): compact tables, built on first use
    int sm_n = 0, sm_cstart[9] = { 0 }; int *sm_s = nullptr; unsigned short *sm_nbr = nullptr;
    bool have_K = false, have_flags = false, have_tw = false, have_p = false, have_xx = false, have_evi = false;
    size_t vlen() const { return 3 * (size_t)g.NS; }
};

// duplicated-node tables of one level on the device (ccu_comm.cu)
struct CcuHalo
{
    std::vector<int> nb_rank, nb_off, nb_cnt;
    int n_send = 0, n_shared = 0;
    int *sh_s = nullptr, *sh_n = nullptr, *sh_ptr = nullptr, *sh_src = nullptr;   // per duplicated node: storage / natural index, CSR of contributions
    int *send_s = nullptr, *send_n = nullptr, *send_t = nullptr;                     // per packed node: storage / natural / compact index
    unsigned char *bits = nullptr;   // [NS] bit0 = owned for dot products, bit1 = duplicated (OFFSIDE)
    double *face = nullptr;          // [3][n_shared] partial rows of the duplicated nodes (smoother)
    int4 *p2p_seg = nullptr;         // [segments] {neighbour rank, my send/recv offset, node count, the neighbour's offset for my data} (peer-memory exchange)
};
struct CcuComm
{
    int nranks = 1, rank = 0, nproc[3] = { 1, 1, 1 }, me[3] = { 0, 0, 0 };
    void *nccl = nullptr;            // ncclComm_t
    CcuHalo halo[CCU_MAX_LEVELS];
    void *sendbuf = nullptr, *recvbuf = nullptr;
    // peer-memory halo exchange over NVLink (ccu_comm.cu): every rank exports two landing buffers and a flag word per sender
    // through CUDA IPC; a push kernel stores the packed faces straight into the neighbours' landing buffers and raises their
    // flags, a wait kernel on the receiving side spins on its flags and moves the data into recvbuf.  p2p = false: NCCL send/recv.
    bool p2p = false;
    int opt_p2p = 0;                 // measured on 8 x B200 (r02): 0.265 s/step against 0.245 with NCCL's fused send/recv kernel -> off by default
    size_t land_half = 0;            // bytes of one landing buffer
    char *land = nullptr;            // [2][land_half] own landing buffers (parity of the exchange sequence number)
    unsigned *flags = nullptr;       // [nranks] sequence number of the last exchange rank r has delivered
    unsigned *seq = nullptr;         // device: exchange sequence number (bumped by a kernel: graph replay safe)
    unsigned *done = nullptr;        // [32] blocks finished per segment (push kernel)
    unsigned *p2p_err = nullptr;     // set when a wait gives up
    char **peer_land = nullptr;      // device [nranks]: the ranks' landing buffers as mapped here (null: not a neighbour)
    unsigned **peer_flags = nullptr; // device [nranks]
    std::vector<void *> ipc_open;    // mapped peer allocations (closed at destroy)
    double *dotstage = nullptr;
    int *mk_counts = nullptr;        // [27] own + [nranks][27] gathered per-direction counts of migrating markers
    long long gneq = 0, gnpno = 0;   // global equation / pressure counts (E->mesh.neq, E->mesh.npno)
};

struct CcuOutput;
struct ccu_ctx
{
    ccu_config cfg;
    CcuOutput *out = nullptr;          // output staging (ccu_output.cu)
    cudaStream_t st = 0, own_stream = 0;
    // second stream for the duplicated-node exchange of a smoother sweep while the far shells relax on `st` (d_relax_sweeps)
    cudaStream_t st2 = 0; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int opt_halo_overlap = 0;      // measured on 2 x B200 (r02): 0.514 s/step with the overlap against 0.435 without -> off by default
    bool bottom_attr_set = false;
    double *sdepv_oldU = nullptr, *sdepv_dU = nullptr; double sdepv_last_misfit = 0.0; int sdepv_last_count = 0;     // SDEPV outer loop (ccu_general_stokes_solver)
    int launch_err = 0; const char *launch_err_kernel = "";     // first failed kernel launch since the last check (LAUNCH / CK_LAUNCHES)
    struct GraphSeg { cudaGraphExec_t exec = nullptr; long long launches = 0; };
    GraphSeg seg[4];
    bool use_graphs = true;
    // kernel selection by level size (lanes per node), ccu_set_option
    int opt_small_nodes = 600, opt_warp_nodes = 30000, opt_quad_nodes = 500000, opt_lanes_large = 1;
    int opt_matvec_tab = 4, opt_relax_tab = 2, opt_smem_nodes = 434, opt_matvec_tab_nodes = 10000, opt_bottom_cluster = 1;   // table-driven row kernels on the large levels (ccu_kernels.cuh)
    // column-resident smoother / matvec (ccu_col.cuh) on levels above opt_col_nodes nodes
    int opt_col_nodes = 2000000, opt_relax_col = 0, opt_matvec_col = 1, opt_col_shape = 0, opt_col_wf = 1;
    int opt_full_nodes = 500000, opt_relax_full = 0, opt_matvec_full = 0;   // full-row copy of K on levels above opt_full_nodes nodes
    Level L[CCU_MAX_LEVELS];
    double *scal = nullptr;        // device scalars
    double *partial = nullptr;     // dot partials
    void *stage = nullptr;         // upload/download staging
    size_t stage_bytes = 0;
    // finest-level Uzawa work space (solve_Ahat_p_fhat's statics, Stokes_flow_Incomp.c:325-336)
    double *uzAh = nullptr, *uzU1 = nullptr;
    double *P = nullptr, *r0 = nullptr, *r1 = nullptr, *r2 = nullptr, *z0 = nullptr, *z1 = nullptr, *s1 = nullptr, *s2 = nullptr, *pAh = nullptr;
    // operator construction state
    CcuViscParams visc;
    int *mat = nullptr;            // [nel] material group per element (E->mat), finest level
    float *T = nullptr;            // [nno] temperature, natural order, finest level
    float *buoy = nullptr;         // [nno]
    float *nodal_tmp = nullptr, *nodal_tmp2 = nullptr;   // [nno finest] scratch for project_viscosity
    double *forceEF = nullptr;     // [8][nel] element force contributions (assemble_forces); [24][nel] for Rsphere
    float *Cnode = nullptr;        // [nno] nodal composition handed in by the host (ccu_set_composition) when no marker set lives on the device
    bool rsphere = false;          // regional-spherical geometry (ccu_set_spherical_coordinates)
    // imposed non-zero boundary velocities (E->VB): the K.VB term of get_elt_f and velocities_conform_bcs
    float *VB[3] = { nullptr, nullptr, nullptr };   // [nno] natural order, finest level
    bool have_vb = false, vb_dirty = false;
    int *vb_slot = nullptr;        // [nel] slot of the element in vbEF, -1 = no flagged node with a non-zero VB
    int *vb_elems = nullptr;       // [n_vb] the elements that have one
    int n_vb = 0;
    double *vbEF = nullptr;        // [24][n_vb] their -K.VB contributions
    double *eltK = nullptr;        // element-block scratch for the stiffness build
    size_t eltK_elems = 0;
    // energy step state (ccu_build_exact.cu, PG_timestep): finest level, natural node order
    struct Energy
    {
        float *Tdot = nullptr, *DTdot = nullptr, *V = nullptr;     // [nno], [nno], [3][nno]
        float *T1 = nullptr, *Tdot1 = nullptr;                     // saved fields of the Tmax safeguard
        float *diffusivity = nullptr, *expansivity = nullptr;      // [noz]
        float *heat_adi = nullptr, *heat_visc = nullptr, *heat_latent = nullptr;   // [nel] process_heating; null = 0, 0, 1
        int adi_heating = 0, visc_heating = 0; float disptn = 0.0f, surf_temp = 0.0f, Atemp_heat = 1.0f;
        // phase changes (Phase_change.c): rescaled parameters, nodal phase functions, transition temperatures {670, 410}
        struct { float zlm = 0, z410 = 0, Ra670 = 0, clap670 = 0, width670 = 0, Ra410 = 0, clap410 = 0, width410 = 0; } ph;
        bool phase_on = false; float *Fas670 = nullptr, *Fas410 = nullptr, *transT = nullptr;
        int step = 0;                                              // E->monitor.solution_cycles (ccu_set_step)
        double *Eres = nullptr;                                    // [nel][8] element residuals
        double *layer = nullptr;                                   // [2][noz] layer sums of remove_horiz_ave
        float *hf = nullptr, *hf_area = nullptr; double *hf_sums = nullptr;   // heat_flux diagnostics
        double *layer_tab = nullptr;                               // [nprocz][2][noz] allreduce table of multi-subdomain runs
        double *transT_tab = nullptr;      // [2 nprocx nprocy] transition temperatures summed over the z subdomains of a column of ranks
        float *red = nullptr;                                      // device scalars: [0] min, [1] max
        float fine_tune_dt = 0.9f, fixed_timestep = 0.0f, gamma = 0.5f, Q0 = 0.0f, diff_timestep = -1.0f;
        int temp_iterations = 2;
        bool have_params = false, have_v = false;
    } en;
    // marker (tracer) state of the compositional field (ccu_build_exact.cu; Composition_adv.c), single subdomain
    struct Markers
    {
        int n = 0, cap = 0, markers_per_ele = 0, rnoz = 0;
        double *X = nullptr, *Xpred = nullptr;          // [3][cap] XMC, XMCpred
        float *VO = nullptr, *Vpred = nullptr;           // [3][cap]
        int *C12 = nullptr, *CElement = nullptr;         // [cap]
        int *count = nullptr;                            // [2][nel] regular / dense markers per element
        float *CE = nullptr, *C = nullptr;               // [nel], [nno]
        double *XP = nullptr;                            // XP[1] | XP[2] | XP[3]: 1-D node coordinates (nox + noy + noz)
        int *RG3 = nullptr;                              // [rnoz+1] z lookup table
        unsigned *Element = nullptr;                     // [nel] element flags (SIDEE)
        int *err = nullptr;                              // device error counter (markers that fell out of the z table)
        double XG1[3] = { 0, 0, 0 }, XG2[3] = { 0, 0, 0 };
        float Acomp = 0.0f;
        // markers changing subdomain (transfer_markers_processors): decomposition, classification / partition scratch,
        // second set of marker arrays (the stayers are gathered into it and the sets swapped), record buffers
        int np[3] = { 1, 1, 1 }, me[3] = { 0, 0, 0 };
        unsigned char *code = nullptr, *stayf = nullptr; int *perm = nullptr, *lv_idx = nullptr, *lv_code = nullptr, *nsel = nullptr;
        void *cub_tmp = nullptr; size_t cub_bytes = 0;
        double *sX = nullptr, *sXpred = nullptr; float *sVO = nullptr, *sVpred = nullptr; int *sC12 = nullptr, *sCElement = nullptr;
        double *sendbuf = nullptr, *recvbuf = nullptr;
        bool ready = false;
    } mk;
    long long launches = 0;
    CcuComm *comm = nullptr;       // null = single subdomain
    // Replicated coarse levels (multi-subdomain runs): levels <= agg_lev of the multigrid hierarchy live in `coarse`, a
    // single-subdomain context holding the GLOBAL mesh of those levels on every GPU.  One all-gather hands the
    // restricted right-hand side to all ranks, every rank runs the identical coarse part of the cycle without any
    // further communication, and takes its own piece of the correction back (ccu_stokes.cu, mg_inner).
    ccu_ctx *coarse = nullptr;
    int agg_lev = -1;
    bool replica = false;          // this context IS such a coarse replica (its top level is not the finest level)
    void *agg_buf = nullptr;       // all-gather landing zone
    size_t agg_bytes = 0;
    bool multi() const { return comm && comm->nranks > 1; }
    // CUDA-event profiling of kernel classes (ccu_profile_*)
    bool prof_on = false;
    struct ProfRec { int cls; cudaEvent_t e0, e1; long long n; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[CCU_PROF_COUNT] = { 0 };
    long long prof_n[CCU_PROF_COUNT] = { 0 };
};

// scoped event pair around a group of launches of one class
struct CcuProfScope
{
    ccu_ctx *c; int cls; long long units; cudaEvent_t e0 = nullptr;
    // `units` = what the class counts per scope (colour-pass launches of the smoother, products of the matvec, ...)
    CcuProfScope(ccu_ctx *ctx, int cls_, bool active, long long units_ = 1) : c(active && ctx->prof_on ? ctx : nullptr), cls(cls_), units(units_)
    {
        if(!c) return;
        e0 = get(); cudaEventRecord(e0, c->st);
    }
    ~CcuProfScope()
    {
        if(!c) return;
        cudaEvent_t e1 = get(); cudaEventRecord(e1, c->st);
        c->prof_recs.push_back({ cls, e0, e1, units });
    }
    cudaEvent_t get()
    {
        if(!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};

// A launch over zero items is skipped (a grid of 0 is an invalid configuration); a launch that fails is remembered with the
// kernel's name and reported by the next CK_LAUNCHES (every synchronising entry point), not at some unrelated later call.
#define LAUNCH(ctx, kern, grid, block, ...) do { const dim3 g_(grid); if(g_.x && g_.y && g_.z) { \
    kern<<<g_, (block), 0, (ctx)->st>>>(__VA_ARGS__); (ctx)->launches++; \
    if(!(ctx)->launch_err) { const cudaError_t le_ = cudaPeekAtLastError(); if(le_ != cudaSuccess) { (ctx)->launch_err = (int)le_; (ctx)->launch_err_kernel = #kern; } } } } while(0)
#define SYNC(ctx) do { CK(cudaStreamSynchronize((ctx)->st)); CK_LAUNCHES(ctx); } while(0)
#define CK_LAUNCHES(ctx) do { if((ctx)->launch_err) { g_ccu_err = std::string("kernel launch failed: ") + (ctx)->launch_err_kernel + ": " + \
    cudaGetErrorString((cudaError_t)(ctx)->launch_err); (ctx)->launch_err = 0; cudaGetLastError(); return 1; } } while(0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
int ccu_ensure_stage(ccu_ctx *c, size_t bytes);
void ccu_output_destroy(ccu_ctx *c);      // ccu_output.cu
void ccu_drop_graphs(ccu_ctx *c);
// ccu_comm.cu
void ccu_comm_destroy(ccu_ctx *c);
int ccu_halo_sum_vec(ccu_ctx *c, int lev, double *vec);       // exchange_id_d20
int ccu_halo_sum_face(ccu_ctx *c, int lev);                   // partial rows of the duplicated nodes -> recvbuf
int ccu_halo_sum_nodal(ccu_ctx *c, int lev, float *field);    // exchange_node_f20
int ccu_allreduce_dots(ccu_ctx *c, int count, double *o0, double *o1, double *o2);
int ccu_allreduce_buffer(ccu_ctx *c, double *buf, int count, int op_max);
int ccu_damp_face_BI(ccu_ctx *c, int lev);
int ccu_allgather(ccu_ctx *c, const void *send, void *recv, size_t bytes_per_rank);
// variable-size exchange of records (`rec` doubles each, grouped by neighbour code in sendbuf) with the up to 26 neighbours
int ccu_marker_exchange(ccu_ctx *c, const int sendcnt[27], const double *sendbuf, int rec, int recvcnt[27], double *recvbuf, size_t cap_records, int *nrecv, int n_resident);
int ccu_agg_gather_evi(ccu_ctx *c);                          // EVI[agg_lev] of all subdomains -> coarse replica (ccu_stokes.cu)                  // rebuild_BI_on_boundary (ccu_stokes.cu)
int ccu_check_lev(ccu_ctx *c, int lev);
int ccu_col_refresh(ccu_ctx *c, int lev);
int ccu_col_refresh_all(ccu_ctx *c);
void ccu_fork_stream(ccu_ctx *c);                             // ccu_stokes.cu: st2 starts after what is queued on st
void ccu_join_stream(ccu_ctx *c);                             //                st continues after what is queued on st2
void ccu_eco_changed(ccu_ctx *c, int lev);                    // ccu_stokes.cu: call after every write of L.eco
void ccu_elt_del_changed(ccu_ctx *c, int lev);                // ccu_stokes.cu                   // ccu_stokes.cu
