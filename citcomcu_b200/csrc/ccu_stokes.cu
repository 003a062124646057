// ccu_stokes.cu -- host side of libcitcomcu_b200.so: device context, operator upload, and the
// solver drivers (gauss_seidel, multi_grid, solve_del2_u, solve_Ahat_p_fhat) re-stated around the
// kernels of ccu_kernels.cuh.  Scalars of the iteration (dot products, line-search alpha, CG alpha
// and delta) stay in device memory; the host reads back only the convergence monitors the
// reference's control flow branches on.
#include "ccu_ctx.cuh"
#include "ccu_kernels.cuh"
#include "ccu_col.cuh"
#include "ccu_comm.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

thread_local std::string g_ccu_err;
const char *ccu_last_error(void) { return g_ccu_err.c_str(); }


int ccu_ensure_stage(ccu_ctx *c, size_t bytes)
{
    if(bytes <= c->stage_bytes) return 0;
    if(c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0;
    CK(cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return 0;
}

int ccu_check_lev(ccu_ctx *c, int lev)
{
    if(!c) FAIL("null context");
    if(lev < c->cfg.levmin || lev > c->cfg.levmax) FAIL("level out of range");
    return 0;
}

static int ensure_smem_tables(ccu_ctx *c, Level &L);

int ccu_create(const ccu_config *cfg, ccu_ctx **out)
{
    if(!cfg || !out) FAIL("ccu_create: null argument");
    if(cfg->levmin < 0 || cfg->levmax >= CCU_MAX_LEVELS || cfg->levmin > cfg->levmax) FAIL("ccu_create: bad level range");
    int ndev = 0;
    if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL("ccu_create: no CUDA device (this library has no CPU fallback)");
    CK(cudaSetDevice(cfg->device));
    ccu_ctx *c = new ccu_ctx();
    c->cfg = *cfg;
    CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));   // own stream: the legacy stream cannot be captured
    c->own_stream = c->st;
    {   // lowest priority: the colour passes queued here must not starve the short exchange kernels on the context's stream
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, lo));
    }
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for(int lev = cfg->levmin; lev <= cfg->levmax; lev++)
    {
        Level &L = c->L[lev];
        L.g = ccu_make_geom(cfg->nox[lev], cfg->noy[lev], cfg->noz[lev]);
        if(lev > cfg->levmin)
        {
            const CcuGeom &gc = c->L[lev - 1].g;
            if(L.g.elx != 2 * gc.elx || L.g.ely != 2 * gc.ely || L.g.elz != 2 * gc.elz) { delete c; FAIL("ccu_create: levels must double"); }
        }
        const size_t NS = (size_t)L.g.NS;
        CK(cudaMalloc(&L.K, sizeof(float) * 14 * 9 * NS));
        CK(cudaMemsetAsync(L.K, 0, sizeof(float) * 14 * 9 * NS, c->st));
        CK(cudaMalloc(&L.BI, sizeof(double) * 3 * NS));
        CK(cudaMemsetAsync(L.BI, 0, sizeof(double) * 3 * NS, c->st));
        CK(cudaMalloc(&L.flags, NS));
        CK(cudaMemsetAsync(L.flags, 0, NS, c->st));
        CK(cudaMalloc(&L.MASS, sizeof(float) * L.g.nno));
        CK(cudaMalloc(&L.TWW, sizeof(float) * 8 * (size_t)L.g.nel));
        CK(cudaMalloc(&L.eco, sizeof(float) * 3 * (size_t)L.g.nel));
        CK(cudaMalloc(&L.ecoT, sizeof(float) * 3 * (size_t)L.g.nel));
        CK(cudaMalloc(&L.elt_del, sizeof(float) * 24 * (size_t)L.g.nel));
        CK(cudaMalloc(&L.elt_delT, sizeof(float) * 24 * (size_t)L.g.nel));
        CK(cudaMalloc(&L.BPI, sizeof(double) * (size_t)L.g.npno));
        const int nvec = (lev == cfg->levmax) ? CCU_VEC_COUNT : CCU_VEC_U;   // U, F, T* only at the top level
        for(int v = 0; v < nvec; v++)
        {
            CK(cudaMalloc(&L.vec[v], sizeof(double) * 3 * NS));
            CK(cudaMemsetAsync(L.vec[v], 0, sizeof(double) * 3 * NS, c->st));
        }
    }
    {
        Level &L = c->L[cfg->levmax];
        const size_t NS = (size_t)L.g.NS, np = (size_t)L.g.npno;
        CK(cudaMalloc(&c->uzAh, sizeof(double) * 3 * NS)); CK(cudaMemsetAsync(c->uzAh, 0, sizeof(double) * 3 * NS, c->st));
        CK(cudaMalloc(&c->uzU1, sizeof(double) * 3 * NS)); CK(cudaMemsetAsync(c->uzU1, 0, sizeof(double) * 3 * NS, c->st));
        double **pv[] = { &c->P, &c->r0, &c->r1, &c->r2, &c->z0, &c->z1, &c->s1, &c->s2, &c->pAh };
        for(auto p : pv) { CK(cudaMalloc(p, sizeof(double) * np)); CK(cudaMemsetAsync(*p, 0, sizeof(double) * np, c->st)); }
    }
    for(int lev = cfg->levmin; lev <= cfg->levmax; lev++)
        if(c->L[lev].g.nno <= 434 && ensure_smem_tables(c, c->L[lev])) return 1;
    CK(cudaMalloc(&c->scal, sizeof(double) * S_COUNT));
    CK(cudaMemsetAsync(c->scal, 0, sizeof(double) * S_COUNT, c->st));
    { const double one = 1.0; CK(cudaMemcpyAsync(c->scal + S_ONE, &one, sizeof(double), cudaMemcpyHostToDevice, c->st)); }
    CK(cudaMalloc(&c->partial, sizeof(double) * 3 * CCU_DOT_BLOCKS));
    SYNC(c);
    *out = c;
    return 0;
}

void ccu_drop_graphs(ccu_ctx *c)
{
    for(auto &g : c->seg) { if(g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.launches = 0; }
}
static void drop_graphs(ccu_ctx *c) { ccu_drop_graphs(c); }

void ccu_destroy(ccu_ctx *c)
{
    if(c) ccu_output_destroy(c);
    if(!c) return;
    if(c->coarse) { ccu_destroy(c->coarse); c->coarse = nullptr; }
    cudaFree(c->agg_buf);
    for(int lev = 0; lev < CCU_MAX_LEVELS; lev++)
    {
        Level &L = c->L[lev];
        cudaFree(L.K); cudaFree(L.KT); cudaFree(L.Kc); cudaFree(L.colofs); cudaFree(L.col_sync); cudaFree(L.col_inv); cudaFree(L.BI); cudaFree(L.flags); cudaFree(L.MASS); cudaFree(L.TWW); cudaFree(L.eco); cudaFree(L.ecoT); cudaFree(L.elt_del); cudaFree(L.elt_delT); cudaFree(L.BPI);
        cudaFree(L.XX); cudaFree(L.SXX); cudaFree(L.EVI); cudaFree(L.node); cudaFree(L.sm_s); cudaFree(L.sm_nbr);
        for(auto v : L.vec) cudaFree(v);
    }
    cudaFree(c->en.Tdot); cudaFree(c->en.DTdot); cudaFree(c->en.V); cudaFree(c->en.T1); cudaFree(c->en.Tdot1); cudaFree(c->en.diffusivity);
    cudaFree(c->en.hf); cudaFree(c->en.hf_area); cudaFree(c->en.hf_sums); cudaFree(c->en.layer_tab); cudaFree(c->en.transT_tab); cudaFree(c->en.Fas670); cudaFree(c->en.Fas410); cudaFree(c->en.transT); cudaFree(c->en.heat_adi); cudaFree(c->en.heat_visc); cudaFree(c->en.heat_latent);
    cudaFree(c->en.expansivity); cudaFree(c->en.Eres); cudaFree(c->en.layer); cudaFree(c->en.red);
    { auto &M = c->mk; cudaFree(M.X); cudaFree(M.Xpred); cudaFree(M.VO); cudaFree(M.Vpred); cudaFree(M.C12); cudaFree(M.CElement); cudaFree(M.count);
      cudaFree(M.CE); cudaFree(M.C); cudaFree(M.XP); cudaFree(M.RG3); cudaFree(M.Element); cudaFree(M.err); }
    cudaFree(c->mat); cudaFree(c->T); cudaFree(c->buoy); cudaFree(c->nodal_tmp); cudaFree(c->nodal_tmp2); cudaFree(c->eltK);
    cudaFree(c->scal); cudaFree(c->partial); cudaFree(c->stage); cudaFree(c->uzAh); cudaFree(c->uzU1);
    drop_graphs(c);
    ccu_comm_destroy(c);
    if(c->own_stream) cudaStreamDestroy(c->own_stream);
    if(c->st2) cudaStreamDestroy(c->st2);
    if(c->ev_fork) cudaEventDestroy(c->ev_fork);
    if(c->ev_join) cudaEventDestroy(c->ev_join);
    for(auto &r : c->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for(auto e : c->prof_pool) cudaEventDestroy(e);
    {   // marker migration scratch
        auto &M = c->mk;
        cudaFree(M.code); cudaFree(M.stayf); cudaFree(M.perm); cudaFree(M.lv_idx); cudaFree(M.lv_code); cudaFree(M.nsel); cudaFree(M.cub_tmp);
        cudaFree(M.sX); cudaFree(M.sXpred); cudaFree(M.sVO); cudaFree(M.sVpred); cudaFree(M.sC12); cudaFree(M.sCElement);
        cudaFree(M.sendbuf); cudaFree(M.recvbuf);
    }
    cudaFree(c->forceEF); cudaFree(c->sdepv_oldU); cudaFree(c->sdepv_dU);
    for(int d = 0; d < 3; d++) cudaFree(c->VB[d]);
    cudaFree(c->vb_slot); cudaFree(c->vb_elems); cudaFree(c->vbEF); cudaFree(c->Cnode);
    cudaFree(c->P); cudaFree(c->r0); cudaFree(c->r1); cudaFree(c->r2); cudaFree(c->z0); cudaFree(c->z1); cudaFree(c->s1); cudaFree(c->s2); cudaFree(c->pAh);
    delete c;
}

int ccu_set_stream(ccu_ctx *c, void *s)
{
    if(!c) FAIL("null context");
    SYNC(c);
    c->st = s ? (cudaStream_t)s : c->own_stream;     // NULL selects the context's own stream again
    if(c->coarse) { drop_graphs(c); c->coarse->st = c->st; }
    return 0;
}
static int col_refresh_all(ccu_ctx *c);
int ccu_set_option(ccu_ctx *c, int option, int value)
{
    if(!c) FAIL("null context");
    switch(option)
    {
    case CCU_OPT_GRAPHS: c->use_graphs = value != 0; drop_graphs(c); return 0;
    case CCU_OPT_SMALL_NODES: c->opt_small_nodes = value; drop_graphs(c); return 0;
    case CCU_OPT_WARP_NODES: c->opt_warp_nodes = value; drop_graphs(c); return 0;
    case CCU_OPT_QUAD_NODES: c->opt_quad_nodes = value; drop_graphs(c); return 0;
    case CCU_OPT_LANES_LARGE: if(value != 1 && value != 4) FAIL("lanes must be 1 or 4"); c->opt_lanes_large = value; drop_graphs(c); return 0;
    case CCU_OPT_SMEM_NODES: c->opt_smem_nodes = value > 434 ? 434 : value; drop_graphs(c); return 0;
    case CCU_OPT_MATVEC_TAB: c->opt_matvec_tab = value; drop_graphs(c); return 0;
    case CCU_OPT_RELAX_TAB: c->opt_relax_tab = value; drop_graphs(c); return 0;
    case CCU_OPT_COL_NODES: c->opt_col_nodes = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_RELAX_COL: c->opt_relax_col = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_MATVEC_COL: c->opt_matvec_col = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_FULL_NODES: c->opt_full_nodes = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_RELAX_FULL: c->opt_relax_full = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_MATVEC_FULL: c->opt_matvec_full = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_P2P_HALO: if(c->comm) c->comm->opt_p2p = value != 0; drop_graphs(c); return 0;
    case CCU_OPT_HALO_OVERLAP: c->opt_halo_overlap = value; drop_graphs(c); return 0;
    case CCU_OPT_COL_WF: c->opt_col_wf = value != 0; if(c->coarse) c->coarse->opt_col_wf = c->opt_col_wf; drop_graphs(c); return 0;
    case CCU_OPT_COL_SHAPE: if(value < 0 || value > 2) FAIL("column shape must be 0..2"); c->opt_col_shape = value; drop_graphs(c); return col_refresh_all(c);
    case CCU_OPT_BOTTOM_CLUSTER: c->opt_bottom_cluster = value; if(c->coarse) c->coarse->opt_bottom_cluster = value; drop_graphs(c); return 0;
    default: FAIL("set_option: unknown option");
    }
}
int ccu_get_sdepv_iterations(ccu_ctx *c, int *count_out, double *misfit_out)
{
    if(!c) FAIL("null context");
    if(count_out) *count_out = c->sdepv_last_count;
    if(misfit_out) *misfit_out = c->sdepv_last_misfit;
    return 0;
}
int ccu_get_option(ccu_ctx *c, int option, int lev, int *value)
{
    if(!c || !value) FAIL("get_option: null argument");
    if(lev < c->cfg.levmin || lev > c->cfg.levmax) FAIL("get_option: level out of range");
    const bool col = c->L[lev].col_shape >= 0;
    switch(option)
    {
    case CCU_OPT_GRAPHS: *value = c->use_graphs; return 0;
    case CCU_OPT_RELAX_COL: *value = col && c->opt_relax_col; return 0;
    case CCU_OPT_MATVEC_COL: *value = col && c->opt_matvec_col; return 0;
    case CCU_OPT_FULL_NODES: *value = c->opt_full_nodes; return 0;
    case CCU_OPT_RELAX_FULL: *value = c->L[lev].have_KT && c->opt_relax_full; return 0;
    case CCU_OPT_MATVEC_FULL: *value = c->L[lev].have_KT && c->opt_matvec_full; return 0;
    case CCU_OPT_P2P_HALO: *value = c->comm && c->comm->p2p && c->comm->opt_p2p; return 0;
    case CCU_OPT_HALO_OVERLAP: *value = c->opt_halo_overlap; return 0;
    case CCU_OPT_COL_WF: *value = c->opt_col_wf; return 0;
    case CCU_OPT_COL_SHAPE: *value = c->L[lev].col_shape; return 0;
    case CCU_OPT_COL_NODES: *value = c->opt_col_nodes; return 0;
    case CCU_OPT_RELAX_TAB: *value = c->opt_relax_tab; return 0;
    case CCU_OPT_MATVEC_TAB: *value = c->opt_matvec_tab; return 0;
    default: FAIL("get_option: option not readable");
    }
}
int ccu_synchronize(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    SYNC(c);
    if(c->comm && c->comm->p2p)
    {
        unsigned e = 0;
        CK(cudaMemcpy(&e, c->comm->p2p_err, sizeof e, cudaMemcpyDeviceToHost));
        if(e) FAIL("peer-memory halo exchange: a wait for a neighbour's data gave up (a rank stopped, or the ranks ran different exchange sequences)");
    }
    return 0;
}
long long ccu_launch_count(ccu_ctx *c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------ replicated coarse levels
int ccu_agglomerate(ccu_ctx *c, int agg_lev, ccu_ctx **coarse_out)
{
    if(!c) FAIL("null context");
    if(!c->multi()) FAIL("agglomerate: only meaningful after ccu_comm_init with more than one subdomain");
    if(c->coarse) FAIL("agglomerate: already set up");
    if(agg_lev < c->cfg.levmin || agg_lev >= c->cfg.levmax) FAIL("agglomerate: level must be below the finest level");
    ccu_config cfg = c->cfg;
    cfg.levmax = agg_lev;
    const int *np = c->comm->nproc;
    for(int lev = cfg.levmin; lev <= agg_lev; lev++)
    {
        cfg.nox[lev] = c->L[lev].g.elx * np[0] + 1;
        cfg.noy[lev] = c->L[lev].g.ely * np[1] + 1;
        cfg.noz[lev] = c->L[lev].g.elz * np[2] + 1;
    }
    ccu_ctx *g = nullptr;
    if(ccu_create(&cfg, &g)) return 1;
    g->replica = true;
    g->use_graphs = false;                       // its kernels are captured into the owner's graphs
    cudaStreamDestroy(g->own_stream);
    g->own_stream = 0; g->st = c->st;
    g->visc = c->visc;
    Level &Ll = c->L[agg_lev];
    const size_t need = std::max(sizeof(double) * Ll.vlen(), sizeof(float) * 8 * (size_t)Ll.g.nel) * (size_t)c->comm->nranks;
    CK(cudaMalloc(&c->agg_buf, need));
    c->agg_bytes = need;
    c->coarse = g; c->agg_lev = agg_lev;
    drop_graphs(c);
    if(coarse_out) *coarse_out = g;
    return 0;
}

// ------------------------------------------------------------------ uploads
int ccu_set_node_flags(ccu_ctx *c, int lev, const unsigned *node)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    if(ccu_ensure_stage(c, sizeof(unsigned) * L.g.nno)) return 1;
    CK(cudaMemcpyAsync(c->stage, node, sizeof(unsigned) * L.g.nno, cudaMemcpyHostToDevice, c->st));
    LAUNCH(c, ccu_k_flags_to_dev, cdiv(L.g.nno, 256), 256, L.g, (const unsigned *)c->stage, L.flags);
    if(!L.node) CK(cudaMalloc(&L.node, sizeof(unsigned) * L.g.nno));        // raw bits (temperature BCs of the energy step)
    CK(cudaMemcpyAsync(L.node, c->stage, sizeof(unsigned) * L.g.nno, cudaMemcpyDeviceToDevice, c->st));
    SYNC(c);
    L.have_flags = true;
    if(lev == c->cfg.levmax) c->vb_dirty = true;      // which elements carry the K.VB force term depends on the flags
    return 0;
}

static int vec_h2d(ccu_ctx *c, Level &L, const double *host, double *dev)
{
    if(ccu_ensure_stage(c, sizeof(double) * L.g.neq)) return 1;
    CK(cudaMemcpyAsync(c->stage, host, sizeof(double) * L.g.neq, cudaMemcpyHostToDevice, c->st));
    LAUNCH(c, ccu_k_vec_to_dev, cdiv(L.g.nno, 256), 256, L.g, (const double *)c->stage, dev);
    return 0;
}
static int vec_d2h(ccu_ctx *c, Level &L, const double *dev, double *host)
{
    if(ccu_ensure_stage(c, sizeof(double) * L.g.neq)) return 1;
    LAUNCH(c, ccu_k_vec_to_nat, cdiv(L.g.nno, 256), 256, L.g, dev, (double *)c->stage);
    CK(cudaMemcpyAsync(host, c->stage, sizeof(double) * L.g.neq, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}

int ccu_set_stiffness(ccu_ctx *c, int lev, const float *k1, const float *k2, const float *k3, const double *BI)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    const size_t n42 = (size_t)L.g.nno * 42;
    if(ccu_ensure_stage(c, sizeof(float) * 3 * n42)) return 1;
    float *s = (float *)c->stage;
    CK(cudaMemcpyAsync(s, k1, sizeof(float) * n42, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(s + n42, k2, sizeof(float) * n42, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(s + 2 * n42, k3, sizeof(float) * n42, cudaMemcpyHostToDevice, c->st));
    LAUNCH(c, ccu_k_stiffness_to_dev, cdiv(L.g.nno, 128), 128, L.g, s, s + n42, s + 2 * n42, L.K);
    SYNC(c);
    if(vec_h2d(c, L, BI, L.BI)) return 1;
    SYNC(c);
    L.have_K = true;
    return ccu_col_refresh(c, lev);
}

int ccu_set_pressure_ops(ccu_ctx *c, int lev, const float *elt_del, const double *BPI)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    CK(cudaMemcpyAsync(L.elt_del, elt_del, sizeof(float) * 24 * (size_t)L.g.nel, cudaMemcpyHostToDevice, c->st));
    ccu_elt_del_changed(c, lev);
    CK(cudaMemcpyAsync(L.BPI, BPI, sizeof(double) * L.g.npno, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    L.have_p = true;
    return 0;
}

int ccu_set_transfer_weights(ccu_ctx *c, int lev, const float *TWW, const float *MASS, const float *eco)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    CK(cudaMemcpyAsync(L.TWW, TWW, sizeof(float) * 8 * (size_t)L.g.nel, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(L.MASS, MASS, sizeof(float) * L.g.nno, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(L.eco, eco, sizeof(float) * 3 * (size_t)L.g.nel, cudaMemcpyHostToDevice, c->st));
    ccu_eco_changed(c, lev);
    SYNC(c);
    L.have_tw = true;
    return 0;
}

// coefficient-major copy of elt_del for div_u / grad_p; call after every write of L.elt_del
void ccu_fork_stream(ccu_ctx *c)
{
    cudaEventRecord(c->ev_fork, c->st);
    cudaStreamWaitEvent(c->st2, c->ev_fork, 0);
}
void ccu_join_stream(ccu_ctx *c)
{
    cudaEventRecord(c->ev_join, c->st2);
    cudaStreamWaitEvent(c->st, c->ev_join, 0);
}
void ccu_eco_changed(ccu_ctx *c, int lev)
{
    Level &L = c->L[lev];
    LAUNCH(c, ccu_k_eco_transpose, cdiv((size_t)L.g.nel * 3, 256), 256, L.g.nel, L.eco, L.ecoT);
}
void ccu_elt_del_changed(ccu_ctx *c, int lev)
{
    Level &L = c->L[lev];
    LAUNCH(c, ccu_k_elt_del_transpose, cdiv((size_t)L.g.nel * 24, 256), 256, L.g.nel, L.elt_del, L.elt_delT);
}

// ------------------------------------------------------------------ device-side building blocks
static const CcuCoef C_ONE = { nullptr, nullptr, 1.0 };
static const CcuCoef C_ZERO = { nullptr, nullptr, 0.0 };
static const CcuCoef C_MINUS = { nullptr, nullptr, -1.0 };
static inline CcuCoef coef(const double *num, const double *den, double scale) { return CcuCoef{ num, den, scale }; }

static void d_copy(ccu_ctx *c, double *dst, const double *src, size_t n)
{
    cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->st);
}
static void d_zero(ccu_ctx *c, double *dst, size_t n) { cudaMemsetAsync(dst, 0, sizeof(double) * n, c->st); }
static void d_axpby(ccu_ctx *c, size_t n, double *y, const double *x, CcuCoef a, CcuCoef b)
{
    LAUNCH(c, ccu_k_axpby, min(cdiv(n, 256), 148u * 16u), 256, n, y, x, a, b);
}
static void d_waxpby(ccu_ctx *c, size_t n, double *z, const double *x, const double *y, CcuCoef a, CcuCoef b)
{
    LAUNCH(c, ccu_k_waxpby, min(cdiv(n, 256), 148u * 16u), 256, n, z, x, y, a, b);
}
// up to three dots in one pass over the vectors; results land in scal[slot*]
// `maskL` = the level whose nodal vectors are being dotted (ownership mask in multi-subdomain runs), null for
// element (pressure) vectors.  Multi-subdomain: local sums -> one ncclAllReduce of the (<= 3) scalars -> slots.
static void d_dot3m(ccu_ctx *c, const Level *maskL, size_t n, const double *a0, const double *b0, int s0, const double *a1 = nullptr,
                    const double *b1 = nullptr, int s1 = -1, const double *a2 = nullptr, const double *b2 = nullptr, int s2 = -1)
{
    const int nb = (int)min((size_t)CCU_DOT_BLOCKS, (size_t)cdiv(n, 256));
    double *o0 = c->scal + s0, *o1 = s1 >= 0 ? c->scal + s1 : nullptr, *o2 = s2 >= 0 ? c->scal + s2 : nullptr;
    if(!c->multi())
    {
        LAUNCH(c, ccu_k_dot_partial, nb, 256, n, a0, b0, a1, b1, a2, b2, c->partial, (const unsigned char *)nullptr, (size_t)1);
        LAUNCH(c, ccu_k_dot_final, 1, 256, c->partial, nb, o0, o1, o2);
        return;
    }
    const unsigned char *own = nullptr; size_t ns = 1;
    if(maskL) { own = c->comm->halo[maskL - c->L].bits; ns = (size_t)maskL->g.NS; }
    double *st = c->comm->dotstage;
    LAUNCH(c, ccu_k_dot_partial, nb, 256, n, a0, b0, a1, b1, a2, b2, c->partial, own, ns);
    LAUNCH(c, ccu_k_dot_final, 1, 256, c->partial, nb, st, o1 ? st + 1 : nullptr, o2 ? st + 2 : nullptr);
    ccu_allreduce_dots(c, 1 + (o1 ? 1 : 0) + (o2 ? 1 : 0), o0, o1, o2);
}
static void d_dot3(ccu_ctx *c, size_t n, const double *a0, const double *b0, int s0, const double *a1 = nullptr, const double *b1 = nullptr,
                   int s1 = -1, const double *a2 = nullptr, const double *b2 = nullptr, int s2 = -1)
{
    d_dot3m(c, nullptr, n, a0, b0, s0, a1, b1, s1, a2, b2, s2);
}
static int read_scal(ccu_ctx *c, int first, int count, double *out)
{
    CK(cudaMemcpyAsync(out, c->scal + first, sizeof(double) * count, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}

static void d_strip(ccu_ctx *c, Level &L, double *v) { LAUNCH(c, ccu_k_strip, cdiv(L.g.NS, 256), 256, L.g, L.flags, v); }

// ---- column-resident kernels (ccu_col.cuh)
typedef CcuColShape<6, 8, 2> ColA;      // 128 threads, 93 KB: two CTAs per SM, 19 % halo blocks
typedef CcuColShape<12, 8, 1> ColB;     // 256 threads, 182 KB: one CTA per SM, 13 % halo blocks
typedef CcuColShape<4, 8, 3> ColC;      // 96 threads, 67 KB: three CTAs per SM (small subdomains: more columns)
template <class SH, int MODE, int WF>
static int launch_col(ccu_ctx *c, Level &L, int cc, const double *F, double *x, double *out, int strip)
{
    CcuColArgs A;
    A.g = L.g; A.Kc = L.Kc; A.colofs = L.colofs; A.F = F; A.x = x; A.out = out;
    A.nI = L.col_nI; A.nJ = L.col_nJ; A.cc = cc; A.strip = strip;
    A.ticket = nullptr; A.progress = nullptr;
    for(int q = 0; q < 5; q++) A.cstart[q] = 0;
    unsigned grid = (unsigned)(A.nI * A.nJ);
    if(MODE == 0 && !WF) grid = (unsigned)(((A.nI - (cc >> 1) + 1) / 2) * ((A.nJ - (cc & 1) + 1) / 2));
    if(MODE == 0 && WF)
    {   // tickets in colour order 3, 2, 1, 0
        for(int grp = 0; grp < 4; grp++)
        {
            const int col = 3 - grp;
            A.cstart[grp + 1] = A.cstart[grp] + ((A.nI - (col >> 1) + 1) / 2) * ((A.nJ - (col & 1) + 1) / 2);
        }
        A.ticket = L.col_sync; A.progress = L.col_sync + 4;
    }
    if(!grid) return 0;
    ccu_k_col<SH, MODE, WF><<<grid, SH::THREADS, SH::SMEM, c->st>>>(A);
    c->launches++;
    return 0;
}
template <int MODE, int WF>
static int launch_col_shape(ccu_ctx *c, Level &L, int cc, const double *F, double *x, double *out, int strip)
{
    if(L.col_shape == 1) return launch_col<ColB, MODE, WF>(c, L, cc, F, x, out, strip);
    if(L.col_shape == 2) return launch_col<ColC, MODE, WF>(c, L, cc, F, x, out, strip);
    return launch_col<ColA, MODE, WF>(c, L, cc, F, x, out, strip);
}
template <class SH>
static int col_attrs()
{   // more than 48 KB of dynamic shared memory needs the opt-in, per device; set for the current device of the caller
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SMEM));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SMEM));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SMEM));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SMEM));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 0, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 0, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 1, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(ccu_k_col<SH, 2, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
template <int TI, int TJ>
static int col_relayout(ccu_ctx *c, Level &L, int lev)
{
    const CcuGeom &g = L.g;
    if((unsigned)g.noz + 8u >= CCU_COL_EPOCH) FAIL("column kernels: more z layers than the progress words can count");
    const int nI = (g.noy + TI - 1) / TI, nJ = (g.nox + TJ - 1) / TJ;
    std::vector<size_t> ofs((size_t)nI * nJ);
    size_t total = 0;
    for(int I = 0; I < nI; I++)
        for(int J = 0; J < nJ; J++)
        {
            ofs[(size_t)I * nJ + J] = total;
            total += (size_t)g.noz * ccu_col_dims(std::min(TI, g.noy - I * TI), std::min(TJ, g.nox - J * TJ)).cb;
        }
    if(total > L.Kc_bytes || L.col_nI != nI || L.col_nJ != nJ)
    {
        cudaFree(L.Kc); cudaFree(L.colofs); cudaFree(L.col_sync);
        L.Kc = nullptr; L.colofs = nullptr; L.col_sync = nullptr; L.Kc_bytes = 0;
        CK(cudaMalloc(&L.Kc, total));
        CK(cudaMemsetAsync(L.Kc, 0, total, c->st));      // block positions no block maps to (ccu_col_pos) are never written again
        CK(cudaMalloc(&L.colofs, sizeof(size_t) * ofs.size()));
        CK(cudaMalloc(&L.col_sync, sizeof(unsigned) * (4 + ofs.size())));

        CK(cudaMemsetAsync(L.col_sync, 0, sizeof(unsigned) * (4 + ofs.size()), c->st));
        L.Kc_bytes = total;
    }
    if(L.col_inv_key != TI * 100 + TJ)
    {   // position -> block id of the full column shape (the inverse of ccu_col_pos), -1 where no block sits
        const CcuColDims cf = ccu_col_dims(TI, TJ);
        std::vector<short> inv((size_t)cf.nbp, (short)-1);
        for(int id = 0; id < cf.nb; id++) inv[(size_t)ccu_col_pos(cf, id)] = (short)id;
        cudaFree(L.col_inv); L.col_inv = nullptr;
        CK(cudaMalloc(&L.col_inv, sizeof(short) * inv.size()));
        CK(cudaMemcpy(L.col_inv, inv.data(), sizeof(short) * inv.size(), cudaMemcpyHostToDevice));
        L.col_inv_key = TI * 100 + TJ;
    }
    CK(cudaMemcpyAsync(L.colofs, ofs.data(), sizeof(size_t) * ofs.size(), cudaMemcpyHostToDevice, c->st));
    SYNC(c);               // `ofs` leaves scope
    L.col_nI = nI; L.col_nJ = nJ;
    const unsigned char *bits = c->multi() ? c->comm->halo[lev].bits : nullptr;
    LAUNCH(c, (ccu_k_col_relayout<TI, TJ>), dim3((unsigned)(nI * nJ), (unsigned)((g.noz + 31) / 32)), 256, g, nJ, L.colofs, L.K, L.BI, L.flags, bits,
           L.col_inv, L.Kc);
    return 0;
}
// The column kernels read a column-major copy of the level's stiffness, inverse diagonal and flags: (re)made whenever
// they change (ccu_set_stiffness, ccu_construct_stiffness_B_matrix, ccu_comm_init) or the column options do.
// Never called inside a graph capture.
int ccu_col_refresh(ccu_ctx *c, int lev)
{
    Level &L = c->L[lev];
    // full-row copy (ccu_k_build_KT)
    L.have_KT = false;
    if((c->opt_relax_full || c->opt_matvec_full) && L.g.nno > c->opt_full_nodes && L.have_K)
    {
        if(!L.KT)
        {
            CK(cudaMalloc(&L.KT, sizeof(float) * 117 * (size_t)L.g.NS));
            CK(cudaMemsetAsync(L.KT, 0, sizeof(float) * 117 * (size_t)L.g.NS, c->st));
        }
        LAUNCH(c, ccu_k_build_KT, dim3(cdiv(L.g.NC, 256), 8), 256, L.g, ccu_make_stencil(L.g), L.K, L.KT);
        L.have_KT = true;
    }
    L.col_shape = -1;
    if(!(c->opt_relax_col || c->opt_matvec_col) || L.g.nno <= c->opt_col_nodes || !L.have_K || !L.have_flags) return 0;
    int rc;
    if(c->opt_col_shape == 1) rc = col_attrs<ColB>() || col_relayout<ColB::TI, ColB::TJ>(c, L, lev);
    else if(c->opt_col_shape == 2) rc = col_attrs<ColC>() || col_relayout<ColC::TI, ColC::TJ>(c, L, lev);
    else rc = col_attrs<ColA>() || col_relayout<ColA::TI, ColA::TJ>(c, L, lev);
    if(rc) return rc;
    L.col_shape = c->opt_col_shape;
    return 0;
}
int ccu_col_refresh_all(ccu_ctx *c) { return col_refresh_all(c); }
static int col_refresh_all(ccu_ctx *c)
{
    for(int lev = c->cfg.levmin; lev <= c->cfg.levmax; lev++)
        if(ccu_col_refresh(c, lev)) return 1;
    if(c->coarse)
    {
        c->coarse->opt_col_nodes = c->opt_col_nodes; c->coarse->opt_relax_col = c->opt_relax_col; c->coarse->opt_matvec_col = c->opt_matvec_col;
        c->coarse->opt_col_shape = c->opt_col_shape; c->coarse->opt_col_wf = c->opt_col_wf;
        c->coarse->opt_full_nodes = c->opt_full_nodes; c->coarse->opt_relax_full = c->opt_relax_full; c->coarse->opt_matvec_full = c->opt_matvec_full;
        return col_refresh_all(c->coarse);
    }
    return 0;
}
static bool use_col(const ccu_ctx *c, const Level &L, int on) { (void)c; return on && L.col_shape >= 0; }

// Lanes per node by level size: the smaller the level, the more the per-node chain of dependent loads is the
// whole kernel time, so it is split over more lanes (ccu_kernels.cuh).  Tunable through ccu_set_option.
static int lanes_for(const ccu_ctx *c, const Level &L)
{
    if(L.g.nno <= c->opt_small_nodes) return c->multi() ? 32 : 0;          // single-CTA fused kernel (one exchange per sweep rules it out)
    if(L.g.nno <= c->opt_warp_nodes) return 32;
    if(L.g.nno <= c->opt_quad_nodes) return 4;
    return c->opt_lanes_large;
}
static void d_matvec(ccu_ctx *c, Level &L, const double *u, double *Au, int strip)
{
    CcuProfScope ps(c, CCU_PROF_MATVEC_FINE, &L == &c->L[c->cfg.levmax]);
    CcuProfScope pl(c, CCU_PROF_LEVEL0 + (int)(&L - c->L), true, 0);
    const int T = (c->opt_matvec_tab && L.g.nno > c->opt_matvec_tab_nodes) ? 1 : lanes_for(c, L);   // the table-driven kernel wins from ~1e4 nodes up
    if(c->opt_matvec_full && L.have_KT) LAUNCH(c, (ccu_k_matvec_full<0, 2>), cdiv(L.g.NC, 32), 256, L.g, ccu_make_stencil(L.g), L.K, L.KT, L.flags, u, nullptr, Au, strip);
    else if(use_col(c, L, c->opt_matvec_col)) launch_col_shape<1, 0>(c, L, 0, nullptr, const_cast<double *>(u), Au, strip);
    else if(T == 0 || T == 32) LAUNCH(c, (ccu_k_matvec_lanes<32, 0>), L.g.NC, 256, L.g, L.K, L.flags, u, nullptr, Au, strip);
    else if(T == 4) LAUNCH(c, (ccu_k_matvec_lanes<4, 0>), cdiv(L.g.NC, 8), 256, L.g, L.K, L.flags, u, nullptr, Au, strip);
    else if(c->opt_matvec_tab >= 4) LAUNCH(c, (ccu_k_matvec_tab<0, 4>), cdiv(L.g.NC, 32), 256, L.g, ccu_make_stencil(L.g), L.K, L.flags, u, nullptr, Au, strip);
    else if(c->opt_matvec_tab) LAUNCH(c, (ccu_k_matvec_tab<0, 2>), cdiv(L.g.NC, 32), 256, L.g, ccu_make_stencil(L.g), L.K, L.flags, u, nullptr, Au, strip);
    else LAUNCH(c, ccu_k_matvec<0>, cdiv(L.g.NC, 32), 256, L.g, L.K, L.flags, u, nullptr, Au, strip);
    if(c->multi()) ccu_halo_sum_vec(c, (int)(&L - c->L), Au);      // exchange_id_d20 (Element_calculations.c:612)
}
// out = rhs - K u, boundary rows of K u stripped first (the reference's res = rhs - AU with AU stripped)
static void d_residual(ccu_ctx *c, Level &L, const double *u, const double *rhs, double *out)
{
    if(c->multi())
    {   // the halo sum sits between the product and the subtraction
        d_matvec(c, L, u, out, 1);
        d_axpby(c, L.vlen(), out, rhs, C_ONE, C_MINUS);
        return;
    }
    CcuProfScope ps(c, CCU_PROF_MATVEC_FINE, &L == &c->L[c->cfg.levmax]);
    CcuProfScope pl(c, CCU_PROF_LEVEL0 + (int)(&L - c->L), true, 0);
    const int T = (c->opt_matvec_tab && L.g.nno > c->opt_matvec_tab_nodes) ? 1 : lanes_for(c, L);
    if(c->opt_matvec_full && L.have_KT) { LAUNCH(c, (ccu_k_matvec_full<1, 2>), cdiv(L.g.NC, 32), 256, L.g, ccu_make_stencil(L.g), L.K, L.KT, L.flags, u, rhs, out, 1); return; }
    if(use_col(c, L, c->opt_matvec_col)) { launch_col_shape<2, 0>(c, L, 0, rhs, const_cast<double *>(u), out, 1); return; }
    if(T == 0 || T == 32) { LAUNCH(c, (ccu_k_matvec_lanes<32, 1>), L.g.NC, 256, L.g, L.K, L.flags, u, rhs, out, 1); return; }
    if(T == 4) { LAUNCH(c, (ccu_k_matvec_lanes<4, 1>), cdiv(L.g.NC, 8), 256, L.g, L.K, L.flags, u, rhs, out, 1); return; }
    if(c->opt_matvec_tab >= 4) { LAUNCH(c, (ccu_k_matvec_tab<1, 4>), cdiv(L.g.NC, 32), 256, L.g, ccu_make_stencil(L.g), L.K, L.flags, u, rhs, out, 1); return; }
    if(c->opt_matvec_tab) { LAUNCH(c, (ccu_k_matvec_tab<1, 2>), cdiv(L.g.NC, 32), 256, L.g, ccu_make_stencil(L.g), L.K, L.flags, u, rhs, out, 1); return; }
    LAUNCH(c, ccu_k_matvec<1>, cdiv(L.g.NC, 32), 256, L.g, L.K, L.flags, u, rhs, out, 1);
}

// compact colour-sorted tables of a tiny level for ccu_k_relax_smem
static int ensure_smem_tables(ccu_ctx *c, Level &L)
{
    if(L.sm_s) return 0;
    const CcuGeom &g = L.g;
    const int LO[13][3] = CCU_LO_INIT;
    const int n = g.nno;
    std::vector<int> s(n), compact((size_t)g.NS, n);
    int t = 0;
    for(int col = 0; col < 8; col++)
    {
        L.sm_cstart[col] = t;
        for(int cell = 0; cell < g.NC; cell++)
        {
            int i, j, k;
            if(!ccu_decode(g, col, cell, i, j, k)) continue;
            s[t] = col * g.NC + cell;
            compact[s[t]] = t;
            t++;
        }
    }
    L.sm_cstart[8] = t;
    if(t != n) FAIL("smem tables: node count mismatch");
    for(int col = 0; col < 8; col++)
        if(L.sm_cstart[col + 1] - L.sm_cstart[col] > 128) return 0;      // more nodes per colour than the kernel's 128 groups: not eligible
    std::vector<unsigned short> nbr((size_t)27 * n);
    for(int tt = 0; tt < n; tt++)
    {
        const int col = s[tt] / g.NC, cell = s[tt] % g.NC;
        int i, j, k;
        ccu_decode(g, col, cell, i, j, k);
        for(int b = 0; b < 27; b++)
        {
            int di = 0, dj = 0, dk = 0;
            if(b >= 1 && b <= 13) { di = LO[b - 1][0]; dj = LO[b - 1][1]; dk = LO[b - 1][2]; }
            if(b >= 14) { di = -LO[b - 14][0]; dj = -LO[b - 14][1]; dk = -LO[b - 14][2]; }
            const int ii = i + di, jj = j + dj, kk = k + dk;
            const bool in = ii >= 0 && ii < g.noy && jj >= 0 && jj < g.nox && kk >= 0 && kk < g.noz;
            nbr[(size_t)b * n + tt] = (unsigned short)(in ? compact[ccu_sidx(g, ii, jj, kk)] : n);
        }
    }
    CK(cudaMalloc(&L.sm_s, sizeof(int) * n));
    CK(cudaMalloc(&L.sm_nbr, sizeof(unsigned short) * 27 * n));
    CK(cudaMemcpy(L.sm_s, s.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.sm_nbr, nbr.data(), sizeof(unsigned short) * 27 * n, cudaMemcpyHostToDevice));
    L.sm_n = n;
    static bool attr_set = false;
    if(!attr_set) { CK(cudaFuncSetAttribute(ccu_k_relax_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); attr_set = true; }
    return 0;
}


// distance window of colour `col` (passes run 7 .. 0) in the three modes of a sweep: 0 = every node that is not duplicated,
// 1 = "far" (more than p + 1 nodes away from a duplicated node, p = 7 - col the pass number), 2 = "near" (the rest).  The far part of
// pass p only reads nodes at distance > p, i.e. far parts of earlier passes and values no near pass or face update of this sweep has
// touched yet, so the eight far passes can run while the duplicated nodes are exchanged; the near parts follow in pass order and
// see exactly the values they see in the unsplit sweep (bitwise the same result, tests/test_gpu_multi.py).
static inline int pass_window(int col, int mode)
{
    const int p = 7 - col;
    return mode == 0 ? CCU_D_ALL : (mode == 1 ? CCU_D_RANGE(p + 2, 15) : CCU_D_RANGE(1, p + 1));
}
template <int T>
static void launch_relax_lanes(ccu_ctx *c, Level &L, double *x, const double *F, const unsigned char *bits, int mode)
{
    const unsigned grid = cdiv((size_t)L.g.NC * T, 128);
    LAUNCH(c, (ccu_k_relax_lanes<T, 7>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(7, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 6>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(6, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 5>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(5, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 4>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(4, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 3>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(3, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 2>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(2, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 1>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(1, mode));
    LAUNCH(c, (ccu_k_relax_lanes<T, 0>), grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(0, mode));
}
// Multi-subdomain sweeps start with the duplicated face nodes: every owner computes its part of their rows, one halo
// round sums the parts, and all owners apply the same damped-Jacobi update (the reference's OFFSIDE treatment,
// General_matrix_functions.c:1218-1230, 1262-1284: one exchange per sweep).  The colour passes then skip those nodes.
static void relax_faces(ccu_ctx *c, Level &L, double *x, const double *F)
{
    const int lev = (int)(&L - c->L);
    const CcuHalo &H = c->comm->halo[lev];
    if(H.n_shared == 0) return;
    CcuProfScope pf(c, CCU_PROF_FACES_FINE, lev == c->cfg.levmax);
    LAUNCH(c, ccu_k_face_rows<0>, cdiv((size_t)H.n_shared * 32, 128), 128, L.g, L.K, x, H.n_shared, H.sh_s, H.face);
    ccu_halo_sum_face(c, lev);
    LAUNCH(c, ccu_k_face_update, cdiv(H.n_shared, 128), 128, L.g, H.n_shared, H.sh_s, H.sh_ptr, H.sh_src, H.face,
           (const double *)c->comm->recvbuf, L.BI, F, x);
}
static void d_relax_sweeps(ccu_ctx *c, Level &L, double *x, const double *F, int cycles)
{
    CcuProfScope ps(c, CCU_PROF_RELAX_FINE, &L == &c->L[c->cfg.levmax], (use_col(c, L, c->opt_relax_col) ? (c->opt_col_wf ? 1LL : 4LL) : 8LL) * cycles);
    CcuProfScope pl(c, CCU_PROF_LEVEL0 + (int)(&L - c->L), true, cycles);
    if(use_col(c, L, c->opt_relax_col))
    {   // column order: column colours 3..0, z ascending inside a column, (y, x)-parity colours 3..0 inside a layer (ccu_col.cuh)
        const bool multi = c->multi();
        for(int s = 0; s < cycles; s++)
        {
            if(multi) relax_faces(c, L, x, F);
            if(c->opt_col_wf) launch_col_shape<0, 1>(c, L, 0, F, x, nullptr, 0);
            else
                for(int cc = 3; cc >= 0; cc--) launch_col_shape<0, 0>(c, L, cc, F, x, nullptr, 0);
        }
        return;
    }
    const int T = lanes_for(c, L);
    if(T == 0)
    {   // one CTA does every sweep and colour of a tiny level in a single launch
        if(L.g.nno <= c->opt_smem_nodes && L.sm_s && c->opt_bottom_cluster)
        {   // ... spread over an 8-CTA cluster, fp64 rows resident in distributed shared memory (ccu_k_relax_bottom)
            CcuSmemLevel sl; sl.n = L.sm_n; sl.s = L.sm_s; sl.nbr = L.sm_nbr;
            for(int q = 0; q < 9; q++) sl.cstart[q] = L.sm_cstart[q];
            int maxrows = 0;
            for(int q = 0; q < 8; q++) maxrows += (L.sm_cstart[q + 1] - L.sm_cstart[q] + CCU_BOT_CTAS - 1) / CCU_BOT_CTAS;
            const size_t bytes = sizeof(double) * (3 * (size_t)(L.sm_n + 1) + (size_t)maxrows * 9 * CCU_BOT_LD + 6 * (size_t)maxrows)
                                 + sizeof(unsigned short) * (size_t)maxrows * CCU_BOT_LD + 16;
            if(!c->bottom_attr_set)
            {   // the opt-in is per device: once per context (a context never changes device)
                if(cudaFuncSetAttribute(ccu_k_relax_bottom, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess && !c->launch_err)
                { c->launch_err = (int)cudaGetLastError(); c->launch_err_kernel = "cudaFuncSetAttribute(ccu_k_relax_bottom)"; }
                c->bottom_attr_set = true;
            }
            ccu_k_relax_bottom<<<CCU_BOT_CTAS, CCU_BOT_THREADS, bytes, c->st>>>(L.g, sl, maxrows, L.K, L.BI, F, x, cycles, 0);
            c->launches++;
            if(!c->launch_err) { const cudaError_t le_ = cudaPeekAtLastError(); if(le_ != cudaSuccess) { c->launch_err = (int)le_; c->launch_err_kernel = "ccu_k_relax_bottom"; } }
            return;
        }
        if(L.g.nno <= c->opt_smem_nodes && L.sm_s)
        {   // ... out of shared memory when the whole half-matrix fits
            CcuSmemLevel sl; sl.n = L.sm_n; sl.s = L.sm_s; sl.nbr = L.sm_nbr;
            for(int q = 0; q < 9; q++) sl.cstart[q] = L.sm_cstart[q];
            const size_t bytes = (size_t)(L.sm_n + 1) * (3 * sizeof(double) + 126 * sizeof(float)) + (size_t)L.sm_n * 6 + 16;
            ccu_k_relax_smem<<<1, 1024, bytes, c->st>>>(L.g, sl, L.K, L.BI, F, x, cycles, 0);
            c->launches++;
            return;
        }
        LAUNCH(c, ccu_k_relax_small, 1, 1024, L.g, L.K, L.BI, F, x, cycles, 0);
        return;
    }
    const unsigned grid = cdiv(L.g.NC, 128);
    const unsigned char *bits = c->multi() ? c->comm->halo[&L - c->L].bits : nullptr;
    // the eight colour passes 7 .. 0 (odd-odd-odd nodes first, the coarse-grid nodes last) over the distance window of `mode`
    auto passes = [&](int mode) {
        if(T == 32) { launch_relax_lanes<32>(c, L, x, F, bits, mode); return; }
        if(T == 4) { launch_relax_lanes<4>(c, L, x, F, bits, mode); return; }      // mid levels: four lanes per node
        if(c->opt_relax_full && L.have_KT)
        {   // full rows: every coefficient streams once per sweep
            const CcuStencil st = ccu_make_stencil(L.g);
            for(int col = 7; col >= 0; col--) LAUNCH(c, ccu_k_relax_full<2>, grid, 128, L.g, st, col, L.K, L.KT, L.BI, F, x, bits, pass_window(col, mode));
            return;
        }
        if(c->opt_relax_tab)
        {
            const CcuStencil st = ccu_make_stencil(L.g);
            for(int col = 7; col >= 0; col--)
            {
                if(c->opt_relax_tab >= 7) LAUNCH(c, ccu_k_relax_tab<7>, grid, 128, L.g, st, col, L.K, L.BI, F, x, bits, pass_window(col, mode));
                else if(c->opt_relax_tab >= 4) LAUNCH(c, ccu_k_relax_tab<4>, grid, 128, L.g, st, col, L.K, L.BI, F, x, bits, pass_window(col, mode));
                else LAUNCH(c, ccu_k_relax_tab<2>, grid, 128, L.g, st, col, L.K, L.BI, F, x, bits, pass_window(col, mode));
            }
            return;
        }
        LAUNCH(c, ccu_k_relax<7>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(7, mode));
        LAUNCH(c, ccu_k_relax<6>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(6, mode));
        LAUNCH(c, ccu_k_relax<5>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(5, mode));
        LAUNCH(c, ccu_k_relax<4>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(4, mode));
        LAUNCH(c, ccu_k_relax<3>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(3, mode));
        LAUNCH(c, ccu_k_relax<2>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(2, mode));
        LAUNCH(c, ccu_k_relax<1>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(1, mode));
        LAUNCH(c, ccu_k_relax<0>, grid, 128, L.g, L.K, L.BI, F, x, bits, pass_window(0, mode));
    };
    // overlap pays where the far shells hold most of the level; tiny subdomains are all "near"
    const bool overlap = bits && c->opt_halo_overlap && c->st2 && c->comm->halo[&L - c->L].n_shared > 0 &&
                         (c->opt_halo_overlap > 1 || L.g.nno > 400000);
    for(int s = 0; s < cycles; s++)
    {
        if(!bits) { passes(0); continue; }
        if(!overlap) { relax_faces(c, L, x, F); passes(0); continue; }
        // duplicated nodes (partial rows, exchange, damped-Jacobi update) stay on the context's stream -- every NCCL call of the
        // library is issued there -- and are queued first; the far shells relax meanwhile on the second stream
        cudaStream_t st = c->st;
        ccu_fork_stream(c);
        relax_faces(c, L, x, F);
        c->st = c->st2;
        passes(1);
        c->st = st;
        ccu_join_stream(c);
        passes(2);
    }
}

// rebuild_BI_on_boundary (Construct_arrays.c:892-952): damped inverse diagonal on the duplicated nodes
int ccu_damp_face_BI(ccu_ctx *c, int lev)
{
    if(!c->multi()) return 0;
    Level &L = c->L[lev];
    const CcuHalo &H = c->comm->halo[lev];
    if(H.n_shared == 0) return 0;
    LAUNCH(c, ccu_k_face_rows<1>, cdiv((size_t)H.n_shared * 32, 128), 128, L.g, L.K, (const double *)nullptr, H.n_shared, H.sh_s, H.face);
    if(ccu_halo_sum_face(c, lev)) return 1;
    LAUNCH(c, ccu_k_face_damp_BI, cdiv(H.n_shared, 128), 128, L.g, H.n_shared, H.sh_s, H.sh_ptr, H.sh_src, H.face,
           (const double *)c->comm->recvbuf, L.BI);
    return 0;
}

// gauss_seidel (General_matrix_functions.c:1160): d0, Ad = K d0
static void d_gauss_seidel(ccu_ctx *c, Level &L, double *d0, const double *F, double *Ad, int cycles, int guess)
{
    if(!guess) d_zero(c, d0, L.vlen());
    d_relax_sweeps(c, L, d0, F, cycles);
    d_matvec(c, L, d0, Ad, 1);
}

static void d_project(ccu_ctx *c, int lev, const double *fine, double *coarse, int strip)
{
    Level &Lf = c->L[lev], &Lc = c->L[lev - 1];
    CcuProfScope ps(c, CCU_PROF_TRANSFER_FINE, lev == c->cfg.levmax);
    CcuProfScope pl(c, CCU_PROF_LEVEL0 + lev, true, 0);
    const int multi = c->multi() ? 1 : 0;
    LAUNCH(c, ccu_k_project, cdiv(Lc.g.nno, 128), 128, Lc.g, Lf.g, Lc.TWW, Lc.MASS, fine, coarse, multi ? 0 : 1);
    if(multi)
    {   // exchange_id_d20 before the mass factor (Solver_multigrid.c:150-157)
        ccu_halo_sum_vec(c, lev - 1, coarse);
        LAUNCH(c, ccu_k_mass_mul, cdiv(8 * (size_t)Lc.g.NC, 128), 128, Lc.g, Lc.MASS, coarse);
    }
    if(strip) d_strip(c, Lc, coarse);
}
static void d_interp(ccu_ctx *c, int lev, const double *coarse, double *fine, int strip)
{
    Level &Lc = c->L[lev], &Lf = c->L[lev + 1];
    CcuProfScope ps(c, CCU_PROF_TRANSFER_FINE, lev + 1 == c->cfg.levmax);
    CcuProfScope pl(c, CCU_PROF_LEVEL0 + lev + 1, true, 0);
    LAUNCH(c, ccu_k_interp, cdiv(8 * (size_t)Lf.g.NC, 128), 128, Lc.g, Lf.g, Lf.ecoT, Lf.flags, coarse, fine, strip);
}

// A fixed launch sequence captured once into a CUDA graph and replayed: the coarse levels of the
// multigrid cycle are hundreds of microsecond-scale kernels, bound by launch latency when issued one by one.
template <class F>
static int run_segment(ccu_ctx *c, int id, F body)
{
    CcuProfScope ps(c, CCU_PROF_COARSE, true);
    if(!c->use_graphs) { body(); return 0; }
    ccu_ctx::GraphSeg &s = c->seg[id];
    if(!s.exec)
    {
        const long long l0 = c->launches, g0 = c->coarse ? c->coarse->launches : 0;
        const bool prof = c->prof_on;
        c->prof_on = false;
        CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
        body();
        cudaGraph_t g = nullptr;
        CK(cudaStreamEndCapture(c->st, &g));
        c->prof_on = prof;
        s.launches = c->launches - l0 + (c->coarse ? c->coarse->launches - g0 : 0);
        c->launches = l0;
        if(c->coarse) c->coarse->launches = g0;
        CK(cudaGraphInstantiate(&s.exec, g, 0));
        cudaGraphDestroy(g);
    }
    CK(cudaGraphLaunch(s.exec, c->st));
    c->launches += s.launches;
    return 0;
}

// multi_grid (General_matrix_functions.c:525-653), one V-cycle on level `lev` split in three so that the part
// below the finest level can be replayed as a graph:
//   mg_down_top : smooth on lev (warm start), res = rhs - K vel, restrict to rhs[lev-1]            (:590-605, dlev == lev)
//   mg_inner    : down-stroke lev-1 .. levmin+1 from zero, bottom solve, up-stroke levmin+1 .. lev-1  (:590-636)
//   mg_up_top   : interpolate the correction, smooth, line search alpha, update vel (and res on levmax) (:618-636, ulev == lev)
static void mg_down(ccu_ctx *c, int dlev, bool warm)
{
    Level *L = c->L;
    Level &D = L[dlev];
    const int cycles = (dlev == c->cfg.levmax && !c->replica) ? c->cfg.v_steps_high : c->cfg.down_heavy;
    if(!warm) d_zero(c, D.vec[CCU_VEC_VEL], D.vlen());
    d_relax_sweeps(c, D, D.vec[CCU_VEC_VEL], D.vec[CCU_VEC_RHS], cycles);
    d_residual(c, D, D.vec[CCU_VEC_VEL], D.vec[CCU_VEC_RHS], D.vec[CCU_VEC_RES]);   // res = rhs - AU
    d_project(c, dlev, D.vec[CCU_VEC_RES], L[dlev - 1].vec[CCU_VEC_RHS], 1);
}
static void mg_up(ccu_ctx *c, int ulev)
{
    Level *L = c->L;
    Level &U = L[ulev];
    const int cycles = (ulev == c->cfg.levmax && !c->replica) ? c->cfg.v_steps_high : c->cfg.up_heavy;
    d_interp(c, ulev - 1, L[ulev - 1].vec[CCU_VEC_VEL], U.vec[CCU_VEC_DEL_VEL], 1);
    d_gauss_seidel(c, U, U.vec[CCU_VEC_DEL_VEL], U.vec[CCU_VEC_RES], U.vec[CCU_VEC_AU], cycles, 1);
    // alpha = <AU,res>/<AU,AU>  (line search, :626-627); both dots share one pass
    d_dot3m(c, &U, U.vlen(), U.vec[CCU_VEC_AU], U.vec[CCU_VEC_AU], S_DOT1, U.vec[CCU_VEC_AU], U.vec[CCU_VEC_RES], S_DOT2);
    d_axpby(c, U.vlen(), U.vec[CCU_VEC_VEL], U.vec[CCU_VEC_DEL_VEL], coef(c->scal + S_DOT2, c->scal + S_DOT1, 1.0), C_ONE);
    if(ulev == c->cfg.levmax && !c->replica)
        d_axpby(c, U.vlen(), U.vec[CCU_VEC_RES], U.vec[CCU_VEC_AU], coef(c->scal + S_DOT2, c->scal + S_DOT1, -1.0), C_ONE);
}
static void mg_bottom(ccu_ctx *c)
{
    Level &B = c->L[c->cfg.levmin];
    d_gauss_seidel(c, B, B.vec[CCU_VEC_VEL], B.vec[CCU_VEC_RHS], B.vec[CCU_VEC_AU], c->cfg.v_steps_low, 0);
}
// replicated coarse levels: vector `v` of level agg_lev, subdomain pieces -> global copy in the replica, and back
static CcuAgg agg_info(const ccu_ctx *c)
{
    const CcuComm *m = c->comm;
    return CcuAgg{ m->nproc[0], m->nproc[1], m->nproc[2], m->me[0], m->me[1], m->me[2] };
}
static int agg_gather(ccu_ctx *c, int v)
{
    ccu_ctx *g = c->coarse;
    Level &Ll = c->L[c->agg_lev], &Lg = g->L[c->agg_lev];
    if(ccu_allgather(c, Ll.vec[v], c->agg_buf, sizeof(double) * Ll.vlen())) return 1;
    LAUNCH(g, ccu_k_agg_scatter_vec, cdiv(Lg.g.nno, 128), 128, Ll.g, Lg.g, agg_info(c), (const double *)c->agg_buf, Lg.vec[v]);
    return 0;
}
static void agg_extract(ccu_ctx *c, int v)
{
    ccu_ctx *g = c->coarse;
    Level &Ll = c->L[c->agg_lev], &Lg = g->L[c->agg_lev];
    LAUNCH(g, ccu_k_agg_extract_vec, cdiv(Ll.g.nno, 128), 128, Ll.g, Lg.g, agg_info(c), (const double *)Lg.vec[v], Ll.vec[v]);
}
int ccu_agg_gather_evi(ccu_ctx *c)
{
    ccu_ctx *g = c->coarse;
    Level &Ll = c->L[c->agg_lev], &Lg = g->L[c->agg_lev];
    if(!Ll.EVI || !Lg.EVI) FAIL("agglomeration: element viscosity arrays missing");
    if(ccu_allgather(c, Ll.EVI, c->agg_buf, sizeof(float) * 8 * (size_t)Ll.g.nel)) return 1;
    LAUNCH(g, ccu_k_agg_scatter_evi, cdiv(Lg.g.nel, 128), 128, Ll.g, Lg.g, agg_info(c), (const float *)c->agg_buf, Lg.EVI);
    Lg.have_evi = true;
    return 0;
}

static void mg_inner(ccu_ctx *c, int lev)       // everything of a V-cycle on `lev` that lives below it
{
    ccu_ctx *g = c->coarse;
    if(!g)
    {
        for(int dlev = lev - 1; dlev >= c->cfg.levmin + 1; dlev--) mg_down(c, dlev, false);
        mg_bottom(c);
        for(int ulev = c->cfg.levmin + 1; ulev <= lev - 1; ulev++) mg_up(c, ulev);
        return;
    }
    const int La = c->agg_lev;                   // lev > La: levels lev-1 .. La+1 are distributed, La .. levmin replicated
    for(int dlev = lev - 1; dlev > La; dlev--) mg_down(c, dlev, false);
    agg_gather(c, CCU_VEC_RHS);
    for(int dlev = La; dlev >= g->cfg.levmin + 1; dlev--) mg_down(g, dlev, false);
    mg_bottom(g);
    for(int ulev = g->cfg.levmin + 1; ulev <= La; ulev++) mg_up(g, ulev);
    agg_extract(c, CCU_VEC_VEL);
    for(int ulev = La + 1; ulev <= lev - 1; ulev++) mg_up(c, ulev);
}

// F: in rhs, out residual; d1: out correction.  Leaves the squared residual norm in scal[S_DOT0].
static int d_multi_grid(ccu_ctx *c, double *d1, double *F)
{
    const int levmin = c->cfg.levmin, levmax = c->cfg.levmax;
    Level *L = c->L;
    d_copy(c, L[levmax].vec[CCU_VEC_FL], F, L[levmax].vlen());
    if(levmax > levmin)
    {
        d_project(c, levmax, L[levmax].vec[CCU_VEC_FL], L[levmax - 1].vec[CCU_VEC_FL], 1);
        // full multigrid below the finest level: restrict fl, bottom solve, nested V-cycles (:559-640)
        if(run_segment(c, 0, [&]() {
            // levels <= La are replicated in `g` (g == c, La == levmin - 1 ... without agglomeration everything is `c`)
            ccu_ctx *g = c->coarse ? c->coarse : c;
            const int La = c->coarse ? c->agg_lev : levmax;
            auto X = [&](int lev) { return lev <= La ? g : c; };
            for(int lev = levmax - 1; lev > levmin; lev--)
            {
                if(c->coarse && lev == La) agg_gather(c, CCU_VEC_FL);
                ccu_ctx *x = X(lev);
                d_project(x, lev, x->L[lev].vec[CCU_VEC_FL], x->L[lev - 1].vec[CCU_VEC_FL], 1);
            }
            if(c->coarse && La == levmin) agg_gather(c, CCU_VEC_FL);
            {
                Level &B = g->L[levmin];
                d_gauss_seidel(g, B, B.vec[CCU_VEC_VEL], B.vec[CCU_VEC_FL], B.vec[CCU_VEC_AU], c->cfg.v_steps_low, 0);
            }
            if(c->coarse && La == levmin) agg_extract(c, CCU_VEC_VEL);
            for(int lev = levmin + 1; lev < levmax; lev++)
            {
                ccu_ctx *x = X(lev);
                Level *XL = x->L;
                d_interp(x, lev - 1, XL[lev - 1].vec[CCU_VEC_VEL], XL[lev].vec[CCU_VEC_VEL], 1);
                d_copy(x, XL[lev].vec[CCU_VEC_RHS], XL[lev].vec[CCU_VEC_FL], XL[lev].vlen());
                for(int Vn = 1; Vn <= c->cfg.mg_cycle; Vn++)
                {
                    mg_down(x, lev, true);
                    mg_inner(x, lev);
                    mg_up(x, lev);
                }
                if(c->coarse && lev == La) agg_extract(c, CCU_VEC_VEL);
            }
        })) return 1;
        d_interp(c, levmax - 1, L[levmax - 1].vec[CCU_VEC_VEL], L[levmax].vec[CCU_VEC_VEL], 1);
        d_copy(c, L[levmax].vec[CCU_VEC_RHS], L[levmax].vec[CCU_VEC_FL], L[levmax].vlen());
        for(int Vn = 1; Vn <= c->cfg.mg_cycle; Vn++)
        {
            mg_down(c, levmax, true);
            if(run_segment(c, 1, [&]() { mg_inner(c, levmax); })) return 1;
            mg_up(c, levmax);
        }
    }
    else
    {   // single level: the "multigrid" is the bottom smoother (:572-574) and res = F - AU
        Level &B = L[levmin];
        d_gauss_seidel(c, B, B.vec[CCU_VEC_VEL], B.vec[CCU_VEC_FL], B.vec[CCU_VEC_AU], c->cfg.v_steps_low, 0);
        d_waxpby(c, B.vlen(), B.vec[CCU_VEC_RES], B.vec[CCU_VEC_FL], B.vec[CCU_VEC_AU], C_ONE, C_MINUS);
    }
    d_copy(c, F, L[levmax].vec[CCU_VEC_RES], L[levmax].vlen());
    d_copy(c, d1, L[levmax].vec[CCU_VEC_VEL], L[levmax].vlen());
    d_dot3m(c, &L[levmax], L[levmax].vlen(), F, F, S_DOT0);
    return 0;
}

// solve_del2_u (General_matrix_functions.c:368-520), multigrid branch.  d0 out, F in (device).
static int d_solve_del2_u(ccu_ctx *c, double *d0, const double *F, double acc, int *valid, int *cycles_out)
{
    Level &L = c->L[c->cfg.levmax];
    double *r = L.vec[CCU_VEC_T0], *D1 = L.vec[CCU_VEC_T1];
    const double gneq = c->comm ? (double)c->comm->gneq : (double)L.g.neq;
    d_copy(c, r, F, L.vlen());
    d_zero(c, d0, L.vlen());
    d_dot3m(c, &L, L.vlen(), r, r, S_DOT0);
    double rr;
    if(read_scal(c, S_DOT0, 1, &rr)) return 1;
    double residual = sqrt(rr / gneq);
    const double r0 = residual;
    acc = fmax(acc, r0 * c->cfg.accuracy);
    *valid = (residual < acc) ? 0 : 1;
    int count = 0;
    while(residual > acc)
    {
        if(d_multi_grid(c, D1, r)) return 1;
        d_axpby(c, L.vlen(), d0, D1, C_ONE, C_ONE);
        if(read_scal(c, S_DOT0, 1, &rr)) return 1;
        residual = sqrt(rr / gneq);
        count++;
        if(!(residual == residual) || count > 500) { g_ccu_err = "solve_del2_u: multigrid diverged or stalled"; return 3; }
    }
    if(cycles_out) *cycles_out = count;
    return 0;
}

// conj_grad (General_matrix_functions.c:661-770): Jacobi-preconditioned CG on K at level `lev` (the reference's
// Solver=cgrad path).  The loop condition and alpha's zero test branch on global reductions, so the scalars come back
// to the host each iteration exactly where the reference's control flow needs them.  Work vectors: the level's
// multigrid slots (d0 = VEL, r = RES, z = FL, p = DEL_VEL, Ap = AU; the reference's r0/r2, z0, p1 shuffles are in place).
static int d_conj_grad(ccu_ctx *c, int lev, const double *F, double acc, int *cycles, double *residual_out)
{
    if(c->multi()) FAIL("conj_grad: single-subdomain contexts only");
    Level &L = c->L[lev];
    const size_t nv = L.vlen();
    double *d0 = L.vec[CCU_VEC_VEL], *r = L.vec[CCU_VEC_RES], *z = L.vec[CCU_VEC_FL], *p = L.vec[CCU_VEC_DEL_VEL], *Ap = L.vec[CCU_VEC_AU];
    const double neq = (double)L.g.neq;
    const int steps = *cycles;
    double h[2], r1z1, r0z0 = 0.0;
    d_copy(c, r, F, nv);
    d_zero(c, d0, nv);
    d_dot3m(c, &L, nv, r, r, S_DOT0);
    if(read_scal(c, S_DOT0, 1, h)) return 1;
    double residual = sqrt(h[0] / neq);
    if(residual == 0.0) FAIL("conj_grad: initial residual is zero");        // the reference asserts (:707)
    int count = 0;
    while((residual > acc && count < steps) || count == 0)
    {
        LAUNCH(c, ccu_k_mul, min(cdiv(nv, 256), 148u * 16u), 256, nv, z, L.BI, r);
        d_dot3m(c, &L, nv, r, z, S_DOT0);
        if(read_scal(c, S_DOT0, 1, h)) return 1;
        r1z1 = h[0];
        if(count == 0) d_copy(c, p, z, nv);
        else d_axpby(c, nv, p, z, C_ONE, coef(nullptr, nullptr, r1z1 / r0z0));     // p = z + beta p
        r0z0 = r1z1;
        d_matvec(c, L, p, Ap, 1);
        d_dot3m(c, &L, nv, p, Ap, S_DOT0);
        if(read_scal(c, S_DOT0, 1, h)) return 1;
        const double alpha = (h[0] == 0.0) ? 1.0e-3 : r1z1 / h[0];
        d_axpby(c, nv, d0, p, coef(nullptr, nullptr, alpha), C_ONE);
        d_axpby(c, nv, r, Ap, coef(nullptr, nullptr, -alpha), C_ONE);
        d_dot3m(c, &L, nv, r, r, S_DOT0);
        if(read_scal(c, S_DOT0, 1, h)) return 1;
        residual = sqrt(h[0] / neq);
        count++;
    }
    *cycles = count;
    d_strip(c, L, d0);
    if(residual_out) *residual_out = residual;
    return 0;
}

static void d_div_u(ccu_ctx *c, Level &L, const double *U, double *divU)
{
    LAUNCH(c, ccu_k_div_u, cdiv(L.g.nel, 128), 128, L.g, L.elt_delT, U, divU);
}
static void d_grad_p(ccu_ctx *c, Level &L, const double *P, double *gradP)
{
    LAUNCH(c, ccu_k_grad_p, cdiv(8 * (size_t)L.g.NC, 128), 128, L.g, L.elt_delT, L.flags, P, gradP);
    if(c->multi()) ccu_halo_sum_vec(c, (int)(&L - c->L), gradP);      // exchange_id_d20 (Element_calculations.c:764)
}

// solve_Ahat_p_fhat (Stokes_flow_Incomp.c:295-497) on resident V (= vec U), P, F.
static int d_solve_Ahat_p_fhat(ccu_ctx *c, double imp, int *steps_max, float *residual_out, double *hist)
{
    Level &L = c->L[c->cfg.levmax];
    const size_t nv = L.vlen(), np = (size_t)L.g.npno;
    const long long ineq = c->comm ? c->comm->gneq : (long long)L.g.neq, inpno = c->comm ? c->comm->gnpno : (long long)L.g.npno;
    const double gneq = (double)ineq, gnpno = (double)inpno;
    double *V = L.vec[CCU_VEC_U], *F = L.vec[CCU_VEC_F], *Ah = c->uzAh, *u1 = c->uzU1;
    double *r0 = c->r0, *r1 = c->r1, *r2 = c->r2, *z0 = c->z0, *z1 = c->z1, *s1 = c->s1, *s2 = c->s2, *P = c->P, *pAh = c->pAh;
    double h[8];
    int valid = 1;

    d_dot3m(c, &L, nv, F, F, S_DOT0);
    if(read_scal(c, S_DOT0, 1, h)) return 1;
    const double v_res = sqrt(h[0] / gneq);

    d_grad_p(c, L, P, Ah);
    d_matvec(c, L, V, u1, 1);
    d_waxpby(c, nv, Ah, F, Ah, C_ONE, C_MINUS);         // Ah = F - gradP
    d_axpby(c, nv, Ah, u1, C_MINUS, C_ONE);             //    - K V
    d_strip(c, L, Ah);
    if(d_solve_del2_u(c, u1, Ah, imp * v_res, &valid, nullptr)) return 1;
    d_strip(c, L, u1);
    d_axpby(c, nv, V, u1, C_ONE, C_ONE);
    d_div_u(c, L, V, r1);
    d_dot3(c, np, r1, r1, S_DOT0);
    if(read_scal(c, S_DOT0, 1, h)) return 1;
    const double residual = sqrt(h[0] / gnpno);
    if(residual_out) *residual_out = (float)residual;

    int count = 0;
    float dpressure = 1.0f, dvelocity = 1.0f;
    while(count < *steps_max && (dpressure >= imp || dvelocity >= imp))
    {
        LAUNCH(c, ccu_k_mul, min(cdiv(np, 256), 148u * 16u), 256, np, z1, L.BPI, r1);
        if(count == 0)
        {
            d_dot3(c, np, r1, z1, S_R1Z1);
            d_copy(c, s2, z1, np);
        }
        else
        {
            d_dot3(c, np, r1, z1, S_R1Z1, r0, z0, S_R0Z0);
            d_waxpby(c, np, s2, z1, s1, C_ONE, coef(c->scal + S_R1Z1, c->scal + S_R0Z0, 1.0));   // s2 = z1 + delta*s1
        }
        d_grad_p(c, L, s2, Ah);
        if(d_solve_del2_u(c, u1, Ah, imp * v_res, &valid, nullptr)) return 1;
        d_strip(c, L, u1);
        d_div_u(c, L, u1, pAh);
        d_dot3(c, np, s2, pAh, S_S2AH);
        // alpha = r1dotz1 / s2dotAhat, or 0 when the velocity solve was a no-op (:422-425)
        const CcuCoef alpha = valid ? coef(c->scal + S_R1Z1, c->scal + S_S2AH, 1.0) : C_ZERO;
        const CcuCoef malpha = valid ? coef(c->scal + S_R1Z1, c->scal + S_S2AH, -1.0) : C_ZERO;
        d_waxpby(c, np, r2, r1, pAh, C_ONE, malpha);
        d_axpby(c, np, P, s2, alpha, C_ONE);
        d_axpby(c, nv, V, u1, malpha, C_ONE);
        d_div_u(c, L, V, pAh);
        d_dot3m(c, &L, nv, V, V, S_VDOTV, u1, u1, S_U1U1);
        d_dot3(c, np, P, P, S_PDOTP, pAh, pAh, S_AHAH, s2, s2, S_S2S2);
        if(read_scal(c, S_R1Z1, S_U1U1 - S_R1Z1 + 1, h)) return 1;
        const double r1z1 = h[0], s2ah = h[S_S2AH - S_R1Z1], vdotv = h[S_VDOTV - S_R1Z1], pdotp = h[S_PDOTP - S_R1Z1];
        const double ahah = h[S_AHAH - S_R1Z1], s2s2 = h[S_S2S2 - S_R1Z1], u1u1 = h[S_U1U1 - S_R1Z1];
        const double al = valid ? r1z1 / s2ah : 0.0;
        const float incomp = (float)sqrt((double)(ineq / inpno) * (1.0e-32 + ahah / (1.0e-32 + (double)(float)vdotv)));
        dpressure = (float)(al * sqrt(s2s2 / (1.0e-32 + (double)(float)pdotp)));
        dvelocity = (float)(al * sqrt(u1u1 / (1.0e-32 + (double)(float)vdotv)));
        if(hist)
        {
            hist[5 * count + 0] = sqrt((double)(float)vdotv / gneq); hist[5 * count + 1] = dvelocity; hist[5 * count + 2] = incomp;
            hist[5 * count + 3] = sqrt((double)(float)pdotp / gnpno); hist[5 * count + 4] = dpressure;
        }
        count++;
        double *sh;
        sh = s1; s1 = s2; s2 = sh;
        sh = r0; r0 = r1; r1 = r2; r2 = sh;
        sh = z0; z0 = z1; z1 = sh;
    }
    c->s1 = s1; c->s2 = s2; c->r0 = r0; c->r1 = r1; c->r2 = r2; c->z0 = z0; c->z1 = z1;
    *steps_max = count;
    return 0;
}

// ------------------------------------------------------------------ profiling
int ccu_profile_enable(ccu_ctx *c, int on) { if(!c) FAIL("null context"); c->prof_on = on != 0; return 0; }
static int prof_fold(ccu_ctx *c)
{
    SYNC(c);
    for(auto &r : c->prof_recs)
    {
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
        c->prof_ms[r.cls] += ms; c->prof_n[r.cls] += r.n;
        c->prof_pool.push_back(r.e0); c->prof_pool.push_back(r.e1);
    }
    c->prof_recs.clear();
    return 0;
}
int ccu_profile_read(ccu_ctx *c, int cls, double *ms_total, long long *launches)
{
    if(!c) FAIL("null context");
    if(cls < 0 || cls >= CCU_PROF_COUNT) FAIL("profile_read: bad class");
    if(prof_fold(c)) return 1;
    if(ms_total) *ms_total = c->prof_ms[cls];
    if(launches) *launches = c->prof_n[cls];
    return 0;
}
int ccu_profile_reset(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    if(prof_fold(c)) return 1;
    for(int i = 0; i < CCU_PROF_COUNT; i++) { c->prof_ms[i] = 0.0; c->prof_n[i] = 0; }
    return 0;
}

// ------------------------------------------------------------------ general_stokes_solver (Drive_solvers.c:45-162)
int ccu_general_stokes_solver(ccu_ctx *c, const float *T, const float *buoyancy, int rebuild, int augmented_Lagr, double augmented,
                              int precondition, int guess, double *U, double *P, int *iterations_out, float *residual_out)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(T && ccu_set_temperature(c, T)) return 1;
    // the K.VB force term reads the viscosity as it stands BEFORE this call's update (Drive_solvers.c:107 comes ahead of :124);
    // at the very first call that is the one common_initial_fields evaluated from the initial state (Instructions.c:1233)
    if(c->have_vb && !L.have_evi) { if(ccu_get_system_viscosity(c)) return 1; }
    if(ccu_assemble_forces(c, buoyancy, nullptr)) return 1;
    if(rebuild)
    {
        CcuProfScope ps(c, CCU_PROF_BUILD, true);
        if(c->visc.tdepv || c->visc.sdepv || c->visc.cdepv || c->visc.bdepv || !L.have_evi) { if(ccu_get_system_viscosity(c)) return 1; }
        if(ccu_construct_stiffness_B_matrix(c, augmented_Lagr, augmented, precondition)) return 1;
    }
    if(!L.have_K || !L.have_flags || !L.have_p) FAIL("general_stokes_solver: operator not built");
    if(guess == 0)
    {
        CK(cudaMemsetAsync(L.vec[CCU_VEC_U], 0, sizeof(double) * L.vlen(), c->st));
        CK(cudaMemsetAsync(c->P, 0, sizeof(double) * L.g.npno, c->st));
    }
    else if(guess == 1)
    {
        if(!U || !P) FAIL("general_stokes_solver: guess == 1 needs U and P");
        if(vec_h2d(c, L, U, L.vec[CCU_VEC_U])) return 1;
        CK(cudaMemcpyAsync(c->P, P, sizeof(double) * L.g.npno, cudaMemcpyHostToDevice, c->st));
    }
    if(c->have_vb) { if(ccu_conform_velocity_bcs(c)) return 1; }   // velocities_conform_bcs (Boundary_conditions.c:993)
    else d_strip(c, L, L.vec[CCU_VEC_U]);       // the same with zero imposed velocities
    int steps = c->cfg.p_iterations;
    if(d_solve_Ahat_p_fhat(c, c->cfg.accuracy, &steps, residual_out, nullptr)) return 1;
    if(c->visc.sdepv || c->visc.bdepv)      // need_to_iterate (Drive_solvers.c)
    {   // E->V of the solve just done (solve_constrained_flow_iterative ends with v_from_vector, before any damping): what the next
        // visc_from_S / visc_from_B call reads
        if(ccu_v_from_vector(c, nullptr)) return 1;
        // stress-dependent viscosity: viscosity <-> velocity iteration (Drive_solvers.c:120-159) with the damping of :137-141
        if(!c->sdepv_oldU)
        {   // oldU persists from timestep to timestep (a static of the reference's, Drive_solvers.c:53), zero at the first call
            CK(cudaMalloc(&c->sdepv_oldU, sizeof(double) * L.vlen())); CK(cudaMalloc(&c->sdepv_dU, sizeof(double) * L.vlen()));
            CK(cudaMemsetAsync(c->sdepv_oldU, 0, sizeof(double) * L.vlen(), c->st));
        }
        double *oldU = c->sdepv_oldU, *dU = c->sdepv_dU;
        const double alpha = (double)c->visc.sdepv_iter_damp;
        const bool damp = fabs(alpha - 1.0) > 1e-7;
        int count = 1, total = steps;
        for(;;)
        {
            if(damp) d_axpby(c, L.vlen(), L.vec[CCU_VEC_U], oldU, coef(nullptr, nullptr, 1.0 - alpha), coef(nullptr, nullptr, alpha));   // U = alpha U + (1 - alpha) oldU
            d_waxpby(c, L.vlen(), dU, L.vec[CCU_VEC_U], oldU, C_ONE, C_MINUS);
            d_copy(c, oldU, L.vec[CCU_VEC_U], L.vlen());
            d_dot3m(c, &L, L.vlen(), L.vec[CCU_VEC_U], L.vec[CCU_VEC_U], S_DOT1, dU, dU, S_DOT2);
            double uu[2];
            if(read_scal(c, S_DOT1, 2, uu)) return 1;
            const double Umag = sqrt(uu[0]);
            double dmag = sqrt(uu[1]);
            if(Umag != 0.0) dmag /= Umag;
            c->sdepv_last_misfit = dmag; c->sdepv_last_count = count;
            if(!(dmag > (double)c->visc.sdepv_misfit) || count >= c->visc.sdepv_max_iter) break;
            {
                CcuProfScope ps(c, CCU_PROF_BUILD, true);
                if(ccu_get_system_viscosity(c)) return 1;
                if(ccu_construct_stiffness_B_matrix(c, augmented_Lagr, augmented, precondition)) return 1;
            }
            steps = c->cfg.p_iterations;
            if(d_solve_Ahat_p_fhat(c, c->cfg.accuracy, &steps, residual_out, nullptr)) return 1;
            if(ccu_v_from_vector(c, nullptr)) return 1;
            total += steps; count++;
        }
        steps = total;
    }
    if(iterations_out) *iterations_out = steps;
    if(P) CK(cudaMemcpyAsync(P, c->P, sizeof(double) * L.g.npno, cudaMemcpyDeviceToHost, c->st));
    if(U) return vec_d2h(c, L, L.vec[CCU_VEC_U], U);
    SYNC(c);
    return 0;
}

// ------------------------------------------------------------------ C ABI: device-resident forms
static double *vecp(ccu_ctx *c, int lev, int v)
{
    if(v < 0 || v >= CCU_VEC_COUNT) return nullptr;
    return c->L[lev].vec[v];
}
#define VEC(ptr, lev, v) double *ptr = vecp(c, lev, v); if(!ptr) FAIL("bad vector id for this level")

int ccu_vec_upload(ccu_ctx *c, int lev, int v, const double *host)
{
    if(ccu_check_lev(c, lev)) return 2;
    VEC(p, lev, v);
    if(vec_h2d(c, c->L[lev], host, p)) return 1;
    SYNC(c);
    return 0;
}
int ccu_vec_download(ccu_ctx *c, int lev, int v, double *host)
{
    if(ccu_check_lev(c, lev)) return 2;
    VEC(p, lev, v);
    return vec_d2h(c, c->L[lev], p, host);
}
int ccu_pvec_upload(ccu_ctx *c, const double *host)
{
    if(!c) FAIL("null context");
    CK(cudaMemcpyAsync(c->P, host, sizeof(double) * c->L[c->cfg.levmax].g.npno, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    return 0;
}
int ccu_pvec_download(ccu_ctx *c, double *host)
{
    if(!c) FAIL("null context");
    CK(cudaMemcpyAsync(host, c->P, sizeof(double) * c->L[c->cfg.levmax].g.npno, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
int ccu_dev_matvec(ccu_ctx *c, int lev, int vu, int vAu, int strip)
{
    if(ccu_check_lev(c, lev)) return 2;
    VEC(u, lev, vu); VEC(Au, lev, vAu);
    d_matvec(c, c->L[lev], u, Au, strip);
    CK(cudaGetLastError());
    return 0;
}
int ccu_dev_gauss_seidel(ccu_ctx *c, int lev, int vd, int vF, int vAd, int cycles, int guess)
{
    if(ccu_check_lev(c, lev)) return 2;
    VEC(d0, lev, vd); VEC(F, lev, vF); VEC(Ad, lev, vAd);
    d_gauss_seidel(c, c->L[lev], d0, F, Ad, cycles, guess);
    CK(cudaGetLastError());
    return 0;
}
int ccu_dev_relax_sweeps(ccu_ctx *c, int lev, int vd, int vF, int cycles)
{
    if(ccu_check_lev(c, lev)) return 2;
    VEC(d0, lev, vd); VEC(F, lev, vF);
    d_relax_sweeps(c, c->L[lev], d0, F, cycles);
    CK(cudaGetLastError());
    return 0;
}
int ccu_dev_multi_grid(ccu_ctx *c, int vd1, int vF, double *residual_out)
{
    if(!c) FAIL("null context");
    const int lev = c->cfg.levmax;
    VEC(d1, lev, vd1); VEC(F, lev, vF);
    if(d_multi_grid(c, d1, F)) return 1;
    double rr;
    if(read_scal(c, S_DOT0, 1, &rr)) return 1;
    if(residual_out) *residual_out = sqrt(rr / (double)c->L[lev].g.neq);
    return 0;
}
int ccu_dev_solve_Ahat_p_fhat(ccu_ctx *c, double imp, int *steps_max, float *residual_out, double *hist)
{
    if(!c) FAIL("null context");
    return d_solve_Ahat_p_fhat(c, imp, steps_max, residual_out, hist);
}

// ------------------------------------------------------------------ C ABI: host-vector forms
int ccu_n_assemble_del2_u(ccu_ctx *c, int lev, const double *u, double *Au, int strip)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    if(vec_h2d(c, L, u, L.vec[CCU_VEC_VEL])) return 1;
    d_matvec(c, L, L.vec[CCU_VEC_VEL], L.vec[CCU_VEC_AU], strip);
    return vec_d2h(c, L, L.vec[CCU_VEC_AU], Au);
}
int ccu_gauss_seidel(ccu_ctx *c, int lev, double *d0, const double *F, double *Ad, int cycles, int guess)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    if(guess && vec_h2d(c, L, d0, L.vec[CCU_VEC_VEL])) return 1;
    if(vec_h2d(c, L, F, L.vec[CCU_VEC_RHS])) return 1;
    d_gauss_seidel(c, L, L.vec[CCU_VEC_VEL], L.vec[CCU_VEC_RHS], L.vec[CCU_VEC_AU], cycles, guess);
    if(vec_d2h(c, L, L.vec[CCU_VEC_VEL], d0)) return 1;
    return vec_d2h(c, L, L.vec[CCU_VEC_AU], Ad);
}
int ccu_project_vector(ccu_ctx *c, int lev, const double *AU, double *AD)
{
    if(ccu_check_lev(c, lev) || ccu_check_lev(c, lev - 1)) return 2;
    if(vec_h2d(c, c->L[lev], AU, c->L[lev].vec[CCU_VEC_RES])) return 1;
    d_project(c, lev, c->L[lev].vec[CCU_VEC_RES], c->L[lev - 1].vec[CCU_VEC_RHS], 0);
    return vec_d2h(c, c->L[lev - 1], c->L[lev - 1].vec[CCU_VEC_RHS], AD);
}
int ccu_interp_vector(ccu_ctx *c, int lev, const double *AD, double *AU)
{
    if(ccu_check_lev(c, lev) || ccu_check_lev(c, lev + 1)) return 2;
    if(vec_h2d(c, c->L[lev], AD, c->L[lev].vec[CCU_VEC_VEL])) return 1;
    d_interp(c, lev, c->L[lev].vec[CCU_VEC_VEL], c->L[lev + 1].vec[CCU_VEC_DEL_VEL], 0);
    return vec_d2h(c, c->L[lev + 1], c->L[lev + 1].vec[CCU_VEC_DEL_VEL], AU);
}
int ccu_strip_bcs_from_residual(ccu_ctx *c, int lev, double *res)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    if(vec_h2d(c, L, res, L.vec[CCU_VEC_RES])) return 1;
    d_strip(c, L, L.vec[CCU_VEC_RES]);
    return vec_d2h(c, L, L.vec[CCU_VEC_RES], res);
}
int ccu_assemble_div_u(ccu_ctx *c, int lev, const double *U, double *divU)
{
    if(ccu_check_lev(c, lev)) return 2;
    if(lev != c->cfg.levmax) FAIL("div_u: only the finest level is resident");
    Level &L = c->L[lev];
    if(vec_h2d(c, L, U, L.vec[CCU_VEC_VEL])) return 1;
    d_div_u(c, L, L.vec[CCU_VEC_VEL], c->pAh);
    CK(cudaMemcpyAsync(divU, c->pAh, sizeof(double) * L.g.npno, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
int ccu_assemble_grad_p(ccu_ctx *c, int lev, const double *P, double *gradP)
{
    if(ccu_check_lev(c, lev)) return 2;
    if(lev != c->cfg.levmax) FAIL("grad_p: only the finest level is resident");
    Level &L = c->L[lev];
    CK(cudaMemcpyAsync(c->pAh, P, sizeof(double) * L.g.npno, cudaMemcpyHostToDevice, c->st));
    d_grad_p(c, L, c->pAh, L.vec[CCU_VEC_AU]);
    return vec_d2h(c, L, L.vec[CCU_VEC_AU], gradP);
}
int ccu_global_vdot(ccu_ctx *c, int lev, const double *A, const double *B, double *out)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    if(vec_h2d(c, L, A, L.vec[CCU_VEC_VEL])) return 1;
    SYNC(c);
    if(vec_h2d(c, L, B, L.vec[CCU_VEC_RES])) return 1;
    d_dot3m(c, &L, L.vlen(), L.vec[CCU_VEC_VEL], L.vec[CCU_VEC_RES], S_TMP);
    return read_scal(c, S_TMP, 1, out);
}
int ccu_global_pdot(ccu_ctx *c, int lev, const double *A, const double *B, double *out)
{
    if(ccu_check_lev(c, lev)) return 2;
    if(lev != c->cfg.levmax) FAIL("pdot: only the finest level is resident");
    const size_t np = (size_t)c->L[lev].g.npno;
    CK(cudaMemcpyAsync(c->pAh, A, sizeof(double) * np, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->z1, B, sizeof(double) * np, cudaMemcpyHostToDevice, c->st));
    d_dot3(c, np, c->pAh, c->z1, S_TMP);
    return read_scal(c, S_TMP, 1, out);
}
int ccu_multi_grid(ccu_ctx *c, double *d1, double *F, double *residual_out)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(vec_h2d(c, L, F, L.vec[CCU_VEC_T0])) return 1;
    if(ccu_dev_multi_grid(c, CCU_VEC_T1, CCU_VEC_T0, residual_out)) return 1;
    if(vec_d2h(c, L, L.vec[CCU_VEC_T0], F)) return 1;
    return vec_d2h(c, L, L.vec[CCU_VEC_T1], d1);
}
int ccu_conj_grad(ccu_ctx *c, int lev, double *d0, const double *F, double acc, int *cycles, double *residual_out)
{
    if(ccu_check_lev(c, lev)) return 2;
    if(!cycles) FAIL("conj_grad: cycles is null");
    Level &L = c->L[lev];
    if(vec_h2d(c, L, F, L.vec[CCU_VEC_RHS])) return 1;
    if(d_conj_grad(c, lev, L.vec[CCU_VEC_RHS], acc, cycles, residual_out)) return 1;
    return vec_d2h(c, L, L.vec[CCU_VEC_VEL], d0);
}
int ccu_e_assemble_del2_u(ccu_ctx *c, int lev, const double *u, double *Au, int strip_bcs)
{   // the element-by-element product equals the node-stored one (assemble_del2_u dispatch, Element_calculations.c:480-488)
    return ccu_n_assemble_del2_u(c, lev, u, Au, strip_bcs);
}
int ccu_solve_del2_u(ccu_ctx *c, double *d0, const double *F, double acc, int *valid_out, int *cycles_out)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(vec_h2d(c, L, F, L.vec[CCU_VEC_T2])) return 1;
    int valid = 0, cyc = 0;
    if(d_solve_del2_u(c, c->uzU1, L.vec[CCU_VEC_T2], acc, &valid, &cyc)) return 1;
    if(valid_out) *valid_out = valid;
    if(cycles_out) *cycles_out = cyc;
    return vec_d2h(c, L, c->uzU1, d0);
}
int ccu_solve_Ahat_p_fhat(ccu_ctx *c, double *V, double *P, const double *F, double imp, int *steps_max, float *residual_out, double *hist)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(!L.have_K || !L.have_flags || !L.have_p) FAIL("solve_Ahat_p_fhat: operator not uploaded");
    if(vec_h2d(c, L, V, L.vec[CCU_VEC_U])) return 1;
    SYNC(c);
    if(vec_h2d(c, L, F, L.vec[CCU_VEC_F])) return 1;
    CK(cudaMemcpyAsync(c->P, P, sizeof(double) * L.g.npno, cudaMemcpyHostToDevice, c->st));
    if(d_solve_Ahat_p_fhat(c, imp, steps_max, residual_out, hist)) return 1;
    CK(cudaMemcpyAsync(P, c->P, sizeof(double) * L.g.npno, cudaMemcpyDeviceToHost, c->st));
    return vec_d2h(c, L, L.vec[CCU_VEC_U], V);
}
