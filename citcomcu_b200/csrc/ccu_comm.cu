// ccu_comm.cu -- subdomain-per-GPU communication: duplicated-node tables, NCCL binding, halo-sum and
// allreduce primitives (see ccu_comm.cuh).
#include "ccu_ctx.cuh"
#include "ccu_comm.cuh"
#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

// ------------------------------------------------------------------ host tables
namespace {
struct Region { int lo[3], n[3]; };   // per axis (y=i, x=j, z=k): first index and extent of the shared region
// region of local nodes shared with the neighbour at block offset o = (ox, oy, oz)
Region region_for(const int o[3], int nox, int noy, int noz)
{
    Region r;
    const int ext[3] = { noy, nox, noz };           // axis order i (y), j (x), k (z)
    const int off[3] = { o[1], o[0], o[2] };
    for(int a = 0; a < 3; a++)
    {
        if(off[a] == 0) { r.lo[a] = 0; r.n[a] = ext[a]; }
        else { r.lo[a] = off[a] < 0 ? 0 : ext[a] - 1; r.n[a] = 1; }
    }
    return r;
}
}

void ccu_build_halo_host(const int nproc[3], const int me[3], int nox, int noy, int noz, CcuHaloHost &h)
{
    h = CcuHaloHost();
    const int my_rank = ccu_rank_of(nproc, me[0], me[1], me[2]);
    struct Nb { int rank, o[3]; };
    std::vector<Nb> nbs;
    for(int oy = -1; oy <= 1; oy++)
        for(int ox = -1; ox <= 1; ox++)
            for(int oz = -1; oz <= 1; oz++)
            {
                if(!ox && !oy && !oz) continue;
                const int x = me[0] + ox, y = me[1] + oy, z = me[2] + oz;
                if(x < 0 || x >= nproc[0] || y < 0 || y >= nproc[1] || z < 0 || z >= nproc[2]) continue;
                nbs.push_back({ ccu_rank_of(nproc, x, y, z), { ox, oy, oz } });
            }
    std::sort(nbs.begin(), nbs.end(), [](const Nb &a, const Nb &b) { return a.rank < b.rank; });
    const int nno = nox * noy * noz;
    // compact numbering of the duplicated (OFFSIDE) nodes, natural order
    std::vector<int> compact(nno, -1);
    auto on_shared_face = [&](int i, int j, int k) {
        return (i == 0 && me[1] > 0) || (i == noy - 1 && me[1] < nproc[1] - 1) || (j == 0 && me[0] > 0) ||
               (j == nox - 1 && me[0] < nproc[0] - 1) || (k == 0 && me[2] > 0) || (k == noz - 1 && me[2] < nproc[2] - 1);
    };
    for(int i = 0; i < noy; i++)
        for(int j = 0; j < nox; j++)
            for(int k = 0; k < noz; k++)
                if(on_shared_face(i, j, k))
                {
                    const int n = k + noz * (j + nox * i);
                    compact[n] = (int)h.sh_n.size();
                    h.sh_n.push_back(n);
                }
    // send segments: the region shared with each neighbour, natural order (both sides enumerate the same
    // physical nodes in the same order because the free axes run in the same order on both)
    int off = 0;
    for(const Nb &nb : nbs)
    {
        const Region r = region_for(nb.o, nox, noy, noz);
        h.nb_rank.push_back(nb.rank);
        h.nb_off.push_back(off);
        const int cnt = r.n[0] * r.n[1] * r.n[2];
        h.nb_cnt.push_back(cnt);
        for(int a = 0; a < r.n[0]; a++)
            for(int b = 0; b < r.n[1]; b++)
                for(int c = 0; c < r.n[2]; c++)
                    h.send_t.push_back(compact[(r.lo[2] + c) + noz * ((r.lo[1] + b) + nox * (r.lo[0] + a))]);
        off += cnt;
    }
    // contributions per duplicated node in ascending rank order
    h.owned.assign(nno, 1);
    h.sh_ptr.push_back(0);
    for(size_t t = 0; t < h.sh_n.size(); t++)
    {
        const int n = h.sh_n[t];
        const int k = n % noz, j = (n / noz) % nox, i = n / (noz * nox);
        bool self_done = false;
        for(size_t q = 0; q < nbs.size(); q++)
        {
            const Region r = region_for(nbs[q].o, nox, noy, noz);
            const int c[3] = { i, j, k };
            bool in = true;
            for(int a = 0; a < 3; a++) in = in && c[a] >= r.lo[a] && c[a] < r.lo[a] + r.n[a];
            if(!in) continue;
            if(!self_done && nbs[q].rank > my_rank) { h.sh_src.push_back(-1); self_done = true; }
            if(nbs[q].rank < my_rank) h.owned[n] = 0;
            const int pos = ((i - r.lo[0]) * r.n[1] + (j - r.lo[1])) * r.n[2] + (k - r.lo[2]);
            h.sh_src.push_back(h.nb_off[q] + pos);
        }
        if(!self_done) h.sh_src.push_back(-1);
        h.sh_ptr.push_back((int)h.sh_src.size());
    }
}

// host-only C entry points (tests exercise the tables without a GPU)
extern "C" int ccu_halo_sizes(const int nproc[3], const int me[3], int nox, int noy, int noz, int sizes[4])
{
    CcuHaloHost h;
    ccu_build_halo_host(nproc, me, nox, noy, noz, h);
    sizes[0] = (int)h.nb_rank.size(); sizes[1] = (int)h.send_t.size(); sizes[2] = (int)h.sh_n.size(); sizes[3] = (int)h.sh_src.size();
    return 0;
}
extern "C" int ccu_halo_tables(const int nproc[3], const int me[3], int nox, int noy, int noz, int *nb_rank, int *nb_off, int *nb_cnt,
                               int *send_n, int *sh_n, int *sh_ptr, int *sh_src, unsigned char *owned)
{
    CcuHaloHost h;
    ccu_build_halo_host(nproc, me, nox, noy, noz, h);
    std::copy(h.nb_rank.begin(), h.nb_rank.end(), nb_rank);
    std::copy(h.nb_off.begin(), h.nb_off.end(), nb_off);
    std::copy(h.nb_cnt.begin(), h.nb_cnt.end(), nb_cnt);
    for(size_t e = 0; e < h.send_t.size(); e++) send_n[e] = h.sh_n[h.send_t[e]];
    std::copy(h.sh_n.begin(), h.sh_n.end(), sh_n);
    std::copy(h.sh_ptr.begin(), h.sh_ptr.end(), sh_ptr);
    std::copy(h.sh_src.begin(), h.sh_src.end(), sh_src);
    std::copy(h.owned.begin(), h.owned.end(), owned);
    return 0;
}

// ------------------------------------------------------------------ NCCL, bound at run time
namespace {
struct NcclApi
{
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if(g_nccl.handle) return 0;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for(const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if(h) break; }
    if(!h) FAIL(std::string("cannot load libnccl.so.2: ") + dlerror());
#define BIND(field, sym) do { *(void **)(&g_nccl.field) = dlsym(h, sym); if(!g_nccl.field) FAIL(std::string("libnccl lacks ") + sym); } while(0)
    BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
    BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv"); BIND(AllReduce, "ncclAllReduce"); BIND(AllGather, "ncclAllGather");
    BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    g_nccl.handle = h;
    return 0;
}
}
#define NK(call) do { ncclResult_t r_ = (call); if(r_ != ncclSuccess) { \
    g_ccu_err = std::string(#call) + ": " + g_nccl.GetErrorString(r_) + " @" + __FILE__ + ":" + std::to_string(__LINE__); return 1; } } while(0)

int ccu_comm_unique_id(char *out128)
{
    if(load_nccl()) return 1;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return 0;
}

// ------------------------------------------------------------------ kernels
// pack: out[e*ND + d] = src[d*stride + idx[e]]
template <class T, int ND>
__global__ void ccu_k_halo_pack(const int n, const int *__restrict__ idx, const T *__restrict__ src, const size_t stride, T *out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    const int s = idx[e];
#pragma unroll
    for(int d = 0; d < ND; d++) out[(size_t)e * ND + d] = src[(size_t)d * stride + s];
}
// unpack: every duplicated node := sum of all owners' values in ascending rank order (the local value read from
// `own`, laid out like the destination unless own_compact, then own[d*n + t])
template <class T, int ND>
__global__ void ccu_k_halo_unpack(const int n, const int *__restrict__ idx, const int *__restrict__ ptr, const int *__restrict__ src,
                                  const T *__restrict__ recv, T *vec, const size_t stride)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n) return;
    const int s = idx[t];
#pragma unroll
    for(int d = 0; d < ND; d++)
    {
        T acc = 0;
        for(int e = ptr[t]; e < ptr[t + 1]; e++)
        {
            const int q = src[e];
            const T v = (q < 0) ? vec[(size_t)d * stride + s] : recv[(size_t)q * ND + d];
            acc = (e == ptr[t]) ? v : acc + v;
        }
        vec[(size_t)d * stride + s] = acc;
    }
}
__global__ void ccu_k_scatter_scal(const double *__restrict__ stage, double *o0, double *o1, double *o2)
{
    if(o0) *o0 = stage[0];
    if(o1) *o1 = stage[1];
    if(o2) *o2 = stage[2];
}

// ------------------------------------------------------------------ peer-memory exchange (NVLink / NVSwitch, CUDA IPC)
__global__ void ccu_k_p2p_bump(unsigned *seq) { *seq += 1u; }
// grid (segments, slices): slice `y` of the packed values for neighbour segment `x` goes straight into that neighbour's
// landing buffer (parity = sequence number & 1); the last slice of a segment to finish raises the neighbour's flag.
__global__ void __launch_bounds__(256) ccu_k_p2p_push(const int4 *__restrict__ seg, const int bytes_per_node, const char *__restrict__ sendbuf,
                                                      char *const *__restrict__ peer_land, unsigned *const *__restrict__ peer_flags,
                                                      const unsigned *__restrict__ seq, unsigned *done, const int myrank, const size_t land_half)
{
    const int4 sg = seg[blockIdx.x];
    const unsigned seqv = *seq;
    const char *srcb = sendbuf + (size_t)sg.y * bytes_per_node;
    char *dstb = peer_land[sg.x] + (size_t)(seqv & 1u) * land_half + (size_t)sg.w * bytes_per_node;
    const size_t bytes = (size_t)sg.z * bytes_per_node;
    if((((size_t)srcb | (size_t)dstb | bytes) & 15) == 0)
    {   // 128-bit stores over NVLink
        const uint4 *src = (const uint4 *)srcb; uint4 *dst = (uint4 *)dstb;
        for(size_t w = (size_t)blockIdx.y * blockDim.x + threadIdx.x; w < bytes / 16; w += (size_t)gridDim.y * blockDim.x) dst[w] = src[w];
    }
    else
    {
        const unsigned *src = (const unsigned *)srcb; unsigned *dst = (unsigned *)dstb;
        for(size_t w = (size_t)blockIdx.y * blockDim.x + threadIdx.x; w < bytes / 4; w += (size_t)gridDim.y * blockDim.x) dst[w] = src[w];
    }
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        if(atomicAdd(done + blockIdx.x, 1u) + 1u == gridDim.y)
        {
            done[blockIdx.x] = 0u;
            __threadfence_system();
            *(volatile unsigned *)(peer_flags[sg.x] + myrank) = seqv;
        }
    }
}
// every block waits until all neighbours of this level have delivered exchange `seq`, then moves its slice of the landing
// buffer into recvbuf (where the unpack / face kernels read)
__global__ void __launch_bounds__(256) ccu_k_p2p_wait(const int nseg, const int4 *__restrict__ seg, const int bytes_per_node, const int n_send,
                                                      const char *__restrict__ land, const size_t land_half, const unsigned *flags,
                                                      const unsigned *__restrict__ seq, char *recvbuf, unsigned *err)
{
    const unsigned seqv = *seq;
    if(threadIdx.x < nseg)
    {
        const volatile unsigned *f = flags + seg[threadIdx.x].x;
        unsigned spins = 0;
        while((int)(*f - seqv) < 0)
        {
            __nanosleep(32);
            if(++spins > (1u << 25)) { *err = 1u; break; }        // never in a healthy run: do not hang the device
        }
    }
    __syncthreads();
    __threadfence_system();
    const unsigned *src = (const unsigned *)(land + (size_t)(seqv & 1u) * land_half);
    unsigned *dst = (unsigned *)recvbuf;
    const size_t words = (size_t)n_send * bytes_per_node / 4;
    for(size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) dst[w] = __ldcg(src + w);
}

// Exports this rank's landing buffers and flags through CUDA IPC, gathers everybody's handles over NCCL, maps the
// neighbours' and builds the per-level segment tables (the neighbour's receive offset for my data comes from building ITS
// duplicated-node tables here: they are pure index arithmetic).  Any failure leaves p2p = false (NCCL send/recv is used).
static int ccu_p2p_setup(ccu_ctx *c, size_t max_send)
{
    CcuComm *m = c->comm;
    const int nranks = m->nranks;
    m->land_half = (sizeof(double) * 3 * std::max<size_t>(max_send, 1) + 255) & ~(size_t)255;
    CK(cudaMalloc(&m->land, 2 * m->land_half));
    CK(cudaMalloc(&m->flags, sizeof(unsigned) * nranks));
    CK(cudaMalloc(&m->seq, sizeof(unsigned)));
    CK(cudaMalloc(&m->done, sizeof(unsigned) * 32));
    CK(cudaMalloc(&m->p2p_err, sizeof(unsigned)));
    CK(cudaMemset(m->land, 0, 2 * m->land_half));
    CK(cudaMemset(m->flags, 0, sizeof(unsigned) * nranks));
    CK(cudaMemset(m->seq, 0, sizeof(unsigned)));
    CK(cudaMemset(m->done, 0, sizeof(unsigned) * 32));
    CK(cudaMemset(m->p2p_err, 0, sizeof(unsigned)));
    struct Handles { cudaIpcMemHandle_t land, flags; int ok; int pad[3]; };
    Handles mine;
    memset(&mine, 0, sizeof mine);
    mine.ok = cudaIpcGetMemHandle(&mine.land, m->land) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, m->flags) == cudaSuccess;
    cudaGetLastError();
    Handles *dev = nullptr;
    CK(cudaMalloc(&dev, sizeof(Handles) * (size_t)(nranks + 1)));
    CK(cudaMemcpy(dev, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NK(g_nccl.AllGather(dev, dev + 1, sizeof(Handles), ncclChar, (ncclComm_t)m->nccl, c->st));
    std::vector<Handles> all((size_t)nranks);
    CK(cudaMemcpyAsync(all.data(), dev + 1, sizeof(Handles) * (size_t)nranks, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    cudaFree(dev);
    bool ok = true;
    for(int r = 0; r < nranks; r++) ok = ok && all[(size_t)r].ok;
    std::vector<char *> pl((size_t)nranks, nullptr);
    std::vector<unsigned *> pf((size_t)nranks, nullptr);
    std::vector<char> is_nb((size_t)nranks, 0);
    for(int lev = c->cfg.levmin; lev <= c->cfg.levmax; lev++)
        for(int r : m->halo[lev].nb_rank) is_nb[(size_t)r] = 1;
    for(int r = 0; r < nranks && ok; r++)
    {
        if(!is_nb[(size_t)r] || r == m->rank) continue;
        void *a = nullptr, *b = nullptr;
        if(cudaIpcOpenMemHandle(&a, all[(size_t)r].land, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
           cudaIpcOpenMemHandle(&b, all[(size_t)r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
        m->ipc_open.push_back(a); m->ipc_open.push_back(b);
        pl[(size_t)r] = (char *)a; pf[(size_t)r] = (unsigned *)b;
    }
    // everybody must agree (a rank that falls back to NCCL while its neighbour pushes would wait forever)
    {
        double *flag = m->dotstage;
        const double v = ok ? 0.0 : 1.0;
        CK(cudaMemcpy(flag, &v, sizeof v, cudaMemcpyHostToDevice));
        NK(g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclMax, (ncclComm_t)m->nccl, c->st));
        double w = 0.0;
        CK(cudaMemcpyAsync(&w, flag, sizeof w, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        ok = (w == 0.0);
    }
    if(!ok) { m->p2p = false; return 0; }
    CK(cudaMalloc(&m->peer_land, sizeof(char *) * (size_t)nranks));
    CK(cudaMalloc(&m->peer_flags, sizeof(unsigned *) * (size_t)nranks));
    CK(cudaMemcpy(m->peer_land, pl.data(), sizeof(char *) * (size_t)nranks, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m->peer_flags, pf.data(), sizeof(unsigned *) * (size_t)nranks, cudaMemcpyHostToDevice));
    for(int lev = c->cfg.levmin; lev <= c->cfg.levmax; lev++)
    {
        Level &L = c->L[lev];
        CcuHalo &H = m->halo[lev];
        std::vector<int4> seg(H.nb_rank.size());
        for(size_t q = 0; q < H.nb_rank.size(); q++)
        {
            const int r = H.nb_rank[q];
            const int me_r[3] = { (r / m->nproc[2]) % m->nproc[0], r / (m->nproc[2] * m->nproc[0]), r % m->nproc[2] };
            CcuHaloHost hr;
            ccu_build_halo_host(m->nproc, me_r, L.g.nox, L.g.noy, L.g.noz, hr);
            int peer_off = -1;
            for(size_t t = 0; t < hr.nb_rank.size(); t++)
                if(hr.nb_rank[t] == m->rank) { peer_off = hr.nb_off[t]; if(hr.nb_cnt[t] != H.nb_cnt[q]) peer_off = -1; }
            if(peer_off < 0) FAIL("peer-memory exchange: inconsistent duplicated-node tables between neighbours");
            seg[q] = make_int4(r, H.nb_off[q], H.nb_cnt[q], peer_off);
        }
        if(seg.size() > 32) FAIL("peer-memory exchange: more than 32 neighbour segments");
        CK(cudaMalloc(&H.p2p_seg, sizeof(int4) * std::max<size_t>(seg.size(), 1)));
        if(!seg.empty()) CK(cudaMemcpy(H.p2p_seg, seg.data(), sizeof(int4) * seg.size(), cudaMemcpyHostToDevice));
    }
    m->p2p = true;
    // nobody may push into a landing buffer before its owner has zeroed the flags: one barrier
    NK(g_nccl.AllReduce(m->dotstage, m->dotstage, 1, ncclDouble, ncclMax, (ncclComm_t)m->nccl, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ------------------------------------------------------------------ communicator
int ccu_comm_init(ccu_ctx *c, int nprocx, int nprocy, int nprocz, int me_x, int me_y, int me_z, const char *unique_id128)
{
    if(!c) FAIL("null context");
    if(c->comm) FAIL("comm_init: already initialised");
    const int nproc[3] = { nprocx, nprocy, nprocz }, me[3] = { me_x, me_y, me_z };
    const int nranks = nprocx * nprocy * nprocz;
    if(nranks < 1 || me_x < 0 || me_x >= nprocx || me_y < 0 || me_y >= nprocy || me_z < 0 || me_z >= nprocz) FAIL("comm_init: bad processor grid");
    CcuComm *m = new CcuComm();
    m->nranks = nranks; m->rank = ccu_rank_of(nproc, me_x, me_y, me_z);
    for(int a = 0; a < 3; a++) { m->nproc[a] = nproc[a]; m->me[a] = me[a]; }
    if(nranks > 1)
    {
        if(!unique_id128) { delete m; FAIL("comm_init: unique id missing"); }
        if(load_nccl()) { delete m; return 1; }
        ncclUniqueId id;
        memcpy(&id, unique_id128, 128);
        ncclComm_t comm;
        CK(cudaSetDevice(c->cfg.device));
        NK(g_nccl.CommInitRank(&comm, nranks, id, m->rank));
        m->nccl = (void *)comm;
    }
    size_t max_send = 0;
    long long gnno = 0, gnel = 0;
    for(int lev = c->cfg.levmin; lev <= c->cfg.levmax; lev++)
    {
        Level &L = c->L[lev];
        CcuHaloHost h;
        ccu_build_halo_host(nproc, me, L.g.nox, L.g.noy, L.g.noz, h);
        CcuHalo &H = m->halo[lev];
        H.nb_rank = h.nb_rank; H.nb_off = h.nb_off; H.nb_cnt = h.nb_cnt;
        H.n_send = (int)h.send_t.size(); H.n_shared = (int)h.sh_n.size();
        max_send = std::max(max_send, h.send_t.size());
        std::vector<int> sh_s(h.sh_n.size()), send_s(h.send_t.size()), send_n(h.send_t.size());
        for(size_t t = 0; t < h.sh_n.size(); t++)
        {
            const int n = h.sh_n[t];
            sh_s[t] = ccu_sidx(L.g, n / (L.g.noz * L.g.nox), (n / L.g.noz) % L.g.nox, n % L.g.noz);
        }
        for(size_t e = 0; e < h.send_t.size(); e++) { send_s[e] = sh_s[h.send_t[e]]; send_n[e] = h.sh_n[h.send_t[e]]; }
        auto up = [&](int *&dst, const std::vector<int> &v) -> int {
            CK(cudaMalloc(&dst, sizeof(int) * std::max<size_t>(v.size(), 1)));
            if(!v.empty()) CK(cudaMemcpy(dst, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
            return 0;
        };
        if(up(H.sh_s, sh_s) || up(H.sh_n, h.sh_n) || up(H.sh_ptr, h.sh_ptr) || up(H.sh_src, h.sh_src) || up(H.send_s, send_s) ||
           up(H.send_n, send_n) || up(H.send_t, h.send_t)) return 1;
        // ownership / duplicated-node bits into the device flag byte (colour layout)
        std::vector<unsigned char> bits((size_t)L.g.NS, 0);
        for(int n = 0; n < L.g.nno; n++)
        {
            const int s = ccu_sidx(L.g, n / (L.g.noz * L.g.nox), (n / L.g.noz) % L.g.nox, n % L.g.noz);
            bits[s] = h.owned[n] ? 1 : 0;
        }
        for(size_t t = 0; t < sh_s.size(); t++) bits[sh_s[t]] |= 2;
        {   // bits 4-7: distance to the nearest duplicated node, capped at 15 (the split colour passes of the overlapped sweep).  The
            // duplicated nodes of a block decomposition are whole faces of the subdomain: the faces that have a neighbour.
            const int nox = L.g.nox, noy = L.g.noy, noz = L.g.noz;
            const bool lo_i = me[1] > 0, hi_i = me[1] < nproc[1] - 1, lo_j = me[0] > 0, hi_j = me[0] < nproc[0] - 1, lo_k = me[2] > 0, hi_k = me[2] < nproc[2] - 1;
            size_t zero = 0;
            for(int i = 0; i < noy; i++)
                for(int j = 0; j < nox; j++)
                    for(int k = 0; k < noz; k++)
                    {
                        int d = 15;
                        if(lo_i) d = std::min(d, i);
                        if(hi_i) d = std::min(d, noy - 1 - i);
                        if(lo_j) d = std::min(d, j);
                        if(hi_j) d = std::min(d, nox - 1 - j);
                        if(lo_k) d = std::min(d, k);
                        if(hi_k) d = std::min(d, noz - 1 - k);
                        const int s = ccu_sidx(L.g, i, j, k);
                        if((d == 0) != ((bits[s] & 2) != 0)) { delete m; FAIL("comm_init: duplicated-node table and subdomain faces disagree"); }
                        zero += d == 0;
                        bits[s] |= (unsigned char)(d << 4);
                    }
            if(zero != sh_s.size()) { delete m; FAIL("comm_init: duplicated-node count and subdomain faces disagree"); }
        }
        CK(cudaMalloc(&H.bits, bits.size()));
        CK(cudaMemcpy(H.bits, bits.data(), bits.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&H.face, sizeof(double) * 3 * std::max<size_t>(sh_s.size(), 1)));
        if(lev == c->cfg.levmax)
        {
            const long long gx = (long long)L.g.elx * nprocx, gy = (long long)L.g.ely * nprocy, gz = (long long)L.g.elz * nprocz;
            gnno = (gx + 1) * (gy + 1) * (gz + 1); gnel = gx * gy * gz;
        }
    }
    m->gneq = 3 * gnno; m->gnpno = gnel;
    CK(cudaMalloc(&m->sendbuf, sizeof(double) * 3 * std::max<size_t>(max_send, 1)));
    CK(cudaMalloc(&m->recvbuf, sizeof(double) * 3 * std::max<size_t>(max_send, 1)));
    CK(cudaMalloc(&m->dotstage, sizeof(double) * 4));
    c->comm = m;
    if(nranks > 1 && ccu_p2p_setup(c, max_send)) return 1;
    ccu_drop_graphs(c);
    return ccu_col_refresh_all(c);      // the column chunks carry BI = 0 on the duplicated nodes (ccu_col.cuh)
}

void ccu_comm_destroy(ccu_ctx *c)
{
    CcuComm *m = c->comm;
    if(!m) return;
    for(auto &H : m->halo)
    {
        cudaFree(H.sh_s); cudaFree(H.sh_n); cudaFree(H.sh_ptr); cudaFree(H.sh_src); cudaFree(H.send_s); cudaFree(H.send_n); cudaFree(H.send_t);
        cudaFree(H.bits); cudaFree(H.face);
    }
    for(void *p : m->ipc_open) cudaIpcCloseMemHandle(p);
    for(auto &H : m->halo) cudaFree(H.p2p_seg);
    cudaFree(m->land); cudaFree(m->flags); cudaFree(m->seq); cudaFree(m->done); cudaFree(m->p2p_err); cudaFree(m->peer_land); cudaFree(m->peer_flags);
    cudaFree(m->sendbuf); cudaFree(m->recvbuf); cudaFree(m->dotstage); cudaFree(m->mk_counts);
    if(m->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)m->nccl);
    delete m;
    c->comm = nullptr;
}

// one grouped send/recv round of `per_node` values of type T per duplicated node with every neighbour
template <class T>
static int sendrecv(ccu_ctx *c, const CcuHalo &H, int per_node)
{
    CcuComm *m = c->comm;
    if(m->p2p && m->opt_p2p)
    {   // peer-memory exchange: no NCCL call, no host involvement, a few microseconds over NVLink
        const int nseg = (int)H.nb_rank.size(), bpn = (int)sizeof(T) * per_node;
        LAUNCH(c, ccu_k_p2p_bump, 1, 1, m->seq);
        const unsigned slices = (unsigned)std::min<size_t>(32, std::max<size_t>(1, (size_t)H.n_send * bpn / nseg / 16384));
        LAUNCH(c, ccu_k_p2p_push, dim3((unsigned)nseg, slices), 256, (const int4 *)H.p2p_seg, bpn, (const char *)m->sendbuf, (char *const *)m->peer_land,
               (unsigned *const *)m->peer_flags, (const unsigned *)m->seq, m->done, m->rank, m->land_half);
        LAUNCH(c, ccu_k_p2p_wait, std::min<unsigned>(64, std::max<unsigned>(1, (unsigned)((size_t)H.n_send * bpn / 16384))), 256, nseg, (const int4 *)H.p2p_seg, bpn, H.n_send, (const char *)m->land, m->land_half, (const unsigned *)m->flags,
               (const unsigned *)m->seq, (char *)m->recvbuf, m->p2p_err);
        return 0;
    }
    const ncclDataType_t dt = sizeof(T) == 8 ? ncclDouble : ncclFloat;
    NK(g_nccl.GroupStart());
    for(size_t q = 0; q < H.nb_rank.size(); q++)
    {
        const size_t off = (size_t)H.nb_off[q] * per_node, cnt = (size_t)H.nb_cnt[q] * per_node;
        NK(g_nccl.Send((const T *)m->sendbuf + off, cnt, dt, H.nb_rank[q], (ncclComm_t)m->nccl, c->st));
        NK(g_nccl.Recv((T *)m->recvbuf + off, cnt, dt, H.nb_rank[q], (ncclComm_t)m->nccl, c->st));
    }
    NK(g_nccl.GroupEnd());
    return 0;
}

// exchange_id_d20 (Parallel_related.c:1181): colour-layout fp64 vector, 3 dofs per node
int ccu_halo_sum_vec(ccu_ctx *c, int lev, double *vec)
{
    CcuComm *m = c->comm;
    if(!m || m->nranks == 1) return 0;
    const CcuHalo &H = m->halo[lev];
    if(H.n_shared == 0) return 0;
    const size_t NS = (size_t)c->L[lev].g.NS;
    LAUNCH(c, (ccu_k_halo_pack<double, 3>), cdiv(H.n_send, 128), 128, H.n_send, H.send_s, (const double *)vec, NS, (double *)m->sendbuf);
    if(sendrecv<double>(c, H, 3)) return 1;
    LAUNCH(c, (ccu_k_halo_unpack<double, 3>), cdiv(H.n_shared, 128), 128, H.n_shared, H.sh_s, H.sh_ptr, H.sh_src, (const double *)m->recvbuf, vec, NS);
    return 0;
}
// the compact per-level face array H.face[d*n_shared + t] (partial rows of the duplicated nodes)
int ccu_halo_sum_face(ccu_ctx *c, int lev)
{
    CcuComm *m = c->comm;
    if(!m || m->nranks == 1) return 0;
    const CcuHalo &H = m->halo[lev];
    if(H.n_shared == 0) return 0;
    LAUNCH(c, (ccu_k_halo_pack<double, 3>), cdiv(H.n_send, 128), 128, H.n_send, H.send_t, (const double *)H.face, (size_t)H.n_shared, (double *)m->sendbuf);
    if(sendrecv<double>(c, H, 3)) return 1;
    return 0;     // the consumer (ccu_k_face_update) folds own + received in rank order itself
}
// exchange_node_f20 (Parallel_related.c:1270): natural-order fp32 nodal field
int ccu_halo_sum_nodal(ccu_ctx *c, int lev, float *field)
{
    CcuComm *m = c->comm;
    if(!m || m->nranks == 1) return 0;
    const CcuHalo &H = m->halo[lev];
    if(H.n_shared == 0) return 0;
    LAUNCH(c, (ccu_k_halo_pack<float, 1>), cdiv(H.n_send, 128), 128, H.n_send, H.send_n, (const float *)field, (size_t)0, (float *)m->sendbuf);
    if(sendrecv<float>(c, H, 1)) return 1;
    LAUNCH(c, (ccu_k_halo_unpack<float, 1>), cdiv(H.n_shared, 128), 128, H.n_shared, H.sh_n, H.sh_ptr, H.sh_src, (const float *)m->recvbuf, field, (size_t)0);
    return 0;
}
// sum `count` (<= 3) doubles at dotstage over all ranks, then scatter to the scalar slots
int ccu_allreduce_dots(ccu_ctx *c, int count, double *o0, double *o1, double *o2)
{
    CcuComm *m = c->comm;
    if(m && m->nranks > 1) NK(g_nccl.AllReduce(m->dotstage, m->dotstage, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)m->nccl, c->st));
    LAUNCH(c, ccu_k_scatter_scal, 1, 1, (const double *)m->dotstage, o0, o1, o2);
    return 0;
}
int ccu_allreduce_buffer(ccu_ctx *c, double *buf, int count, int op_max)
{
    CcuComm *m = c->comm;
    if(m && m->nranks > 1) NK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, op_max ? ncclMax : ncclSum, (ncclComm_t)m->nccl, c->st));
    return 0;
}

// every rank's `bytes_per_rank` block, concatenated in rank order, on every rank
int ccu_allgather(ccu_ctx *c, const void *send, void *recv, size_t bytes_per_rank)
{
    CcuComm *m = c->comm;
    if(!m || m->nranks == 1) { CK(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->st)); return 0; }
    NK(g_nccl.AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)m->nccl, c->st));
    return 0;
}

// host-only routing table of the marker exchange (no GPU needed; tests/test_decomp.py): for the subdomain `me` of an
// nproc[0] x nproc[1] x nproc[2] grid and the gathered per-direction send counts all_counts[rank][27] of every rank,
// nb_rank[code] = rank of the neighbour at offset code (-1: none) and recvcnt[code] = records it sends here
extern "C" int ccu_marker_routes(const int nproc[3], const int me[3], const int *all_counts, int nb_rank[27], int recvcnt[27])
{
    for(int code = 0; code < 27; code++)
    {
        const int ox = code % 3 - 1, oy = (code / 3) % 3 - 1, oz = code / 9 - 1;
        const int x = me[0] + ox, y = me[1] + oy, z = me[2] + oz;
        nb_rank[code] = -1; recvcnt[code] = 0;
        if(code == 13 || x < 0 || y < 0 || z < 0 || x >= nproc[0] || y >= nproc[1] || z >= nproc[2]) continue;
        nb_rank[code] = ccu_rank_of(nproc, x, y, z);
        recvcnt[code] = all_counts[(size_t)nb_rank[code] * 27 + (26 - code)];      // its code for the opposite offset
    }
    return 0;
}

// ------------------------------------------------------------------ markers changing subdomain
// exchange_number_rec_markers + exchange_markers (Composition_adv.c:421-560): every rank learns how many records each
// neighbour sends it (one all-gather of the 27 per-direction counts), then one grouped send/recv round moves the records.
// Neighbour code = (ox+1) + 3 (oy+1) + 9 (oz+1) of the offset (ox, oy, oz); what rank R at offset o sends to me sits
// under R's code for the opposite offset, 26 - code.  Received records are stored in ascending code order.
// Capacity is decided COLLECTIVELY (n_resident = markers that stay on this rank): every rank learns through the same
// all-gather whether any rank would overflow, and all of them fail together before the send/recv group instead of leaving
// the others blocked in it (the reference terminates all ranks here, parallel_process_termination).
int ccu_marker_exchange(ccu_ctx *c, const int sendcnt[27], const double *sendbuf, int rec, int recvcnt[27], double *recvbuf, size_t cap_records, int *nrecv, int n_resident)
{
    CcuComm *m = c->comm;
    for(int q = 0; q < 27; q++) recvcnt[q] = 0;
    *nrecv = 0;
    if(!m || m->nranks == 1) return 0;
    const int W = 29;                      // 27 per-direction counts + markers staying + capacity, per rank
    if(!m->mk_counts) CK(cudaMalloc(&m->mk_counts, sizeof(int) * W * (size_t)(m->nranks + 1)));
    int *mine = m->mk_counts, *all = m->mk_counts + W;
    int mine_h[W];
    for(int q = 0; q < 27; q++) mine_h[q] = sendcnt[q];
    mine_h[27] = n_resident; mine_h[28] = (int)cap_records;
    CK(cudaMemcpyAsync(mine, mine_h, sizeof(int) * W, cudaMemcpyHostToDevice, c->st));
    NK(g_nccl.AllGather(mine, all, W * sizeof(int), ncclChar, (ncclComm_t)m->nccl, c->st));
    std::vector<int> hw(W * (size_t)m->nranks), h(27 * (size_t)m->nranks);
    CK(cudaMemcpyAsync(hw.data(), all, sizeof(int) * hw.size(), cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    for(int r = 0; r < m->nranks; r++)
        for(int q = 0; q < 27; q++) h[27 * (size_t)r + q] = hw[W * (size_t)r + q];
    int nb[27];
    for(int r = 0; r < m->nranks; r++)
    {   // what arrives at every rank: the same verdict everywhere
        const int z = r % m->nproc[2], x = (r / m->nproc[2]) % m->nproc[0], y = r / (m->nproc[2] * m->nproc[0]);
        const int me_r[3] = { x, y, z };
        int nb_r[27], rc_r[27], in = 0;
        ccu_marker_routes(m->nproc, me_r, h.data(), nb_r, rc_r);
        for(int code = 0; code < 27; code++) in += rc_r[code];
        if(in > hw[W * (size_t)r + 28] || hw[W * (size_t)r + 27] + in > hw[W * (size_t)r + 28])
            FAIL("markers: rank " + std::to_string(r) + " would hold more markers than its capacity (markers_uplimit); all ranks stop, as the reference's Composition_adv.c:207");
    }
    ccu_marker_routes(m->nproc, m->me, h.data(), nb, recvcnt);
    for(int code = 0; code < 27; code++) *nrecv += recvcnt[code];
    NK(g_nccl.GroupStart());
    size_t soff = 0, roff = 0;
    for(int code = 0; code < 27; code++)
    {
        if(nb[code] >= 0 && sendcnt[code]) NK(g_nccl.Send(sendbuf + soff * rec, (size_t)sendcnt[code] * rec, ncclDouble, nb[code], (ncclComm_t)m->nccl, c->st));
        if(nb[code] >= 0 && recvcnt[code]) NK(g_nccl.Recv(recvbuf + roff * rec, (size_t)recvcnt[code] * rec, ncclDouble, nb[code], (ncclComm_t)m->nccl, c->st));
        soff += sendcnt[code]; roff += recvcnt[code];
    }
    NK(g_nccl.GroupEnd());
    return 0;
}
