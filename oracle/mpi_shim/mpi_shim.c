/* mpi_shim.c -- process-based MPI subset for the CPU oracle (TEST INFRASTRUCTURE).
 *
 * See mpi.h.  N ranks = N processes forked inside MPI_Init (env CCU_MPI_NP),
 * sharing one anonymous MAP_SHARED arena that holds, for every ordered pair
 * (src,dst), a single-producer/single-consumer channel: a ring of message
 * descriptors plus a byte ring.  Sends are buffered (copied at Isend time);
 * receives match the oldest unconsumed message of the channel with the same
 * tag (MPI's non-overtaking rule).  Collectives are built from point-to-point
 * messages on reserved tags, reduced in rank order so results are
 * deterministic.  Only what the reference calls is implemented.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <sys/wait.h>

#define SHIM_MAX_NP 64
#define SHIM_NDESC 256
#define TAG_COLL 0x7f000001

typedef struct {
    volatile int state;          /* 0 free, 1 full, 2 consumed (awaiting head advance) */
    int tag;
    size_t nbytes;
    size_t off;                  /* offset into channel byte ring */
    uint64_t endabs;             /* producer's byte counter after this payload */
} shim_desc;

typedef struct {
    volatile uint64_t dhead, dtail;   /* descriptor ring: consumer head, producer tail */
    volatile uint64_t bhead, btail;   /* byte ring (monotonic counters) */
    shim_desc d[SHIM_NDESC];
} shim_chan;

static int g_np = 1, g_me = 0;
static size_t g_chbytes = 0;
static shim_chan *g_ch = NULL;    /* [src*np+dst] */
static char *g_bytes = NULL;      /* [src*np+dst][g_chbytes] */
static pid_t g_kids[SHIM_MAX_NP];

typedef struct { int active; void *buf; size_t nbytes; int src, tag; } shim_req;
#define SHIM_MAXREQ 4096
static shim_req g_req[SHIM_MAXREQ];

typedef struct { int n; int ranks[SHIM_MAX_NP]; int used; } shim_group;
#define SHIM_MAXCOMM 64
static shim_group g_comm[SHIM_MAXCOMM];   /* index 0 = world */
static shim_group g_group[SHIM_MAXCOMM];

static size_t dt_size(MPI_Datatype dt)
{
    switch (dt) { case MPI_INT: return sizeof(int); case MPI_FLOAT: return sizeof(float);
                  case MPI_DOUBLE: return sizeof(double); }
    fprintf(stderr, "mpi_shim: unknown datatype %d\n", dt); abort();
}

static inline shim_chan *chan(int src, int dst) { return &g_ch[src * g_np + dst]; }
static inline char *chbytes(int src, int dst) { return g_bytes + (size_t)(src * g_np + dst) * g_chbytes; }

static void chan_push(int dst, int tag, const void *buf, size_t nbytes)
{
    shim_chan *c = chan(g_me, dst);
    char *base = chbytes(g_me, dst);
    if (nbytes > g_chbytes / 2) { fprintf(stderr, "mpi_shim: message of %zu bytes exceeds channel (set CCU_MPI_CHBYTES)\n", nbytes); abort(); }
    for (;;) {
        uint64_t dh = __atomic_load_n(&c->dhead, __ATOMIC_ACQUIRE);
        uint64_t bh = __atomic_load_n(&c->bhead, __ATOMIC_ACQUIRE);
        size_t pos = (size_t)(c->btail % g_chbytes);
        size_t pad = (pos + nbytes > g_chbytes) ? (g_chbytes - pos) : 0;   /* keep payload contiguous */
        if (c->dtail - dh < SHIM_NDESC && (c->btail + pad + nbytes) - bh <= g_chbytes) {
            shim_desc *d = &c->d[c->dtail % SHIM_NDESC];
            if (__atomic_load_n(&d->state, __ATOMIC_ACQUIRE) == 0) {
                c->btail += pad;
                d->off = (size_t)(c->btail % g_chbytes);
                d->tag = tag; d->nbytes = nbytes;
                memcpy(base + d->off, buf, nbytes);
                c->btail += nbytes;
                d->endabs = c->btail;
                __atomic_store_n(&d->state, 1, __ATOMIC_RELEASE);
                __atomic_store_n(&c->dtail, c->dtail + 1, __ATOMIC_RELEASE);
                return;
            }
        }
        sched_yield();
    }
}

static void chan_pop(int src, int tag, void *buf, size_t nbytes)
{
    shim_chan *c = chan(src, g_me);
    char *base = chbytes(src, g_me);
    for (;;) {
        uint64_t dt = __atomic_load_n(&c->dtail, __ATOMIC_ACQUIRE);
        for (uint64_t i = c->dhead; i < dt; i++) {
            shim_desc *d = &c->d[i % SHIM_NDESC];
            if (__atomic_load_n(&d->state, __ATOMIC_ACQUIRE) == 1 && d->tag == tag) {
                if (d->nbytes > nbytes) { fprintf(stderr, "mpi_shim: rank %d truncation from %d tag %d (%zu > %zu)\n", g_me, src, tag, d->nbytes, nbytes); abort(); }
                memcpy(buf, base + d->off, d->nbytes);
                d->state = 2;
                /* advance head over consumed descriptors, releasing their bytes */
                while (c->dhead < dt && c->d[c->dhead % SHIM_NDESC].state == 2) {
                    shim_desc *h = &c->d[c->dhead % SHIM_NDESC];
                    uint64_t endabs = h->endabs;
                    __atomic_store_n(&c->bhead, endabs, __ATOMIC_RELEASE);
                    __atomic_store_n(&h->state, 0, __ATOMIC_RELEASE);
                    __atomic_store_n(&c->dhead, c->dhead + 1, __ATOMIC_RELEASE);
                }
                return;
            }
        }
        sched_yield();
    }
}

int MPI_Init(int *argc, char ***argv)
{
    (void)argc; (void)argv;
    const char *s = getenv("CCU_MPI_NP");
    g_np = s ? atoi(s) : 1;
    if (g_np < 1 || g_np > SHIM_MAX_NP) { fprintf(stderr, "mpi_shim: bad CCU_MPI_NP\n"); exit(2); }
    s = getenv("CCU_MPI_CHBYTES");
    g_chbytes = s ? (size_t)atoll(s) : ((size_t)32 << 20);
    g_comm[0].n = g_np; g_comm[0].used = 1;
    for (int i = 0; i < g_np; i++) g_comm[0].ranks[i] = i;
    g_me = 0;
    if (g_np == 1) return MPI_SUCCESS;
    size_t nch = (size_t)g_np * g_np;
    g_ch = mmap(NULL, nch * sizeof(shim_chan), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    g_bytes = mmap(NULL, nch * g_chbytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_ch == MAP_FAILED || g_bytes == MAP_FAILED) { perror("mpi_shim mmap"); exit(2); }
    memset(g_ch, 0, nch * sizeof(shim_chan));
    fflush(NULL);
    for (int r = 1; r < g_np; r++) {
        pid_t p = fork();
        if (p < 0) { perror("mpi_shim fork"); exit(2); }
        if (p == 0) { g_me = r; break; }
        g_kids[r] = p;
    }
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    fflush(NULL);
    if (g_np > 1) {
        MPI_Barrier(MPI_COMM_WORLD);
        if (g_me == 0)
            for (int r = 1; r < g_np; r++) { int st; waitpid(g_kids[r], &st, 0); }
    }
    return MPI_SUCCESS;
}

static shim_group *comm_of(MPI_Comm c)
{
    if (c < 0 || c >= SHIM_MAXCOMM || !g_comm[c].used) { fprintf(stderr, "mpi_shim: bad comm %d\n", c); abort(); }
    return &g_comm[c];
}

int MPI_Comm_rank(MPI_Comm comm, int *rank)
{
    shim_group *g = comm_of(comm);
    for (int i = 0; i < g->n; i++) if (g->ranks[i] == g_me) { *rank = i; return MPI_SUCCESS; }
    *rank = -1; return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm_of(comm)->n; return MPI_SUCCESS; }

static void reduce_into(void *acc, const void *in, int count, MPI_Datatype dt, MPI_Op op)
{
#define RED(T) { T *a = (T *)acc; const T *b = (const T *)in; for (int i = 0; i < count; i++) { \
        if (op == MPI_SUM) a[i] += b[i]; else if (op == MPI_MAX) { if (b[i] > a[i]) a[i] = b[i]; } \
        else if (op == MPI_MIN) { if (b[i] < a[i]) a[i] = b[i]; } } }
    if (dt == MPI_INT) RED(int) else if (dt == MPI_FLOAT) RED(float) else RED(double)
#undef RED
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
    shim_group *g = comm_of(comm);
    size_t nb = dt_size(dt) * (size_t)count;
    if (g->n == 1) { if (recvbuf != sendbuf) memmove(recvbuf, sendbuf, nb); return MPI_SUCCESS; }
    int root = g->ranks[0];
    if (g_me == root) {
        void *tmp = malloc(nb);
        if (recvbuf != sendbuf) memmove(recvbuf, sendbuf, nb);
        for (int i = 1; i < g->n; i++) { chan_pop(g->ranks[i], TAG_COLL, tmp, nb); reduce_into(recvbuf, tmp, count, dt, op); }
        for (int i = 1; i < g->n; i++) chan_push(g->ranks[i], TAG_COLL, recvbuf, nb);
        free(tmp);
    } else {
        chan_push(root, TAG_COLL, sendbuf, nb);
        chan_pop(root, TAG_COLL, recvbuf, nb);
    }
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{
    shim_group *g = comm_of(comm);
    size_t nb = dt_size(dt) * (size_t)count;
    int rr = g->ranks[root];
    if (g->n == 1) return MPI_SUCCESS;
    if (g_me == rr) { for (int i = 0; i < g->n; i++) if (g->ranks[i] != rr) chan_push(g->ranks[i], TAG_COLL, buf, nb); }
    else chan_pop(rr, TAG_COLL, buf, nb);
    return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) { int a = 1, b; return MPI_Allreduce(&a, &b, 1, MPI_INT, MPI_SUM, comm); }

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm)
{
    shim_group *g = comm_of(comm);
    chan_push(g->ranks[dest], tag, buf, dt_size(dt) * (size_t)count);
    return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st)
{
    shim_group *g = comm_of(comm);
    chan_pop(g->ranks[src], tag, buf, dt_size(dt) * (size_t)count);
    if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
    return MPI_SUCCESS;
}
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
    MPI_Send(buf, count, dt, dest, tag, comm);
    *req = -1;                       /* already complete (buffered) */
    return MPI_SUCCESS;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req)
{
    shim_group *g = comm_of(comm);
    for (int i = 0; i < SHIM_MAXREQ; i++)
        if (!g_req[i].active) {
            g_req[i].active = 1; g_req[i].buf = buf; g_req[i].nbytes = dt_size(dt) * (size_t)count;
            g_req[i].src = g->ranks[src]; g_req[i].tag = tag; *req = i; return MPI_SUCCESS;
        }
    fprintf(stderr, "mpi_shim: out of requests\n"); abort();
}
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts)
{
    (void)sts;
    for (int i = 0; i < n; i++) {
        int r = reqs[i];
        if (r < 0) continue;
        chan_pop(g_req[r].src, g_req[r].tag, g_req[r].buf, g_req[r].nbytes);
        g_req[r].active = 0;
    }
    return MPI_SUCCESS;
}

int MPI_Comm_group(MPI_Comm comm, MPI_Group *gout)
{
    shim_group *g = comm_of(comm);
    for (int i = 0; i < SHIM_MAXCOMM; i++) if (!g_group[i].used) { g_group[i] = *g; g_group[i].used = 1; *gout = i; return MPI_SUCCESS; }
    fprintf(stderr, "mpi_shim: out of groups\n"); abort();
}
int MPI_Group_incl(MPI_Group gin, int n, const int *ranks, MPI_Group *gout)
{
    for (int i = 0; i < SHIM_MAXCOMM; i++) if (!g_group[i].used) {
        g_group[i].used = 1; g_group[i].n = n;
        for (int k = 0; k < n; k++) g_group[i].ranks[k] = g_group[gin].ranks[ranks[k]];
        *gout = i; return MPI_SUCCESS;
    }
    fprintf(stderr, "mpi_shim: out of groups\n"); abort();
}
int MPI_Comm_create(MPI_Comm comm, MPI_Group gin, MPI_Comm *newcomm)
{
    (void)comm;
    int member = 0;
    for (int k = 0; k < g_group[gin].n; k++) if (g_group[gin].ranks[k] == g_me) member = 1;
    if (!member) { *newcomm = MPI_COMM_NULL; return MPI_SUCCESS; }
    for (int i = 1; i < SHIM_MAXCOMM; i++) if (!g_comm[i].used) { g_comm[i] = g_group[gin]; g_comm[i].used = 1; *newcomm = i; return MPI_SUCCESS; }
    fprintf(stderr, "mpi_shim: out of comms\n"); abort();
}
int MPI_Comm_free(MPI_Comm *c) { if (*c > 0 && *c < SHIM_MAXCOMM) g_comm[*c].used = 0; *c = MPI_COMM_NULL; return MPI_SUCCESS; }
int MPI_Group_free(MPI_Group *g) { if (*g >= 0 && *g < SHIM_MAXCOMM) g_group[*g].used = 0; *g = -1; return MPI_SUCCESS; }

double MPI_Wtime(void)
{
    struct timeval tv; gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}
