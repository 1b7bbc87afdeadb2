/*
 * mpi_shim.c -- fork + shared-memory implementation of the MPI subset declared
 * in mpi.h.  TEST / BASELINE INFRASTRUCTURE; see mpi.h for the rationale.
 *
 * Layout of the shared mapping:
 *   [ control block | result area (CHUNK bytes) | P slots of CHUNK bytes ]
 * Collectives move data in CHUNK-sized pieces:
 *   bcast     : root -> result area, barrier, everyone copies out, barrier
 *   scatterv  : per destination rank, chunked through the result area
 *   reduce    : every rank -> its slot, barrier, rank r reduces strip r of the
 *               chunk over all slots (rank order) into the result area,
 *               barrier, consumers copy out, barrier
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <sched.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define SHIM_CHUNK ((size_t)8 << 20) /* bytes staged per step */
#define SHIM_MAX_RANKS 256

typedef struct {
    atomic_int arrived;
    atomic_int generation;
    atomic_int failed;
    char pad[64 - 3 * sizeof(atomic_int)];
} shim_ctrl;

static int g_rank = 0;
static int g_size = 1;
static shim_ctrl* g_ctrl = NULL;
static unsigned char* g_result = NULL;
static unsigned char* g_slots = NULL;
static pid_t g_children[SHIM_MAX_RANKS];

static size_t dt_size(MPI_Datatype dt)
{
    switch (dt) {
    case MPI_INT: return sizeof(int);
    case MPI_FLOAT: return sizeof(float);
    case MPI_DOUBLE: return sizeof(double);
    default:
        fprintf(stderr, "mpi_shim: unsupported datatype %d\n", dt);
        exit(1);
    }
}

static void shim_barrier(void)
{
    if (g_size == 1) return;
    const int gen = atomic_load_explicit(&g_ctrl->generation, memory_order_acquire);
    if (atomic_fetch_add_explicit(&g_ctrl->arrived, 1, memory_order_acq_rel) == g_size - 1) {
        atomic_store_explicit(&g_ctrl->arrived, 0, memory_order_relaxed);
        atomic_store_explicit(&g_ctrl->generation, gen + 1, memory_order_release);
        return;
    }
    unsigned spins = 0;
    while (atomic_load_explicit(&g_ctrl->generation, memory_order_acquire) == gen) {
        if (++spins > 2000) {
            sched_yield();
            if (atomic_load_explicit(&g_ctrl->failed, memory_order_relaxed)) _exit(3);
            if ((spins & 0xfffff) == 0 && g_rank != 0 && getppid() == 1) _exit(3);
        } else {
            __builtin_ia32_pause();
        }
    }
}

/* MPI_SHIM_CPUS="c0,c1,...": rank r is pinned to CPU c[r % count] (one rank per physical core, as mpirun --bind-to core
 * would do); without it the ranks float and the kernel may stack two of them on one core's hyperthreads. */
static void shim_pin(int rank)
{
    const char* list = getenv("MPI_SHIM_CPUS");
    if (!list || !*list) return;
    int cpus[1024], n = 0;
    const char* p = list;
    while (*p && n < 1024) {
        char* end = NULL;
        long v = strtol(p, &end, 10);
        if (end == p) break;
        cpus[n++] = (int)v;
        p = (*end == ',') ? end + 1 : end;
        if (*end != ',') break;
    }
    if (n == 0) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(cpus[rank % n], &set);
    sched_setaffinity(0, sizeof(set), &set); /* best effort */
}

int MPI_Init(int* argc, char*** argv)
{
    (void)argc; (void)argv;
    const char* np = getenv("MPI_SHIM_NP");
    g_size = np ? atoi(np) : 1;
    if (g_size < 1) g_size = 1;
    if (g_size > SHIM_MAX_RANKS) g_size = SHIM_MAX_RANKS;
    g_rank = 0;
    if (g_size == 1) {
        shim_pin(0);
        return MPI_SUCCESS;
    }

    const size_t bytes = 4096 + SHIM_CHUNK * (size_t)(g_size + 1);
    void* base = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) {
        perror("mpi_shim: mmap");
        exit(1);
    }
    g_ctrl = (shim_ctrl*)base;
    atomic_init(&g_ctrl->arrived, 0);
    atomic_init(&g_ctrl->generation, 0);
    atomic_init(&g_ctrl->failed, 0);
    g_result = (unsigned char*)base + 4096;
    g_slots = g_result + SHIM_CHUNK;

    fflush(stdout);
    fflush(stderr);
    for (int r = 1; r < g_size; ++r) {
        pid_t pid = fork();
        if (pid < 0) {
            perror("mpi_shim: fork");
            atomic_store(&g_ctrl->failed, 1);
            exit(1);
        }
        if (pid == 0) {
            g_rank = r;
            shim_pin(r);
            return MPI_SUCCESS;
        }
        g_children[r] = pid;
    }
    shim_pin(0);
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    if (g_size == 1) return MPI_SUCCESS;
    shim_barrier();
    if (g_rank != 0) {
        fflush(stdout);
        _exit(0);
    }
    for (int r = 1; r < g_size; ++r) {
        int st = 0;
        waitpid(g_children[r], &st, 0);
    }
    return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int* rank) { (void)comm; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int* size) { (void)comm; *size = g_size; return MPI_SUCCESS; }

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{
    (void)comm;
    if (g_size == 1) return MPI_SUCCESS;
    const size_t total = (size_t)count * dt_size(dt);
    unsigned char* p = (unsigned char*)buf;
    for (size_t off = 0; off < total || off == 0; off += SHIM_CHUNK) {
        const size_t len = total - off < SHIM_CHUNK ? total - off : SHIM_CHUNK;
        if (g_rank == root) memcpy(g_result, p + off, len);
        shim_barrier();
        if (g_rank != root) memcpy(p + off, g_result, len);
        shim_barrier();
        if (total == 0) break;
    }
    return MPI_SUCCESS;
}

int MPI_Ibcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm, MPI_Request* req)
{
    *req = 1;
    return MPI_Bcast(buf, count, dt, root, comm);
}

int MPI_Wait(MPI_Request* req, MPI_Status* status)
{
    (void)status;
    if (req) *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs,
                 MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                 int root, MPI_Comm comm)
{
    (void)comm;
    const size_t rsz = dt_size(recvtype);
    if (g_size == 1) {
        const size_t ssz = dt_size(sendtype);
        memcpy(recvbuf, (const unsigned char*)sendbuf + (size_t)displs[0] * ssz,
               (size_t)sendcounts[0] * ssz);
        return MPI_SUCCESS;
    }
    /* Every rank must know each destination's byte count to walk the same
     * chunk schedule: the root publishes the counts first. */
    int* counts_shared = (int*)g_result;
    if (g_rank == root) memcpy(counts_shared, sendcounts, sizeof(int) * (size_t)g_size);
    shim_barrier();
    int counts[SHIM_MAX_RANKS];
    memcpy(counts, counts_shared, sizeof(int) * (size_t)g_size);
    shim_barrier();
    const size_t ssz = (g_rank == root) ? dt_size(sendtype) : rsz;
    for (int r = 0; r < g_size; ++r) {
        const size_t total = (size_t)counts[r] * ssz;
        if (r == root) {
            if (g_rank == root)
                memcpy(recvbuf, (const unsigned char*)sendbuf + (size_t)displs[r] * ssz, total);
            continue;
        }
        for (size_t off = 0; off < total; off += SHIM_CHUNK) {
            const size_t len = total - off < SHIM_CHUNK ? total - off : SHIM_CHUNK;
            if (g_rank == root)
                memcpy(g_result, (const unsigned char*)sendbuf + (size_t)displs[r] * ssz + off, len);
            shim_barrier();
            if (g_rank == r) memcpy((unsigned char*)recvbuf + off, g_result, len);
            shim_barrier();
        }
    }
    (void)recvcount;
    return MPI_SUCCESS;
}

#define REDUCE_STRIP(T)                                                            \
    do {                                                                           \
        T* out = (T*)g_result;                                                     \
        for (size_t e = lo; e < hi; ++e) {                                         \
            T acc = ((const T*)g_slots)[e];                                        \
            for (int r = 1; r < g_size; ++r) {                                     \
                const T v = ((const T*)(g_slots + (size_t)r * SHIM_CHUNK))[e];     \
                if (op == MPI_SUM) acc += v;                                       \
                else if (v > acc) acc = v;                                         \
            }                                                                      \
            out[e] = acc;                                                          \
        }                                                                          \
    } while (0)

static int shim_reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt,
                       MPI_Op op, int root, int all)
{
    const size_t esz = dt_size(dt);
    if (op != MPI_SUM && op != MPI_MAX) {
        fprintf(stderr, "mpi_shim: unsupported op %d\n", op);
        exit(1);
    }
    if (g_size == 1) {
        if (recvbuf != sendbuf) memcpy(recvbuf, sendbuf, (size_t)count * esz);
        return MPI_SUCCESS;
    }
    const size_t per_chunk = SHIM_CHUNK / esz;
    const size_t n = (size_t)count;
    for (size_t base = 0; base < n; base += per_chunk) {
        const size_t cnt = n - base < per_chunk ? n - base : per_chunk;
        memcpy(g_slots + (size_t)g_rank * SHIM_CHUNK, (const unsigned char*)sendbuf + base * esz,
               cnt * esz);
        shim_barrier();
        const size_t strip = (cnt + (size_t)g_size - 1) / (size_t)g_size;
        const size_t lo = strip * (size_t)g_rank < cnt ? strip * (size_t)g_rank : cnt;
        const size_t hi = lo + strip < cnt ? lo + strip : cnt;
        if (dt == MPI_FLOAT) REDUCE_STRIP(float);
        else if (dt == MPI_DOUBLE) REDUCE_STRIP(double);
        else REDUCE_STRIP(int);
        shim_barrier();
        if (all || g_rank == root)
            memcpy((unsigned char*)recvbuf + base * esz, g_result, cnt * esz);
        shim_barrier();
    }
    return MPI_SUCCESS;
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op,
               int root, MPI_Comm comm)
{
    (void)comm;
    return shim_reduce(sendbuf, recvbuf, count, dt, op, root, 0);
}

int MPI_Ireduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op,
                int root, MPI_Comm comm, MPI_Request* req)
{
    (void)comm;
    *req = 1;
    return shim_reduce(sendbuf, recvbuf, count, dt, op, root, 0);
}

int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op,
                   MPI_Comm comm, MPI_Request* req)
{
    (void)comm;
    *req = 1;
    return shim_reduce(sendbuf, recvbuf, count, dt, op, 0, 1);
}
