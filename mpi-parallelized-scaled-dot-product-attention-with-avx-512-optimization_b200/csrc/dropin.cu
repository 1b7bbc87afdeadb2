// dropin.cu -- the reference entry point `attention(...)` (attention-mpi.c:191-192) on top
// of the context API.  Same contract as the reference: blocking, fp64 host arrays valid on
// rank 0, result complete on rank 0 at return, fatal errors -> stderr + exit(1)
// (the harness convention of mpi.c:419-422,436-448 -- the entry point has no error channel).
#include "common.cuh"

#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/stat.h>
#include <time.h>
#include <string>

#include <vector>

using namespace sdpa;

namespace {

struct Cached {
    sdpa_ctx* ctx = nullptr;
    sdpa_config cfg;
    int world = 0, rank = -1;
};
Cached g_cached;
unsigned char g_boot_id[128];
bool g_boot_id_set = false;

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "sdpa_b200: %s: %s\n", what, sdpa_last_error());
    exit(1);
}

int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int env_precision()
{
    const char* v = getenv("SDPA_PRECISION");
    if (!v || !*v || !strcmp(v, "auto")) return SDPA_PREC_AUTO;
    if (!strcmp(v, "f32") || !strcmp(v, "fp32")) return SDPA_PREC_F32;
    if (!strcmp(v, "bf16")) return SDPA_PREC_BF16;
    if (!strcmp(v, "bf16x3") || !strcmp(v, "f32x3")) return SDPA_PREC_BF16X3;
    fprintf(stderr, "sdpa_b200: SDPA_PRECISION must be auto|f32|bf16|bf16x3 (got %s)\n", v);
    exit(1);
}

// File rendezvous for the ncclUniqueId when the launcher did not call sdpa_set_bootstrap_id().  Record = magic, a 64-bit
// launch nonce, the 128-byte id.  Rank 0 removes whatever a previous run left at the path, writes <file>.tmp and renames it
// into place; the others poll and accept a record only if (a) its nonce equals theirs -- the nonce is a hash of the first of
// SDPA_NCCL_ID_NONCE / the launcher's job identifiers (PMIX_NAMESPACE, OMPI_MCA_ess_base_jobid, SLURM_JOB_ID + SLURM_STEP_ID,
// PMI_JOBID) found in the environment, identical on every rank of one launch -- and (b) the file is not older than this
// process (minus a slack for launch skew), which also covers launchers that export none of those variables.  Rank 0 deletes
// the file again once its communicator exists (every rank has read the id by then).
const time_t g_load_time = time(nullptr);
const char kIdMagic[8] = {'S', 'D', 'P', 'A', 'I', 'D', '0', '2'};

uint64_t launch_nonce()
{
    const char* names[] = {"SDPA_NCCL_ID_NONCE", "PMIX_NAMESPACE", "OMPI_MCA_ess_base_jobid", "SLURM_JOB_ID", "SLURM_STEP_ID", "PMI_JOBID"};
    uint64_t h = 1469598103934665603ull;   // FNV-1a over name=value of every identifier present
    for (const char* nm : names) {
        const char* v = getenv(nm);
        if (!v || !*v) continue;
        for (const char* p = nm; *p; ++p) h = (h ^ (unsigned char)*p) * 1099511628211ull;
        for (const char* p = v; *p; ++p) h = (h ^ (unsigned char)*p) * 1099511628211ull;
        if (nm == names[0]) break;   // an explicit nonce overrides the launcher's variables
    }
    return h;
}

void bootstrap_id_via_file(int rank, unsigned char* id)
{
    const char* path = getenv("SDPA_NCCL_ID_FILE");
    if (!path || !*path) {
        fprintf(stderr, "sdpa_b200: mpi_size > 1 needs sdpa_set_bootstrap_id() or env SDPA_NCCL_ID_FILE\n");
        exit(1);
    }
    const uint64_t nonce = launch_nonce();
    unsigned char rec[8 + 8 + 128];
    if (rank == 0) {
        if (sdpa_get_unique_id(id) != SDPA_OK) die("ncclGetUniqueId");
        unlink(path);   // a record left by an earlier run must never be readable while this run's is being written
        memcpy(rec, kIdMagic, 8);
        memcpy(rec + 8, &nonce, 8);
        memcpy(rec + 16, id, 128);
        std::string tmp = std::string(path) + ".tmp";
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(rec, 1, sizeof(rec), f) != sizeof(rec)) {
            fprintf(stderr, "sdpa_b200: cannot write %s\n", tmp.c_str());
            exit(1);
        }
        fclose(f);
        rename(tmp.c_str(), path);
    } else {
        const int slack_s = env_int("SDPA_NCCL_ID_MAX_SKEW_S", 30);
        for (int tries = 0; tries < 120000; ++tries) {
            struct stat sb;
            FILE* f = fopen(path, "rb");
            if (f) {
                const size_t got = fread(rec, 1, sizeof(rec), f);
                const bool fresh = fstat(fileno(f), &sb) == 0 && sb.st_mtime + slack_s >= g_load_time;
                fclose(f);
                uint64_t theirs = 0;
                memcpy(&theirs, rec + 8, 8);
                if (got == sizeof(rec) && !memcmp(rec, kIdMagic, 8) && theirs == nonce && fresh) {
                    memcpy(id, rec + 16, 128);
                    return;
                }
            }
            usleep(1000);
        }
        fprintf(stderr, "sdpa_b200: timed out waiting for a fresh id record in %s (stale file of another launch?)\n", path);
        exit(1);
    }
}

sdpa_ctx* get_ctx(int mpi_rank, int mpi_size)
{
    sdpa_config cfg;
    sdpa_config_init(&cfg);
    cfg.precision = env_precision();
    cfg.q_batch = env_int("SDPA_Q_BATCH", 0);
    cfg.kv_splits = env_int("SDPA_KV_SPLITS", 0);
    const char* mg = getenv("SDPA_MERGE");
    cfg.merge = (mg && !strcmp(mg, "peer")) ? SDPA_MERGE_PEER : (mg && !strcmp(mg, "nccl3")) ? SDPA_MERGE_NCCL : SDPA_MERGE_NCCL2;
    if (mpi_size <= 1) {
        const char* dist = getenv("SDPA_DISTRIBUTION");   // kv (default, the reference's sharding) | q | auto
        cfg.distribution = (dist && !strcmp(dist, "q")) ? SDPA_DIST_Q : (dist && !strcmp(dist, "auto")) ? SDPA_DIST_AUTO : SDPA_DIST_KV;
        cfg.num_local = env_int("SDPA_NGPUS", 1);
        cfg.first_device = env_int("SDPA_FIRST_DEVICE", 0);
        cfg.world_size = cfg.num_local;
        cfg.rank_base = 0;
    } else {
        const int ndev = sdpa_device_count();
        cfg.num_local = 1;
        cfg.first_device = env_int("SDPA_FIRST_DEVICE", -1);
        if (cfg.first_device < 0) cfg.first_device = ndev > 0 ? env_int("LOCAL_RANK", mpi_rank) % ndev : 0;
        cfg.world_size = mpi_size;
        cfg.rank_base = mpi_rank;
    }
    if (g_cached.ctx && g_cached.world == cfg.world_size && g_cached.rank == cfg.rank_base &&
        !memcmp(&g_cached.cfg, &cfg, sizeof(cfg)))
        return g_cached.ctx;
    if (g_cached.ctx) {
        sdpa_ctx_destroy(g_cached.ctx);
        g_cached.ctx = nullptr;
    }
    const void* id = nullptr;
    unsigned char idbuf[128];
    if (mpi_size > 1) {
        if (g_boot_id_set) memcpy(idbuf, g_boot_id, 128);
        else bootstrap_id_via_file(mpi_rank, idbuf);
        id = idbuf;
    }
    sdpa_ctx* ctx = nullptr;
    if (sdpa_ctx_create(&ctx, &cfg, id) != SDPA_OK) die("context creation");
    if (mpi_size > 1 && mpi_rank == 0 && !g_boot_id_set) {   // every rank has joined the communicator: the record is spent
        const char* path = getenv("SDPA_NCCL_ID_FILE");
        if (path && *path) unlink(path);
    }
    g_cached.ctx = ctx;
    g_cached.cfg = cfg;
    g_cached.world = cfg.world_size;
    g_cached.rank = cfg.rank_base;
    return ctx;
}

}  // namespace

extern "C" {

sdpa_status sdpa_set_bootstrap_id(const void* id128)
{
    if (!id128) {
        g_boot_id_set = false;
        return SDPA_OK;
    }
    memcpy(g_boot_id, id128, 128);
    g_boot_id_set = true;
    return SDPA_OK;
}

/* Creates the cached context (CUDA contexts, streams, NCCL communicator) ahead of the first
 * attention() call -- the analogue of MPI_Init (mpi.c:504), which the reference also keeps
 * outside its timed region (mpi.c:519-522). */
sdpa_status sdpa_ctx_prewarm(sdpa_ctx* ctx, size_t pool_bytes);

sdpa_status sdpa_runtime_init(int mpi_rank, int mpi_size)
{
    sdpa_ctx* ctx = get_ctx(mpi_rank, mpi_size);
    // memory pool, kernel code and host staging threads ahead of the timed call; SDPA_PREALLOC_MB sizes the pool (default 1024)
    return sdpa_ctx_prewarm(ctx, (size_t)env_int("SDPA_PREALLOC_MB", 1024) << 20);
}

sdpa_status sdpa_ctx_max(sdpa_ctx* ctx, double* value);

sdpa_status sdpa_runtime_max(double* value)
{
    if (!g_cached.ctx) {
        set_error("sdpa_runtime_max: no runtime context (call sdpa_runtime_init or attention first)");
        return SDPA_ERR_INVALID;
    }
    return sdpa_ctx_max(g_cached.ctx, value);
}

void sdpa_runtime_shutdown(void)
{
    if (g_cached.ctx) sdpa_ctx_destroy(g_cached.ctx);
    g_cached.ctx = nullptr;
}

sdpa_status sdpa_scatter_attention(sdpa_ctx* ctx, const double* Q, const double* K, const double* V,
                                   double* result, int m, int n, int dk, int dv);

void attention(double* Q, double* K, double* V, double* result, int m, int n, int dk, int dv, int mpi_rank,
               int mpi_size)
{
    if (mpi_size < 1 || mpi_rank < 0 || mpi_rank >= mpi_size) {
        fprintf(stderr, "sdpa_b200: attention(): bad mpi_rank/mpi_size %d/%d\n", mpi_rank, mpi_size);
        exit(1);
    }
    sdpa_ctx* ctx = get_ctx(mpi_rank, mpi_size);
    if (mpi_size == 1) {
        if (m < 0 || n < 0 || dk < 1 || dv < 1) {
            fprintf(stderr, "sdpa_b200: attention(): bad dimensions m=%d n=%d dk=%d dv=%d\n", m, n, dk, dv);
            exit(1);
        }
        if (sdpa_load_kv_host_full(ctx, K, V, n, dk, dv) != SDPA_OK) die("K/V upload");
        if (sdpa_attention_host(ctx, Q, result, m) != SDPA_OK) die("attention");
        return;
    }
    // one process per GPU: dimensions and data live on rank 0 (mpi.c:193-197, 508-517)
    if (sdpa_scatter_attention(ctx, Q, K, V, result, m, n, dk, dv) != SDPA_OK) die("sharded attention");
}

}  // extern "C"
