// merge_kernels.cu -- merging partial softmax states (split-KV inside a GPU and K/V shards
// across GPUs).  All of them evaluate the merge of attention-mpi.c:340-362 / SURVEY 3.3,
//     gmax = max_r lmax_r ; c_r = e^(lmax_r - gmax) ; gsum = sum_r lsum_r c_r ;
//     out  = sum_r contrib_r c_r / gsum          (gsum == 0 -> 0, mpi.c:359)
// with the max kept in the log2 domain (tmax = lmax * log2 e, so c_r = 2^(tmax_r - gmax)).
// These are O(rows * dv) HBM-bound passes; one warp owns one row, lanes stride over dv with
// 16-byte vectors, and the row statistics are combined with warp shuffles.
#include "common.cuh"

#include <math_constants.h>

namespace sdpa {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr float kLn2 = 0.6931471805599453f;

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Generic combine over `count` partial states addressed through per-state base pointers.
// Each state s: o_s[row*dv + d], tmax_s[row], lsum_s[row].
struct StatePtrs {
    const float* o[64];
    const float* tmax[64];
    const float* lsum[64];
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded spin (a protocol bug traps instead of hanging the GPU).  Epochs only grow; the signed
// difference tolerates wrap-around.  The bound is generous (default 120 s, SDPA_FLAG_TIMEOUT_S): ranks of one job may be
// seconds apart (a paused process, a first-call module load, a slow pageable upload on one rank) without that being an error.
__device__ long long g_flag_timeout_cycles = 240000000000LL;

__device__ __forceinline__ void spin_until(const unsigned int* flag, unsigned int epoch, int tag)
{
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(200);
        if (clock64() - t0 > g_flag_timeout_cycles) {
            printf("sdpa_b200: peer flag timeout tag=%d block=%d want=%u have=%u\n", tag, blockIdx.x, epoch, ld_acquire_sys(flag));
            __trap();
        }
    }
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// trace (SDPA_EXCHANGE_TRACE, developer aid): 4 x u64 per epoch -- [0] this rank published its state, and on the root
// [1] the merge kernel started, [2] every rank's flag had been seen, [3] the merge was complete (%globaltimer, ns).
__global__ void signal_flag_kernel(unsigned int* flag, unsigned int epoch, unsigned long long* trace)
{
    __threadfence_system();   // everything this GPU wrote before (earlier kernels on the stream) is visible system-wide
    st_release_sys(flag, epoch);
    if (trace) trace[0] = global_ns();
}
__global__ void wait_flag_kernel(const unsigned int* flag, unsigned int epoch) { spin_until(flag, epoch, 1); }

template <bool FINAL>
__device__ __forceinline__ void merge_one_row(const struct StatePtrs& st, int count, int row, int dv, double* __restrict__ out64_row,
                                              float* __restrict__ contrib_row, float* __restrict__ tmax_dst,
                                              float* __restrict__ lsum_dst, float max_unit, bool vec_ok);
template <bool FINAL>
__device__ __forceinline__ void merge_rows(const struct StatePtrs& st, int count, int rows, int dv, double* __restrict__ out64,
                                           float* __restrict__ contrib, float* __restrict__ tmax_out,
                                           float* __restrict__ lsum_out, float max_unit, bool vec_ok);

struct SyncArgs {
    const unsigned int* ready[64];
    unsigned int* consumed;
    unsigned int* block_counter;
    unsigned int epoch;        // value released into `consumed` when the last block is done
    unsigned int spin_epoch;   // value awaited in every `ready` flag at entry
    int enabled;               // bit 0: spin on the ready flags at entry; bit 1: release `consumed` at the end
    int nready;                // flags to spin on (0 = the kernel's state count)
    int root;                  // this is the root's cross-GPU merge (timeline trace slots 1-3 and 4+r)
    unsigned long long* trace;
};

// Entry / exit protocol shared by the merge kernels (sync.enabled bit 0 / bit 1):
//   entry: spin until every `ready` flag carries spin_epoch -- the shards' "state published" flags (cross-GPU merge on the
//          root), or the root's "slot consumed" flag (a shard's own split merge about to overwrite its exchange slot);
//   exit : the last block releases `consumed` = epoch -- "peers' slots may be reused" / "this shard's state is published".
__device__ __forceinline__ void sync_enter(const SyncArgs& sync, int count)
{
    if (!(sync.enabled & 1)) return;
    const int nready = sync.nready > 0 ? sync.nready : count;
    const bool root_merge = sync.root != 0;
    if (sync.trace && root_merge && blockIdx.x == 0 && threadIdx.x == 0) sync.trace[1] = global_ns();
    if ((int)threadIdx.x < nready) {
        spin_until(sync.ready[threadIdx.x], sync.spin_epoch, 2);
        if (sync.trace && root_merge && blockIdx.x == 0 && threadIdx.x < 8) sync.trace[4 + threadIdx.x] = global_ns();   // flag r seen (root's clock)
    }
    __syncthreads();
    if (sync.trace && root_merge && blockIdx.x == 0 && threadIdx.x == 0) sync.trace[2] = global_ns();
}
__device__ __forceinline__ void sync_exit(const SyncArgs& sync)
{
    if (!(sync.enabled & 2)) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // the rows just written may sit in another GPU's memory / are read by another GPU next
        if (atomicAdd(sync.block_counter, 1u) == gridDim.x - 1) {
            *sync.block_counter = 0;
            __threadfence_system();
            st_release_sys(sync.consumed, sync.epoch);
            if (sync.trace) sync.trace[sync.root ? 3 : 0] = global_ns();   // [3] root merge done, [0] state published
        }
    }
}

template <bool FINAL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
merge_states_kernel(StatePtrs st, int count, int rows, int dv, double* __restrict__ out64,
                    float* __restrict__ contrib, float* __restrict__ tmax_out,
                    float* __restrict__ lsum_out, float max_unit, bool vec_ok, const SyncArgs sync)
{
    sync_enter(sync, count);
    merge_rows<FINAL>(st, count, rows, dv, out64, contrib, tmax_out, lsum_out, max_unit, vec_ok);
    sync_exit(sync);
}

template <bool FINAL>
__device__ __forceinline__ void merge_rows(const StatePtrs& st, int count, int rows, int dv, double* __restrict__ out64,
                                           float* __restrict__ contrib, float* __restrict__ tmax_out,
                                           float* __restrict__ lsum_out, float max_unit, bool vec_ok)
{
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    merge_one_row<FINAL>(st, count, row, dv, FINAL ? out64 + (size_t)row * dv : nullptr, FINAL ? nullptr : contrib + (size_t)row * dv,
                         FINAL ? nullptr : tmax_out + row, FINAL ? nullptr : lsum_out + row, max_unit, vec_ok);
}

// One warp merges `count` states of row `row`; the destinations are already resolved to the row.
template <bool FINAL>
__device__ __forceinline__ void merge_one_row(const StatePtrs& st, int count, int row, int dv, double* __restrict__ out64_row,
                                              float* __restrict__ contrib_row, float* __restrict__ tmax_dst,
                                              float* __restrict__ lsum_dst, float max_unit, bool vec_ok)
{
    const int lane = threadIdx.x & 31;

    // Up to 8 states (the pieces of a row block, the shards of a box) and 16-byte vectors: the o vectors of this lane's first
    // four columns are requested BEFORE the statistics are reduced, so the row costs one memory round trip (local L2 or NVLink)
    // instead of two dependent ones (statistics -> shuffles -> vectors).
    const bool vec = (dv & 3) == 0 && vec_ok;
    const bool pre = vec && count <= 8 && lane * 4 < dv;
    float4 pv[8];
    if (pre) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < count) pv[s] = *reinterpret_cast<const float4*>(st.o[s] + (size_t)row * dv + lane * 4);
    }

    // lanes hold the per-state statistics (count <= 64: two per lane)
    float t0 = lane < count ? st.tmax[lane][row] : -CUDART_INF_F;
    float t1 = lane + 32 < count ? st.tmax[lane + 32][row] : -CUDART_INF_F;
    const float l0 = lane < count ? st.lsum[lane][row] : 0.f;
    const float l1 = lane + 32 < count ? st.lsum[lane + 32][row] : 0.f;
    float gmax = fmaxf(t0, t1);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
    // all states empty (-inf): weights 0, output 0
    const float w0 = (t0 == -CUDART_INF_F) ? 0.f : exp2f(t0 - gmax);
    const float w1 = (t1 == -CUDART_INF_F) ? 0.f : exp2f(t1 - gmax);
    float gsum = l0 * w0 + l1 * w1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, off);
    const float inv = (gsum == 0.f) ? 0.f : 1.f / gsum;

    // weights of the states, one per lane (two if count > 32), broadcast with shuffles below
    float wv[8];
    if (count <= 8) {   // warp-uniform: every lane takes part in these shuffles
#pragma unroll
        for (int s = 0; s < 8; ++s) wv[s] = __shfl_sync(0xffffffffu, w0, s);
    }
    if (vec) {
        // 16-byte vectors: lane owns 4 consecutive output columns per step
        for (int d = lane * 4; d < dv; d += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pre && d == lane * 4) {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    if (s < count) {
                        acc.x = fmaf(pv[s].x, wv[s], acc.x);
                        acc.y = fmaf(pv[s].y, wv[s], acc.y);
                        acc.z = fmaf(pv[s].z, wv[s], acc.z);
                        acc.w = fmaf(pv[s].w, wv[s], acc.w);
                    }
                }
            } else if (count <= 8) {
                for (int s = 0; s < count; ++s) {
                    const float4 v = *reinterpret_cast<const float4*>(st.o[s] + (size_t)row * dv + d);
                    const float w = wv[s & 7];
                    acc.x = fmaf(v.x, w, acc.x);
                    acc.y = fmaf(v.y, w, acc.y);
                    acc.z = fmaf(v.z, w, acc.z);
                    acc.w = fmaf(v.w, w, acc.w);
                }
            } else {
#pragma unroll 4
                for (int s = 0; s < count; ++s) {
                    const float w = __shfl_sync(0xffffffffu, s < 32 ? w0 : w1, s & 31);
                    const float4 v = *reinterpret_cast<const float4*>(st.o[s] + (size_t)row * dv + d);
                    acc.x = fmaf(v.x, w, acc.x);
                    acc.y = fmaf(v.y, w, acc.y);
                    acc.z = fmaf(v.z, w, acc.z);
                    acc.w = fmaf(v.w, w, acc.w);
                }
            }
            if (FINAL) {
                double2* dst = reinterpret_cast<double2*>(out64_row + d);
                dst[0] = make_double2((double)(acc.x * inv), (double)(acc.y * inv));
                dst[1] = make_double2((double)(acc.z * inv), (double)(acc.w * inv));
            } else {
                *reinterpret_cast<float4*>(contrib_row + d) = acc;
            }
        }
    } else {
        for (int d = lane; d < dv; d += 32) {
            float acc = 0.f;
            for (int s = 0; s < count; ++s) {
                const float w = __shfl_sync(0xffffffffu, s < 32 ? w0 : w1, s & 31);
                acc = fmaf(st.o[s][(size_t)row * dv + d], w, acc);
            }
            if (FINAL) out64_row[d] = (double)(acc * inv);
            else contrib_row[d] = acc;
        }
    }
    if (!FINAL && lane == 0) {
        *tmax_dst = gmax * max_unit;  // max_unit = 1 (log2 domain) or ln2 (reference's lmax)
        *lsum_dst = gsum;
    }
}

// contrib *= 2^(tmax - gmax), lsum *= 2^(tmax - gmax)        (mpi.c:346-351)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rescale_kernel(float* __restrict__ contrib, float* __restrict__ lsum, const float* __restrict__ tmax,
               const float* __restrict__ gmax, int rows, int dv)
{
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + warp;
    if (row >= rows) return;
    const float t = tmax[row];
    const float c = (t == -CUDART_INF_F) ? 0.f : exp2f(t - gmax[row]);
    for (int d = lane; d < dv; d += 32) contrib[(size_t)row * dv + d] *= c;
    if (lane == 0) lsum[row] *= c;
}

// contrib *= (gsum == 0 ? 0 : 1/gsum)                         (mpi.c:358-362)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_kernel(float* __restrict__ contrib, const float* __restrict__ gsum, int rows, int dv)
{
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + warp;
    if (row >= rows) return;
    const float g = gsum[row];
    const float inv = (g == 0.f) ? 0.f : 1.f / g;
    for (int d = lane; d < dv; d += 32) contrib[(size_t)row * dv + d] *= inv;
}

// out64 = reduced / gsum  (mpi.c:358-362 applied after the reduce instead of before it, plus cvt_f2d mpi.c:373)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
finalize_reduced_kernel(double* __restrict__ out64, const float* __restrict__ reduced, const float* __restrict__ gsum,
                        int rows, int dv)
{
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + warp;
    if (row >= rows) return;
    const float g = gsum[row];
    const float inv = (g == 0.f) ? 0.f : 1.f / g;
    for (int d = lane; d < dv; d += 32) out64[(size_t)row * dv + d] = (double)(reduced[(size_t)row * dv + d] * inv);
}

// Split merge with routing (push form of the cross-GPU exchange): the merged state of row `row` is written straight
// into the inbox of the rank that owns the row's slice -- a posted NVLink store instead of a remote read later -- and the
// last block raises this rank's "delivered" flag at every peer.
struct RouteArgs {
    float* o[64];               // per destination rank: this source's segment of its inbox (rows of its slice)
    float* tmax[64];
    float* lsum[64];
    unsigned int* flag[64];     // per destination rank: "source delivered epoch e"
    unsigned int* block_counter;
    unsigned int epoch;
    int base, rem, world;       // slice r has base + (r < rem) rows, slices are consecutive
    // states per row: `count`, or (persistent fused kernel) the pieces of the row's block / all max_pieces after the exact twin
    WorkMap wm;
    int max_pieces;             // 0: every row has `count` states
    const unsigned int* guard;
    unsigned int guard_epoch;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
merge_route_kernel(StatePtrs st, int count, int rows, int dv, RouteArgs rt, bool vec_ok)
{
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row < rows) {
        const int boundary = rt.rem * (rt.base + 1);
        int r, local;
        if (row < boundary) {
            r = row / (rt.base + 1);
            local = row - r * (rt.base + 1);
        } else {
            r = rt.rem + (row - boundary) / rt.base;
            local = (row - boundary) - (r - rt.rem) * rt.base;
        }
        const int n = rt.max_pieces > 0 ? ((*rt.guard == rt.guard_epoch) ? rt.max_pieces : wm_pieces(rt.wm, row / 256)) : count;
        merge_one_row<false>(st, n, row, dv, nullptr, rt.o[r] + (size_t)local * dv, rt.tmax[r] + local, rt.lsum[r] + local,
                             1.f, vec_ok);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(rt.block_counter, 1u) == gridDim.x - 1) {
            *rt.block_counter = 0;
            __threadfence_system();
            for (int r = 0; r < rt.world; ++r) st_release_sys(rt.flag[r], rt.epoch);
        }
    }
}

// Split merge behind the persistent fused kernel (attn_umma_kernel_v8): the number of partial states of a row is the
// number of pieces its row block was cut into (wm_pieces), unless the overflow guard handed the launch to the SAFE
// kernel, which fills all `max_pieces` slots with equal splits.
template <bool FINAL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
merge_pieces_kernel(StatePtrs st, WorkMap wm, int max_pieces, int rows, int dv, double* __restrict__ out64, float* __restrict__ contrib,
                    float* __restrict__ tmax_out, float* __restrict__ lsum_out, bool vec_ok, const unsigned int* __restrict__ guard,
                    unsigned int epoch, const SyncArgs sync, int base, int fixed_own)
{
    // states [0, base): other shards' published states (root's in-stream merge); then this shard's own partial states:
    // fixed_own split states, or (persistent fused kernel) the pieces of the row's block / all max_pieces after the exact twin
    sync_enter(sync, 0);
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row < rows) {
        const int count = base + (fixed_own > 0 ? fixed_own : ((*guard == epoch) ? max_pieces : wm_pieces(wm, row / 256)));
        merge_one_row<FINAL>(st, count, row, dv, FINAL ? out64 + (size_t)row * dv : nullptr, FINAL ? nullptr : contrib + (size_t)row * dv,
                             FINAL ? nullptr : tmax_out + row, FINAL ? nullptr : lsum_out + row, 1.f, vec_ok);
    }
    sync_exit(sync);
}

// Root GPU, sliced merge: wait until every rank has delivered its rows of the batch into the staging buffer,
// move them to their destination (the caller's result array or the D2H buffer) and release the slot.
__global__ void __launch_bounds__(256)
collect_slices_kernel(double2* __restrict__ dst, const double2* __restrict__ src, size_t units, SyncArgs sync, int count)
{
    if (threadIdx.x < count) spin_until(sync.ready[threadIdx.x], sync.spin_epoch, 3);
    __syncthreads();
    const size_t total = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += total) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(sync.block_counter, 1u) == gridDim.x - 1) {
            *sync.block_counter = 0;
            __threadfence_system();
            st_release_sys(sync.consumed, sync.epoch);   // staging slot and every rank's state slot may be reused
        }
    }
}

// Root GPU, push form of the exchange: every shard has written its state into segment r of the root's LOCAL inbox (posted
// stores over NVLink) and raised ready[r] in the root's memory.  This kernel merges the `world` states of every row and writes
// the normalised fp64 rows; the last block hands the slot back by writing `consumed` into every rank's own flag block (again
// posted stores: nobody polls over NVLink).  It is a BACKGROUND kernel: sm_count CTAs of 128 threads and <= 32 registers, no
// shared memory, so that it runs beside the persistent fused kernel of the next queued pass instead of between two passes.
struct InboxArgs {
    const float* inbox;          // segment r at r * seg floats: [o cap_rows*dv | tmax cap_rows | lsum cap_rows]
    size_t seg;
    int cap_rows, world;
    const unsigned int* ready;   // ready[r], local
    unsigned int* consumed[64];  // per rank (own block for the root)
    unsigned int* block_counter;
    unsigned int epoch;
    unsigned long long* trace;
};

__global__ void __launch_bounds__(128, 16)
merge_inbox_bg_kernel(const InboxArgs a, double* __restrict__ out64, int rows, int dv, bool vec_ok)
{
    const int lane = threadIdx.x & 31;
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[1] = global_ns();
    if ((int)threadIdx.x < a.world) {
        spin_until(a.ready + threadIdx.x, a.epoch, 4);
        if (a.trace && blockIdx.x == 0 && threadIdx.x < 8) a.trace[4 + threadIdx.x] = global_ns();
    }
    __syncthreads();
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[2] = global_ns();
    const float* tmax0 = a.inbox + (size_t)a.cap_rows * dv;
    const float* lsum0 = tmax0 + a.cap_rows;
    const bool vec = vec_ok && (dv & 3) == 0;
    for (int row = blockIdx.x * 4 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 4) {
        const float t0 = lane < a.world ? tmax0[(size_t)lane * a.seg + row] : -CUDART_INF_F;
        const float t1 = lane + 32 < a.world ? tmax0[(size_t)(lane + 32) * a.seg + row] : -CUDART_INF_F;
        const float l0 = lane < a.world ? lsum0[(size_t)lane * a.seg + row] : 0.f;
        const float l1 = lane + 32 < a.world ? lsum0[(size_t)(lane + 32) * a.seg + row] : 0.f;
        float gmax = fmaxf(t0, t1);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
        const float w0 = (t0 == -CUDART_INF_F) ? 0.f : exp2f(t0 - gmax);
        const float w1 = (t1 == -CUDART_INF_F) ? 0.f : exp2f(t1 - gmax);
        float gsum = l0 * w0 + l1 * w1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, off);
        const float inv = (gsum == 0.f) ? 0.f : 1.f / gsum;
        const float* orow = a.inbox + (size_t)row * dv;
        double* dst = out64 + (size_t)row * dv;
        if (vec) {
            for (int d0 = 0; d0 < dv; d0 += 128) {   // warp-uniform trip count: the shuffles below need every lane
                const int d = d0 + lane * 4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s = 0; s < a.world; ++s) {
                    const float w = __shfl_sync(0xffffffffu, s < 32 ? w0 : w1, s & 31);
                    if (d < dv) {
                        const float4 v = *reinterpret_cast<const float4*>(orow + (size_t)s * a.seg + d);
                        acc.x = fmaf(v.x, w, acc.x);
                        acc.y = fmaf(v.y, w, acc.y);
                        acc.z = fmaf(v.z, w, acc.z);
                        acc.w = fmaf(v.w, w, acc.w);
                    }
                }
                if (d < dv) {
                    reinterpret_cast<double2*>(dst + d)[0] = make_double2((double)(acc.x * inv), (double)(acc.y * inv));
                    reinterpret_cast<double2*>(dst + d)[1] = make_double2((double)(acc.z * inv), (double)(acc.w * inv));
                }
            }
        } else {
            for (int d0 = 0; d0 < dv; d0 += 32) {
                const int d = d0 + lane;
                float acc = 0.f;
                for (int s = 0; s < a.world; ++s) {
                    const float w = __shfl_sync(0xffffffffu, s < 32 ? w0 : w1, s & 31);
                    if (d < dv) acc = fmaf(orow[(size_t)s * a.seg + d], w, acc);
                }
                if (d < dv) dst[d] = (double)(acc * inv);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.block_counter, 1u) == gridDim.x - 1) {
            *a.block_counter = 0;
            __threadfence_system();
            for (int r = 0; r < a.world; ++r) st_release_sys(a.consumed[r], a.epoch);
            if (a.trace) a.trace[3] = global_ns();
        }
    }
}

}  // namespace


void preload_merge_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, merge_pieces_kernel<true>);
    cudaFuncGetAttributes(&a, merge_pieces_kernel<false>);
    cudaFuncGetAttributes(&a, merge_states_kernel<true>);
    cudaFuncGetAttributes(&a, merge_states_kernel<false>);
    cudaFuncGetAttributes(&a, merge_inbox_bg_kernel);
    cudaFuncGetAttributes(&a, signal_flag_kernel);
    cudaFuncGetAttributes(&a, wait_flag_kernel);
    cudaGetLastError();
}

sdpa_status launch_finalize_reduced(double* out64, const float* reduced, const float* gsum, int rows, int dv,
                                    cudaStream_t stream)
{
    if (rows <= 0) return SDPA_OK;
    finalize_reduced_kernel<<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(out64, reduced, gsum, rows, dv);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

// `publish` (exchange slot hand-over of one shard): wait until *publish->wait_flag >= wait_epoch (the root has consumed the
// slot's previous content; NULL = no wait), merge, then release publish->flag = epoch -- all inside the merge kernel.
static SyncArgs publish_args(const PublishSync* publish)
{
    SyncArgs sa{};
    if (!publish) return sa;
    sa.enabled = 2 | (publish->wait_flag ? 1 : 0);
    sa.ready[0] = publish->wait_flag;
    sa.nready = 1;
    sa.spin_epoch = publish->wait_epoch;
    sa.consumed = publish->flag;
    sa.block_counter = publish->block_counter;
    sa.epoch = publish->epoch;
    sa.trace = publish->trace;
    return sa;
}

sdpa_status launch_merge_splits(Partials part, int rows, int dv, double* out64, float* contrib,
                                float* tmax_out, float* lsum_out, bool natural_log_max,
                                cudaStream_t stream, const PublishSync* publish)
{
    if (rows <= 0 && !publish) return SDPA_OK;
    if (rows < 0) rows = 0;
    if (part.splits < 1 || part.splits > 64) {
        set_error("merge supports 1..64 split states (got %d)", part.splits);
        return SDPA_ERR_INVALID;
    }
    StatePtrs st;
    for (int s = 0; s < part.splits; ++s) {
        st.o[s] = part.o + (size_t)s * part.rows_capacity * dv;
        st.tmax[s] = part.tmax + (size_t)s * part.rows_capacity;
        st.lsum[s] = part.lsum + (size_t)s * part.rows_capacity;
    }
    const int blocks = std::max(1, ceil_div(rows, kWarpsPerBlock));
    bool vec_ok = al16(out64) && al16(contrib);
    for (int s = 0; s < part.splits; ++s) vec_ok = vec_ok && al16(st.o[s]);
    const SyncArgs sa = publish_args(publish);
    if (out64 != nullptr)
        merge_states_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, part.splits, rows, dv, out64, nullptr, nullptr, nullptr, 1.f,
                                                                             vec_ok, sa);
    else
        merge_states_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, part.splits, rows, dv, nullptr, contrib, tmax_out, lsum_out,
                                                                              natural_log_max ? kLn2 : 1.f, vec_ok, sa);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_merge_peers(const float* const* contrib_ptrs, const float* const* tmax_ptrs,
                               const float* const* lsum_ptrs, int shards, int rows, int dv,
                               double* out64, cudaStream_t stream)
{
    if (rows <= 0) return SDPA_OK;
    if (shards < 1 || shards > 64) {
        set_error("peer merge supports 1..64 shards (got %d)", shards);
        return SDPA_ERR_INVALID;
    }
    StatePtrs st;
    for (int s = 0; s < shards; ++s) {
        st.o[s] = contrib_ptrs[s];
        st.tmax[s] = tmax_ptrs[s];
        st.lsum[s] = lsum_ptrs[s];
    }
    const int blocks = ceil_div(rows, kWarpsPerBlock);
    bool vec_ok = al16(out64);
    for (int s = 0; s < shards; ++s) vec_ok = vec_ok && al16(st.o[s]);
    merge_states_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, shards, rows, dv, out64, nullptr,
                                                                         nullptr, nullptr, 1.f, vec_ok, SyncArgs{});
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_merge_peers_synced(const float* const* contrib_ptrs, const float* const* tmax_ptrs,
                                      const float* const* lsum_ptrs, int shards, int rows, int dv, double* out64,
                                      const PeerSync& sync, cudaStream_t stream)
{
    if (rows < 0) rows = 0;   // a rank without rows still has to raise its flag: one block, no row work
    if (shards < 1 || shards > 64) {
        set_error("peer merge supports 1..64 shards (got %d)", shards);
        return SDPA_ERR_INVALID;
    }
    StatePtrs st;
    SyncArgs sa{};
    for (int s = 0; s < shards; ++s) {
        st.o[s] = contrib_ptrs[s];
        st.tmax[s] = tmax_ptrs[s];
        st.lsum[s] = lsum_ptrs[s];
        sa.ready[s] = sync.ready[s];
    }
    sa.consumed = sync.consumed;
    sa.block_counter = sync.block_counter;
    sa.epoch = sa.spin_epoch = sync.epoch;
    sa.enabled = 3;
    sa.nready = 0;
    sa.root = sync.trace ? 1 : 0;
    sa.trace = sync.trace;
    const int blocks = std::max(1, ceil_div(rows, kWarpsPerBlock));
    bool vec_ok = al16(out64);
    for (int s = 0; s < shards; ++s) vec_ok = vec_ok && al16(st.o[s]);
    merge_states_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, shards, rows, dv, out64, nullptr, nullptr,
                                                                         nullptr, 1.f, vec_ok, sa);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_merge_pieces(Partials part, const WorkMap& wm, int max_pieces, int rows, int dv, double* out64, float* contrib,
                                float* tmax_out, float* lsum_out, const unsigned int* guard, unsigned int epoch, cudaStream_t stream,
                                const PublishSync* publish)
{
    if (rows <= 0 && !publish) return SDPA_OK;
    if (rows < 0) rows = 0;
    if (max_pieces < 1 || max_pieces > 64 || max_pieces > part.splits || !guard || (!out64 && !(contrib && tmax_out && lsum_out))) {
        set_error("merge_pieces: bad arguments (pieces=%d, partial slots=%d)", max_pieces, part.splits);
        return SDPA_ERR_INVALID;
    }
    StatePtrs st;
    bool vec_ok = al16(out64) && al16(contrib);
    for (int s = 0; s < max_pieces; ++s) {
        st.o[s] = part.o + (size_t)s * part.rows_capacity * dv;
        st.tmax[s] = part.tmax + (size_t)s * part.rows_capacity;
        st.lsum[s] = part.lsum + (size_t)s * part.rows_capacity;
        vec_ok = vec_ok && al16(st.o[s]);
    }
    const int blocks = std::max(1, ceil_div(rows, kWarpsPerBlock));
    const SyncArgs sa = publish_args(publish);
    if (out64) merge_pieces_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, wm, max_pieces, rows, dv, out64, nullptr, nullptr, nullptr,
                                                                                    vec_ok, guard, epoch, sa, 0, 0);
    else merge_pieces_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, wm, max_pieces, rows, dv, nullptr, contrib, tmax_out, lsum_out,
                                                                               vec_ok, guard, epoch, sa, 0, 0);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

// Root GPU, in-stream form of the cross-GPU merge: ONE kernel merges the root's own partial states (pieces of the persistent
// kernel when wm != NULL, else part.splits split states) with the other shards' published states -- it waits for their flags
// at entry, writes the normalised fp64 rows and releases "consumed".  The root neither publishes a state of its own nor runs a
// second merge.
sdpa_status launch_merge_root_instream(Partials part, const WorkMap* wm, int max_pieces, const unsigned int* guard, unsigned int guard_epoch,
                                       const float* const* peer_c, const float* const* peer_t, const float* const* peer_l, int npeers,
                                       int rows, int dv, double* out64, const PeerSync& sync, cudaStream_t stream)
{
    const int own = wm ? max_pieces : part.splits;
    if (npeers < 1 || own < 1 || npeers + own > 64 || !out64) {
        set_error("in-stream root merge: %d peers + %d own states (at most 64 together)", npeers, own);
        return SDPA_ERR_INVALID;
    }
    if (rows < 0) rows = 0;
    StatePtrs st;
    bool vec_ok = al16(out64);
    for (int r = 0; r < npeers; ++r) {
        st.o[r] = peer_c[r];
        st.tmax[r] = peer_t[r];
        st.lsum[r] = peer_l[r];
        vec_ok = vec_ok && al16(st.o[r]);
    }
    for (int k = 0; k < own; ++k) {
        st.o[npeers + k] = part.o + (size_t)k * part.rows_capacity * dv;
        st.tmax[npeers + k] = part.tmax + (size_t)k * part.rows_capacity;
        st.lsum[npeers + k] = part.lsum + (size_t)k * part.rows_capacity;
        vec_ok = vec_ok && al16(st.o[npeers + k]);
    }
    SyncArgs sa{};
    for (int r = 0; r < npeers; ++r) sa.ready[r] = sync.ready[r];
    sa.nready = npeers;
    sa.consumed = sync.consumed;
    sa.block_counter = sync.block_counter;
    sa.epoch = sa.spin_epoch = sync.epoch;
    sa.enabled = 3;
    sa.root = 1;
    sa.trace = sync.trace;
    const int blocks = std::max(1, ceil_div(rows, kWarpsPerBlock));
    merge_pieces_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(st, wm ? *wm : WorkMap{1, 1, 1}, max_pieces, rows, dv, out64, nullptr, nullptr,
                                                                         nullptr, vec_ok, wm ? guard : nullptr, guard_epoch, sa, npeers,
                                                                         wm ? 0 : part.splits);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_merge_inbox_background(const float* inbox, size_t seg_floats, int cap_rows, int world, const unsigned int* ready,
                                          unsigned int* const* consumed, unsigned int* block_counter, unsigned int epoch,
                                          unsigned long long* trace, int rows, int dv, double* out64, int ctas, cudaStream_t stream)
{
    if (world < 1 || world > 64 || !inbox || !out64 || ctas < 1) {
        set_error("inbox merge: 1..64 shards, an inbox and a destination");
        return SDPA_ERR_INVALID;
    }
    if (rows < 0) rows = 0;
    InboxArgs a{};
    a.inbox = inbox;
    a.seg = seg_floats;
    a.cap_rows = cap_rows;
    a.world = world;
    a.ready = ready;
    for (int r = 0; r < world; ++r) a.consumed[r] = consumed[r];
    a.block_counter = block_counter;
    a.epoch = epoch;
    a.trace = trace;
    const bool vec_ok = al16(out64) && al16(inbox) && (seg_floats % 4 == 0) && (((size_t)cap_rows * dv) % 4 == 0);
    {
        static bool carve_done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !carve_done[dev]) {   // never makes an SM re-partition its shared memory under the fused kernel
            cudaFuncSetAttribute(merge_inbox_bg_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve_done[dev] = true;
        }
    }
    merge_inbox_bg_kernel<<<ctas, 128, 0, stream>>>(a, out64, rows, dv, vec_ok);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_merge_splits_routed(Partials part, int rows, int dv, const RouteTargets& to, cudaStream_t stream, const WorkMap* wm,
                                       int max_pieces, const unsigned int* guard, unsigned int guard_epoch)
{
    if (part.splits < 1 || part.splits > 64 || to.world < 1 || to.world > 64) {
        set_error("routed merge supports 1..64 split states and ranks");
        return SDPA_ERR_INVALID;
    }
    if (rows < 0) rows = 0;
    StatePtrs st;
    bool vec_ok = true;
    for (int s = 0; s < part.splits; ++s) {
        st.o[s] = part.o + (size_t)s * part.rows_capacity * dv;
        st.tmax[s] = part.tmax + (size_t)s * part.rows_capacity;
        st.lsum[s] = part.lsum + (size_t)s * part.rows_capacity;
        vec_ok = vec_ok && al16(st.o[s]);
    }
    RouteArgs rt;
    for (int r = 0; r < to.world; ++r) {
        rt.o[r] = to.o[r];
        rt.tmax[r] = to.tmax[r];
        rt.lsum[r] = to.lsum[r];
        rt.flag[r] = to.flag[r];
        vec_ok = vec_ok && al16(to.o[r]);
    }
    rt.block_counter = to.block_counter;
    rt.epoch = to.epoch;
    rt.world = to.world;
    rt.base = rows / to.world;
    rt.rem = rows % to.world;
    rt.wm = wm ? *wm : WorkMap{0, 0, 0};
    rt.max_pieces = wm ? max_pieces : 0;
    rt.guard = guard;
    rt.guard_epoch = guard_epoch;
    merge_route_kernel<<<std::max(1, ceil_div(rows, kWarpsPerBlock)), kWarpsPerBlock * 32, 0, stream>>>(st, part.splits, rows, dv, rt,
                                                                                                       vec_ok);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_collect_slices(double* dst, const double* staged, int rows, int dv, const PeerSync& sync, int ranks,
                                  cudaStream_t stream)
{
    if (ranks < 1 || ranks > 64) {
        set_error("collect_slices supports 1..64 ranks (got %d)", ranks);
        return SDPA_ERR_INVALID;
    }
    if (((size_t)rows * dv) % 2 != 0 || !al16(dst) || !al16(staged)) {
        set_error("collect_slices needs 16-byte aligned buffers and an even element count");
        return SDPA_ERR_INVALID;
    }
    SyncArgs sa{};
    for (int r = 0; r < ranks; ++r) sa.ready[r] = sync.ready[r];
    sa.consumed = sync.consumed;
    sa.block_counter = sync.block_counter;
    sa.epoch = sa.spin_epoch = sync.epoch;
    sa.enabled = 3;
    const size_t units = (size_t)rows * dv / 2;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(148 * 4, (units + 255) / 256));
    collect_slices_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<double2*>(dst), reinterpret_cast<const double2*>(staged),
                                                      units, sa, ranks);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

// Sets the spin bound of the flag waits on the current device (seconds at a nominal 2 GHz).
sdpa_status set_flag_timeout_seconds(double seconds)
{
    const long long cycles = (long long)(seconds * 2.0e9);
    SDPA_CUDA_TRY(cudaMemcpyToSymbol(g_flag_timeout_cycles, &cycles, sizeof(cycles)));
    return SDPA_OK;
}

sdpa_status launch_signal_flag(unsigned int* flag, unsigned int epoch, cudaStream_t stream, unsigned long long* trace)
{
    signal_flag_kernel<<<1, 1, 0, stream>>>(flag, epoch, trace);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_wait_flag(const unsigned int* flag, unsigned int epoch, cudaStream_t stream)
{
    wait_flag_kernel<<<1, 1, 0, stream>>>(flag, epoch);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_rescale_to_gmax(float* contrib, float* lsum, const float* tmax, const float* gmax,
                                   int rows, int dv, cudaStream_t stream)
{
    if (rows <= 0) return SDPA_OK;
    rescale_kernel<<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(contrib, lsum, tmax,
                                                                                       gmax, rows, dv);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

sdpa_status launch_normalize(float* contrib, const float* gsum, int rows, int dv, cudaStream_t stream)
{
    if (rows <= 0) return SDPA_OK;
    normalize_kernel<<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(contrib, gsum, rows, dv);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

}  // namespace sdpa
