// attn_umma_bf16.cu -- fused QK^T -> online softmax -> .V on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA), bf16 operands / fp32 accumulate.
//
// Three kernel generations live here behind one launcher (launch_attn_umma); all produce the same
// partial softmax state and pass the same parity tests:
//   attn_umma_kernel_v7  DEFAULT.  Cluster of two CTAs forming one M=256 tcgen05.mma.cta_group::2 (each CTA keeps
//                        half of every K/V tile), S and P double-buffered in TMEM (no MMA waits for the softmax of
//                        its own S), two softmax groups ping-ponged over alternating key tiles.  See its banner.
//   attn_umma_kernel     v5 (SDPA_UMMA_V7=0): one CTA, two Q tiles ping-ponged, P aliases S.  Its SAFE variant is
//                        the overflow-guard fallback of every generation.  Described right below.
//   attn_umma_kernel_v8  (SDPA_UMMA_V8=1, EXPERIMENTAL, single GPU) v7 made persistent: one cluster per SM pair walks a
//                        contiguous range of (row block, key tile) units; written after the round's GPU budget was
//                        spent -- its barrier protocol is checked by tools/v8_protocol_sim.py, not yet by hardware.
//
// The B200 counterpart of online_softmax_attention (attention-mpi.c:168-189): where the
// reference does one AVX-512 dot (dot_avx512, :103-121) and one axpy (axpy_avx512, :123-140)
// per (query, key) pair, this kernel does two 128x128x128 tensor-core GEMMs per
// (128-query tile, 128-key tile) pair; the running max / running sum of mpi.c:177-180 live in
// registers, one query row per thread.
//
// CTA = 256 query rows (two 128-row tiles A and B, ping-ponged) x one contiguous range of
// 128-key tiles (split-KV over blockIdx.y).  20 warps:
//   warp 0      TMA producer   : Q tiles once, then K/V tiles through 2-stage mbarrier rings
//                                (cp.async.bulk.tensor, 128-byte swizzle)
//   warp 1      MMA issuer     : one thread issues   S_t = Q_t K^T   (SS, both K-major)
//                                and                 O_t += P_t V    (TS: P from TMEM, V MN-major)
//                                order  S_A(0) S_B(0) | PV_A(j) S_A(j+1) PV_B(j) S_B(j+1) | ...
//                                so the softmax of one tile overlaps the MMAs of the other
//   warps 2-3   idle (keep the softmax warpgroups aligned to the TMEM lane quadrants)
//   warps 4-11  softmax of tile A, warps 12-19 softmax of tile B.  A tile has TWO warpgroups:
//               a thread owns one query row and one 64-key half of it, so every SM sub-partition
//               runs two warps of the same tile and the MUFU (ex2) pipe of one is fed while the
//               other issues its FFMA2/FADD2/F2FP (measured on v1: one warp per sub-partition
//               reaches only ~0.3 IPC and the per-tile chain softmax -> PV -> next S is serial,
//               see profiles/r01).  Per tile: tcgen05.ld the half row, 8-chain FMNMX3 max, the two
//               halves' maxima are exchanged through shared memory (named barrier, 256 threads),
//               lazy rescale of O (only when the max grew by > 2^8, warp vote), exp2 with the
//               1/sqrt(dk)*log2(e) scale folded into packed FFMA2, packed FADD2 row sums, bf16 P
//               written back into the TMEM columns of S (tcgen05.st); finally the epilogue.
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512); P_t aliases the
// first 64 columns of S_t (the tensor pipe executes MMAs in issue order, so S_t(j+1) cannot
// overwrite P_t(j) before PV_t(j) has consumed it).
// Shared memory: Q_A, Q_B, 2 x K, 2 x V tiles of 32 KiB = 192 KiB.
// Roofline: tensor pipe, 4*128^3 flops per tile pair; algorithmic HBM bytes are the bf16
// Q/K/V and the fp32 partial outputs.
#include "umma_ptx.cuh"
#include "umma_general.h"

#include <type_traits>
#include <vector>
#include <stdlib.h>

namespace sdpa {

namespace {

using namespace umma;

constexpr int HEAD = 128;            // dk == dv
constexpr int BLOCK_ROWS = 2 * TILE; // Q rows per CTA
constexpr int NTHREADS = 640;
constexpr int STAGES = 2;
constexpr float kLazyThreshold = 8.0f;                // safe mode: rescale O only if the max grew by > 2^8
constexpr float kGuardThreshold = 64.0f;              // fast mode: exponents beyond 2^64 hand the launch to the safe kernel
constexpr int kDefaultPoly = 0;                       // of every 16 exponentials, this many run on the FMA pipe (0, 4 or 8)

constexpr uint32_t TMEM_S = 0;    // + 128 * tile
constexpr uint32_t TMEM_O = 256;  // + 128 * tile

struct __align__(1024) SharedStorage {
    uint8_t q[2][TILE_BYTES];
    uint8_t k[STAGES][TILE_BYTES];
    uint8_t v[STAGES][TILE_BYTES];
    uint64_t q_full[2];
    uint64_t k_full[STAGES], k_empty[STAGES];
    uint64_t v_full[STAGES], v_empty[STAGES];
    uint64_t s_full[2], p_ready[2], o_done[2];
    uint32_t tmem_base;
    float xchg[2][2][2][TILE];   // [tile][parity][column half][row]: row-max / row-sum exchange
};


struct KernelParams {
    int rows;            // valid Q rows
    int n;               // keys in the shard
    int tiles_total;     // ceil(n / 128)
    int splits;
    float scale_log2;    // 1/sqrt(dk) * log2(e)
    float* part_o;
    float* part_tmax;
    float* part_lsum;
    int rows_capacity;
    double* out64;       // non-null (splits == 1): normalised fp64 output
    long long* trace;    // TRACE build only: clock64 stamps of CTA (0,0), [role][iteration][event]
    unsigned int* guard; // fast mode writes `epoch` here when an exponent would overflow; the safe kernel runs iff *guard == epoch
    unsigned int epoch;
    WorkMap wm;          // v8 only: the persistent kernel's work decomposition
};

constexpr int TRACE_ITERS = 24, TRACE_EVENTS = 8, TRACE_ROLES = 6;

// SAFE = the always-correct variant (row max agreed every tile, lazy rescale).  The default launch is
// the fast variant (reference fixed after the first tile) followed by the SAFE variant, which exits
// immediately unless the fast one raised the overflow guard.  POLY: see kDefaultPoly.
template <bool TRACE, bool SAFE, int POLY, bool CHUNKED = true>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_umma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const KernelParams prm)
{
    if constexpr (SAFE) {
        if (*prm.guard != prm.epoch) return;   // nothing overflowed in the fast pass: the whole grid leaves at once
    }
    extern __shared__ uint8_t smem_raw[];
    SharedStorage& sm = *reinterpret_cast<SharedStorage*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int row_block = blockIdx.x;
    const int split = blockIdx.y;

    // balanced contiguous partition of the key tiles over the splits (same rule as owner_count/owner_disp)
    const int tq = prm.tiles_total / prm.splits, tr = prm.tiles_total % prm.splits;
    const int tile_begin = split * tq + min(split, tr);
    const int num_tiles = tq + (split < tr ? 1 : 0);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_k);
        prefetch_tensormap(&map_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.q_full[i], 1);
            mbar_init(&sm.s_full[i], 1);
            mbar_init(&sm.p_ready[i], 256);
            mbar_init(&sm.o_done[i], 1);
        }
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&sm.k_full[i], 1);
            mbar_init(&sm.k_empty[i], 1);
            mbar_init(&sm.v_full[i], 1);
            mbar_init(&sm.v_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&sm.tmem_base, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = sm.tmem_base;
    auto stamp = [&](int role, int j, int ev) {
        if constexpr (TRACE) {
            if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && j < TRACE_ITERS)
                prm.trace[(role * TRACE_ITERS + j) * TRACE_EVENTS + ev] = clock64();
        }
    };

    // Register re-allocation between the warpgroups (each branch is dominated by its own
    // setmaxnreg, so ptxas budgets it separately): the producer / MMA warpgroup needs few
    // registers, each softmax thread holds a 64-column half of an S row (128*64 + 512*104 = 640*96).
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (num_tiles > 0) {
        if (warp == 0) {
            // ================================ TMA producer ================================
            // all 32 lanes walk the loop (uniform control flow); one elected lane issues
            const int qrow = row_block * BLOCK_ROWS;
            for (int j = 0; j < num_tiles; ++j) {
                const int stage = j % STAGES;
                const uint32_t ph = (uint32_t)(j / STAGES) & 1u;
                const int key0 = (tile_begin + j) * TILE;
                if (j == 0 && elect_one_sync()) {
                    mbar_arrive_expect_tx(&sm.q_full[0], TILE_BYTES);
                    tma_load_2d(sm.q[0], &map_q, &sm.q_full[0], 0, qrow);
                    tma_load_2d(sm.q[0] + HALF_BYTES, &map_q, &sm.q_full[0], 64, qrow);
                }
                mbar_wait(&sm.k_empty[stage], ph ^ 1u, 100 + stage);
                stamp(5, j, 0);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&sm.k_full[stage], TILE_BYTES);
                    tma_load_2d(sm.k[stage], &map_k, &sm.k_full[stage], 0, key0);
                    tma_load_2d(sm.k[stage] + HALF_BYTES, &map_k, &sm.k_full[stage], 64, key0);
                }
                if (j == 0 && elect_one_sync()) {
                    mbar_arrive_expect_tx(&sm.q_full[1], TILE_BYTES);
                    tma_load_2d(sm.q[1], &map_q, &sm.q_full[1], 0, qrow + TILE);
                    tma_load_2d(sm.q[1] + HALF_BYTES, &map_q, &sm.q_full[1], 64, qrow + TILE);
                }
                mbar_wait(&sm.v_empty[stage], ph ^ 1u, 110 + stage);
                stamp(5, j, 1);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&sm.v_full[stage], TILE_BYTES);
                    tma_load_2d(sm.v[stage], &map_v, &sm.v_full[stage], 0, key0);
                    tma_load_2d(sm.v[stage] + HALF_BYTES, &map_v, &sm.v_full[stage], 64, key0);
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // ================================ MMA issuer ==================================
            // Uniform control flow for the whole warp; descriptors are base + constant, the
            // tcgen05.mma / tcgen05.commit instructions sit under elect_one_sync().
            constexpr uint32_t idesc_qk = make_idesc(TILE, TILE, 0);
            constexpr uint32_t idesc_pv = make_idesc(TILE, HEAD, 1);
            const uint64_t dq[2] = {desc_kmajor(smem_u32(sm.q[0]), 0), desc_kmajor(smem_u32(sm.q[1]), 0)};
            const uint64_t dkk[STAGES] = {desc_kmajor(smem_u32(sm.k[0]), 0), desc_kmajor(smem_u32(sm.k[1]), 0)};
            const uint64_t dvv[STAGES] = {desc_mnmajor(smem_u32(sm.v[0]), 0), desc_mnmajor(smem_u32(sm.v[1]), 0)};

            // S_t = Q_t K(stage)^T ; commit -> s_full[t]  (+ optionally release the K stage)
            auto issue_s = [&](int t, int stage, bool release_k) {
                if (elect_one_sync()) {
                    const uint64_t a0 = dq[t], b0 = dkk[stage];
                    const uint32_t d = tmem + TMEM_S + 128u * t;
#pragma unroll
                    for (int kk = 0; kk < HEAD / 16; ++kk) {
                        // 16-column slice kk: (kk%4)*32 B into box kk/4 -> +((kk>>2)*HALF_BYTES + (kk&3)*32) >> 4
                        const uint64_t off = (uint64_t)(((kk >> 2) * HALF_BYTES + (kk & 3) * 32u) >> 4);
                        umma_ss(d, a0 + off, b0 + off, idesc_qk, kk > 0 ? 1u : 0u);
                    }
                    umma_commit(&sm.s_full[t]);
                    if (release_k) umma_commit(&sm.k_empty[stage]);
                }
                __syncwarp();
            };
            // O_t (+)= P_t V(stage) ; optionally release the V stage / signal o_done
            auto issue_pv = [&](int t, int stage, bool accumulate_first, bool release_v, bool last) {
                if (elect_one_sync()) {
                    const uint64_t b0 = dvv[stage];
                    const uint32_t d = tmem + TMEM_O + 128u * t;
                    const uint32_t a = tmem + TMEM_S + 128u * t;
#pragma unroll
                    for (int kk = 0; kk < TILE / 16; ++kk)   // P of keys 0-63 at S+0.., of keys 64-127 at S+64..
                        umma_ts(d, a + (kk < 4 ? 8u * kk : 64u + 8u * (kk - 4)), b0 + (uint64_t)((kk * 2048u) >> 4),
                                idesc_pv, (accumulate_first || kk > 0) ? 1u : 0u);
                    if (release_v) umma_commit(&sm.v_empty[stage]);
                    if (last) umma_commit(&sm.o_done[t]);
                }
                __syncwarp();
            };

            // prologue: S_A(0), S_B(0)
            mbar_wait(&sm.k_full[0], 0, 200);
            mbar_wait(&sm.q_full[0], 0, 201);
            tcgen05_fence_after();
            issue_s(0, 0, false);
            mbar_wait(&sm.q_full[1], 0, 202);
            tcgen05_fence_after();
            issue_s(1, 0, true);   // K(0) is free once S_A(0), S_B(0) have completed

            for (int j = 0; j < num_tiles; ++j) {
                const int stage = j % STAGES;
                const uint32_t ph = (uint32_t)(j / STAGES) & 1u;
                const int nstage = (j + 1) % STAGES;
                const uint32_t nph = (uint32_t)((j + 1) / STAGES) & 1u;
                const bool more = (j + 1) < num_tiles;

                mbar_wait(&sm.v_full[stage], ph, 210);
                stamp(4, j, 0);
                // ---- tile A ----
                mbar_wait(&sm.p_ready[0], (uint32_t)j & 1u, 211);
                stamp(4, j, 1);
                tcgen05_fence_after();
                issue_pv(0, stage, j > 0, false, !more);
                stamp(4, j, 2);
                if (more) {
                    mbar_wait(&sm.k_full[nstage], nph, 212);
                    tcgen05_fence_after();
                    issue_s(0, nstage, false);
                }
                stamp(4, j, 3);
                // ---- tile B ----
                mbar_wait(&sm.p_ready[1], (uint32_t)j & 1u, 213);
                stamp(4, j, 4);
                tcgen05_fence_after();
                issue_pv(1, stage, j > 0, true, !more);
                if (more) issue_s(1, nstage, true);
                stamp(4, j, 5);
            }
        }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        if (num_tiles > 0) {
            // ================================ softmax + epilogue ==========================
            const int sw = warp - 4;                       // 0..15
            const int t = sw >> 3;                         // which Q tile
            const int half = (sw >> 2) & 1;                // which 64-key half of the row
            const int quad = warp & 3;                     // TMEM lane quadrant of this warp
            const int row_in_tile = quad * 32 + lane;
            const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
            const uint32_t s_addr = tmem + lane_base + TMEM_S + 128u * t + 64u * half;   // my 64 S columns;
            const uint32_t p_addr = s_addr;   // my bf16 P (32 packed columns) overwrites the start of my own S columns
            const uint32_t o_addr = tmem + lane_base + TMEM_O + 128u * t + 64u * half;   // my O columns
            const float scale = prm.scale_log2;
            const uint64_t scale2 = pack_f32x2(scale, scale);
            const int bar_id = 1 + 4 * t + quad;           // pair barrier: the two warps that share these 32 rows

            float m_ref = -CUDART_INF_F;   // raw-score reference max used by every exponent so far
            float lsum = 0.f;

            // exp2(s*scale - ref*scale) of 16 keys: packed FFMA2, 2 x ex2 per pair, packed FADD2 row sum, F2FP pack
            auto exp_chunk = [&](const uint32_t* sv, uint64_t neg_ref2, uint64_t& acc0, uint64_t& acc1, uint32_t* pr) {
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const uint64_t x2 = pack_f32x2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
                    const uint64_t t2 = fma_f32x2(x2, scale2, neg_ref2);
                    float p0, p1;
                    // pairs 1 and 5 (POLY=4) or 1,3,5,7 (POLY=8) of the eight pairs use the polynomial
                    const bool poly = (POLY == 4 && (c == 2 || c == 10)) || (POLY == 8 && (c & 2));
                    if (poly) {
                        exp2_poly_x2(t2, p0, p1);
                    } else {
                        float t0, t1;
                        unpack_f32x2(t2, t0, t1);
                        p0 = fast_exp2(t0);
                        p1 = fast_exp2(t1);
                    }
                    const uint64_t p2 = pack_f32x2(p0, p1);
                    if (c & 4) acc1 = add_f32x2(acc1, p2);
                    else acc0 = add_f32x2(acc0, p2);
                    pr[c / 2] = pack_bf16x2(p0, p1);
                }
            };

            // One key tile.  The exponentials are SPECULATIVE on the reference max of the previous tiles
            // so they can start as soon as the first 16 columns arrive from TMEM (the remaining
            // tcgen05.ld's stream behind the MUFU work); the tile's own max is computed alongside and
            // exchanged with the other half of the row at the end.  Only if the max grew by more than
            // 2^kLazyThreshold (or on the very first tile, reference = -inf) is the tile redone with the
            // new reference and O rescaled -- the lazy rescale, decided per warp with a vote.
            // MASKED = the last tile of the shard when n is not a multiple of 128.
            auto tile_step = [&](int j, auto masked_tag, auto first_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                constexpr bool AGREE = SAFE || decltype(first_tag)::value;   // exchange the row max on this tile?
                mbar_wait(&sm.s_full[t], (uint32_t)j & 1u, 300 + t);
                if (quad == 0) stamp(sw >> 2, j, 0);
                tcgen05_fence_after();

                uint32_t sr[64];
                const uint64_t neg_ref2 = pack_f32x2(-m_ref * scale, -m_ref * scale);
                uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
                float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
                int keys_left = 64;
                if constexpr (MASKED) keys_left = prm.n - (tile_begin + j) * TILE - 64 * half;   // valid keys in my half

                if constexpr (CHUNKED) {
                SDPA_TMEM_LD16(s_addr, sr);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    tmem_wait_ld();
                    if (ch == 0 && quad == 0) stamp(sw >> 2, j, 1);
                    if (ch < 3) SDPA_TMEM_LD16(s_addr + 16 * (ch + 1), (sr + 16 * (ch + 1)));   // in flight during this chunk
                    uint32_t* sv = sr + 16 * ch;
                    if constexpr (MASKED) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (16 * ch + c >= keys_left) sv[c] = 0xff800000u;  // -inf
                    }
#pragma unroll
                    for (int c = 0; c < 16; c += 8) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[c + 0]), __uint_as_float(sv[c + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[c + 4]), __uint_as_float(sv[c + 5])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[c + 6]), __uint_as_float(sv[c + 7])));
                    }
                    uint32_t pr[8];
                    exp_chunk(sv, neg_ref2, acc0, acc1, pr);
                    SDPA_TMEM_ST8(p_addr + 8 * ch, pr);   // columns [8ch, 8ch+8) of my region: S values already in registers
                }
                } else {
                    // whole half row at once: one wait, no scheduling barriers between the 16-key groups
                    SDPA_TMEM_LD32(s_addr, sr);
                    SDPA_TMEM_LD32(s_addr + 32, (sr + 32));
                    tmem_wait_ld();
                    if (quad == 0) stamp(sw >> 2, j, 1);
                    if constexpr (MASKED) {
#pragma unroll
                        for (int c = 0; c < 64; ++c)
                            if (c >= keys_left) sr[c] = 0xff800000u;  // -inf
                    }
#pragma unroll
                    for (int c = 0; c < 64; c += 8) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7])));
                    }
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t pr[8];
                        exp_chunk(sr + 16 * ch, neg_ref2, acc0, acc1, pr);
                        SDPA_TMEM_ST8(p_addr + 8 * ch, pr);
                    }
                }
                if (quad == 0) stamp(sw >> 2, j, 2);

                const float my_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if constexpr (AGREE) {
                // agree on the tile max with the thread that owns the other half of this row
                sm.xchg[t][j & 1][half][row_in_tile] = my_max;
                named_barrier_sync(bar_id, 64);
                const float tile_max = fmaxf(my_max, sm.xchg[t][j & 1][half ^ 1][row_in_tile]);
                if (quad == 0) stamp(sw >> 2, j, 3);

                const bool need = (tile_max - m_ref) * scale > kLazyThreshold;   // first tile: -inf reference -> true
                if (__any_sync(0xffffffffu, need)) {
                    // Rare path.  O_t is stable here: PV_t(j-1) completed before s_full(j) fired, and
                    // PV_t(j) is not issued until all 256 threads of the tile signal p_ready(j).
                    const float new_ref = need ? tile_max : m_ref;
                    const float corr = need ? fast_exp2((m_ref - new_ref) * scale) : 1.f;   // -inf reference -> 0
                    m_ref = new_ref;
                    lsum *= corr;
#pragma unroll 1
                    for (int c0 = 0; c0 < 64; c0 += 16) {   // small chunks: keep this path out of the register budget
                        uint32_t orr[16];
                        SDPA_TMEM_LD16(o_addr + c0, orr);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 16; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * corr);
                        SDPA_TMEM_ST16(o_addr + c0, orr);
                    }
                    const uint64_t neg_new2 = pack_f32x2(-new_ref * scale, -new_ref * scale);
                    acc0 = pack_f32x2(0.f, 0.f);
                    acc1 = acc0;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t pr[8];
                        exp_chunk(sr + 16 * ch, neg_new2, acc0, acc1, pr);
                        SDPA_TMEM_ST8(p_addr + 8 * ch, pr);
                    }
                }
                } else {
                    // Fast mode after the first tile: the reference stays where the first tile put it
                    // (any reference gives the same quotient o/lsum; fp32 has the range for 2^64 growth).
                    // If a score outgrows it by more than 2^kGuardThreshold the launch is handed to the
                    // SAFE kernel, which recomputes everything with per-tile agreement.
                    if (quad == 0) stamp(sw >> 2, j, 3);
                    if (__any_sync(0xffffffffu, (my_max - m_ref) * scale > kGuardThreshold)) {
                        if (lane == 0) atomicExch(prm.guard, prm.epoch);
                    }
                }
                float a0, a1, a2, a3;
                unpack_f32x2(acc0, a0, a1);
                unpack_f32x2(acc1, a2, a3);
                lsum += (a0 + a1) + (a2 + a3);

                if (quad == 0) stamp(sw >> 2, j, 4);
                tmem_wait_st();
                tcgen05_fence_before();
                mbar_arrive(&sm.p_ready[t]);
                if (quad == 0) stamp(sw >> 2, j, 5);
            };
            const bool ragged = (prm.n % TILE) != 0 && (tile_begin + num_tiles) == prm.tiles_total;
            const int full_tiles = ragged ? num_tiles - 1 : num_tiles;
            if (full_tiles > 0) tile_step(0, std::false_type{}, std::true_type{});
            for (int j = 1; j < full_tiles; ++j) tile_step(j, std::false_type{}, std::false_type{});
            if (ragged) {
                if (num_tiles == 1) tile_step(0, std::true_type{}, std::true_type{});
                else tile_step(num_tiles - 1, std::true_type{}, std::false_type{});
            }

            // ---------------- epilogue: O_t, reference max, row sum ----------------
            sm.xchg[t][num_tiles & 1][half][row_in_tile] = lsum;
            named_barrier_sync(bar_id, 64);
            lsum += sm.xchg[t][num_tiles & 1][half ^ 1][row_in_tile];

            mbar_wait(&sm.o_done[t], 0, 320 + t);
            tcgen05_fence_after();
            const int grow = row_block * BLOCK_ROWS + t * TILE + row_in_tile;
            const bool valid = grow < prm.rows;
            const float inv = (lsum == 0.f) ? 0.f : 1.f / lsum;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t orr[32];
                SDPA_TMEM_LD32(o_addr + c0, orr);
                tmem_wait_ld();
                if (valid) {
                    const int col = 64 * half + c0;
                    if (prm.out64 != nullptr) {
                        double2* dst = reinterpret_cast<double2*>(prm.out64 + (size_t)grow * HEAD + col);
#pragma unroll
                        for (int c = 0; c < 32; c += 2)
                            dst[c / 2] = make_double2((double)(__uint_as_float(orr[c]) * inv),
                                                      (double)(__uint_as_float(orr[c + 1]) * inv));
                    } else {
                        float4* dst = reinterpret_cast<float4*>(prm.part_o + ((size_t)split * prm.rows_capacity + grow) * HEAD + col);
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            dst[c / 4] = make_float4(__uint_as_float(orr[c]), __uint_as_float(orr[c + 1]),
                                                     __uint_as_float(orr[c + 2]), __uint_as_float(orr[c + 3]));
                    }
                }
            }
            if (valid && half == 0 && prm.out64 == nullptr) {
                prm.part_tmax[(size_t)split * prm.rows_capacity + grow] = m_ref * scale;
                prm.part_lsum[(size_t)split * prm.rows_capacity + grow] = lsum;
            }
        } else {
        // empty key range (n == 0 or more splits than tiles): the neutral state (0, -inf, 0), mpi.c:172,188
        const int sw = warp - 4;
        const int t = sw >> 3, half = (sw >> 2) & 1;
        const int grow = row_block * BLOCK_ROWS + t * TILE + (warp & 3) * 32 + lane;
        if (grow < prm.rows) {
            if (prm.out64 != nullptr) {
                for (int c = 0; c < 64; ++c) prm.out64[(size_t)grow * HEAD + 64 * half + c] = 0.0;
            } else {
                float* dst = prm.part_o + ((size_t)split * prm.rows_capacity + grow) * HEAD + 64 * half;
                for (int c = 0; c < 64; ++c) dst[c] = 0.f;
                if (half == 0) {
                    prm.part_tmax[(size_t)split * prm.rows_capacity + grow] = -CUDART_INF_F;
                    prm.part_lsum[(size_t)split * prm.rows_capacity + grow] = 0.f;
                }
            }
        }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// (v6 -- the cluster-of-two, multicast stepping stone between v5 and v7 -- was removed in round 2: measured slower than both,
//  profiles/README.md keeps its numbers.)
constexpr uint32_t V6_S = 0, V6_O = 256, V6_P = 384;   // TMEM map of the chain-free kernels: S0 S1 | O | P0 P1

// =====================================================================================
// v7: v6 with the 2-CTA tensor-core instruction (tcgen05.mma.cta_group::2).  The two CTAs of a cluster
// form ONE M=256 MMA: each holds its own 128-row Q tile (A) and its own accumulators/P in its TMEM, and
// each keeps only HALF of every K tile (64 keys) and V tile (64 value columns) in shared memory -- the
// B operand is split across the pair -- so per SM the shared-memory traffic per key tile drops from
// 160 KB (v6) to 96 KB and no multicast is needed.  Only the leader CTA (cluster rank 0) issues MMAs;
// its barriers collect both CTAs' TMA bytes (2-SM TMA form: the mbarrier address is mapped to the
// leader) and both CTAs' softmax arrivals (remote mbarrier.arrive through mapa); tcgen05.commit
// multicasts completions to both CTAs.
// =====================================================================================
constexpr int V7_KSTAGES = 4, V7_VSTAGES = 3;
constexpr uint32_t HALF_TILE_BYTES = TILE_BYTES / 2;   // 64 keys x 128 (K half) or 128 keys x 64 columns (V half)

struct __align__(1024) SharedV7 {
    uint8_t q[TILE_BYTES];
    uint8_t k[V7_KSTAGES][HALF_TILE_BYTES];
    uint8_t v[V7_VSTAGES][HALF_TILE_BYTES];
    uint64_t q_full;
    uint64_t k_full[V7_KSTAGES], k_empty[V7_KSTAGES];
    uint64_t v_full[V7_VSTAGES], v_empty[V7_VSTAGES];
    uint64_t s_full[2], p_ready[2], s_free[2], pv_done[2], o_done;
    uint32_t tmem_base;
    float xchg[2][4][TILE];   // [first-tile max | final sum][group*NPARTS + part][row]
    float mref[TILE];         // the agreed reference, handed from softmax group 0 to the others
};


// GROUPS = softmax groups ping-ponged over alternating key tiles (group g owns tiles j = g mod GROUPS and the
// S/P buffers g): with the reference fixed by the first tile the tiles are independent, so while one
// group sits in its fixed latencies (barrier wake-up, TMEM store drain, remote arrive) the other keeps the
// MUFU / FMA pipes busy.  The groups' row sums are added in the epilogue.
template <bool TRACE, int POLY, int NPARTS, int GROUPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 128 * NPARTS * GROUPS, 1)
attn_umma_kernel_v7(const __grid_constant__ CUtensorMap map_khalf, const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const KernelParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    SharedV7& sm = *reinterpret_cast<SharedV7*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int row_block = blockIdx.x;             // 128 rows; blockIdx.x = 2*cluster + rank
    const int split = blockIdx.y;
    const uint32_t rank = cluster_cta_rank();      // 0 = leader (issues the MMAs); rank r keeps keys [64r,64r+64) of K and columns [64r,64r+64) of V
    const bool leader = rank == 0;

    const int tq = prm.tiles_total / prm.splits, tr = prm.tiles_total % prm.splits;
    const int tile_begin = split * tq + min(split, tr);
    const int num_tiles = tq + (split < tr ? 1 : 0);   // identical in both CTAs of the cluster

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_khalf);
        prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_k);
        prefetch_tensormap(&map_v);
        mbar_init(&sm.q_full, 2);          // leader's copy is the one used: one arrival per CTA's producer + all bytes
        mbar_init(&sm.o_done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.s_full[i], 1);
            mbar_init(&sm.p_ready[i], 2 * 4 * NPARTS);   // leader's copy: one arrival per softmax warp of BOTH CTAs
            mbar_init(&sm.s_free[i], 2 * 4 * NPARTS);    // leader's copy: S buffer i has been read into registers
            mbar_init(&sm.pv_done[i], 1);                // both CTAs: P buffer i may be overwritten
        }
        for (int i = 0; i < V7_KSTAGES; ++i) {
            mbar_init(&sm.k_full[i], 2);
            mbar_init(&sm.k_empty[i], 1);  // the leader's commit, multicast to both CTAs
        }
        for (int i = 0; i < V7_VSTAGES; ++i) {
            mbar_init(&sm.v_full[i], 2);
            mbar_init(&sm.v_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2cta(&sm.tmem_base, 512);   // the same warp in both CTAs
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to them
    tcgen05_fence_after();
    const uint32_t tmem = sm.tmem_base;
    auto stamp = [&](int role, int j, int ev) {
        if constexpr (TRACE) {
            if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && j < TRACE_ITERS)
                prm.trace[(role * TRACE_ITERS + j) * TRACE_EVENTS + ev] = clock64();
        }
    };

    if (warp < 4) {
        if constexpr (NPARTS * GROUPS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (num_tiles > 0) {
            if (warp == 0) {
                // ================================ TMA producer ================================
                // every load lands in this CTA's shared memory and reports its bytes to the LEADER's barrier;
                // the leader arms the barrier with the bytes of both CTAs, the peer adds its plain arrival.
                const int qrow = row_block * TILE;
                const uint32_t leader_qfull = map_to_cta(&sm.q_full, 0);
                for (int j = 0; j < num_tiles; ++j) {
                    const int ks = j % V7_KSTAGES, vs = j % V7_VSTAGES;
                    const uint32_t kph = (uint32_t)(j / V7_KSTAGES) & 1u, vph = (uint32_t)(j / V7_VSTAGES) & 1u;
                    const int key0 = (tile_begin + j) * TILE;
                    if (j == 0 && elect_one_sync()) {
                        if (leader) mbar_arrive_expect_tx(&sm.q_full, 2 * TILE_BYTES);
                        else mbar_arrive_cluster(leader_qfull);
                        tma_load_2d_2sm(sm.q, &map_q, &sm.q_full, 0, qrow);
                        tma_load_2d_2sm(sm.q + HALF_BYTES, &map_q, &sm.q_full, 64, qrow);
                    }
                    mbar_wait(&sm.k_empty[ks], kph ^ 1u, 100 + ks);
                    stamp(5, j, 0);
                    if (elect_one_sync()) {
                        if (leader) mbar_arrive_expect_tx(&sm.k_full[ks], TILE_BYTES);
                        else mbar_arrive_cluster(map_to_cta(&sm.k_full[ks], 0));
                        // my 64 keys of the tile: two boxes of 64 columns x 64 rows (8 KiB each)
                        tma_load_2d_2sm(sm.k[ks], &map_khalf, &sm.k_full[ks], 0, key0 + 64 * (int)rank);
                        tma_load_2d_2sm(sm.k[ks] + HALF_TILE_BYTES / 2, &map_khalf, &sm.k_full[ks], 64, key0 + 64 * (int)rank);
                    }
                    mbar_wait(&sm.v_empty[vs], vph ^ 1u, 110 + vs);
                    stamp(5, j, 1);
                    if (elect_one_sync()) {
                        if (leader) mbar_arrive_expect_tx(&sm.v_full[vs], TILE_BYTES);
                        else mbar_arrive_cluster(map_to_cta(&sm.v_full[vs], 0));
                        // my 64 value columns of the tile: one box of 64 columns x 128 keys (16 KiB)
                        tma_load_2d_2sm(sm.v[vs], &map_v, &sm.v_full[vs], 64 * (int)rank, key0);
                    }
                    __syncwarp();
                }
            } else if (warp == 1) {
                // ================================ MMA issuer (leader CTA only) ==================
                if (leader) {
                constexpr uint32_t idesc_qk = make_idesc(2 * TILE, TILE, 0);   // M = 256 over the CTA pair
                constexpr uint32_t idesc_pv = make_idesc(2 * TILE, HEAD, 1);
                const uint64_t dq = desc_kmajor(smem_u32(sm.q), 0);
                uint64_t dkk[V7_KSTAGES], dvv[V7_VSTAGES];
#pragma unroll
                for (int i = 0; i < V7_KSTAGES; ++i) dkk[i] = make_desc(smem_u32(sm.k[i]), 16u, 1024u);
#pragma unroll
                for (int i = 0; i < V7_VSTAGES; ++i) dvv[i] = make_desc(smem_u32(sm.v[i]), HALF_TILE_BYTES, 1024u);
                const uint16_t both = 0x3;

                auto issue_s = [&](int j) {
                    const int sb = j & 1, ks = j % V7_KSTAGES;
                    mbar_wait(&sm.k_full[ks], (uint32_t)(j / V7_KSTAGES) & 1u, 200 + ks);   // (already tested by the scheduler below)
                    tcgen05_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t b0 = dkk[ks];
                        const uint32_t d = tmem + V6_S + 128u * sb;
#pragma unroll
                        for (int kk = 0; kk < HEAD / 16; ++kk) {
                            const uint64_t offa = (uint64_t)(((kk >> 2) * HALF_BYTES + (kk & 3) * 32u) >> 4);            // Q: boxes of 128 rows
                            const uint64_t offb = (uint64_t)(((kk >> 2) * (HALF_TILE_BYTES / 2) + (kk & 3) * 32u) >> 4);   // K half: boxes of 64 rows
                            umma_ss_2cta(d, dq + offa, b0 + offb, idesc_qk, kk > 0 ? 1u : 0u);
                        }
                        umma_commit_2cta(&sm.s_full[sb], both);
                        umma_commit_2cta(&sm.k_empty[ks], both);
                    }
                    __syncwarp();
                };
                auto issue_pv = [&](int j, bool last) {
                    const int pb = j & 1, vs = j % V7_VSTAGES;
                    mbar_wait(&sm.v_full[vs], (uint32_t)(j / V7_VSTAGES) & 1u, 210 + vs);
                    mbar_wait(&sm.p_ready[pb], (uint32_t)(j >> 1) & 1u, 212 + pb);
                    stamp(4, j, 1);
                    tcgen05_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t b0 = dvv[vs];
                        const uint32_t d = tmem + V6_O;
                        const uint32_t a = tmem + V6_P + 64u * pb;
#pragma unroll
                        for (int kk = 0; kk < TILE / 16; ++kk)
                            umma_ts_2cta(d, a + 8u * kk, b0 + (uint64_t)((kk * 2048u) >> 4), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                        umma_commit_2cta(&sm.pv_done[pb], both);
                        umma_commit_2cta(&sm.v_empty[vs], both);
                        if (last) umma_commit_2cta(&sm.o_done, both);
                    }
                    __syncwarp();
                };

                mbar_wait(&sm.q_full, 0, 201);
                issue_s(0);
                if (num_tiles > 1) issue_s(1);
                // Fixed order per tile: S(j+2) as soon as S(j) sits in the softmax registers (s_free), then PV(j) once
                // P(j) is ready.  (A scheduler issuing "whichever stream is ready" was measured slower -- polling two
                // barriers costs more than the occasional head-of-line wait: profiles/r01/sweep_v7_3_testwait_scheduler.txt.)
                for (int j = 0; j < num_tiles; ++j) {
                    stamp(4, j, 0);
                    if (j + 2 < num_tiles) {
                        mbar_wait(&sm.s_free[j & 1], (uint32_t)(j >> 1) & 1u, 214 + (j & 1));
                        issue_s(j + 2);
                    }
                    stamp(4, j, 2);
                    issue_pv(j, j + 1 == num_tiles);
                    stamp(4, j, 3);
                }
                }
            }
        }
    } else {
        if constexpr (NPARTS * GROUPS == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        constexpr int COLS = TILE / NPARTS;            // keys (S columns) per thread: 64 or 32
        constexpr int OCOLS = HEAD / (NPARTS * GROUPS); // output columns per thread in the epilogue
        const int sw = warp - 4;
        const int group = sw / (4 * NPARTS);           // which softmax group (owns tiles j = group mod GROUPS)
        const int half = (sw % (4 * NPARTS)) >> 2;     // which COLS-wide part of the row (0..NPARTS-1)
        const int gp = group * NPARTS + half;          // 0 .. NPARTS*GROUPS-1
        const int quad = warp & 3;                     // TMEM lane quadrant of this warp
        const int row_in_tile = quad * 32 + lane;
        const int grow = row_block * TILE + row_in_tile;
        if (num_tiles > 0) {
            // ================================ softmax + epilogue ==========================
            const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
            const uint32_t o_addr = tmem + lane_base + V6_O + (uint32_t)OCOLS * gp;
            const float scale = prm.scale_log2;
            const uint64_t scale2 = pack_f32x2(scale, scale);
            const int bar_id = 1 + quad;               // the warps of ONE group that share these 32 rows
            const int bar_all = 5 + quad;              // the warps of ALL groups that share these 32 rows
            const uint32_t leader_pready[2] = {map_to_cta(&sm.p_ready[0], 0), map_to_cta(&sm.p_ready[1], 0)};
            const uint32_t leader_sfree[2] = {map_to_cta(&sm.s_free[0], 0), map_to_cta(&sm.s_free[1], 0)};

            float m_ref = -CUDART_INF_F;
            float lsum = 0.f;

            auto exp_chunk = [&](const uint32_t* sv, uint64_t neg_ref2, uint64_t& acc0, uint64_t& acc1, uint32_t* pr) {
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const uint64_t x2 = pack_f32x2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
                    const uint64_t t2 = fma_f32x2(x2, scale2, neg_ref2);
                    float p0, p1;
                    const bool poly = (POLY == 4 && (c == 2 || c == 10)) || (POLY == 8 && (c & 2));
                    if (poly) {
                        exp2_poly_x2(t2, p0, p1);
                    } else {
                        float t0, t1;
                        unpack_f32x2(t2, t0, t1);
                        p0 = fast_exp2(t0);
                        p1 = fast_exp2(t1);
                    }
                    const uint64_t p2 = pack_f32x2(p0, p1);
                    if (c & 4) acc1 = add_f32x2(acc1, p2);
                    else acc0 = add_f32x2(acc0, p2);
                    pr[c / 2] = pack_bf16x2(p0, p1);
                }
            };

            auto tile_step = [&](int j, auto masked_tag, auto first_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                constexpr bool FIRST = decltype(first_tag)::value;
                const int sb = j & 1;
                const uint32_t s_addr = tmem + lane_base + V6_S + 128u * sb + (uint32_t)COLS * half;
                const uint32_t p_addr = tmem + lane_base + V6_P + 64u * sb + (uint32_t)(COLS / 2) * half;
                mbar_wait(&sm.s_full[sb], (uint32_t)(j >> 1) & 1u, 300 + sb);
                if (quad == 0 && half == 0) stamp(group, j, 0);
                tcgen05_fence_after();

                uint32_t sr[COLS];
                SDPA_TMEM_LD32(s_addr, sr);
                if constexpr (COLS == 64) SDPA_TMEM_LD32(s_addr + 32, (sr + 32));
                tmem_wait_ld();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_sfree[sb]);   // S buffer sb may be overwritten by S(j+2)
                if (quad == 0 && half == 0) stamp(group, j, 1);
                if constexpr (MASKED) {
                    const int keys_left = prm.n - (tile_begin + j) * TILE - COLS * half;
#pragma unroll
                    for (int c = 0; c < COLS; ++c)
                        if (c >= keys_left) sr[c] = 0xff800000u;  // -inf
                }
                float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
                for (int c = 0; c < COLS; c += 8) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])));
                    mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5])));
                    mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7])));
                }
                const float my_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if constexpr (FIRST) {
                    // the reference of the whole key range: the first tile's row max, agreed by the two halves
                    sm.xchg[0][half][row_in_tile] = my_max;
                    named_barrier_sync(bar_id, 32 * NPARTS);
                    m_ref = my_max;
#pragma unroll
                    for (int p = 0; p < NPARTS; ++p) m_ref = fmaxf(m_ref, sm.xchg[0][p][row_in_tile]);
                    if constexpr (GROUPS > 1) {
                        if (half == 0) sm.mref[row_in_tile] = m_ref;
                        named_barrier_sync(bar_all, 32 * NPARTS * GROUPS);   // the other groups pick the reference up
                    }
                }
                // No wait is needed before overwriting P buffer sb: PV(j-2), its last reader, was issued before
                // S(j), and the commit behind s_full(j) covers every MMA issued before it.
                if (quad == 0 && half == 0) stamp(group, j, 2);

                const uint64_t neg_ref2 = pack_f32x2(-m_ref * scale, -m_ref * scale);
                uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
                uint32_t pr[COLS / 2];   // the packed bf16 P of the whole part; stored at the end, once PV(j-2) has left the buffer
#pragma unroll
                for (int ch = 0; ch < COLS / 16; ++ch) exp_chunk(sr + 16 * ch, neg_ref2, acc0, acc1, pr + 8 * ch);
                if (j >= 2) mbar_wait(&sm.pv_done[sb], (uint32_t)((j >> 1) - 1) & 1u, 310 + sb);
                tcgen05_fence_after();
#pragma unroll
                for (int ch = 0; ch < COLS / 16; ++ch) SDPA_TMEM_ST8(p_addr + 8 * ch, (pr + 8 * ch));
                float a0, a1, a2, a3;
                unpack_f32x2(acc0, a0, a1);
                unpack_f32x2(acc1, a2, a3);
                lsum += (a0 + a1) + (a2 + a3);
                if constexpr (!FIRST) {
                    // overflow guard, off the critical path (the max chain overlaps the exponentials)
                    if (__any_sync(0xffffffffu, (my_max - m_ref) * scale > kGuardThreshold)) {
                        if (lane == 0) atomicExch(prm.guard, prm.epoch);   // hand the launch to the SAFE kernel
                    }
                }
                if (quad == 0 && half == 0) stamp(group, j, 4);
                tmem_wait_st();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_pready[sb]);   // one arrival per warp, on the leader's barrier
                if (quad == 0 && half == 0) stamp(group, j, 5);
            };

            const bool ragged = (prm.n % TILE) != 0 && (tile_begin + num_tiles) == prm.tiles_total;
            if (group == 0) {
                if (ragged && num_tiles == 1) tile_step(0, std::true_type{}, std::true_type{});
                else tile_step(0, std::false_type{}, std::true_type{});
            } else {
                named_barrier_sync(bar_all, 32 * NPARTS * GROUPS);   // wait for group 0's reference
                m_ref = sm.mref[row_in_tile];
            }
            for (int j = (group == 0 ? GROUPS : group); j < num_tiles; j += GROUPS) {
                if (ragged && j == num_tiles - 1) tile_step(j, std::true_type{}, std::false_type{});
                else tile_step(j, std::false_type{}, std::false_type{});
            }

            // ---------------- epilogue ----------------
            sm.xchg[1][gp][row_in_tile] = lsum;
            named_barrier_sync(bar_all, 32 * NPARTS * GROUPS);
            lsum = 0.f;
#pragma unroll
            for (int p = 0; p < NPARTS * GROUPS; ++p) lsum += sm.xchg[1][p][row_in_tile];
            mbar_wait(&sm.o_done, 0, 320);
            tcgen05_fence_after();
            const bool valid = grow < prm.rows;
            const float inv = (lsum == 0.f) ? 0.f : 1.f / lsum;
#pragma unroll
            for (int c0 = 0; c0 < OCOLS; c0 += 32) {
                uint32_t orr[32];
                SDPA_TMEM_LD32(o_addr + c0, orr);
                tmem_wait_ld();
                if (valid) {
                    const int col = OCOLS * gp + c0;
                    if (prm.out64 != nullptr) {
                        double2* dst = reinterpret_cast<double2*>(prm.out64 + (size_t)grow * HEAD + col);
#pragma unroll
                        for (int c = 0; c < 32; c += 2)
                            dst[c / 2] = make_double2((double)(__uint_as_float(orr[c]) * inv),
                                                      (double)(__uint_as_float(orr[c + 1]) * inv));
                    } else {
                        float4* dst = reinterpret_cast<float4*>(prm.part_o + ((size_t)split * prm.rows_capacity + grow) * HEAD + col);
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            dst[c / 4] = make_float4(__uint_as_float(orr[c]), __uint_as_float(orr[c + 1]),
                                                     __uint_as_float(orr[c + 2]), __uint_as_float(orr[c + 3]));
                    }
                }
            }
            if (valid && gp == 0 && prm.out64 == nullptr) {
                prm.part_tmax[(size_t)split * prm.rows_capacity + grow] = m_ref * scale;
                prm.part_lsum[(size_t)split * prm.rows_capacity + grow] = lsum;
            }
        } else if (grow < prm.rows) {
            // empty key range: the neutral state (0, -inf, 0), mpi.c:172,188
            if (prm.out64 != nullptr) {
                for (int c = 0; c < OCOLS; ++c) prm.out64[(size_t)grow * HEAD + OCOLS * gp + c] = 0.0;
            } else {
                float* dst = prm.part_o + ((size_t)split * prm.rows_capacity + grow) * HEAD + OCOLS * gp;
                for (int c = 0; c < OCOLS; ++c) dst[c] = 0.f;
                if (gp == 0) {
                    prm.part_tmax[(size_t)split * prm.rows_capacity + grow] = -CUDART_INF_F;
                    prm.part_lsum[(size_t)split * prm.rows_capacity + grow] = 0.f;
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();   // neither CTA leaves while the other may still multicast into it or arrive on its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2cta(tmem, 512);
    }
}

// =====================================================================================
// v8 (SDPA_UMMA_V8=1, EXPERIMENTAL: written after the round's GPU budget was spent, not yet run on hardware).
// v7's pipeline made PERSISTENT: one cluster per SM pair walks a contiguous range of the linear work space
// (row block of 256 rows, key tile) -- "stream-K" over the key axis.  With 74 clusters every cluster gets
// RB*T/74 key tiles (+-1): no wave quantisation (v7 on c3: 288 equal units on 74 slots = 97.3 %), the prologue
// (barrier init, TMEM alloc, cluster sync, first loads) runs once per SM instead of once per 57-tile unit, and a
// row block is cut into ~T*74/(RB*T) + 1 pieces instead of 9 splits, so the split merge reads a third of the partials.
// A cluster's range crosses row-block boundaries: it is processed as SEGMENTS (row block, first tile, tile count).
// What v7 keeps per launch is carried across segments here:
//   * K/V rings and the S/P double buffers run on ONE tile counter g over all segments of the cluster, so the
//     S MMAs of the next segment's first tiles are issued while the current segment drains (no pipeline refill);
//   * Q has two shared-memory slots (segment parity); q_free (a commit behind the segment's last S MMA) lets the
//     producer overwrite a slot two segments later;
//   * the single O accumulator is handed back by o_free: every softmax warp of both CTAs arrives after it has read
//     its share of O in the segment's epilogue; the first PV of the next segment waits for it;
//   * the softmax reference is per segment: the group owning the segment's first tile (g & 1) fixes and publishes it.
// The partial state of segment (rb, piece) goes to partial slot `piece` = cluster - first cluster touching rb; the
// merge reads pieces(rb) states per row (merge_pieces_kernel), or all max_pieces after the SAFE twin ran.
// =====================================================================================
struct __align__(1024) SharedV8 {
    uint8_t q[2][TILE_BYTES];
    uint8_t k[V7_KSTAGES][HALF_TILE_BYTES];
    uint8_t v[V7_VSTAGES][HALF_TILE_BYTES];
    uint64_t q_full[2], q_free[2];
    uint64_t k_full[V7_KSTAGES], k_empty[V7_KSTAGES];
    uint64_t v_full[V7_VSTAGES], v_empty[V7_VSTAGES];
    uint64_t s_full[2], p_ready[2], s_free[2], pv_done[2], o_done, o_free;
    uint32_t tmem_base;
    float xchg[2][4][TILE];   // [first-tile max | final sum][group*2 + part][row]
    float mref[TILE];
};

// Walks the segments of the unit range [begin, end) of one cluster.
struct SegCursor {
    int u, uend, T;
    int seg, rb, t0, nt;
    __device__ __forceinline__ void init(int begin, int end, int tiles_per_row_block)
    {
        u = begin;
        uend = end;
        T = tiles_per_row_block;
        seg = -1;
        nt = 0;
        rb = t0 = 0;
    }
    __device__ __forceinline__ bool next()
    {
        u += nt;
        if (u >= uend) {
            nt = 0;
            return false;
        }
        rb = u / T;
        t0 = u - rb * T;
        nt = min(T - t0, uend - u);
        ++seg;
        return true;
    }
};

template <int POLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
attn_umma_kernel_v8(const __grid_constant__ CUtensorMap map_khalf, const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const KernelParams prm)
{
    constexpr int NPARTS = 2, GROUPS = 2;
    extern __shared__ uint8_t smem_raw[];
    SharedV8& sm = *reinterpret_cast<SharedV8*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int cluster = blockIdx.x >> 1;
    const uint32_t rank = cluster_cta_rank();      // 0 = leader (issues the MMAs); rank r keeps keys [64r,64r+64) of K and columns [64r,64r+64) of V
    const bool leader = rank == 0;
    const WorkMap wm = prm.wm;
    const int T = wm.T;
    const int unit_begin = (int)wm_begin(wm, cluster), unit_end = (int)wm_begin(wm, cluster + 1);
    const int total_tiles = unit_end - unit_begin;   // identical in both CTAs of the cluster

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_khalf);
        prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_k);
        prefetch_tensormap(&map_v);
        mbar_init(&sm.o_done, 1);
        mbar_init(&sm.o_free, 2 * 4 * NPARTS * GROUPS);   // leader's copy: every softmax warp of BOTH CTAs has read its share of O
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.q_full[i], 2);       // leader's copy: one arrival per CTA's producer + all bytes
            mbar_init(&sm.q_free[i], 1);       // both CTAs: the commit behind the last S MMA that read the slot
            mbar_init(&sm.s_full[i], 1);
            mbar_init(&sm.p_ready[i], 2 * 4 * NPARTS);
            mbar_init(&sm.s_free[i], 2 * 4 * NPARTS);
            mbar_init(&sm.pv_done[i], 1);
        }
        for (int i = 0; i < V7_KSTAGES; ++i) {
            mbar_init(&sm.k_full[i], 2);
            mbar_init(&sm.k_empty[i], 1);
        }
        for (int i = 0; i < V7_VSTAGES; ++i) {
            mbar_init(&sm.v_full[i], 2);
            mbar_init(&sm.v_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2cta(&sm.tmem_base, 512);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (total_tiles > 0) {
            if (warp == 0) {
                // ================================ TMA producer ================================
                SegCursor cur;
                cur.init(unit_begin, unit_end, T);
                int g = 0;
                while (cur.next()) {
                    const int qs = cur.seg & 1;
                    if (cur.seg >= 2) mbar_wait(&sm.q_free[qs], (uint32_t)((cur.seg >> 1) - 1) & 1u, 120 + qs);
                    if (elect_one_sync()) {
                        const int qrow = (2 * cur.rb + (int)rank) * TILE;
                        if (leader) mbar_arrive_expect_tx(&sm.q_full[qs], 2 * TILE_BYTES);
                        else mbar_arrive_cluster(map_to_cta(&sm.q_full[qs], 0));
                        tma_load_2d_2sm(sm.q[qs], &map_q, &sm.q_full[qs], 0, qrow);
                        tma_load_2d_2sm(sm.q[qs] + HALF_BYTES, &map_q, &sm.q_full[qs], 64, qrow);
                    }
                    __syncwarp();
                    for (int j = 0; j < cur.nt; ++j, ++g) {
                        const int ks = g % V7_KSTAGES, vs = g % V7_VSTAGES;
                        const uint32_t kph = (uint32_t)(g / V7_KSTAGES) & 1u, vph = (uint32_t)(g / V7_VSTAGES) & 1u;
                        const int key0 = (cur.t0 + j) * TILE;
                        mbar_wait(&sm.k_empty[ks], kph ^ 1u, 100 + ks);
                        if (elect_one_sync()) {
                            if (leader) mbar_arrive_expect_tx(&sm.k_full[ks], TILE_BYTES);
                            else mbar_arrive_cluster(map_to_cta(&sm.k_full[ks], 0));
                            tma_load_2d_2sm(sm.k[ks], &map_khalf, &sm.k_full[ks], 0, key0 + 64 * (int)rank);
                            tma_load_2d_2sm(sm.k[ks] + HALF_TILE_BYTES / 2, &map_khalf, &sm.k_full[ks], 64, key0 + 64 * (int)rank);
                        }
                        mbar_wait(&sm.v_empty[vs], vph ^ 1u, 110 + vs);
                        if (elect_one_sync()) {
                            if (leader) mbar_arrive_expect_tx(&sm.v_full[vs], TILE_BYTES);
                            else mbar_arrive_cluster(map_to_cta(&sm.v_full[vs], 0));
                            tma_load_2d_2sm(sm.v[vs], &map_v, &sm.v_full[vs], 64 * (int)rank, key0);
                        }
                        __syncwarp();
                    }
                }
            } else if (warp == 1) {
                // ================================ MMA issuer (leader CTA only) ==================
                if (leader) {
                constexpr uint32_t idesc_qk = make_idesc(2 * TILE, TILE, 0);
                constexpr uint32_t idesc_pv = make_idesc(2 * TILE, HEAD, 1);
                const uint64_t dq[2] = {desc_kmajor(smem_u32(sm.q[0]), 0), desc_kmajor(smem_u32(sm.q[1]), 0)};
                uint64_t dkk[V7_KSTAGES], dvv[V7_VSTAGES];
#pragma unroll
                for (int i = 0; i < V7_KSTAGES; ++i) dkk[i] = make_desc(smem_u32(sm.k[i]), 16u, 1024u);
#pragma unroll
                for (int i = 0; i < V7_VSTAGES; ++i) dvv[i] = make_desc(smem_u32(sm.v[i]), HALF_TILE_BYTES, 1024u);
                const uint16_t both = 0x3;

                SegCursor cs, cp;   // the S stream runs two tiles ahead of the PV stream, possibly in the next segment
                cs.init(unit_begin, unit_end, T);
                cp.init(unit_begin, unit_end, T);
                cs.next();
                cp.next();
                int js = 0, jp = 0;

                auto issue_s = [&](int g) {
                    const int sb = g & 1, ks = g % V7_KSTAGES, qs = cs.seg & 1;
                    if (js == 0) mbar_wait(&sm.q_full[qs], (uint32_t)(cs.seg >> 1) & 1u, 201 + qs);
                    mbar_wait(&sm.k_full[ks], (uint32_t)(g / V7_KSTAGES) & 1u, 204 + ks);
                    tcgen05_fence_after();
                    const bool last_of_segment = (js == cs.nt - 1);
                    if (elect_one_sync()) {
                        const uint64_t a0 = dq[qs], b0 = dkk[ks];
                        const uint32_t d = tmem + V6_S + 128u * sb;
#pragma unroll
                        for (int kk = 0; kk < HEAD / 16; ++kk) {
                            const uint64_t offa = (uint64_t)(((kk >> 2) * HALF_BYTES + (kk & 3) * 32u) >> 4);
                            const uint64_t offb = (uint64_t)(((kk >> 2) * (HALF_TILE_BYTES / 2) + (kk & 3) * 32u) >> 4);
                            umma_ss_2cta(d, a0 + offa, b0 + offb, idesc_qk, kk > 0 ? 1u : 0u);
                        }
                        umma_commit_2cta(&sm.s_full[sb], both);
                        umma_commit_2cta(&sm.k_empty[ks], both);
                        if (last_of_segment) umma_commit_2cta(&sm.q_free[qs], both);   // no later MMA reads this Q slot
                    }
                    __syncwarp();
                    if (++js == cs.nt) {
                        cs.next();
                        js = 0;
                    }
                };
                auto issue_pv = [&](int g) {
                    const int pb = g & 1, vs = g % V7_VSTAGES;
                    const bool first = (jp == 0), last = (jp == cp.nt - 1);
                    // the accumulator still holds the previous segment until all of its epilogue reads are done
                    if (first && cp.seg >= 1) mbar_wait(&sm.o_free, (uint32_t)(cp.seg - 1) & 1u, 220);
                    mbar_wait(&sm.v_full[vs], (uint32_t)(g / V7_VSTAGES) & 1u, 210 + vs);
                    mbar_wait(&sm.p_ready[pb], (uint32_t)(g >> 1) & 1u, 214 + pb);
                    tcgen05_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t b0 = dvv[vs];
                        const uint32_t d = tmem + V6_O;
                        const uint32_t a = tmem + V6_P + 64u * pb;
#pragma unroll
                        for (int kk = 0; kk < TILE / 16; ++kk)
                            umma_ts_2cta(d, a + 8u * kk, b0 + (uint64_t)((kk * 2048u) >> 4), idesc_pv, (!first || kk > 0) ? 1u : 0u);
                        umma_commit_2cta(&sm.pv_done[pb], both);
                        umma_commit_2cta(&sm.v_empty[vs], both);
                        if (last) umma_commit_2cta(&sm.o_done, both);
                    }
                    __syncwarp();
                    if (++jp == cp.nt) {
                        cp.next();
                        jp = 0;
                    }
                };

                issue_s(0);
                if (total_tiles > 1) issue_s(1);
                for (int g = 0; g < total_tiles; ++g) {
                    if (g + 2 < total_tiles) {
                        mbar_wait(&sm.s_free[g & 1], (uint32_t)(g >> 1) & 1u, 216 + (g & 1));
                        issue_s(g + 2);
                    }
                    issue_pv(g);
                }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        constexpr int COLS = TILE / NPARTS;             // 64 keys (S columns) per thread
        constexpr int OCOLS = HEAD / (NPARTS * GROUPS);  // 32 output columns per thread in the epilogue
        const int sw = warp - 4;
        const int group = sw / (4 * NPARTS);
        const int half = (sw % (4 * NPARTS)) >> 2;
        const int gp = group * NPARTS + half;
        const int quad = warp & 3;
        const int row_in_tile = quad * 32 + lane;
        if (total_tiles > 0) {
            const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
            const uint32_t o_addr = tmem + lane_base + V6_O + (uint32_t)OCOLS * gp;
            const float scale = prm.scale_log2;
            const uint64_t scale2 = pack_f32x2(scale, scale);
            const int bar_id = 1 + quad;
            const int bar_all = 5 + quad;
            const uint32_t leader_pready[2] = {map_to_cta(&sm.p_ready[0], 0), map_to_cta(&sm.p_ready[1], 0)};
            const uint32_t leader_sfree[2] = {map_to_cta(&sm.s_free[0], 0), map_to_cta(&sm.s_free[1], 0)};
            const uint32_t leader_ofree = map_to_cta(&sm.o_free, 0);

            float m_ref = -CUDART_INF_F;
            float lsum = 0.f;

            auto exp_chunk = [&](const uint32_t* sv, uint64_t neg_ref2, uint64_t& acc0, uint64_t& acc1, uint32_t* pr) {
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const uint64_t x2 = pack_f32x2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
                    const uint64_t t2 = fma_f32x2(x2, scale2, neg_ref2);
                    float p0, p1;
                    const bool poly = (POLY == 4 && (c == 2 || c == 10)) || (POLY == 8 && (c & 2));
                    if (poly) {
                        exp2_poly_x2(t2, p0, p1);
                    } else {
                        float t0, t1;
                        unpack_f32x2(t2, t0, t1);
                        p0 = fast_exp2(t0);
                        p1 = fast_exp2(t1);
                    }
                    const uint64_t p2 = pack_f32x2(p0, p1);
                    if (c & 4) acc1 = add_f32x2(acc1, p2);
                    else acc0 = add_f32x2(acc0, p2);
                    pr[c / 2] = pack_bf16x2(p0, p1);
                }
            };

            // g: tile counter of the cluster (buffers and barrier phases), key_tile: index of the tile in the shard
            auto tile_step = [&](int g, int key_tile, auto masked_tag, auto first_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                constexpr bool FIRST = decltype(first_tag)::value;
                const int sb = g & 1;
                const uint32_t s_addr = tmem + lane_base + V6_S + 128u * sb + (uint32_t)COLS * half;
                const uint32_t p_addr = tmem + lane_base + V6_P + 64u * sb + (uint32_t)(COLS / 2) * half;
                mbar_wait(&sm.s_full[sb], (uint32_t)(g >> 1) & 1u, 300 + sb);
                tcgen05_fence_after();

                uint32_t sr[COLS];
                SDPA_TMEM_LD32(s_addr, sr);
                SDPA_TMEM_LD32(s_addr + 32, (sr + 32));
                tmem_wait_ld();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_sfree[sb]);
                if constexpr (MASKED) {
                    const int keys_left = prm.n - key_tile * TILE - COLS * half;
#pragma unroll
                    for (int c = 0; c < COLS; ++c)
                        if (c >= keys_left) sr[c] = 0xff800000u;  // -inf
                }
                float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
                for (int c = 0; c < COLS; c += 8) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])));
                    mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5])));
                    mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7])));
                }
                const float my_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if constexpr (FIRST) {
                    // the reference of the segment: its first tile's row max, agreed by the two halves, published to the other group
                    sm.xchg[0][half][row_in_tile] = my_max;
                    named_barrier_sync(bar_id, 32 * NPARTS);
                    m_ref = fmaxf(sm.xchg[0][0][row_in_tile], sm.xchg[0][1][row_in_tile]);
                    if (half == 0) sm.mref[row_in_tile] = m_ref;
                    named_barrier_sync(bar_all, 32 * NPARTS * GROUPS);
                }
                const uint64_t neg_ref2 = pack_f32x2(-m_ref * scale, -m_ref * scale);
                uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
                uint32_t pr[COLS / 2];
#pragma unroll
                for (int ch = 0; ch < COLS / 16; ++ch) exp_chunk(sr + 16 * ch, neg_ref2, acc0, acc1, pr + 8 * ch);
                if (g >= 2) mbar_wait(&sm.pv_done[sb], (uint32_t)((g >> 1) - 1) & 1u, 310 + sb);
                tcgen05_fence_after();
#pragma unroll
                for (int ch = 0; ch < COLS / 16; ++ch) SDPA_TMEM_ST8(p_addr + 8 * ch, (pr + 8 * ch));
                float a0, a1, a2, a3;
                unpack_f32x2(acc0, a0, a1);
                unpack_f32x2(acc1, a2, a3);
                lsum += (a0 + a1) + (a2 + a3);
                if constexpr (!FIRST) {
                    if (__any_sync(0xffffffffu, (my_max - m_ref) * scale > kGuardThreshold)) {
                        if (lane == 0) atomicExch(prm.guard, prm.epoch);
                    }
                }
                tmem_wait_st();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_pready[sb]);
            };

            SegCursor cur;
            cur.init(unit_begin, unit_end, T);
            int g0 = 0;   // tile counter at the start of the segment
            while (cur.next()) {
                const bool ragged = (prm.n % TILE) != 0 && (cur.t0 + cur.nt) == T;   // the segment ends with the shard's partial tile
                m_ref = -CUDART_INF_F;
                lsum = 0.f;
                int g;
                if (group == (g0 & 1)) {   // this group owns the segment's first tile: it fixes the reference
                    if (ragged && cur.nt == 1) tile_step(g0, cur.t0, std::true_type{}, std::true_type{});
                    else tile_step(g0, cur.t0, std::false_type{}, std::true_type{});
                    g = g0 + 2;
                } else {
                    named_barrier_sync(bar_all, 32 * NPARTS * GROUPS);
                    m_ref = sm.mref[row_in_tile];
                    g = g0 + 1;
                }
                for (; g < g0 + cur.nt; g += 2) {
                    if (ragged && g == g0 + cur.nt - 1) tile_step(g, cur.t0 + (g - g0), std::true_type{}, std::false_type{});
                    else tile_step(g, cur.t0 + (g - g0), std::false_type{}, std::false_type{});
                }

                // ---------------- epilogue of the segment ----------------
                sm.xchg[1][gp][row_in_tile] = lsum;
                named_barrier_sync(bar_all, 32 * NPARTS * GROUPS);
                lsum = 0.f;
#pragma unroll
                for (int p = 0; p < NPARTS * GROUPS; ++p) lsum += sm.xchg[1][p][row_in_tile];
                mbar_wait(&sm.o_done, (uint32_t)cur.seg & 1u, 320);
                tcgen05_fence_after();
                const int grow = (2 * cur.rb + (int)rank) * TILE + row_in_tile;
                const int piece = cluster - wm_cluster_of(wm, (long long)cur.rb * T);
                const bool valid = grow < prm.rows;
                {
                    uint32_t orr[32];
                    SDPA_TMEM_LD32(o_addr, orr);
                    tmem_wait_ld();
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(leader_ofree);   // my share of O is in registers: the accumulator may be reused
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(prm.part_o + ((size_t)piece * prm.rows_capacity + grow) * HEAD + OCOLS * gp);
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            dst[c / 4] = make_float4(__uint_as_float(orr[c]), __uint_as_float(orr[c + 1]),
                                                     __uint_as_float(orr[c + 2]), __uint_as_float(orr[c + 3]));
                    }
                }
                if (valid && gp == 0) {
                    prm.part_tmax[(size_t)piece * prm.rows_capacity + grow] = m_ref * scale;
                    prm.part_lsum[(size_t)piece * prm.rows_capacity + grow] = lsum;
                }
                g0 += cur.nt;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2cta(tmem, 512);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D bf16 row-major [rows][cols] tensor, box = 64 columns x box_rows rows, 128-byte swizzle; rows and columns of a box that
// lie outside the tensor are filled with zeros (this is what replaces the reference's masked vector tails, mpi.c:115-119).
sdpa_status encode_map(CUtensorMap* map, const void* base, int rows, int box_rows = TILE, int cols = HEAD)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return SDPA_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (base=%p rows=%d cols=%d)", (int)r, base, rows, cols);
        return SDPA_ERR_CUDA;
    }
    return SDPA_OK;
}

}  // namespace

struct UmmaPlan {
    int dk = 0, dv = 0, hl = 1;      // operand widths; hl = 2: hi/lo split operands (lo arrays lo_off elements behind hi)
    bool general = false;            // the shape goes to attn_umma_general_kernel (anything but dk = dv = 128 bf16)
    CUtensorMap gmaps[2][6];         // general kernel, per Q slot: q_hi q_lo k_hi k_lo v_hi v_lo
    CUtensorMap map_k, map_v, map_q[2];
    CUtensorMap map_khalf;   // K with 64-row boxes (v7: each CTA of a pair keeps 64 keys of a tile)
    int n = 0;
    bool kv_bound = false, q_bound[2] = {false, false};
    const void* q_base[2] = {nullptr, nullptr};
    const void* k_base = nullptr;
    const void* v_base = nullptr;
    int q_rows[2] = {0, 0};
    bool attr_set[64] = {};
    unsigned int* guard = nullptr;   // device word for the overflow guard (on the plan's device)
    unsigned int epoch = 0;
    bool allow_v8 = false;           // set per call by the engine: the caller merges with launch_merge_pieces
    bool last_v8 = false;            // the last launch used the persistent kernel; last_wm / last_pieces describe its partials
    WorkMap last_wm{0, 0, 0};
    int last_pieces = 0;
    bool attr8_set[64] = {};
};

sdpa_status umma_plan_create(UmmaPlan** plan)
{
    *plan = new UmmaPlan();
    return SDPA_OK;
}
void umma_plan_destroy(UmmaPlan* plan)
{
    if (plan && plan->guard) cudaFree(plan->guard);
    delete plan;
}

bool attn_umma_supported(int dk, int dv, int hl) { return attn_umma_general_shape(dk, dv, hl, nullptr); }

static bool force_general()
{
    const char* e = getenv("SDPA_UMMA_GENERAL");   // developer knob: run dk = dv = 128 bf16 on the general kernel too
    return e && *e == '1';
}

sdpa_status umma_plan_bind_kv(UmmaPlan* plan, const __nv_bfloat16* K, const __nv_bfloat16* V, int n, int dk, int dv, int hl,
                              size_t k_lo_off, size_t v_lo_off)
{
    if (!attn_umma_supported(dk, dv, hl)) {
        set_error("tensor-core kernel: dk, dv must be multiples of 8 up to 256 (split precision: up to 128); got dk=%d dv=%d", dk, dv);
        return SDPA_ERR_UNSUPPORTED;
    }
    if (plan->kv_bound && plan->k_base == K && plan->v_base == V && plan->n == n && plan->dk == dk && plan->dv == dv && plan->hl == hl)
        return SDPA_OK;   // descriptors still valid
    if (plan->dk != dk || plan->hl != hl) plan->q_bound[0] = plan->q_bound[1] = false;
    plan->dk = dk;
    plan->dv = dv;
    plan->hl = hl;
    plan->general = !(dk == HEAD && dv == HEAD && hl == 1) || force_general();
    if (plan->general) {
        for (int slot = 0; slot < 2; ++slot) {
            SDPA_TRY(encode_map(&plan->gmaps[slot][2], K, n, 64, dk));
            SDPA_TRY(encode_map(&plan->gmaps[slot][3], K + (hl == 2 ? k_lo_off : 0), n, 64, dk));
            SDPA_TRY(encode_map(&plan->gmaps[slot][4], V, n, TILE, dv));
            SDPA_TRY(encode_map(&plan->gmaps[slot][5], V + (hl == 2 ? v_lo_off : 0), n, TILE, dv));
        }
    }
    if (dk == HEAD && dv == HEAD && hl == 1) {
        SDPA_TRY(encode_map(&plan->map_k, K, n));
        SDPA_TRY(encode_map(&plan->map_v, V, n));
        SDPA_TRY(encode_map(&plan->map_khalf, K, n, 64));
    }
    plan->k_base = K;
    plan->v_base = V;
    plan->n = n;
    plan->kv_bound = true;
    return SDPA_OK;
}

sdpa_status umma_plan_bind_q(UmmaPlan* plan, int slot, const __nv_bfloat16* Q, int rows_capacity, int dk, int hl, size_t q_lo_off)
{
    if (slot < 0 || slot > 1 || dk != plan->dk || hl != plan->hl) {
        set_error("umma_plan_bind_q: bad slot, or dk / precision differ from the bound K/V shard");
        return SDPA_ERR_INVALID;
    }
    if (plan->q_bound[slot] && plan->q_base[slot] == Q && plan->q_rows[slot] == rows_capacity) return SDPA_OK;
    if (plan->general) {
        SDPA_TRY(encode_map(&plan->gmaps[slot][0], Q, rows_capacity, TILE, dk));
        SDPA_TRY(encode_map(&plan->gmaps[slot][1], Q + (hl == 2 ? q_lo_off : 0), rows_capacity, TILE, dk));
    }
    if (dk == HEAD && plan->dv == HEAD && hl == 1) SDPA_TRY(encode_map(&plan->map_q[slot], Q, rows_capacity));
    plan->q_bound[slot] = true;
    plan->q_base[slot] = Q;
    plan->q_rows[slot] = rows_capacity;
    return SDPA_OK;
}

static bool trace_env_set()
{
    const char* t = getenv("SDPA_UMMA_TRACE");
    return t && *t;
}

// Kernel generation: v7 (cluster of two 128-row CTAs, 2-CTA MMA) unless SDPA_UMMA_V7=0 asks for v5.
static bool use_v7()
{
    const char* e7 = getenv("SDPA_UMMA_V7");
    return e7 ? (*e7 != '0') : true;
}

// Persistent kernel (EXPERIMENTAL, SDPA_UMMA_V8=1): its work map for (rows, n) on sm_count SMs, or false when the shape
// does not suit it (too little work per cluster, or more pieces per row block than the merge takes).
static bool v8_work_map(int rows, int n, int sm_count, WorkMap* wm, int* max_pieces)
{
    const char* e = getenv("SDPA_UMMA_V8");
    if (!(e && *e == '1') || rows <= 0 || n <= 0) return false;
    WorkMap w;
    w.T = ceil_div(n, TILE);
    w.RB = ceil_div(rows, 2 * TILE);
    w.C = std::max(1, sm_count / 2);
    if (wm_total(w) < 4LL * w.C || wm_total(w) > 0x3fffffffLL) return false;
    int mp = 0;
    for (int rb = 0; rb < w.RB; ++rb) mp = std::max(mp, wm_pieces(w, rb));
    if (mp < 1 || mp > 64) return false;
    *wm = w;
    *max_pieces = mp;
    return true;
}

int attn_umma_v8_pieces(int rows, int n, int sm_count)
{
    WorkMap w;
    int mp = 0;
    return v8_work_map(rows, n, sm_count, &w, &mp) ? mp : 0;
}

void umma_plan_allow_v8(UmmaPlan* plan, bool allow)
{
    if (plan) plan->allow_v8 = allow;
}

bool umma_plan_last_v8(const UmmaPlan* plan, WorkMap* wm, int* max_pieces, const unsigned int** guard, unsigned int* epoch)
{
    if (!plan || !plan->last_v8) return false;
    *wm = plan->last_wm;
    *max_pieces = plan->last_pieces;
    *guard = plan->guard;
    *epoch = plan->epoch;
    return true;
}

int attn_umma_pick_splits(int rows, int n, int sm_count)
{
    const int row_blocks = use_v7() ? 2 * ceil_div(ceil_div(rows, TILE), 2) : ceil_div(rows, BLOCK_ROWS);
    const int tiles = ceil_div(n, TILE);
    if (row_blocks <= 0 || tiles <= 1) return 1;
    // choose the split count (<= 64, >= 4 key tiles each) with the best wave efficiency of the
    // grid row_blocks x splits over sm_count CTAs-at-a-time; prefer fewer splits on ties.
    const int max_splits = std::min(64, std::max(1, tiles / 4));
    auto efficiency = [&](int s) {
        const int ctas = row_blocks * s;
        const int waves = ceil_div(ctas, sm_count);
        const int tiles_per = ceil_div(tiles, s);  // the slowest CTA of a wave sets its length
        return (double)row_blocks * tiles / ((double)waves * sm_count * tiles_per);
    };
    double best_eff = 0.0;
    for (int s = 1; s <= max_splits; ++s) best_eff = std::max(best_eff, efficiency(s));
    // fewest splits within 4% of the best: every extra split costs rows*dv*8 bytes of partial traffic
    int best = 1;
    for (int s = 1; s <= max_splits; ++s)
        if (efficiency(s) >= 0.96 * best_eff) {
            best = s;
            break;
        }
    return best;
}

sdpa_status launch_attn_umma(UmmaPlan* plan, int q_slot, int rows, int splits, Partials part, double* out64,
                             int sm_count, cudaStream_t stream)
{
    if (!plan || !plan->kv_bound || q_slot < 0 || q_slot > 1 || !plan->q_bound[q_slot]) {
        set_error("launch_attn_umma: plan is not bound");
        return SDPA_ERR_INVALID;
    }
    if (rows <= 0) return SDPA_OK;
    if (splits < 1) splits = 1;
    if (out64 != nullptr && splits != 1) {
        set_error("direct fp64 output requires splits == 1");
        return SDPA_ERR_INVALID;
    }
    if (plan->general) {
        // every shape but dk = dv = 128 bf16: the general kernel, then its exact twin (leaves at once unless the guard fired)
        if (!plan->guard) {
            SDPA_CUDA_TRY(cudaMalloc(&plan->guard, sizeof(unsigned int)));
            SDPA_CUDA_TRY(cudaMemset(plan->guard, 0, sizeof(unsigned int)));
        }
        GeneralLaunch L;
        L.rows = rows;
        L.n = plan->n;
        L.dk = plan->dk;
        L.dv = plan->dv;
        L.hl = plan->hl;
        L.splits = splits;
        L.part = part;
        L.out64 = out64;
        L.guard = plan->guard;
        L.epoch = ++plan->epoch;
        if (plan->epoch == 0) L.epoch = ++plan->epoch;   // 0 is the "never raised" value
        L.maps = plan->gmaps[q_slot];
        plan->last_v8 = false;
        const char* env_safe = getenv("SDPA_UMMA_SAFE");
        if (env_safe && *env_safe == '1') {   // developer knob: only the exact variant, guard forced
            SDPA_CUDA_TRY(cudaMemsetAsync(plan->guard, 0xff, sizeof(unsigned int), stream));
            L.epoch = 0xffffffffu;
        } else {
            L.exact = false;
            SDPA_TRY(launch_attn_umma_general(L, stream));
        }
        L.exact = true;
        return launch_attn_umma_general(L, stream);
    }
    const size_t smem_bytes = sizeof(SharedStorage) + 1024;
    int dev = 0;
    SDPA_CUDA_TRY(cudaGetDevice(&dev));
    // developer knobs: SDPA_UMMA_POLY=0|4|8 exponentials of every 16 on the FMA pipe; SDPA_UMMA_SAFE=1 forces the safe kernel
    const char* env_poly = getenv("SDPA_UMMA_POLY");
    int poly = env_poly ? atoi(env_poly) : kDefaultPoly;
    if (poly != 0 && poly != 4 && poly != 8) poly = kDefaultPoly;
    const char* env_chunk = getenv("SDPA_UMMA_CHUNK");
    const bool chunked = !(env_chunk && *env_chunk == '0');
    const char* env_safe = getenv("SDPA_UMMA_SAFE");
    const bool force_safe = env_safe && *env_safe == '1';
    if (dev < 64 && !plan->attr_set[dev]) {
        const int sb = (int)smem_bytes;
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<false, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<false, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<false, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<false, false, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<false, false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel<true, false, kDefaultPoly>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
        const int sb7 = (int)(sizeof(SharedV7) + 1024);
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v7<false, 4, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb7));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v7<false, 0, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb7));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v7<false, 4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb7));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v7<false, 8, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb7));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v7<true, 4, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb7));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v7<true, 4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb7));
        plan->attr_set[dev] = true;
    }
    if (!plan->guard) {
        SDPA_CUDA_TRY(cudaMalloc(&plan->guard, sizeof(unsigned int)));
        SDPA_CUDA_TRY(cudaMemset(plan->guard, 0, sizeof(unsigned int)));
    }
    KernelParams prm;
    prm.rows = rows;
    prm.n = plan->n;
    prm.tiles_total = ceil_div(plan->n, TILE);
    prm.splits = splits;
    prm.scale_log2 = (1.0f / sqrtf((float)HEAD)) * 1.4426950408889634f;
    prm.part_o = part.o;
    prm.part_tmax = part.tmax;
    prm.part_lsum = part.lsum;
    prm.rows_capacity = part.rows_capacity;
    prm.out64 = out64;
    prm.trace = nullptr;
    prm.guard = plan->guard;
    prm.epoch = ++plan->epoch;
    if (plan->epoch == 0) prm.epoch = ++plan->epoch;   // 0 is the "never raised" value
    dim3 grid(ceil_div(rows, BLOCK_ROWS), splits);                      // v5 / SAFE: 256 rows per CTA
    const bool v7 = use_v7();   // the 2-CTA kernel is the default; SDPA_UMMA_V7=0 selects v5
    const char* env_groups = getenv("SDPA_UMMA_GROUPS");   // v7: softmax groups ping-ponged over key tiles (1 or 2)
    const bool groups2 = env_groups ? (atoi(env_groups) == 2) : true;
    const size_t smem7 = sizeof(SharedV7) + 1024;
    dim3 grid6(2 * ceil_div(ceil_div(rows, TILE), 2), splits);          // v7: 128 rows per CTA, clusters of two along x
    // persistent kernel: only when the engine announced that it merges by pieces, no direct fp64 output, and the caller's
    // split count is the map's piece count (the SAFE twin behind it then fills every partial slot the merge may read)
    WorkMap wm8{0, 0, 0};
    int pieces8 = 0;
    const bool use_v8 = plan->allow_v8 && out64 == nullptr && !force_safe && !(trace_env_set()) &&
                        v8_work_map(rows, plan->n, sm_count, &wm8, &pieces8) && pieces8 == splits;
    plan->last_v8 = false;
    const char* trace_path = getenv("SDPA_UMMA_TRACE");   // developer aid: dump a clock64 timeline of CTA (0,0)
    if (trace_path && *trace_path) {
        const size_t count = (size_t)TRACE_ROLES * TRACE_ITERS * TRACE_EVENTS;
        long long* dtrace = nullptr;
        SDPA_CUDA_TRY(cudaMalloc(&dtrace, count * sizeof(long long)));
        SDPA_CUDA_TRY(cudaMemsetAsync(dtrace, 0, count * sizeof(long long), stream));
        prm.trace = dtrace;
        if (v7 && groups2) attn_umma_kernel_v7<true, 4, 2, 2><<<grid6, 640, smem7, stream>>>(plan->map_khalf, plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else if (v7) attn_umma_kernel_v7<true, 4, 2, 1><<<grid6, 384, smem7, stream>>>(plan->map_khalf, plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else attn_umma_kernel<true, false, kDefaultPoly><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        count_launch();
        std::vector<long long> host(count);
        SDPA_CUDA_TRY(cudaMemcpyAsync(host.data(), dtrace, count * sizeof(long long), cudaMemcpyDeviceToHost, stream));
        SDPA_CUDA_TRY(cudaStreamSynchronize(stream));
        SDPA_CUDA_TRY(cudaFree(dtrace));
        if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "role iteration event clock\n");
            for (int r = 0; r < TRACE_ROLES; ++r)
                for (int j = 0; j < TRACE_ITERS; ++j)
                    for (int e = 0; e < TRACE_EVENTS; ++e)
                        if (host[(r * TRACE_ITERS + j) * TRACE_EVENTS + e])
                            fprintf(f, "%d %d %d %lld\n", r, j, e, host[(r * TRACE_ITERS + j) * TRACE_EVENTS + e]);
            fclose(f);
        }
        prm.trace = nullptr;
    } else if (force_safe) {
        SDPA_CUDA_TRY(cudaMemsetAsync(plan->guard, 0xff, sizeof(unsigned int), stream));
        prm.epoch = 0xffffffffu;
    } else if (use_v8) {
        const int dev8 = dev;
        const size_t smem8 = sizeof(SharedV8) + 1024;
        if (dev8 < 64 && !plan->attr8_set[dev8]) {
            SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v8<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
            SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_kernel_v8<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
            plan->attr8_set[dev8] = true;
        }
        prm.wm = wm8;
        const dim3 grid8(2 * wm8.C, 1);
        if (poly == 4) attn_umma_kernel_v8<4><<<grid8, 640, smem8, stream>>>(plan->map_khalf, plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else attn_umma_kernel_v8<0><<<grid8, 640, smem8, stream>>>(plan->map_khalf, plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        count_launch();
        plan->last_v8 = true;
        plan->last_wm = wm8;
        plan->last_pieces = pieces8;
    } else if (v7) {
#define SDPA_LAUNCH_V7(P, G) attn_umma_kernel_v7<false, P, 2, G><<<grid6, 128 + 256 * G, smem7, stream>>>(plan->map_khalf, plan->map_q[q_slot], plan->map_k, plan->map_v, prm)
        if (!groups2) SDPA_LAUNCH_V7(4, 1);
        else if (poly == 0) SDPA_LAUNCH_V7(0, 2);
        else if (poly == 8) SDPA_LAUNCH_V7(8, 2);
        else SDPA_LAUNCH_V7(4, 2);
#undef SDPA_LAUNCH_V7
        count_launch();
    } else {
        if (!chunked && poly == 0) attn_umma_kernel<false, false, 0, false><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else if (!chunked) attn_umma_kernel<false, false, 4, false><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else if (poly == 0) attn_umma_kernel<false, false, 0><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else if (poly == 8) attn_umma_kernel<false, false, 8><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        else attn_umma_kernel<false, false, 4><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
        count_launch();
    }
    // the safe variant: leaves immediately unless the guard was raised for this epoch
    attn_umma_kernel<false, true, 0><<<grid, NTHREADS, smem_bytes, stream>>>(plan->map_q[q_slot], plan->map_k, plan->map_v, prm);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

}  // namespace sdpa
