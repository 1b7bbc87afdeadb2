// attn_umma_general.cu -- the GENERAL tcgen05 attention kernel: any dk, dv (multiples of 8, up to 256), bf16 operands or
// the fp32-accurate "bf16x3" split, and the exact (two-phase) variant that repairs a launch whose overflow guard fired.
//
// The B200 counterpart of online_softmax_attention (attention-mpi.c:168-189) for every shape the reference accepts on its
// masked-tail path (attention-mpi.c:115-119,134-138,159-165): where the reference masks the tail of a 16-lane vector, this
// kernel lets TMA zero-fill the part of a 64-column box that lies beyond dk / dv, so no padded copy of K or V exists.
//
// Structure (shared with attn_umma_kernel_v7, attn_umma_bf16.cu): a cluster of two CTAs forms ONE M=256
// tcgen05.mma.cta_group::2; each CTA owns a 128-row Q tile, its accumulators and its P in its own TMEM and keeps HALF of
// every K tile (64 keys) and V tile (dv_pad/2 value columns) in shared memory.  Differences:
//   * P ALIASES S.  TMEM = S0 S1 (2 x 128 columns) + O (dv_pad = 128 or 256 columns): a softmax thread overwrites the S
//     columns it has just read with its packed P, so dv up to 256 fits (v7's separate P buffers need 128 more columns) and
//     the split precision has room for P_hi and P_lo.  The tensor pipe executes MMAs in issue order, so S(j+2) -- issued
//     behind PV(j) -- cannot overwrite P(j) early; the MMA order is  S0 S1 | PV0 S2 | PV1 S3 | ...
//   * runtime dk / dv: ceil(dk/16) MMA k-steps per S, idesc N = dv_pad for PV; the shared-memory carve-up and the ring
//     depths are computed at launch (KernelShape).
//   * HL = 2 ("bf16x3", the fp32-accurate path): every fp32 operand x is held as hi = bf16(x), lo = bf16(x - hi) and each
//     contraction runs three MMAs into the same fp32 accumulator,  S = Ql Kh + Qh Kl + Qh Kh,  O += Pl Vh + Ph Vl + Ph Vh
//     (the lo*lo term is below 2^-17 relative and is dropped): fp32-class accuracy (<= 1e-5 on N(0,1) inputs, the
//     reference's own arithmetic is fp32, attention-mpi.c:168-189) at a third of the bf16 tensor rate.
//   * EXACT = the repair variant.  The fast variant fixes the softmax reference by the first key tile and raises a guard
//     when a later score exceeds it by more than 2^64 (attn_umma_bf16.cu); the EXACT variant is launched behind it, exits
//     at once unless the guard carries this launch's epoch, and otherwise walks its key range TWICE: phase 1 computes only
//     S = Q K^T and the exact row maximum of the CTA's own range, phase 2 is the normal pass with that maximum as the
//     reference (every exponent <= 0).  No cross-CTA dependency: each split's partial state carries its own reference.
// Warp roles: 0 = TMA producer of Q and K, 2 = TMA producer of V (separate warps: a K load never queues behind a V slot),
// 1 = MMA issuer (leader CTA), 3 idle, 4-19 = two softmax groups x (row, 64-key half) as in v7.
#include "umma_ptx.cuh"
#include "umma_general.h"

#include <type_traits>
#include <stdlib.h>

namespace sdpa {

namespace {

using namespace umma;

constexpr float kGuardThresholdG = 64.0f;   // fast mode: exponents beyond 2^64 hand the launch to the EXACT variant
constexpr int G_MAXST = 4;                  // ring depth limit
constexpr uint32_t G_TMEM_O = 256;

struct GenParams {
    int rows;            // valid Q rows
    int n;               // keys in the shard
    int tiles_total;     // ceil(n / 128)
    int splits;
    int dk16;            // ceil(dk / 16): MMA k-steps of S
    int dkb;             // ceil(dk / 64): 64-column boxes per operand row
    int dv;              // value width (output row pitch)
    int dv_pad;          // 128 or 256: MMA N of PV, TMEM columns of O
    int kst, vst;        // ring depths
    float scale_log2;    // 1/sqrt(dk) * log2(e)
    float* part_o;
    float* part_tmax;
    float* part_lsum;
    int rows_capacity;
    double* out64;       // non-null (splits == 1): normalised fp64 output
    unsigned int* guard;
    unsigned int epoch;
};

struct GenBarriers {
    uint64_t q_full;
    uint64_t k_full[G_MAXST], k_empty[G_MAXST];
    uint64_t v_full[G_MAXST], v_empty[G_MAXST];
    uint64_t s_full[2], p_ready[2], o_done;
    uint32_t tmem_base;
    uint32_t pad;
    float xmax[4][TILE];   // [group*2 + half][row]: first-tile max (fast) / range max (exact)
    float xsum[4][TILE];   // final row sums
    float mref[TILE];      // the agreed reference, handed from softmax group 0 to group 1 (fast mode)
};

// HL = 1: bf16 operands.  HL = 2: hi/lo split operands (lo arrays behind the *_lo maps).
template <int HL, bool EXACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
attn_umma_general_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                         const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                         const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                         const GenParams prm)
{
    if constexpr (EXACT) {
        if (*prm.guard != prm.epoch) return;   // nothing overflowed in the fast pass: the whole grid leaves at once
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t* const base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int dkb = prm.dkb, vb = prm.dv_pad / 128;   // boxes per operand row: K/Q (64 columns each), this CTA's V half
    const uint32_t q_bytes = (uint32_t)(HL * dkb) * 16384u;   // [HL][dkb] boxes of 128 rows x 64 columns
    const uint32_t k_stage = (uint32_t)(HL * dkb) * 8192u;    // [HL][dkb] boxes of  64 keys x 64 columns
    const uint32_t v_stage = (uint32_t)(HL * vb) * 16384u;    // [HL][vb]  boxes of 128 keys x 64 columns
    uint8_t* const sq = base;
    uint8_t* const sk = sq + q_bytes;
    uint8_t* const sv = sk + (uint32_t)prm.kst * k_stage;
    GenBarriers& sm = *reinterpret_cast<GenBarriers*>(sv + (uint32_t)prm.vst * v_stage);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int row_block = blockIdx.x;             // 128 rows; blockIdx.x = 2*cluster + rank
    const int split = blockIdx.y;
    const uint32_t rank = cluster_cta_rank();      // rank r keeps keys [64r,64r+64) of K and columns [dv_pad/2*r, ..) of V
    const bool leader = rank == 0;

    const int tq = prm.tiles_total / prm.splits, tr = prm.tiles_total % prm.splits;
    const int tile_begin = split * tq + min(split, tr);
    const int nt = tq + (split < tr ? 1 : 0);      // identical in both CTAs of the cluster
    const int g_main = EXACT ? nt : 0;             // global S index of the first tile of the main phase
    const int g_total = g_main + nt;               // S tiles issued in all (exact: every key tile twice)

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_q_hi);
        prefetch_tensormap(&map_k_hi);
        prefetch_tensormap(&map_v_hi);
        if (HL == 2) {
            prefetch_tensormap(&map_q_lo);
            prefetch_tensormap(&map_k_lo);
            prefetch_tensormap(&map_v_lo);
        }
        mbar_init(&sm.q_full, 2);          // the leader's copy is the one used: one arrival per CTA's producer + all bytes
        mbar_init(&sm.o_done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.s_full[i], 1);
            mbar_init(&sm.p_ready[i], 2 * 8);   // leader's copy: one arrival per softmax warp of the owning group, BOTH CTAs
        }
        for (int i = 0; i < G_MAXST; ++i) {
            mbar_init(&sm.k_full[i], 2);
            mbar_init(&sm.k_empty[i], 1);  // the leader's commit, multicast to both CTAs
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

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (nt > 0) {
            if (warp == 0) {
                // ================================ TMA producer: Q once, then K tiles ============================
                // every load lands in this CTA's shared memory and reports its bytes to the LEADER's barrier; the leader arms
                // the barrier with the bytes of both CTAs, the peer adds its plain arrival.
                const int qrow = row_block * TILE;
                if (elect_one_sync()) {
                    if (leader) mbar_arrive_expect_tx(&sm.q_full, 2 * q_bytes);
                    else mbar_arrive_cluster(map_to_cta(&sm.q_full, 0));
                    for (int h = 0; h < HL; ++h)
                        for (int b = 0; b < dkb; ++b)
                            tma_load_2d_2sm(sq + (uint32_t)(h * dkb + b) * 16384u, h ? &map_q_lo : &map_q_hi, &sm.q_full, 64 * b, qrow);
                }
                __syncwarp();
                int ks = 0;
                uint32_t kph = 0;
                for (int g = 0; g < g_total; ++g) {
                    const int j = g >= g_main ? g - g_main : g;
                    const int key0 = (tile_begin + j) * TILE + 64 * (int)rank;   // my 64 keys of the tile
                    mbar_wait(&sm.k_empty[ks], kph ^ 1u, 100 + ks);
                    if (elect_one_sync()) {
                        if (leader) mbar_arrive_expect_tx(&sm.k_full[ks], 2 * k_stage);
                        else mbar_arrive_cluster(map_to_cta(&sm.k_full[ks], 0));
                        uint8_t* dst = sk + (uint32_t)ks * k_stage;
                        for (int h = 0; h < HL; ++h)
                            for (int b = 0; b < dkb; ++b)
                                tma_load_2d_2sm(dst + (uint32_t)(h * dkb + b) * 8192u, h ? &map_k_lo : &map_k_hi, &sm.k_full[ks], 64 * b, key0);
                    }
                    __syncwarp();
                    if (++ks == prm.kst) {
                        ks = 0;
                        kph ^= 1u;
                    }
                }
            } else if (warp == 2) {
                // ================================ TMA producer: V tiles (main phase only) ======================
                int vs = 0;
                uint32_t vph = 0;
                const int col0 = (prm.dv_pad / 2) * (int)rank;   // my value columns
                for (int j = 0; j < nt; ++j) {
                    const int key0 = (tile_begin + j) * TILE;
                    mbar_wait(&sm.v_empty[vs], vph ^ 1u, 110 + vs);
                    if (elect_one_sync()) {
                        if (leader) mbar_arrive_expect_tx(&sm.v_full[vs], 2 * v_stage);
                        else mbar_arrive_cluster(map_to_cta(&sm.v_full[vs], 0));
                        uint8_t* dst = sv + (uint32_t)vs * v_stage;
                        for (int h = 0; h < HL; ++h)
                            for (int b = 0; b < vb; ++b)
                                tma_load_2d_2sm(dst + (uint32_t)(h * vb + b) * 16384u, h ? &map_v_lo : &map_v_hi, &sm.v_full[vs], col0 + 64 * b, key0);
                    }
                    __syncwarp();
                    if (++vs == prm.vst) {
                        vs = 0;
                        vph ^= 1u;
                    }
                }
            } else if (warp == 1 && leader) {
                // ================================ MMA issuer (leader CTA only) ==================================
                const uint32_t idesc_qk = make_idesc(2 * TILE, TILE, 0);          // M = 256 over the CTA pair, N = 128 keys
                const uint32_t idesc_pv = make_idesc(2 * TILE, prm.dv_pad, 1);    // N = dv_pad value columns, V MN-major
                const uint32_t q_addr = smem_u32(sq), k_addr = smem_u32(sk), v_addr = smem_u32(sv);
                const uint16_t both = 0x3;
                int ks = 0, vs = 0;
                uint32_t kph = 0, vph = 0;
                bool o_started = false;

                // S(g) = Q K^T into S buffer g & 1.  Split precision: the two cross terms first, the large term last.
                auto issue_s = [&](int g) {
                    const int sb = g & 1;
                    mbar_wait(&sm.k_full[ks], kph, 200 + ks);
                    tcgen05_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t d = tmem + 128u * sb;
                        const uint32_t kb = k_addr + (uint32_t)ks * k_stage;
                        uint32_t acc = 0;
#pragma unroll 1
                        for (int term = (HL == 2 ? 0 : 2); term < 3; ++term) {
                            // term 0: Q_lo K_hi, term 1: Q_hi K_lo, term 2: Q_hi K_hi
                            const uint32_t qa = q_addr + (term == 0 ? (uint32_t)dkb * 16384u : 0u);
                            const uint32_t ka = kb + (term == 1 ? (uint32_t)dkb * 8192u : 0u);
                            for (int kk = 0; kk < prm.dk16; ++kk) {
                                const uint32_t in = (uint32_t)(kk & 3) * 32u;
                                const uint64_t da = make_desc(qa + (uint32_t)(kk >> 2) * 16384u + in, 16u, 1024u);   // Q: boxes of 128 rows
                                const uint64_t db = make_desc(ka + (uint32_t)(kk >> 2) * 8192u + in, 16u, 1024u);    // K half: boxes of 64 rows
                                umma_ss_2cta(d, da, db, idesc_qk, acc);
                                acc = 1;
                            }
                        }
                        umma_commit_2cta(&sm.s_full[sb], both);
                        umma_commit_2cta(&sm.k_empty[ks], both);
                    }
                    __syncwarp();
                    if (++ks == prm.kst) {
                        ks = 0;
                        kph ^= 1u;
                    }
                };
                // O += P(g) V(j): P sits in the S columns of buffer g & 1 (keys 0-63 at +0, keys 64-127 at +64; lo parts 32 further)
                auto issue_pv = [&](int g, bool last) {
                    const int sb = g & 1;
                    mbar_wait(&sm.v_full[vs], vph, 210 + vs);
                    mbar_wait(&sm.p_ready[sb], (uint32_t)(g >> 1) & 1u, 212 + sb);
                    tcgen05_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t d = tmem + G_TMEM_O;
                        const uint32_t vbase = v_addr + (uint32_t)vs * v_stage;
#pragma unroll 1
                        for (int term = (HL == 2 ? 0 : 2); term < 3; ++term) {
                            // term 0: P_lo V_hi, term 1: P_hi V_lo, term 2: P_hi V_hi
                            const uint32_t pa = tmem + 128u * sb + (term == 0 ? 32u : 0u);
                            const uint32_t va = vbase + (term == 1 ? (uint32_t)vb * 16384u : 0u);
#pragma unroll
                            for (int kk = 0; kk < TILE / 16; ++kk) {
                                const uint64_t db = make_desc(va + (uint32_t)kk * 2048u, 16384u, 1024u);   // 16 keys further: 2 KiB; 64-column groups 16 KiB apart
                                umma_ts_2cta(d, pa + 64u * (uint32_t)(kk >> 2) + 8u * (uint32_t)(kk & 3), db, idesc_pv, o_started ? 1u : 0u);
                                o_started = true;
                            }
                        }
                        umma_commit_2cta(&sm.v_empty[vs], both);
                        if (last) umma_commit_2cta(&sm.o_done, both);
                    }
                    o_started = true;   // warp-uniform copy of the elected lane's flag
                    __syncwarp();
                    if (++vs == prm.vst) {
                        vs = 0;
                        vph ^= 1u;
                    }
                };

                mbar_wait(&sm.q_full, 0, 201);
                if constexpr (EXACT) {
                    // phase 1: S only; p_ready(g) here means "S(g) has been read", so its buffer may be overwritten
                    issue_s(0);
                    if (nt > 1) issue_s(1);
                    for (int g = 0; g < nt; ++g) {
                        mbar_wait(&sm.p_ready[g & 1], (uint32_t)(g >> 1) & 1u, 214 + (g & 1));
                        if (g + 2 < nt) issue_s(g + 2);
                    }
                }
                issue_s(g_main);
                if (nt > 1) issue_s(g_main + 1);
                for (int j = 0; j < nt; ++j) {
                    issue_pv(g_main + j, j + 1 == nt);
                    if (j + 2 < nt) issue_s(g_main + j + 2);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int sw = warp - 4;
        const int group = sw >> 3;                 // softmax group = S/P buffer it owns (S tiles with g & 1 == group)
        const int half = (sw >> 2) & 1;            // which 64-key half of the row
        const int gp = group * 2 + half;           // 0..3
        const int quad = warp & 3;                 // TMEM lane quadrant of this warp
        const int row_in_tile = quad * 32 + lane;
        const int grow = row_block * TILE + row_in_tile;
        const bool valid = grow < prm.rows;
        const int ocols = prm.dv_pad / 4;          // output columns per thread in the epilogue: 32 or 64
        if (nt > 0) {
            // ================================ softmax + epilogue ============================================
            const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
            const float scale = prm.scale_log2;
            const uint64_t scale2 = pack_f32x2(scale, scale);
            const int bar_pair = 1 + quad;             // the two warps of group 0 that share these 32 rows (first tile)
            const int bar_all = 5 + quad;              // the four warps (both groups) that share these 32 rows
            const uint32_t leader_pready = map_to_cta(&sm.p_ready[group], 0);
            const uint32_t s_addr = tmem + lane_base + 128u * (uint32_t)group + 64u * (uint32_t)half;   // my 64 S columns = my P columns

            float m_ref = -CUDART_INF_F;
            float lsum = 0.f;

            auto load_s = [&](int g, uint32_t* sr, int keys_left, bool masked) {
                mbar_wait(&sm.s_full[group], (uint32_t)(g >> 1) & 1u, 300 + group);
                tcgen05_fence_after();
                SDPA_TMEM_LD32(s_addr, sr);
                SDPA_TMEM_LD32(s_addr + 32, (sr + 32));
                tmem_wait_ld();
                if (masked) {
#pragma unroll
                    for (int c = 0; c < 64; ++c)
                        if (c >= keys_left) sr[c] = 0xff800000u;  // -inf: a key beyond n (TMA zero-filled its K row)
                }
            };
            auto row_max = [&](const uint32_t* sr) {
                float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])));
                    mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5])));
                    mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7])));
                }
                return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            };
            auto signal_p = [&]() {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_pready);   // one arrival per warp, on the leader's barrier
            };

            // exact variant, phase 1: the row maximum of this CTA's key range
            auto max_step = [&](int g, bool masked) {
                uint32_t sr[64];
                const int keys_left = prm.n - (tile_begin + g) * TILE - 64 * half;
                load_s(g, sr, keys_left, masked);
                m_ref = fmaxf(m_ref, row_max(sr));
                signal_p();   // S(g) is in registers: its buffer may be overwritten
            };

            auto tile_step = [&](int g, int j, bool masked, auto first_tag) {
                constexpr bool FIRST = decltype(first_tag)::value;   // fast variant, tile 0: this tile fixes the reference
                uint32_t sr[64];
                const int keys_left = prm.n - (tile_begin + j) * TILE - 64 * half;
                load_s(g, sr, keys_left, masked);
                float my_max = 0.f;
                if constexpr (!EXACT) my_max = row_max(sr);
                if constexpr (FIRST) {
                    sm.xmax[half][row_in_tile] = my_max;
                    named_barrier_sync(bar_pair, 64);
                    m_ref = fmaxf(my_max, sm.xmax[half ^ 1][row_in_tile]);
                    if (half == 0) sm.mref[row_in_tile] = m_ref;
                    named_barrier_sync(bar_all, 128);   // group 1 picks the reference up
                }
                const uint64_t neg_ref2 = pack_f32x2(-m_ref * scale, -m_ref * scale);
                uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t ph[8];
                    [[maybe_unused]] uint32_t pl[8];
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const uint64_t x2 = pack_f32x2(__uint_as_float(sr[16 * ch + c]), __uint_as_float(sr[16 * ch + c + 1]));
                        const uint64_t t2 = fma_f32x2(x2, scale2, neg_ref2);
                        float t0, t1;
                        unpack_f32x2(t2, t0, t1);
                        const float p0 = fast_exp2(t0), p1 = fast_exp2(t1);
                        const uint64_t p2 = pack_f32x2(p0, p1);
                        if (c & 4) acc1 = add_f32x2(acc1, p2);
                        else acc0 = add_f32x2(acc0, p2);
                        const uint32_t hi = pack_bf16x2(p0, p1);
                        ph[c / 2] = hi;
                        if constexpr (HL == 2) {
                            // residual of the bf16 rounding, itself rounded to bf16: p = hi + lo to ~2^-17 relative
                            const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
                            pl[c / 2] = pack_bf16x2(p0 - h0, p1 - h1);
                        }
                    }
                    // my P columns overwrite my own S columns (all 64 are in registers already)
                    SDPA_TMEM_ST8(s_addr + 8 * ch, ph);
                    if constexpr (HL == 2) SDPA_TMEM_ST8(s_addr + 32 + 8 * ch, pl);
                }
                float a0, a1, a2, a3;
                unpack_f32x2(acc0, a0, a1);
                unpack_f32x2(acc1, a2, a3);
                lsum += (a0 + a1) + (a2 + a3);
                if constexpr (!EXACT && !FIRST) {
                    // overflow guard (valid rows only: rows beyond `rows` hold whatever the Q buffer held before)
                    if (__any_sync(0xffffffffu, valid && (my_max - m_ref) * scale > kGuardThresholdG)) {
                        if (lane == 0) atomicExch(prm.guard, prm.epoch);   // hand the launch to the EXACT variant
                    }
                }
                tmem_wait_st();
                signal_p();
            };

            const bool ragged = (prm.n % TILE) != 0 && (tile_begin + nt) == prm.tiles_total;
            if constexpr (EXACT) {
                for (int g = group; g < nt; g += 2) max_step(g, ragged && g == nt - 1);
                sm.xmax[gp][row_in_tile] = m_ref;
                named_barrier_sync(bar_all, 128);
                m_ref = fmaxf(fmaxf(sm.xmax[0][row_in_tile], sm.xmax[1][row_in_tile]), fmaxf(sm.xmax[2][row_in_tile], sm.xmax[3][row_in_tile]));
                for (int g = g_main + ((g_main + group) & 1); g < g_total; g += 2)
                    tile_step(g, g - g_main, ragged && g == g_total - 1, std::false_type{});
            } else {
                if (group == 0) {
                    tile_step(0, 0, ragged && nt == 1, std::true_type{});
                } else {
                    named_barrier_sync(bar_all, 128);   // wait for group 0's reference
                    m_ref = sm.mref[row_in_tile];
                }
                for (int g = (group == 0 ? 2 : 1); g < nt; g += 2) tile_step(g, g, ragged && g == nt - 1, std::false_type{});
            }

            // ---------------- epilogue ----------------
            sm.xsum[gp][row_in_tile] = lsum;
            named_barrier_sync(bar_all, 128);
            lsum = (sm.xsum[0][row_in_tile] + sm.xsum[1][row_in_tile]) + (sm.xsum[2][row_in_tile] + sm.xsum[3][row_in_tile]);
            mbar_wait(&sm.o_done, 0, 320);
            tcgen05_fence_after();
            const float inv = (lsum == 0.f) ? 0.f : 1.f / lsum;
            const uint32_t o_addr = tmem + lane_base + G_TMEM_O + (uint32_t)(ocols * gp);
            for (int c0 = 0; c0 < ocols; c0 += 32) {
                uint32_t orr[32];
                SDPA_TMEM_LD32(o_addr + c0, orr);
                tmem_wait_ld();
                const int col = ocols * gp + c0;
                if (valid && col < prm.dv) {
                    const int ncol = min(32, prm.dv - col);   // dv is a multiple of 8
                    if (prm.out64 != nullptr) {
                        double2* dst = reinterpret_cast<double2*>(prm.out64 + (size_t)grow * prm.dv + col);
#pragma unroll
                        for (int c = 0; c < 32; c += 2)
                            if (c < ncol)
                                dst[c / 2] = make_double2((double)(__uint_as_float(orr[c]) * inv), (double)(__uint_as_float(orr[c + 1]) * inv));
                    } else {
                        float4* dst = reinterpret_cast<float4*>(prm.part_o + ((size_t)split * prm.rows_capacity + grow) * prm.dv + col);
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            if (c < ncol)
                                dst[c / 4] = make_float4(__uint_as_float(orr[c]), __uint_as_float(orr[c + 1]), __uint_as_float(orr[c + 2]),
                                                         __uint_as_float(orr[c + 3]));
                    }
                }
            }
            if (valid && gp == 0 && prm.out64 == nullptr) {
                prm.part_tmax[(size_t)split * prm.rows_capacity + grow] = m_ref * scale;
                prm.part_lsum[(size_t)split * prm.rows_capacity + grow] = lsum;
            }
        } else if (valid) {
            // empty key range: the neutral state (0, -inf, 0), mpi.c:172,188
            for (int c = ocols * gp; c < min(ocols * (gp + 1), prm.dv); ++c) {
                if (prm.out64 != nullptr) prm.out64[(size_t)grow * prm.dv + c] = 0.0;
                else prm.part_o[((size_t)split * prm.rows_capacity + grow) * prm.dv + c] = 0.f;
            }
            if (gp == 0 && prm.out64 == nullptr) {
                prm.part_tmax[(size_t)split * prm.rows_capacity + grow] = -CUDART_INF_F;
                prm.part_lsum[(size_t)split * prm.rows_capacity + grow] = 0.f;
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

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
void preload_attn_umma_general_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, attn_umma_general_kernel<1, false>);
    cudaFuncGetAttributes(&a, attn_umma_general_kernel<1, true>);
    cudaFuncGetAttributes(&a, attn_umma_general_kernel<2, false>);
    cudaFuncGetAttributes(&a, attn_umma_general_kernel<2, true>);
    cudaGetLastError();
}

static const size_t kGenSmemLimit = 232448 - 1024;   // 227 KiB opt-in limit minus the alignment slack

// Shared-memory need of a shape: Q tile + the K and V rings + barriers.  Ring depths: as deep as fits, at most 4, at least 2.
bool attn_umma_general_shape(int dk, int dv, int hl, GeneralShape* out)
{
    if (dk < 8 || dv < 8 || dk > 256 || dv > 256 || (dk % 8) != 0 || (dv % 8) != 0 || (hl != 1 && hl != 2)) return false;
    GeneralShape s;
    s.dk16 = ceil_div(dk, 16);
    s.dkb = ceil_div(dk, 64);
    s.dv_pad = dv <= 128 ? 128 : 256;
    const size_t q = (size_t)hl * s.dkb * 16384, k = (size_t)hl * s.dkb * 8192, v = (size_t)hl * (s.dv_pad / 128) * 16384;
    const size_t fixed = q + sizeof(GenBarriers);
    int kst = 2, vst = 2;
    if (fixed + kst * k + vst * v > kGenSmemLimit) return false;
    // grow the rings alternately while they fit (K first: a late K tile stalls the tensor pipe at once)
    for (bool grew = true; grew;) {
        grew = false;
        if (kst < G_MAXST && fixed + (kst + 1) * k + vst * v <= kGenSmemLimit) { ++kst; grew = true; }
        if (vst < G_MAXST && fixed + kst * k + (vst + 1) * v <= kGenSmemLimit) { ++vst; grew = true; }
    }
    s.kst = kst;
    s.vst = vst;
    s.smem_bytes = fixed + kst * k + vst * v + 1024;
    if (out) *out = s;
    return true;
}

sdpa_status launch_attn_umma_general(const GeneralLaunch& L, cudaStream_t stream)
{
    GeneralShape shp;
    if (!attn_umma_general_shape(L.dk, L.dv, L.hl, &shp)) {
        set_error("tensor-core kernel: unsupported shape dk=%d dv=%d (multiples of 8 up to 256%s)", L.dk, L.dv,
                  L.hl == 2 ? "; split precision: up to 128" : "");
        return SDPA_ERR_UNSUPPORTED;
    }
    static bool attr_done[64] = {};
    int dev = 0;
    SDPA_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !attr_done[dev]) {
        const int lim = (int)kGenSmemLimit + 1024;
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_general_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_general_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_general_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_umma_general_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        attr_done[dev] = true;
    }
    GenParams prm;
    prm.rows = L.rows;
    prm.n = L.n;
    prm.tiles_total = ceil_div(L.n, TILE);
    prm.splits = L.splits;
    prm.dk16 = shp.dk16;
    prm.dkb = shp.dkb;
    prm.dv = L.dv;
    prm.dv_pad = shp.dv_pad;
    prm.kst = shp.kst;
    prm.vst = shp.vst;
    prm.scale_log2 = (1.0f / sqrtf((float)L.dk)) * 1.4426950408889634f;
    prm.part_o = L.part.o;
    prm.part_tmax = L.part.tmax;
    prm.part_lsum = L.part.lsum;
    prm.rows_capacity = L.part.rows_capacity;
    prm.out64 = L.out64;
    prm.guard = L.guard;
    prm.epoch = L.epoch;
    const dim3 grid(2 * ceil_div(ceil_div(L.rows, TILE), 2), L.splits);
    const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(L.maps);   // q_hi q_lo k_hi k_lo v_hi v_lo
    if (L.hl == 2) {
        if (!L.exact) attn_umma_general_kernel<2, false><<<grid, 640, shp.smem_bytes, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], prm);
        else attn_umma_general_kernel<2, true><<<grid, 640, shp.smem_bytes, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], prm);
    } else {
        if (!L.exact) attn_umma_general_kernel<1, false><<<grid, 640, shp.smem_bytes, stream>>>(m[0], m[0], m[2], m[2], m[4], m[4], prm);
        else attn_umma_general_kernel<1, true><<<grid, 640, shp.smem_bytes, stream>>>(m[0], m[0], m[2], m[2], m[4], m[4], prm);
    }
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

}  // namespace sdpa
