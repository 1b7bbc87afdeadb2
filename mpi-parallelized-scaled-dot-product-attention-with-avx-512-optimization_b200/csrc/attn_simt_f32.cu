// attn_simt_f32.cu -- fused QK^T -> online softmax -> .V in fp32 on the CUDA cores.
//
// This is the B200 counterpart of online_softmax_attention (attention-mpi.c:168-189)
// with its helpers dot_avx512 (:103-121), axpy_avx512 (:123-140) and
// memset_zero_scale (:142-166), in the reference's own arithmetic (fp32 operands,
// fp32 accumulation).  Where the reference streams the whole K/V shard once per Q
// row and rescales the accumulator on every key, this kernel
//   * tiles 64 Q rows x 64 keys per CTA so each K/V element fetched into shared
//     memory is used by 64 rows (K/V tiles staged with 16-byte cp.async),
//   * keeps the running max / sum per row in registers and reduces them across the
//     16 lanes that share a row with warp shuffles,
//   * rescales the output accumulator once per 64-key tile, not per key,
//   * splits the key range over `splits` CTAs per row block (grid = row blocks x
//     splits) so small m still fills the 148 SMs; the split states are merged by
//     merge_kernels.cu with the same arithmetic as the cross-rank merge.
// It accepts any dk, dv <= 256 (the reference accepts any; masked tails there,
// zero-padded tiles here).  Bound: fp32 FFMA issue rate (2*m*n*(dk+dv) flops).
#include "common.cuh"

#include <math_constants.h>

#include <algorithm>

namespace sdpa {

namespace {

constexpr int BM = 64;        // Q rows per CTA
constexpr int BN = 64;        // keys per tile
constexpr int NTHREADS = 256; // 16 row groups (4 rows each) x 16 column lanes
constexpr int PSTRIDE = BN + 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Stage `rows_valid` rows of `width` floats (row stride `width` in global memory) into a
// [ROWS][D+4] shared tile; rows >= rows_valid and columns in [width, width_pad) are zeroed.
template <int ROWS, int D>
__device__ __forceinline__ void stage_tile(float* __restrict__ tile, const float* __restrict__ g,
                                           int rows_valid, int width, int width_pad, bool vec_ok)
{
    constexpr int STRIDE = D + 4;
    if (vec_ok) {
        const int chunks = width >> 2;  // 16-byte chunks per row (width % 4 == 0)
        for (int idx = threadIdx.x; idx < ROWS * chunks; idx += NTHREADS) {
            const int r = idx / chunks;
            const int c = idx - r * chunks;
            float* dst = tile + r * STRIDE + c * 4;
            if (r < rows_valid) cp_async16(dst, g + (size_t)r * width + c * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int idx = threadIdx.x; idx < ROWS * width_pad; idx += NTHREADS) {
            const int r = idx / width_pad;
            const int c = idx - r * width_pad;
            tile[r * STRIDE + c] = (r < rows_valid && c < width) ? g[(size_t)r * width + c] : 0.f;
        }
    }
}

template <int D>
struct Smem {
    static constexpr int STRIDE = D + 4;
    static constexpr size_t bytes = sizeof(float) * ((size_t)BM * STRIDE + (size_t)BN * STRIDE + (size_t)BM * PSTRIDE);
};

// grid = (row blocks, splits).  Split s covers key tiles [s*tiles_per_split, ...).
template <int D>
__global__ void __launch_bounds__(NTHREADS, (D <= 128) ? 2 : 1)
attn_f32_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                int rows, int n, int dk, int dv, float scale_log2, int tiles_per_split,
                float* __restrict__ part_o, float* __restrict__ part_tmax, float* __restrict__ part_lsum,
                int rows_capacity, double* __restrict__ out64, int vec_flags)
{
    constexpr int STRIDE = D + 4;
    constexpr int DC = D / 64;  // float4 column groups of the output owned by a thread
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                  // [BM][STRIDE]
    float* KVs = Qs + BM * STRIDE;     // [BN][STRIDE]  (K tile, then V tile)
    float* Ps = KVs + BN * STRIDE;     // [BM][PSTRIDE]

    const int tid = threadIdx.x;
    const int ti = tid >> 4;  // row group: rows ti*4 .. ti*4+3
    const int tj = tid & 15;  // column lane
    const int row0 = blockIdx.x * BM;
    const int split = blockIdx.y;
    const int rows_here = min(BM, rows - row0);

    const int tiles_total = (n + BN - 1) / BN;
    const int tile_begin = split * tiles_per_split;
    const int tile_end = min(tiles_total, tile_begin + tiles_per_split);

    const int dk_pad = (dk + 3) & ~3;
    const bool k_vec = (vec_flags & 1) != 0;  // dk % 4 == 0 and Q, K 16-byte aligned
    const bool v_vec = (vec_flags & 2) != 0;  // dv % 4 == 0 and V, partials 16-byte aligned

    // Q block -> shared (once).
    stage_tile<BM, D>(Qs, Q + (size_t)row0 * dk, rows_here, dk, dk_pad, k_vec);

    float o[4][DC * 4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < DC * 4; ++c) o[r][c] = 0.f;
    float row_max[4], row_sum[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { row_max[r] = -CUDART_INF_F; row_sum[r] = 0.f; }

    for (int t = tile_begin; t < tile_end; ++t) {
        const int key0 = t * BN;
        const int keys_here = min(BN, n - key0);

        // ---- K tile -> shared -------------------------------------------------
        stage_tile<BN, D>(KVs, K + (size_t)key0 * dk, keys_here, dk, dk_pad, k_vec);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();

        // ---- S = Q K^T : thread owns rows ti*4+r, keys tj+16c -------------------
        float s[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
        {
            const float* qp = Qs + (ti * 4) * STRIDE;
            const float* kp = KVs + tj * STRIDE;
#pragma unroll 4
            for (int k = 0; k < dk_pad; k += 4) {
                float4 qv[4], kv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) qv[r] = *reinterpret_cast<const float4*>(qp + r * STRIDE + k);
#pragma unroll
                for (int c = 0; c < 4; ++c) kv[c] = *reinterpret_cast<const float4*>(kp + (16 * c) * STRIDE + k);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        s[r][c] = fmaf(qv[r].x, kv[c].x, s[r][c]);
                        s[r][c] = fmaf(qv[r].y, kv[c].y, s[r][c]);
                        s[r][c] = fmaf(qv[r].z, kv[c].z, s[r][c]);
                        s[r][c] = fmaf(qv[r].w, kv[c].w, s[r][c]);
                    }
            }
        }
        __syncthreads();  // everyone is done reading the K tile

        // ---- V tile -> the same buffer, overlapped with the softmax -------------
        stage_tile<BN, D>(KVs, V + (size_t)key0 * dv, keys_here, dv, dv, v_vec);
        cp_async_commit();

        // ---- online softmax in the log2 domain ------------------------------------
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float tmax = -CUDART_INF_F;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool valid = (tj + 16 * c) < keys_here;
                s[r][c] = valid ? s[r][c] * scale_log2 : -CUDART_INF_F;
                tmax = fmaxf(tmax, s[r][c]);
            }
            // reduce over the 16 lanes that share this row (contiguous half-warp)
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 8));
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 4));
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
            const float new_max = fmaxf(row_max[r], tmax);      // finite: every tile has >= 1 valid key
            const float corr = exp2f(row_max[r] - new_max);     // first tile: 2^(-inf) = 0
            row_max[r] = new_max;
            float psum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float p = exp2f(s[r][c] - new_max);        // masked keys: 2^(-inf) = 0
                psum += p;
                Ps[(ti * 4 + r) * PSTRIDE + tj + 16 * c] = p;
            }
            row_sum[r] = row_sum[r] * corr + psum;               // lane-partial; reduced at the end
#pragma unroll
            for (int c = 0; c < DC * 4; ++c) o[r][c] *= corr;
        }
        cp_async_wait_all();
        __syncthreads();  // P and the V tile are visible

        // ---- O += P V : thread owns rows ti*4+r, columns tj*4 + 64*g + {0..3} ------
        {
            const float* pp = Ps + (ti * 4) * PSTRIDE;
#pragma unroll 2
            for (int j = 0; j < BN; j += 4) {
                float4 pv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) pv[r] = *reinterpret_cast<const float4*>(pp + r * PSTRIDE + j);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                    for (int g = 0; g < DC; ++g) {
                        const float4 vv = *reinterpret_cast<const float4*>(KVs + (j + jj) * STRIDE + tj * 4 + 64 * g);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float p = jj == 0 ? pv[r].x : jj == 1 ? pv[r].y : jj == 2 ? pv[r].z : pv[r].w;
                            o[r][g * 4 + 0] = fmaf(p, vv.x, o[r][g * 4 + 0]);
                            o[r][g * 4 + 1] = fmaf(p, vv.y, o[r][g * 4 + 1]);
                            o[r][g * 4 + 2] = fmaf(p, vv.z, o[r][g * 4 + 2]);
                            o[r][g * 4 + 3] = fmaf(p, vv.w, o[r][g * 4 + 3]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // done with P and the V tile before the next K tile lands
    }

    // ---- epilogue ----------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float l = row_sum[r];
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        const int row = ti * 4 + r;
        if (row >= rows_here) continue;
        const int grow = row0 + row;
        if (out64 != nullptr) {
            // single split, single shard: normalise and widen in place of
            // mpi.c:358-362 + cvt_f2d (mpi.c:373); gsum == 0 -> 0 as mpi.c:359.
            const float inv = (l == 0.f) ? 0.f : 1.f / l;
#pragma unroll
            for (int g = 0; g < DC; ++g)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int d = tj * 4 + 64 * g + e;
                    if (d < dv) out64[(size_t)grow * dv + d] = (double)(o[r][g * 4 + e] * inv);
                }
        } else {
            float* po = part_o + ((size_t)split * rows_capacity + grow) * dv;
#pragma unroll
            for (int g = 0; g < DC; ++g) {
                const int d = tj * 4 + 64 * g;
                if (v_vec && d + 3 < dv) {
                    *reinterpret_cast<float4*>(po + d) =
                        make_float4(o[r][g * 4 + 0], o[r][g * 4 + 1], o[r][g * 4 + 2], o[r][g * 4 + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (d + e < dv) po[d + e] = o[r][g * 4 + e];
                }
            }
            if (tj == 0) {
                part_tmax[(size_t)split * rows_capacity + grow] = row_max[r];
                part_lsum[(size_t)split * rows_capacity + grow] = l;
            }
        }
    }
}

template <int D>
sdpa_status launch_d(const float* Q, const float* K, const float* V, int rows, int n, int dk, int dv,
                     int splits, Partials part, double* out64, cudaStream_t stream)
{
    static bool configured[64] = {};  // the attribute is per device
    int dev = 0;
    SDPA_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        SDPA_CUDA_TRY(cudaFuncSetAttribute(attn_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Smem<D>::bytes));
        configured[dev] = true;
    }
    const int tiles_total = ceil_div(n, BN);
    const int tiles_per_split = tiles_total > 0 ? ceil_div(tiles_total, splits) : 1;
    const float scale_log2 = (1.0f / sqrtf((float)dk)) * 1.4426950408889634f;
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    int vec_flags = 0;
    if ((dk & 3) == 0 && al16(Q) && al16(K)) vec_flags |= 1;
    if ((dv & 3) == 0 && al16(V) && (out64 != nullptr || al16(part.o))) vec_flags |= 2;
    dim3 grid(ceil_div(rows, BM), splits);
    attn_f32_kernel<D><<<grid, NTHREADS, Smem<D>::bytes, stream>>>(
        Q, K, V, rows, n, dk, dv, scale_log2, tiles_per_split, part.o, part.tmax, part.lsum,
        part.rows_capacity, out64, vec_flags);
    count_launch();
    SDPA_CUDA_TRY(cudaGetLastError());
    return SDPA_OK;
}

}  // namespace

bool attn_f32_supported(int dk, int dv) { return dk >= 1 && dv >= 1 && dk <= 256 && dv <= 256; }

int attn_f32_pick_splits(int rows, int n, int sm_count)
{
    // Fewest splits (>= 4 key tiles each, <= 64) whose grid row_blocks x splits fills whole waves of
    // 2 resident CTAs per SM best; every extra split costs rows*dv*8 bytes of partial traffic.
    const int row_blocks = ceil_div(rows, BM);
    const int tiles = ceil_div(n, BN);
    if (row_blocks <= 0 || tiles <= 1) return 1;
    const int slots = 2 * sm_count;
    const int max_splits = std::min(64, std::max(1, tiles / 4));
    auto efficiency = [&](int s) {
        const int ctas = row_blocks * s;
        const int waves = ceil_div(ctas, slots);
        const int tiles_per = ceil_div(tiles, s);
        return (double)row_blocks * tiles / ((double)waves * slots * tiles_per);
    };
    double best_eff = 0.0;
    for (int s = 1; s <= max_splits; ++s) best_eff = std::max(best_eff, efficiency(s));
    for (int s = 1; s <= max_splits; ++s)
        if (efficiency(s) >= 0.96 * best_eff) return s;
    return 1;
}

void preload_attn_f32_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, attn_f32_kernel<64>);
    cudaFuncGetAttributes(&a, attn_f32_kernel<128>);
    cudaFuncGetAttributes(&a, attn_f32_kernel<256>);
    cudaGetLastError();
}

sdpa_status launch_attn_f32(const float* Q, const float* K, const float* V, int rows, int n, int dk,
                            int dv, int splits, Partials part, double* out64, cudaStream_t stream)
{
    if (!attn_f32_supported(dk, dv)) {
        set_error("fp32 kernel supports 1 <= dk, dv <= 256 (got dk=%d dv=%d)", dk, dv);
        return SDPA_ERR_UNSUPPORTED;
    }
    if (rows <= 0) return SDPA_OK;
    if (splits < 1) splits = 1;
    if (out64 != nullptr && splits != 1) {
        set_error("direct fp64 output requires splits == 1");
        return SDPA_ERR_INVALID;
    }
    const int d = dk > dv ? dk : dv;
    if (d <= 64) return launch_d<64>(Q, K, V, rows, n, dk, dv, splits, part, out64, stream);
    if (d <= 128) return launch_d<128>(Q, K, V, rows, n, dk, dv, splits, part, out64, stream);
    return launch_d<256>(Q, K, V, rows, n, dk, dv, splits, part, out64, stream);
}

}  // namespace sdpa
