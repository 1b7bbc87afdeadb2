// umma_ptx.cuh -- inline-PTX wrappers shared by the tcgen05 attention kernels (sm_100a): mbarrier, TMA
// (cp.async.bulk.tensor, 1-CTA and 2-CTA forms), tcgen05 alloc / mma / commit / ld / st, cluster helpers, packed-fp32 math,
// shared-memory matrix descriptors and the instruction descriptor of kind::f16.
#pragma once

#include "common.cuh"

#include <cuda.h>
#include <math_constants.h>

namespace sdpa {
namespace umma {

constexpr int TILE = 128;                 // rows per Q tile, keys per K/V tile
constexpr uint32_t BOX_COLS = 64;         // bf16 columns of one 128-byte-swizzled TMA box

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("sdpa_b200: mbarrier timeout tag=%d block=(%d,%d) thread=%d parity=%u\n", tag, blockIdx.x, blockIdx.y,
                   threadIdx.x, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// One lane of a converged warp (the instruction is warp-uniform, so everything around it can
// stay on the uniform datapath; a divergent `lane == 0` branch forces R2UR/ELECT traffic per MMA).
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define SDPA_TMEM_LD32(taddr, r)                                                                             \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),            \
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),          \
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),          \
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                               \
        : "r"(taddr)                                                                                         \
        : "memory")

#define SDPA_TMEM_ST32(taddr, r)                                                                             \
    asm volatile(                                                                                            \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                      \
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                           \
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                   \
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),       \
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),     \
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])      \
        : "memory")


#define SDPA_TMEM_ST8(taddr, r)                                                                              \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"             \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),  \
                   "r"(r[7])                                                                                  \
                 : "memory")

#define SDPA_TMEM_LD16(taddr, r)                                                                             \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),            \
          "=r"(r[15])                                                                                         \
        : "r"(taddr)                                                                                         \
        : "memory")

#define SDPA_TMEM_ST16(taddr, r)                                                                             \
    asm volatile(                                                                                            \
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                      \
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"                           \
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])        \
        : "memory")

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// 2^x for two packed values on the FMA/ALU pipes (no MUFU): x = n + f, n = round(x) taken from the
// low mantissa bits of x + 1.5*2^23, f in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial (max
// relative error 7.5e-5, far below the bf16 rounding of P), 2^n by an integer multiply-add into the
// exponent field.  Inputs are clamped at -126 (result ~ 0); callers guarantee x < 64.
__device__ __forceinline__ void exp2_poly_x2(uint64_t x2, float& p0, float& p1)
{
    float x0, x1;
    unpack_f32x2(x2, x0, x1);
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    const uint64_t xc = pack_f32x2(x0, x1);
    const uint64_t magic = pack_f32x2(12582912.f, 12582912.f);
    const uint64_t xr = add_f32x2(xc, magic);                                   // integer part lands in the mantissa
    const uint64_t n2 = add_f32x2(xr, pack_f32x2(-12582912.f, -12582912.f));    // round(x) as a float
    const uint64_t f2 = fma_f32x2(n2, pack_f32x2(-1.f, -1.f), xc);              // x - round(x)
    uint64_t p = fma_f32x2(pack_f32x2(0.0551716685f, 0.0551716685f), f2, pack_f32x2(0.2426111251f, 0.2426111251f));
    p = fma_f32x2(p, f2, pack_f32x2(0.6932609677f, 0.6932609677f));
    p = fma_f32x2(p, f2, pack_f32x2(0.9999280572f, 0.9999280572f));
    float q0, q1, r0, r1;
    unpack_f32x2(p, q0, q1);
    unpack_f32x2(xr, r0, r1);
    p0 = __uint_as_float(__float_as_uint(r0) * 0x800000u + __float_as_uint(q0));   // += n << 23
    p1 = __uint_as_float(__float_as_uint(r1) * 0x800000u + __float_as_uint(q1));
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr uint32_t TILE_BYTES = TILE * 128 * 2;       // one 128 x 128 bf16 tile (32 KiB)
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;        // one [128][64] box

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100): start address [0,14) (>>4), leading byte offset
// [16,30) (>>4), stride byte offset [32,46) (>>4), version = 1 at [46,48), layout type at
// [61,64) (2 = 128-byte swizzle).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major operand tile [128 rows][128 cols] bf16 stored as two [128][64] 128B-swizzled boxes:
// 8-row groups are 1024 B apart (SBO); the k-th 16-column slice starts (k%4)*32 B into box k/4.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int k16)
{
    return make_desc(tile_addr + (uint32_t)(k16 >> 2) * HALF_BYTES + (uint32_t)(k16 & 3) * 32u, 16u, 1024u);
}
// MN-major operand tile (V: [128 keys][128 dv], dv contiguous) stored as two [128 keys][64 dv]
// swizzled boxes: 64-column groups are HALF_BYTES apart (LBO), 8-key groups 1024 B apart (SBO);
// the k-th 16-key slice starts k*16 rows = k*2048 B into the tile.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int k16)
{
    return make_desc(tile_addr + (uint32_t)k16 * 2048u, HALF_BYTES, 1024u);
}
// Instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1),
// b_major at bit 16, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- clusters / 2-CTA forms
// arrive (once the MMAs issued so far complete) on the barrier at this offset in every CTA of the mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// 2-SM TMA load: data into THIS CTA's shared memory, completion bytes on the LEADER CTA's mbarrier
// (bit 24 of the shared::cluster address selects the CTA of the pair; clearing it addresses rank 0).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* local, uint32_t cta_rank)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(cta_rank));
    return ra;
}
// Remote arrive with the default (CTA-scope release) semantics.  The data handed over is in TMEM / is
// TMA traffic, ordered by tcgen05.fence / complete_tx; an explicit .release.cluster here costs ~1100
// cycles per arrive (measured, profiles/r01/timeline_trace_v7_first.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

}  // namespace umma
}  // namespace sdpa
