// umma_general.h -- launch interface of the general tcgen05 attention kernel (attn_umma_general.cu).
#pragma once

#include "common.cuh"

namespace sdpa {

struct GeneralShape {
    int dk16;      // MMA k-steps of S = Q K^T
    int dkb;       // 64-column boxes per Q / K row
    int dv_pad;    // 128 or 256
    int kst, vst;  // K and V ring depths that fit the 227 KiB of shared memory
    size_t smem_bytes;
};
// hl = 1: bf16 operands; hl = 2: hi/lo split operands (fp32-accurate).  False when no tensor-core kernel takes the shape.
bool attn_umma_general_shape(int dk, int dv, int hl, GeneralShape* out);

struct GeneralLaunch {
    int rows, n, dk, dv, hl, splits;
    bool exact;              // the two-phase repair variant (runs only if *guard == epoch)
    Partials part;
    double* out64;
    unsigned int* guard;
    unsigned int epoch;
    const void* maps;        // CUtensorMap[6]: q_hi q_lo k_hi k_lo v_hi v_lo (K maps: 64-row boxes; Q, V: 128-row boxes)
};
sdpa_status launch_attn_umma_general(const GeneralLaunch& launch, cudaStream_t stream);

}  // namespace sdpa
