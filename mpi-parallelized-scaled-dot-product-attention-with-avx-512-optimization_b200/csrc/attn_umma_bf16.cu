// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
namespace sdpa {
struct UmmaPlan { int unused; };
sdpa_status umma_plan_create(UmmaPlan** plan) { *plan = new UmmaPlan(); return SDPA_OK; }
void umma_plan_destroy(UmmaPlan* plan) { delete plan; }
sdpa_status umma_plan_bind_kv(UmmaPlan*, const __nv_bfloat16*, const __nv_bfloat16*, int, int, int) { return SDPA_OK; }
sdpa_status umma_plan_bind_q(UmmaPlan*, int, const __nv_bfloat16*, int, int) { return SDPA_OK; }
sdpa_status launch_attn_umma(UmmaPlan*, int, int, int, Partials, double*, int, cudaStream_t)
{
    set_error("bf16 tcgen05 kernel not built yet");
    return SDPA_ERR_UNSUPPORTED;
}
bool attn_umma_supported(int, int) { return false; }
int attn_umma_pick_splits(int, int, int) { return 1; }
}  // namespace sdpa
