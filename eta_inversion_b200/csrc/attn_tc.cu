// tcgen05 flash self-attention (placeholder until the kernel lands: reports "unsupported" so callers use attn_simt.cu)
#include "ops.cuh"
namespace etai {
bool attention_tc_supported(const SelfAttnArgs&) { return false; }
void attention_tc(const SelfAttnArgs&, cudaStream_t) { throw Error(ETAI_ERR_UNSUPPORTED, "attention_tc not built"); }
}  // namespace etai
