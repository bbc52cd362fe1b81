// clip_logits.cu -- placeholder; the tcgen05/TMA GEMM lands in a later commit.
#include "common.cuh"
extern "C" int ovdet_clip_logits_bf16(const void *, const void *, int, int, int, unsigned, float, float *, void *, float *, void *)
{
    ovdet::set_error("ovdet_clip_logits_bf16 not implemented yet");
    return OVDET_ERR_UNSUPPORTED;
}
