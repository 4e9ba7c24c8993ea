// bf16 tcgen05 implicit-GEMM path (AZP_DTYPE_BF16) -- placeholder until the kernels land.
#include "azp_net.cuh"

int azp_umma_create(azp_engine* e) { e->set_error("bf16 tcgen05 path not built yet"); return AZP_ERR_UNSUPPORTED; }
void azp_umma_destroy(azp_engine*) {}
int azp_umma_build(azp_engine* e, const HostNet&, NetWeights*) { e->set_error("bf16 tcgen05 path not built yet"); return AZP_ERR_UNSUPPORTED; }
int azp_umma_eval_leaves(azp_engine* e, NetWeights*, const int32_t*, int, int, cudaStream_t) { e->set_error("bf16 tcgen05 path not built yet"); return AZP_ERR_UNSUPPORTED; }
int azp_umma_eval_planes(azp_engine* e, NetWeights*, int, const float*, float*, float*, float*, cudaStream_t) { e->set_error("bf16 tcgen05 path not built yet"); return AZP_ERR_UNSUPPORTED; }
