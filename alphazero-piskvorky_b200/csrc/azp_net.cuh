// Network weights / workspace shared by the float32 SIMT path (azp_net.cu) and the bf16
// tcgen05 path (azp_net_umma.cu).
#pragma once
#include "azp_internal.cuh"

// logical layers parsed from a reference state_dict blob (BN already folded in)
struct HostConv {
  int cin, cout;
  bool relu;
  int residual;                 // 1: add the block input before the activation (tower conv2)
  std::vector<float> w;         // PyTorch layout [cout][cin][3][3]
  std::vector<float> b;         // [cout]
};
struct HostNet {
  std::vector<HostConv> convs;
  int C, PC, VC;                // trunk channels, policy-head channels, value-head channels
  std::vector<float> head_w, head_b;      // [PC+VC][C], [PC+VC]   (1x1 convs, policy rows first)
  std::vector<float> pfc_w, pfc_b;        // [A][PC*A], [A]
  std::vector<float> vfc1_w, vfc1_b;      // [64][VC*A], [64]
  std::vector<float> vfc2_w, vfc2_b;      // [64], [1]
};

struct DevConv {
  int cin, cout;
  bool relu;
  int residual;
  float* w = nullptr;           // f32 path: [tap][cin][cout]
  float* b = nullptr;           // [cout]
  __nv_bfloat16* w16 = nullptr; // bf16 path: per tap, [cout][cin] K-major (UMMA B operand), see azp_net_umma.cu
};

struct NetWeights {
  bool loaded = false;
  std::vector<void*> allocs;    // device buffers owned by this weight set
  void (*umma_free)(void*) = nullptr;
  ~NetWeights() {
    for (void* q : allocs) cudaFree(q);
    if (umma && umma_free) umma_free(umma);
  }
  int C = 0, PC = 0, VC = 0;
  size_t dominant_layer = 0;    // index of the conv with the most FLOPs (profiled for the roofline)
  std::vector<DevConv> layers;
  float* head_w = nullptr; float* head_b = nullptr;
  float* pfc_t = nullptr; float* pfc_b = nullptr;       // [PC*A][A]
  float* vfc1_t = nullptr; float* vfc1_b = nullptr;     // [VC*A][64]
  float* vfc2_w = nullptr; float* vfc2_b = nullptr;
  float* conv1_taps = nullptr;  // bf16 path: first conv as [tap][plane(3)][cout] float32
  void* umma = nullptr;         // bf16 path private data (tensor maps etc.)
};

struct NetWorkspace {
  int cmax = 0;
  float* in4 = nullptr;         // f32 path: NHWC input planes [G][A][4]
  float* act[3] = {nullptr, nullptr, nullptr};   // f32 path: NHWC activations [G][A][cmax]
  void* umma = nullptr;         // bf16 path private data
};

int azp_net_parse(const azp_engine* e, const float* blob, size_t n, HostNet& net);

// azp_net_umma.cu
int azp_umma_create(azp_engine* e);
void azp_umma_destroy(azp_engine* e);
int azp_umma_build(azp_engine* e, const HostNet& net, NetWeights* W);
int azp_umma_reload_device(azp_engine* e, NetWeights* W, const float* dev_blob, cudaStream_t s);
int azp_umma_eval_leaves(azp_engine* e, NetWeights* W, const int32_t* count_dev, int count_host, int back,
                         cudaStream_t s);
int azp_umma_debug(azp_engine* e, int layer, int n, float* out, cudaStream_t s);
int azp_umma_eval_planes(azp_engine* e, NetWeights* W, int n, const float* planes, float* logits, float* policy,
                         float* value, cudaStream_t s);
