// Policy/value network, float32 SIMT path (AZP_DTYPE_F32): the <=1e-5 parity path for
// GomokuNet.forward (alphazero/net.py:55-72) and for the residual tower of BASELINE config 4.
// Weight ingestion (state_dict order, PyTorch layouts -> kernel layouts, BN folding) is shared
// with the bf16 tcgen05 path in azp_net_umma.cu.
#include "azp_net.cuh"
#include <math.h>
#include <string.h>

// ---- input encoders: NHWC float32 [row][A][4] ---------------------------------------------------
__global__ void planes_to_nhwc_kernel(int A, int n, const float* __restrict__ planes, float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * A) return;
  size_t b = i / A; int a = i % A;
  const float* p = planes + b * 4 * A + a;
  float4 v = make_float4(p[0], p[A], p[2 * A], p[3 * A]);
  reinterpret_cast<float4*>(out)[i] = v;
}

// leaf batch (bitboards, indexed by slot) -> NHWC planes; same plane meaning as games.py:86-129
__global__ void leaves_to_nhwc_kernel(int A, int NW, int G, const uint64_t* __restrict__ lx,
                                      const uint64_t* __restrict__ lo, const uint8_t* __restrict__ ltm,
                                      const int16_t* __restrict__ llast, float* __restrict__ out,
                                      const int32_t* __restrict__ count_dev, int count_host, int back) {
  const int count = count_dev ? *count_dev : count_host;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)count * A) return;
  int bi = i / A, a = i % A;
  int row = back ? G - 1 - bi : bi;
  uint64_t wx = lx[(size_t)row * NW + (a >> 6)], wo = lo[(size_t)row * NW + (a >> 6)];
  bool isx = (wx >> (a & 63)) & 1, iso = (wo >> (a & 63)) & 1;
  bool xm = ltm[row] == 1;
  float4 v = make_float4((xm ? isx : iso) ? 1.f : 0.f, (xm ? iso : isx) ? 1.f : 0.f,
                         llast[row] == a ? 1.f : 0.f, 0.f);
  reinterpret_cast<float4*>(out)[(size_t)row * A + a] = v;
}

// ---- 3x3 convolution, padding 1, NHWC, implicit GEMM with shared-memory tiles ------------------
// out[row][px][co] = act( bias[co] + sum_{tap,ci} in[row][px+tap][ci] * w[tap][ci][co] (+ residual) )
template <int KC, int TN>
__global__ void __launch_bounds__(256) conv3x3_f32_kernel(
    const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ residual, float* __restrict__ out, int B, int CIN, int COUT, int relu,
    const int32_t* __restrict__ count_dev, int count_host, int back, int G) {
  constexpr int TM = 64, TNT = TN / 16;
  const int count = count_dev ? *count_dev : count_host;
  const int bi = blockIdx.z;
  if (bi >= count) return;
  const int row = back ? G - 1 - bi : bi;
  const int A = B * B;
  const int px0 = blockIdx.x * TM, co0 = blockIdx.y * TN;
  __shared__ float As[KC][TM + 4];
  __shared__ float Bs[KC][TN];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  float acc[4][TNT];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < TNT; j++) acc[i][j] = 0.f;
  const float* inb = in + (size_t)row * A * CIN;
  for (int tap = 0; tap < 9; tap++) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    for (int c0 = 0; c0 < CIN; c0 += KC) {
      for (int idx = tid; idx < TM * KC; idx += 256) {
        int k = idx % KC, m = idx / KC;
        int px = px0 + m;
        float v = 0.f;
        if (px < A) {
          int r = px / B + dy, c = px % B + dx;
          if (r >= 0 && r < B && c >= 0 && c < B) v = inb[(size_t)(r * B + c) * CIN + c0 + k];
        }
        As[k][m] = v;
      }
      for (int idx = tid; idx < KC * TN; idx += 256) {
        int k = idx / TN, nn = idx % TN;
        Bs[k][nn] = w[((size_t)tap * CIN + c0 + k) * COUT + co0 + nn];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KC; k++) {
        float a[4], b[TNT];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < TNT; j++) b[j] = Bs[k][tx * TNT + j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < TNT; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int px = px0 + ty * 4 + i;
    if (px >= A) continue;
#pragma unroll
    for (int j = 0; j < TNT; j++) {
      int co = co0 + tx * TNT + j;
      size_t o = ((size_t)row * A + px) * COUT + co;
      float v = acc[i][j] + bias[co];
      if (residual) v += residual[o];
      if (relu) v = fmaxf(v, 0.f);
      out[o] = v;
    }
  }
}

// ---- heads: 1x1 convs + FCs + softmax / tanh, one block per board --------------------------------
// feat layout is the reference's NCHW flatten (net.py:64,69): index = ch*A + pixel.
__global__ void __launch_bounds__(256) heads_f32_kernel(
    const float* __restrict__ act, int C, int B, int PC, int VC,
    const float* __restrict__ head_w /*[PC+VC][C]*/, const float* __restrict__ head_b,
    const float* __restrict__ pfc_t /*[PC*A][A]*/, const float* __restrict__ pfc_b,
    const float* __restrict__ vfc1_t /*[VC*A][64]*/, const float* __restrict__ vfc1_b,
    const float* __restrict__ vfc2_w, const float* __restrict__ vfc2_b,
    float* __restrict__ logits_out, float* __restrict__ policy_out, float* __restrict__ value_out,
    const int32_t* __restrict__ count_dev, int count_host, int back, int G, int out_by_index) {
  const int count = count_dev ? *count_dev : count_host;
  const int bi = blockIdx.x;
  if (bi >= count) return;
  const int row = back ? G - 1 - bi : bi;
  const int orow = out_by_index ? bi : row;
  const int A = B * B, NJ = PC + VC;
  extern __shared__ float sm[];
  float* feat = sm;                 // [NJ*A]  policy channels first, then value channels
  float* lg = feat + NJ * A;        // [A]
  float* hid = lg + A;              // [64]
  float* red = hid + 64;            // [32]
  const int tid = threadIdx.x;
  const float* ab = act + (size_t)row * A * C;
  for (int t = tid; t < A * NJ; t += 256) {
    int px = t / NJ, j = t % NJ;
    const float* ar = ab + (size_t)px * C;
    const float* wr = head_w + (size_t)j * C;
    float s = 0.f;
    for (int c = 0; c < C; c += 4) {
      float4 a4 = *reinterpret_cast<const float4*>(ar + c);
      float4 w4 = *reinterpret_cast<const float4*>(wr + c);
      s = fmaf(a4.x, w4.x, s); s = fmaf(a4.y, w4.y, s); s = fmaf(a4.z, w4.z, s); s = fmaf(a4.w, w4.w, s);
    }
    feat[j * A + px] = fmaxf(s + head_b[j], 0.f);
  }
  __syncthreads();
  // policy logits
  for (int a = tid; a < A; a += 256) {
    float s = pfc_b[a];
    const int K = PC * A;
    for (int k = 0; k < K; k++) s = fmaf(feat[k], pfc_t[(size_t)k * A + a], s);
    lg[a] = s;
  }
  // value hidden layer
  if (tid < 64) {
    float s = vfc1_b[tid];
    const int K = VC * A;
    const float* vf = feat + PC * A;
    for (int k = 0; k < K; k++) s = fmaf(vf[k], vfc1_t[(size_t)k * 64 + tid], s);
    hid[tid] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (tid < 32) {
    float s = hid[tid] * vfc2_w[tid] + hid[tid + 32] * vfc2_w[tid + 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(AZP_FULL, s, o);
    if (tid == 0 && value_out) value_out[orow] = tanhf(s + vfc2_b[0]);
  }
  // softmax over ALL cells (controller.py:49)
  float m = -INFINITY;
  for (int a = tid; a < A; a += 256) m = fmaxf(m, lg[a]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(AZP_FULL, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; i++) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int a = tid; a < A; a += 256) s += expf(lg[a] - m);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(AZP_FULL, s, o);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += red[i];
  for (int a = tid; a < A; a += 256) {
    if (logits_out) logits_out[(size_t)orow * A + a] = lg[a];
    if (policy_out) policy_out[(size_t)orow * A + a] = expf(lg[a] - m) / s;
  }
}

// ================= weight ingestion ================================================================
static void fold_bn(std::vector<float>& w, std::vector<float>& b, int cout, int per_out, const float* gamma,
                    const float* beta, const float* mean, const float* var) {
  for (int co = 0; co < cout; co++) {
    float sc = gamma[co] / sqrtf(var[co] + 1e-5f);
    for (int i = 0; i < per_out; i++) w[(size_t)co * per_out + i] *= sc;
    b[co] = (b[co] - mean[co]) * sc + beta[co];
  }
}

size_t azp_net_weight_count(const azp_engine* e) {
  const size_t A = e->A;
  if (e->cfg.net_kind == AZP_NET_STEM)
    return 32 * 36 + 32 + 64 * 288 + 64 + 128 * 576 + 128 + 4 * 128 + 4 + A * 4 * A + A + 2 * 128 + 2 +
           64 * 2 * A + 64 + 64 + 1;
  const size_t C = e->cfg.tower_channels, NB = e->cfg.tower_blocks;
  size_t n = C * 36 + C + 4 * C;                       // conv + bn
  n += NB * 2 * (C * C * 9 + 4 * C);                   // res blocks: conv (no bias) + bn
  n += 2 * C + 2 + 4 * 2 + A * 2 * A + A;              // policy_conv, policy_bn, policy_fc
  n += C + 1 + 4 * 1 + 64 * A + 64 + 64 + 1;           // value_conv, value_bn, value_fc1, value_fc2
  return n;
}

// Parse the blob (state_dict key order) into logical layers with BN folded in.
int azp_net_parse(const azp_engine* e, const float* blob, size_t n, HostNet& net) {
  if (n != azp_net_weight_count(e)) return AZP_ERR_INVALID_ARG;
  const int A = e->A;
  const float* p = blob;
  auto take = [&](size_t cnt) { const float* q = p; p += cnt; return q; };
  auto conv = [&](int cin, int cout, bool has_bias, bool relu, int residual) {
    HostConv c; c.cin = cin; c.cout = cout; c.relu = relu; c.residual = residual;
    const float* w = take((size_t)cout * cin * 9);
    c.w.assign(w, w + (size_t)cout * cin * 9);
    if (has_bias) { const float* b = take(cout); c.b.assign(b, b + cout); } else c.b.assign(cout, 0.f);
    return c;
  };
  net.convs.clear();
  if (e->cfg.net_kind == AZP_NET_STEM) {
    net.convs.push_back(conv(4, 32, true, true, 0));
    net.convs.push_back(conv(32, 64, true, true, 0));
    net.convs.push_back(conv(64, 128, true, true, 0));
    net.C = 128; net.PC = 4; net.VC = 2;
    const float* pw = take(4 * 128); const float* pb = take(4);
    const float* fw = take((size_t)A * 4 * A); const float* fb = take(A);
    const float* vw = take(2 * 128); const float* vb = take(2);
    const float* v1w = take((size_t)64 * 2 * A); const float* v1b = take(64);
    const float* v2w = take(64); const float* v2b = take(1);
    net.head_w.assign(pw, pw + 4 * 128); net.head_w.insert(net.head_w.end(), vw, vw + 2 * 128);
    net.head_b.assign(pb, pb + 4); net.head_b.insert(net.head_b.end(), vb, vb + 2);
    net.pfc_w.assign(fw, fw + (size_t)A * 4 * A); net.pfc_b.assign(fb, fb + A);
    net.vfc1_w.assign(v1w, v1w + (size_t)64 * 2 * A); net.vfc1_b.assign(v1b, v1b + 64);
    net.vfc2_w.assign(v2w, v2w + 64); net.vfc2_b.assign(v2b, v2b + 1);
  } else {
    const int C = e->cfg.tower_channels, NB = e->cfg.tower_blocks;
    // key order of the old checkpoints: conv.weight, conv.bias, bn.{weight,bias,running_mean,running_var}
    HostConv c0 = conv(4, C, true, true, 0);
    { const float* g = take(C); const float* b = take(C); const float* m = take(C); const float* v = take(C);
      fold_bn(c0.w, c0.b, C, 36, g, b, m, v); }
    net.convs.push_back(c0);
    for (int i = 0; i < NB; i++) {
      for (int j = 0; j < 2; j++) {
        HostConv c = conv(C, C, false, j == 0, j == 1 ? 1 : 0);
        if (j == 1) c.relu = true;                     // relu(bn2(conv2(y)) + x)
        const float* g = take(C); const float* b = take(C); const float* m = take(C); const float* v = take(C);
        fold_bn(c.w, c.b, C, C * 9, g, b, m, v);
        net.convs.push_back(c);
      }
    }
    net.C = C; net.PC = 2; net.VC = 1;
    const float* pwp = take((size_t)2 * C); const float* pbp = take(2);
    std::vector<float> pw(pwp, pwp + (size_t)2 * C), pb(pbp, pbp + 2);
    { const float* g = take(2); const float* b = take(2); const float* m = take(2); const float* v = take(2);
      fold_bn(pw, pb, 2, C, g, b, m, v); }
    const float* fw = take((size_t)A * 2 * A); const float* fb = take(A);
    const float* vwp = take(C); const float* vbp = take(1);
    std::vector<float> vw(vwp, vwp + C), vb(vbp, vbp + 1);
    { const float* g = take(1); const float* b = take(1); const float* m = take(1); const float* v = take(1);
      fold_bn(vw, vb, 1, C, g, b, m, v); }
    const float* v1w = take((size_t)64 * A); const float* v1b = take(64);
    const float* v2w = take(64); const float* v2b = take(1);
    net.head_w = pw; net.head_w.insert(net.head_w.end(), vw.begin(), vw.end());
    net.head_b = pb; net.head_b.insert(net.head_b.end(), vb.begin(), vb.end());
    net.pfc_w.assign(fw, fw + (size_t)A * 2 * A); net.pfc_b.assign(fb, fb + A);
    net.vfc1_w.assign(v1w, v1w + (size_t)64 * A); net.vfc1_b.assign(v1b, v1b + 64);
    net.vfc2_w.assign(v2w, v2w + 64); net.vfc2_b.assign(v2b, v2b + 1);
  }
  if ((size_t)(p - blob) != n) return AZP_ERR_INVALID_ARG;
  return AZP_OK;
}

template <typename T>
static int upload(azp_engine* e, T** dst, const std::vector<T>& src) {
  int st = e->alloc(dst, src.size());
  if (st) return st;
  AZP_CUDA(e, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return AZP_OK;
}

static int build_f32(azp_engine* e, const HostNet& net, NetWeights* W) {
  const int A = e->A;
  W->layers.clear();
  for (const HostConv& c : net.convs) {
    DevConv d; d.cin = c.cin; d.cout = c.cout; d.relu = c.relu; d.residual = c.residual;
    std::vector<float> w((size_t)9 * c.cin * c.cout);       // [tap][ci][co] from [co][ci][ky][kx]
    for (int co = 0; co < c.cout; co++)
      for (int ci = 0; ci < c.cin; ci++)
        for (int t = 0; t < 9; t++) w[((size_t)t * c.cin + ci) * c.cout + co] = c.w[((size_t)co * c.cin + ci) * 9 + t];
    int st = upload(e, &d.w, w); if (st) return st;
    st = upload(e, &d.b, c.b); if (st) return st;
    W->layers.push_back(d);
  }
  int st;
  if ((st = upload(e, &W->head_w, net.head_w))) return st;
  if ((st = upload(e, &W->head_b, net.head_b))) return st;
  const int PK = net.PC * A, VK = net.VC * A;
  std::vector<float> pt((size_t)PK * A), vt((size_t)VK * 64);
  for (int a = 0; a < A; a++) for (int k = 0; k < PK; k++) pt[(size_t)k * A + a] = net.pfc_w[(size_t)a * PK + k];
  for (int j = 0; j < 64; j++) for (int k = 0; k < VK; k++) vt[(size_t)k * 64 + j] = net.vfc1_w[(size_t)j * VK + k];
  if ((st = upload(e, &W->pfc_t, pt))) return st;
  if ((st = upload(e, &W->pfc_b, net.pfc_b))) return st;
  if ((st = upload(e, &W->vfc1_t, vt))) return st;
  if ((st = upload(e, &W->vfc1_b, net.vfc1_b))) return st;
  if ((st = upload(e, &W->vfc2_w, net.vfc2_w))) return st;
  if ((st = upload(e, &W->vfc2_b, net.vfc2_b))) return st;
  W->C = net.C; W->PC = net.PC; W->VC = net.VC;
  size_t best = 0;
  for (size_t i = 0; i < net.convs.size(); i++)
    if ((size_t)net.convs[i].cin * net.convs[i].cout >= (size_t)net.convs[best].cin * net.convs[best].cout) best = i;
  W->dominant_layer = best;
  return AZP_OK;
}

int azp_net_create(azp_engine* e) {
  e->ws = new NetWorkspace();
  NetWorkspace* ws = e->ws;
  const size_t G = e->slots, A = e->A;
  int cmax = e->cfg.net_kind == AZP_NET_STEM ? 128 : e->cfg.tower_channels;
  ws->cmax = cmax;
  int st;
  if (e->cfg.net_dtype == AZP_DTYPE_F32) {
    if ((st = e->alloc(&ws->in4, G * A * 4))) return st;
    for (int i = 0; i < 3; i++) if ((st = e->alloc(&ws->act[i], G * A * cmax))) return st;
  } else {
    if ((st = azp_umma_create(e))) return st;
  }
  return AZP_OK;
}

void azp_net_destroy(azp_engine* e) {
  for (int i = 0; i < 2; i++) { delete e->weights[i]; e->weights[i] = nullptr; }
  delete e->ws; e->ws = nullptr;
}

int azp_net_load(azp_engine* e, int slot, const float* blob, size_t n, bool on_device, cudaStream_t s) {
  if (slot < 0 || slot > 1) { e->set_error("weight slot must be 0 or 1"); return AZP_ERR_INVALID_ARG; }
  if (n != azp_net_weight_count(e)) {
    e->set_error("weight blob has " + std::to_string(n) + " floats, expected " + std::to_string(azp_net_weight_count(e)));
    return AZP_ERR_INVALID_ARG;
  }
  if (on_device && e->cfg.net_dtype == AZP_DTYPE_BF16 && e->weights[slot] && e->weights[slot]->loaded) {
    // refresh in place on the device: order the repack after every stream that may still run kernels on the old
    // weights (events, no host wait), then enqueue it on `s`
    for (cudaStream_t q : e->net_streams) {
      if (q == s) continue;
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); continue; }
      if (cudaEventRecord(ev, q) == cudaSuccess) cudaStreamWaitEvent(s, ev, 0); else cudaGetLastError();
      cudaEventDestroy(ev);
    }
    const int rst = azp_umma_reload_device(e, e->weights[slot], blob, s);
    if (rst == AZP_OK) {
      if (!e->reload_ev && cudaEventCreateWithFlags(&e->reload_ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError(); e->reload_ev = nullptr;
        AZP_CUDA(e, cudaStreamSynchronize(s));          // no event to order other streams with: finish the refresh now
      } else {
        AZP_CUDA(e, cudaEventRecord(e->reload_ev, s));
      }
      e->reload_stream = s;
      e->note_stream(s);
      return AZP_OK;
    }
    if (rst != AZP_ERR_UNSUPPORTED) return rst;
  }
  std::vector<float> host(n);
  if (on_device) {
    AZP_CUDA(e, cudaStreamSynchronize(s));
    AZP_CUDA(e, cudaMemcpy(host.data(), blob, n * sizeof(float), cudaMemcpyDeviceToHost));
  } else {
    memcpy(host.data(), blob, n * sizeof(float));
  }
  HostNet net;
  int st = azp_net_parse(e, host.data(), n, net);
  if (st) { e->set_error("weight blob does not match the configured network"); return st; }
  // no kernel may still read the weights being replaced: wait for every stream network kernels were launched on
  for (cudaStream_t q : e->net_streams)
    if (cudaStreamSynchronize(q) != cudaSuccess) cudaGetLastError();          // a stream destroyed by its owner has no work left
  AZP_CUDA(e, cudaStreamSynchronize(s));
  // A reload of the same network reuses the previous set's device buffers in allocation order: the kernels'
  // pointer arguments (and with them the captured step graph and the TMA descriptors) stay valid.
  std::vector<void*> old_allocs;
  if (e->weights[slot]) { old_allocs.swap(e->weights[slot]->allocs); delete e->weights[slot]; }
  e->weights[slot] = new NetWeights();
  e->weights[slot]->allocs.swap(old_allocs);
  const size_t reused = e->weights[slot]->allocs.size();
  e->alloc_pool = &e->weights[slot]->allocs;
  e->alloc_reuse = reused > 0; e->alloc_cursor = 0;
  if (e->cfg.net_dtype == AZP_DTYPE_F32) st = build_f32(e, net, e->weights[slot]);
  else st = azp_umma_build(e, net, e->weights[slot]);
  const bool same_layout = reused == 0 || (e->alloc_cursor == reused && e->weights[slot]->allocs.size() == reused);
  e->alloc_pool = nullptr; e->alloc_reuse = false;
  if (!same_layout) e->weights_epoch++;       // addresses changed: the captured graph must be rebuilt
  if (st) { delete e->weights[slot]; e->weights[slot] = nullptr; return st; }
  e->weights[slot]->loaded = true;
  if (reused == 0) e->weights_epoch++;       // first load: (re)capture the simulation-step graph
  return AZP_OK;
}

// ---- forward -----------------------------------------------------------------------------------------
static int run_f32(azp_engine* e, NetWeights* W, const float* in4, const int32_t* count_dev, int count_host,
                   int back, float* logits, float* policy, float* value, int out_by_index, cudaStream_t s) {
  NetWorkspace* ws = e->ws;
  const int B = e->cfg.board, A = e->A, G = e->slots;
  const int zmax = count_dev ? G : count_host;
  if (zmax <= 0) return AZP_OK;
  // activation buffers rotate through ws->act[0..2]; a tower block needs three: x (skip), y, out
  const float* cur = in4;
  int xi = -1, yi = -1;
  for (size_t li = 0; li < W->layers.size(); li++) {
    const DevConv& L = W->layers[li];
    const bool tower_block = e->cfg.net_kind == AZP_NET_TOWER && li >= 1;
    const bool second = tower_block && ((li - 1) % 2 == 1);
    int out_idx;
    const float* residual = nullptr;
    if (!tower_block) { out_idx = (xi + 1) % 3; }
    else if (!second) { out_idx = (xi + 1) % 3; yi = out_idx; cur = ws->act[xi]; }
    else { out_idx = (xi + 2) % 3; cur = ws->act[yi]; residual = ws->act[xi]; }
    float* out = ws->act[out_idx];
    dim3 grid((A + 63) / 64, 1, zmax);
    const bool dominant = li == W->dominant_layer;
    if (dominant) {
      e->prof_flops_per_board = 2.0 * 9.0 * L.cin * L.cout * A;
      azp_prof_begin(e, AZP_PROF_DOMINANT, count_dev, count_host, s);
    }
    if (L.cin % 16 == 0 && L.cout % 64 == 0) {
      grid.y = L.cout / 64;
      conv3x3_f32_kernel<16, 64><<<grid, 256, 0, s>>>(cur, L.w, L.b, residual, out, B, L.cin, L.cout, L.relu,
                                                      count_dev, count_host, back, G);
    } else if (L.cin % 16 == 0 && L.cout % 32 == 0) {
      grid.y = L.cout / 32;
      conv3x3_f32_kernel<16, 32><<<grid, 256, 0, s>>>(cur, L.w, L.b, residual, out, B, L.cin, L.cout, L.relu,
                                                      count_dev, count_host, back, G);
    } else if (L.cin % 4 == 0 && L.cout % 64 == 0) {
      grid.y = L.cout / 64;
      conv3x3_f32_kernel<4, 64><<<grid, 256, 0, s>>>(cur, L.w, L.b, residual, out, B, L.cin, L.cout, L.relu,
                                                     count_dev, count_host, back, G);
    } else if (L.cin % 4 == 0 && L.cout % 32 == 0) {
      grid.y = L.cout / 32;
      conv3x3_f32_kernel<4, 32><<<grid, 256, 0, s>>>(cur, L.w, L.b, residual, out, B, L.cin, L.cout, L.relu,
                                                     count_dev, count_host, back, G);
    } else {
      e->set_error("unsupported conv shape");
      return AZP_ERR_UNSUPPORTED;
    }
    azp_count_launch();
    if (dominant) azp_prof_end(e, AZP_PROF_DOMINANT, s);
    if (!tower_block || second) { xi = out_idx; cur = out; }
  }
  const int NJ = W->PC + W->VC;
  size_t smem = ((size_t)NJ * A + A + 64 + 32) * sizeof(float);
  heads_f32_kernel<<<zmax, 256, smem, s>>>(cur, W->C, B, W->PC, W->VC, W->head_w, W->head_b, W->pfc_t, W->pfc_b,
                                           W->vfc1_t, W->vfc1_b, W->vfc2_w, W->vfc2_b, logits, policy, value,
                                           count_dev, count_host, back, G, out_by_index);
  azp_count_launch();
  return azp_check_launch();
}

int azp_net_eval_leaves(azp_engine* e, int slot, const int32_t* count_dev, int count_host, int back, cudaStream_t s) {
  e->note_stream(s);
  NetWeights* W = e->weights[slot];
  if (!W || !W->loaded) { e->set_error("no weights loaded in slot " + std::to_string(slot)); return AZP_ERR_NO_WEIGHTS; }
  if (e->cfg.net_dtype == AZP_DTYPE_BF16) return azp_umma_eval_leaves(e, W, count_dev, count_host, back, s);
  const int A = e->A, G = e->slots;
  const int zmax = count_dev ? G : count_host;
  if (zmax <= 0) return AZP_OK;
  leaves_to_nhwc_kernel<<<(int)(((size_t)zmax * A + 255) / 256), 256, 0, s>>>(
      A, e->NW, G, e->tv.leaf_x, e->tv.leaf_o, e->tv.leaf_tm, e->tv.leaf_last, e->ws->in4, count_dev, count_host, back);
  azp_count_launch();
  return run_f32(e, W, e->ws->in4, count_dev, count_host, back, e->d_logits, e->d_policy, e->d_value, 0, s);
}

int azp_net_eval_planes(azp_engine* e, int slot, int n, const float* planes, float* logits, float* policy,
                        float* value, cudaStream_t s) {
  e->note_stream(s);
  NetWeights* W = e->weights[slot];
  if (!W || !W->loaded) { e->set_error("no weights loaded in slot " + std::to_string(slot)); return AZP_ERR_NO_WEIGHTS; }
  const int A = e->A, G = e->slots;
  for (int done = 0; done < n; done += G) {
    int m = n - done < G ? n - done : G;
    const float* pin = planes + (size_t)done * 4 * A;
    float* lo = logits ? logits + (size_t)done * A : nullptr;
    float* po = policy ? policy + (size_t)done * A : nullptr;
    float* vo = value ? value + done : nullptr;
    int st;
    if (e->cfg.net_dtype == AZP_DTYPE_BF16) {
      st = azp_umma_eval_planes(e, W, m, pin, lo, po, vo, s);
    } else {
      planes_to_nhwc_kernel<<<(int)(((size_t)m * A + 255) / 256), 256, 0, s>>>(A, m, pin, e->ws->in4);
      azp_count_launch();
      st = run_f32(e, W, e->ws->in4, nullptr, m, 0, lo, po, vo, 1, s);
    }
    if (st) return st;
  }
  return AZP_OK;
}
