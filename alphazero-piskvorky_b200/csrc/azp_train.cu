// Training step of the stem net on GPU-resident examples.
//
// Replaces NeuralNetworkController.train_step (alphazero/controller.py:82-131):
//   forward of GomokuNet (alphazero/net.py:55-72) with the activations kept,
//   loss = cross_entropy(logits, pi) with probability targets + mse(value, z) (controller.py:118-123),
//   backward of every parameter, and -- as a separate stateless entry point -- one AdamW update
//   (torch.optim.AdamW, weight_decay 1e-4, lr set per step: controller.py:75-79, 114-115, 125).
//
// Arithmetic is fp32 throughout (the reference trains in fp32): every contraction -- the three 3x3
// convolutions as implicit-im2col GEMMs (fprop, dgrad, wgrad), the 1x1 head convolutions and the FC
// layers -- runs through ONE hand-written register-tiled SIMT SGEMM (sgemm_kernel) with strided operand
// views, split-K for the weight gradients (K = batch*B*B positions) and a deterministic partial
// reduction (no atomics: two runs of a step give bit-identical gradients).  Activations are "NHWC rows"
// (one row of C channels per board position); the NCHW flattens of the heads (net.py:63, 68) are explicit
// transposes.  All buffers of the ABI are caller-owned device memory; the trainer object owns only
// its activation / im2col / split-K workspaces.
#include "azp_internal.cuh"
#include "../../include/azp_b200.h"

namespace {

// ---- SGEMM ---------------------------------------------------------------------------------------------
// C[M][N] (row-major) = sum_k A(i,k) * B(k,j),  A(i,k) = A[i*sAi + k*sAk],  B(k,j) = B[k*sBk + j*sBj].
// A_KC / B_NC only choose which index runs along consecutive threads of the tile loaders (the contiguous
// one, for coalescing); the strides are general.  blockIdx.z = split-K slice (kChunk wide) writing its own
// [M][N] partial.  Epilogues: 0 none, 1 +bias[j], 2 relu(+bias[j]), 3 * (mask[i][j] > 0)  (ReLU backward).
struct GemmArgs {
  const float* A; const float* B; float* C;
  int M, N, K;
  long long sAi, sAk, sBk, sBj;
  int kChunk;
  int epi;
  const float* bias;
  const float* mask;
};

template <int BM, int BN, int TM, int TN, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(GemmArgs g) {
  constexpr int BK = 16;
  constexpr int NA = BM * BK / 256, NB = BN * BK / 256;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads per tile");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 fragments");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int t = threadIdx.x;
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  const int kBegin = blockIdx.z * g.kChunk;
  const int kEnd = min(g.K, kBegin + g.kChunk);
  float* __restrict__ C = g.C + (size_t)blockIdx.z * (size_t)g.M * (size_t)g.N;

  float ra[NA], rb[NB];
  auto a_coord = [&](int r, int& ii, int& kk) {
    if (A_KC) { kk = t % BK; ii = t / BK + (256 / BK) * r; }
    else { ii = t % BM; kk = t / BM + (256 / BM) * r; }
  };
  auto b_coord = [&](int r, int& jj, int& kk) {
    if (B_NC) { jj = t % BN; kk = t / BN + (256 / BN) * r; }
    else { kk = t % BK; jj = t / BK + (256 / BK) * r; }
  };
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int r = 0; r < NA; r++) {
      int ii, kk; a_coord(r, ii, kk);
      const int i = i0 + ii, k = k0 + kk;
      ra[r] = (i < g.M && k < kEnd) ? __ldg(g.A + (long long)i * g.sAi + (long long)k * g.sAk) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < NB; r++) {
      int jj, kk; b_coord(r, jj, kk);
      const int j = j0 + jj, k = k0 + kk;
      rb[r] = (j < g.N && k < kEnd) ? __ldg(g.B + (long long)k * g.sBk + (long long)j * g.sBj) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int r = 0; r < NA; r++) { int ii, kk; a_coord(r, ii, kk); As[buf][kk][ii] = ra[r]; }
#pragma unroll
    for (int r = 0; r < NB; r++) { int jj, kk; b_coord(r, jj, kk); Bs[buf][kk][jj] = rb[r]; }
  };

  const int tx = t % (BN / TN), ty = t / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int r = 0; r < TM; r++)
#pragma unroll
    for (int c = 0; c < TN; c++) acc[r][c] = 0.f;

  if (kBegin < kEnd) {
    load_tiles(kBegin);
    store_tiles(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kBegin; k0 < kEnd; k0 += BK) {
    const bool more = k0 + BK < kEnd;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < TM; r += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + r]);
        a[r] = v.x; a[r + 1] = v.y; a[r + 2] = v.z; a[r + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < TN; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + c]);
        b[c] = v.x; b[c + 1] = v.y; b[c + 2] = v.z; b[c + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < TM; r++)
#pragma unroll
        for (int c = 0; c < TN; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

#pragma unroll
  for (int r = 0; r < TM; r++) {
    const int i = i0 + ty * TM + r;
    if (i >= g.M) continue;
#pragma unroll
    for (int c = 0; c < TN; c++) {
      const int j = j0 + tx * TN + c;
      if (j >= g.N) continue;
      float v = acc[r][c];
      const size_t o = (size_t)i * g.N + j;
      if (g.epi == 1) v += g.bias[j];
      else if (g.epi == 2) v = fmaxf(v + g.bias[j], 0.f);
      else if (g.epi == 3) v = g.mask[o] > 0.f ? v : 0.f;
      C[o] = v;
    }
  }
}

// out[idx] = sum_s partial[s*stride + src(idx)], s ascending (deterministic).  conv_cin > 0: the partial is a
// conv weight gradient laid out [(tap*Cin+ci)][Cout] and `out` is PyTorch's [Cout][Cin][3][3].
__global__ void reduce_splits_kernel(float* __restrict__ out, const float* __restrict__ partial, int n, long long stride,
                                     int splits, int conv_cin, int conv_cout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int src = idx;
  if (conv_cin > 0) {
    const int tap = idx % 9, ci = (idx / 9) % conv_cin, co = idx / (9 * conv_cin);
    src = (tap * conv_cin + ci) * conv_cout + co;
  }
  float s = 0.f;
  for (int k = 0; k < splits; k++) s += partial[(long long)k * stride + src];
  out[idx] = s;
}

// PyTorch conv weight W[Cout][Cin][3][3] -> fprop operand Wm[(tap*Cin+ci)][Cout] and dgrad operand
// Wd[(tap*Cout+co)][Cin] (the tap flip of the transposed convolution is done by the im2col gather).
__global__ void pack_conv_weights_kernel(const float* __restrict__ W, int cin, int cout, float* __restrict__ Wm,
                                         float* __restrict__ Wd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * cin * 9) return;
  const int tap = idx % 9, ci = (idx / 9) % cin, co = idx / (9 * cin);
  const float v = W[idx];
  Wm[(tap * cin + ci) * cout + co] = v;
  if (Wd) Wd[(tap * cout + co) * cin + ci] = v;
}

// col[(img*A + p)][tap*C + ci] = X(img, p + offset(tap or 8-tap), ci), zero outside the board.
// X element (img, pos, ci) lives at img*sImg + pos*sPos + ci*sC (NHWC rows or the NCHW input planes).
__global__ void im2col_kernel(const float* __restrict__ X, float* __restrict__ col, long long total, int B, int C,
                              long long sImg, long long sPos, long long sC, int flip) {
  const int A = B * B;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % C);
    long long q = idx / C;
    int tap = (int)(q % 9); q /= 9;
    const int p = (int)(q % A);
    const long long img = q / A;
    if (flip) tap = 8 - tap;
    const int r = p / B + tap / 3 - 1, c = p % B + tap % 3 - 1;
    float v = 0.f;
    if (r >= 0 && r < B && c >= 0 && c < B) v = __ldg(X + img * sImg + (long long)(r * B + c) * sPos + ci * sC);
    col[idx] = v;
  }
}

// Same gather for NHWC rows with C = 4 << log2c4 channels: one thread moves one float4 (16 B loads and stores, 32-bit
// index arithmetic, the channel split is a shift) -- the scalar kernel above spent 4.4 ms of a 12 ms step on 64-bit
// divisions.
__global__ void im2col_vec4_kernel(const float4* __restrict__ X, float4* __restrict__ col, unsigned units, int B, int A,
                                   int log2c4, int flip) {
  const unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const unsigned c4 = u & ((1u << log2c4) - 1u);
  const unsigned q = u >> log2c4;
  unsigned tap = q % 9u;
  const unsigned row = q / 9u;
  const unsigned img = row / (unsigned)A;
  const int p = (int)(row - img * (unsigned)A);
  if (flip) tap = 8u - tap;
  const int pr = p / B;
  const int r = pr + (int)(tap / 3u) - 1, c = p - pr * B + (int)(tap % 3u) - 1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= 0 && r < B && c >= 0 && c < B) v = __ldg(X + (((size_t)img * A + (size_t)(r * B + c)) << log2c4) + c4);
  col[u] = v;
}

// heads [n*A][6] (4 policy + 2 value channels, post-ReLU) -> pt [n][4*A], vt [n][2*A] in NCHW-flatten order
__global__ void heads_to_nchw_kernel(const float* __restrict__ hh, int n, int A, float* __restrict__ pt,
                                     float* __restrict__ vt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * A * 6) return;
  const int c = (int)(idx % 6);
  const long long row = idx / 6;
  const int p = (int)(row % A);
  const long long img = row / A;
  const float v = hh[idx];
  if (c < 4) pt[img * 4 * A + (long long)c * A + p] = v;
  else vt[img * 2 * A + (long long)(c - 4) * A + p] = v;
}

// inverse gather with the ReLU mask of the head activations: dhh = d(pt|vt) * (hh > 0)
__global__ void heads_grad_from_nchw_kernel(const float* __restrict__ dpt, const float* __restrict__ dvt,
                                            const float* __restrict__ hh, int n, int A, float* __restrict__ dhh) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * A * 6) return;
  const int c = (int)(idx % 6);
  const long long row = idx / 6;
  const int p = (int)(row % A);
  const long long img = row / A;
  const float d = c < 4 ? dpt[img * 4 * A + (long long)c * A + p] : dvt[img * 2 * A + (long long)(c - 4) * A + p];
  dhh[idx] = hh[idx] > 0.f ? d : 0.f;
}

// Wh[6][128] = {policy_conv.weight [4][128]; value_conv.weight [2][128]}, bh[6] likewise
__global__ void pack_heads_kernel(const float* __restrict__ pw, const float* __restrict__ pb, const float* __restrict__ vw,
                                  const float* __restrict__ vb, float* __restrict__ Wh, float* __restrict__ bh) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 512) Wh[idx] = pw[idx];
  else if (idx < 768) Wh[idx] = vw[idx - 512];
  if (idx < 4) bh[idx] = pb[idx];
  else if (idx < 6) bh[idx] = vb[idx - 4];
}

// column sums of X[M][N] (bias gradients), two deterministic stages
__global__ void colsum_stage1_kernel(const float* __restrict__ X, int M, int N, int rows_per_block,
                                     float* __restrict__ partial) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int j = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (j < N)
    for (int r = r0 + ty; r < r1; r += 8) s += X[(size_t)r * N + j];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < N) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) tot += sm[k][tx];
    partial[(size_t)blockIdx.y * N + j] = tot;
  }
}
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int blocks, int N, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s = 0.f;
  for (int b = 0; b < blocks; b++) s += partial[(size_t)b * N + j];
  out[j] = s;
}

// One warp per example: log-softmax over ALL cells, policy loss -sum(pi*logp) (F.cross_entropy with probability
// targets, controller.py:122), predicted-policy entropy, value = tanh(pre), value loss (v-z)^2 (controller.py:118);
// gradients w.r.t. the logits and the pre-tanh value already carry the 1/n of the batch means.
__global__ void loss_grad_kernel(const float* __restrict__ logits, const float* __restrict__ pi, const float* __restrict__ vpre,
                                 const float* __restrict__ z, int n, int A, float* __restrict__ dlogits,
                                 float* __restrict__ dvpre, float* __restrict__ rowloss) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= n) return;
  const float* l = logits + (size_t)row * A;
  const float* q = pi + (size_t)row * A;
  float m = -INFINITY, sq = 0.f;
  for (int c = lane; c < A; c += 32) { m = fmaxf(m, l[c]); sq += q[c]; }
  for (int o = 16; o; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  float se = 0.f;
  for (int c = lane; c < A; c += 32) se += expf(l[c] - m);
  for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  const float lse = m + logf(se);
  const float inv_n = 1.f / (float)n;
  float pl = 0.f, ent = 0.f;
  for (int c = lane; c < A; c += 32) {
    const float logp = l[c] - lse;
    const float p = expf(logp);
    pl -= q[c] * logp;
    ent -= p * logp;
    if (dlogits) dlogits[(size_t)row * A + c] = (p * sq - q[c]) * inv_n;
  }
  for (int o = 16; o; o >>= 1) {
    pl += __shfl_xor_sync(0xffffffffu, pl, o);
    ent += __shfl_xor_sync(0xffffffffu, ent, o);
  }
  if (lane == 0) {
    const float v = tanhf(vpre[row]);
    const float d = v - z[row];
    rowloss[row] = pl;
    rowloss[n + row] = d * d;
    rowloss[2 * n + row] = ent;
    if (dvpre) dvpre[row] = 2.f * d * inv_n * (1.f - v * v);
  }
}

// losses[0..3] = {policy loss, value loss, policy entropy, policy + value}: batch means, fixed summation order
__global__ void loss_reduce_kernel(const float* __restrict__ rowloss, int n, float* __restrict__ losses) {
  __shared__ float sm[3][256];
  const int t = threadIdx.x;
  float s[3] = {0.f, 0.f, 0.f};
  for (int r = t; r < n; r += 256)
    for (int k = 0; k < 3; k++) s[k] += rowloss[k * n + r];
  for (int k = 0; k < 3; k++) sm[k][t] = s[k];
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (t < o)
      for (int k = 0; k < 3; k++) sm[k][t] += sm[k][t + o];
    __syncthreads();
  }
  if (t == 0) {
    const float inv = 1.f / (float)n;
    losses[0] = sm[0][0] * inv; losses[1] = sm[1][0] * inv; losses[2] = sm[2][0] * inv;
    losses[3] = losses[0] + losses[1];
  }
}

// torch.optim.AdamW, single-tensor path, in its operation order:
//   p *= 1 - lr*wd;  m += (g - m)*(1-b1);  v = v*b2 + (1-b2)*g*g;
//   p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adamw_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, float decay, float w1, float b2, float w2, float step_size,
                             float bc2_sqrt, float eps) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = __fmul_rn(p[i], decay);
    const float mi = __fadd_rn(m[i], __fmul_rn(w1, __fsub_rn(gi, m[i])));
    const float vi = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(w2, gi), gi));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
    pi = __fadd_rn(pi, __fmul_rn(-step_size, __fdiv_rn(mi, denom)));
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace

// ---- trainer object ---------------------------------------------------------------------------------------
struct azp_trainer {
  int board = 0, A = 0, max_batch = 0, device = 0;
  std::string error;
  std::vector<void*> allocs;
  // parameter offsets in the state_dict blob (net.py:37-53 order)
  size_t o_c1w, o_c1b, o_c2w, o_c2b, o_c3w, o_c3b, o_pcw, o_pcb, o_pfw, o_pfb, o_vcw, o_vcb, o_v1w, o_v1b, o_v2w, o_v2b, n_params;
  // activations (NHWC rows) and their gradients
  float *a1, *a2, *a3, *hh, *pt, *vt, *logits, *v1, *v2;
  float *dz1, *dz2, *dz3, *dhh, *dpt, *dvt, *dlogits, *dv1, *dv2;
  float *col; size_t col_floats;
  float *W1m, *W2m, *W3m, *W2d, *W3d, *Wh, *bh, *dWh, *dbh;
  float *partial; size_t partial_floats;
  float *cs_partial; size_t cs_floats;
  float *rowloss;
  void set_error(const std::string& m) { error = m; azp_set_global_error(m); }
  template <typename T> int alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t err = cudaMalloc(&q, count * sizeof(T) + 256);
    if (err != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(err)); return AZP_ERR_CUDA; }
    allocs.push_back(q);
    *p = (T*)q;
    return AZP_OK;
  }
};

namespace {

constexpr int SM_TARGET_CTAS = 2 * 148;      // split-K aims at two CTAs per SM

template <int BM, int BN, int TM, int TN>
void launch_gemm_shape(const GemmArgs& g, bool a_kc, bool b_nc, int splits, cudaStream_t s) {
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, splits);
  if (a_kc && b_nc) sgemm_kernel<BM, BN, TM, TN, true, true><<<grid, 256, 0, s>>>(g);
  else if (a_kc && !b_nc) sgemm_kernel<BM, BN, TM, TN, true, false><<<grid, 256, 0, s>>>(g);
  else if (!a_kc && b_nc) sgemm_kernel<BM, BN, TM, TN, false, true><<<grid, 256, 0, s>>>(g);
  else sgemm_kernel<BM, BN, TM, TN, false, false><<<grid, 256, 0, s>>>(g);
  azp_count_launch();
}

// C = A*B with an epilogue (single K pass)
void gemm(cudaStream_t s, bool a_kc, bool b_nc, const float* A, long long sAi, long long sAk, const float* B,
          long long sBk, long long sBj, float* C, int M, int N, int K, int epi = 0, const float* bias = nullptr,
          const float* mask = nullptr) {
  GemmArgs g{A, B, C, M, N, K, sAi, sAk, sBk, sBj, K, epi, bias, mask};
  // the FC layers have few output tiles (policy_fc forward: 8 x 2 of 128 x 128): the narrow tile gives 4x the CTAs
  const long long wide_tiles = (long long)((M + 127) / 128) * ((N + 127) / 128);
  if (N <= 32 || wide_tiles < 74) launch_gemm_shape<128, 32, 4, 4>(g, a_kc, b_nc, 1, s);
  else if (N <= 64) launch_gemm_shape<128, 64, 8, 4>(g, a_kc, b_nc, 1, s);      // a 128-wide tile would waste half its FMAs
  else launch_gemm_shape<128, 128, 8, 8>(g, a_kc, b_nc, 1, s);
}

// weight gradient: out = A*B reduced over a long K, split over the grid; `out` receives the sum of the partials
// (conv_cin > 0: permuted into PyTorch's conv weight layout)
int gemm_splitk(azp_trainer* t, cudaStream_t s, bool a_kc, bool b_nc, const float* A, long long sAi, long long sAk,
                const float* B, long long sBk, long long sBj, float* out, int M, int N, int K, int conv_cin = 0,
                int conv_cout = 0) {
  const bool narrow = N <= 32, mid = !narrow && N <= 64;
  const int bn = narrow ? 32 : (mid ? 64 : 128);
  const int tiles = ((N + bn - 1) / bn) * ((M + 127) / 128);
  int splits = (SM_TARGET_CTAS + tiles - 1) / tiles;
  const int max_splits = (K + 63) / 64;            // at least four k-tiles per slice
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int chunk = (K + splits - 1) / splits;
  chunk = (chunk + 15) / 16 * 16;
  splits = (K + chunk - 1) / chunk;
  if ((size_t)splits * M * N > t->partial_floats) { t->set_error("split-K workspace too small"); return AZP_ERR_CAPACITY; }
  GemmArgs g{A, B, t->partial, M, N, K, sAi, sAk, sBk, sBj, chunk, 0, nullptr, nullptr};
  if (narrow) launch_gemm_shape<128, 32, 4, 4>(g, a_kc, b_nc, splits, s);
  else if (mid) launch_gemm_shape<128, 64, 8, 4>(g, a_kc, b_nc, splits, s);
  else launch_gemm_shape<128, 128, 8, 8>(g, a_kc, b_nc, splits, s);
  const int n = M * N;
  reduce_splits_kernel<<<(n + 255) / 256, 256, 0, s>>>(out, t->partial, n, (long long)M * N, splits, conv_cin, conv_cout);
  azp_count_launch();
  return AZP_OK;
}

void colsum(azp_trainer* t, cudaStream_t s, const float* X, int M, int N, float* out) {
  const int rpb = 1024;
  const int blocks = (M + rpb - 1) / rpb;
  colsum_stage1_kernel<<<dim3((N + 31) / 32, blocks), dim3(32, 8), 0, s>>>(X, M, N, rpb, t->cs_partial);
  colsum_stage2_kernel<<<(N + 127) / 128, 128, 0, s>>>(t->cs_partial, blocks, N, out);
  azp_count_launch(2);
}

void im2col(cudaStream_t s, const float* X, float* col, int n, int B, int C, bool nchw, bool flip) {
  const int A = B * B;
  const long long total = (long long)n * A * 9 * C;
  if (!nchw && (C == 32 || C == 64 || C == 128) && total / 4 < 0xFFFFFFFFll) {
    const unsigned units = (unsigned)(total / 4);
    const int log2c4 = C == 32 ? 3 : (C == 64 ? 4 : 5);
    im2col_vec4_kernel<<<(units + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(col),
                                                           units, B, A, log2c4, flip ? 1 : 0);
    azp_count_launch();
    return;
  }
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  const long long sImg = (long long)A * C, sPos = nchw ? 1 : C, sC = nchw ? A : 1;
  im2col_kernel<<<(unsigned)blocks, 256, 0, s>>>(X, col, total, B, C, sImg, sPos, sC, flip ? 1 : 0);
  azp_count_launch();
}

}  // namespace

extern "C" int azp_trainer_create(int board, int max_batch, int device, azp_trainer** out) {
  if (!out) return AZP_ERR_INVALID_ARG;
  *out = nullptr;
  if (board < 3 || board > AZP_MAX_BOARD || max_batch < 1) {
    azp_set_global_error("azp_trainer_create: board must be 3..19 and max_batch >= 1");
    return AZP_ERR_INVALID_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    azp_set_global_error("no CUDA device: the trainer has no CPU fallback");
    return AZP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    azp_set_global_error("azp_trainer_create: bad device ordinal");
    return AZP_ERR_INVALID_ARG;
  }
  AzpDeviceGuard dev_guard(device);
  azp_trainer* t = new azp_trainer();
  t->board = board; t->A = board * board; t->max_batch = max_batch; t->device = device;
  const size_t A = t->A, nb = max_batch, M = nb * A;
  size_t o = 0;
  auto take = [&](size_t& off, size_t count) { off = o; o += count; };
  take(t->o_c1w, 32 * 4 * 9); take(t->o_c1b, 32);
  take(t->o_c2w, 64 * 32 * 9); take(t->o_c2b, 64);
  take(t->o_c3w, 128 * 64 * 9); take(t->o_c3b, 128);
  take(t->o_pcw, 4 * 128); take(t->o_pcb, 4);
  take(t->o_pfw, A * 4 * A); take(t->o_pfb, A);
  take(t->o_vcw, 2 * 128); take(t->o_vcb, 2);
  take(t->o_v1w, 64 * 2 * A); take(t->o_v1b, 64);
  take(t->o_v2w, 64); take(t->o_v2b, 1);
  t->n_params = o;
  int st = AZP_OK;
  auto A_ = [&](float** p, size_t count) { if (st == AZP_OK) st = t->alloc(p, count); };
  A_(&t->a1, M * 32); A_(&t->a2, M * 64); A_(&t->a3, M * 128); A_(&t->hh, M * 6);
  A_(&t->pt, nb * 4 * A); A_(&t->vt, nb * 2 * A); A_(&t->logits, nb * A); A_(&t->v1, nb * 64); A_(&t->v2, nb);
  A_(&t->dz1, M * 32); A_(&t->dz2, M * 64); A_(&t->dz3, M * 128); A_(&t->dhh, M * 6);
  A_(&t->dpt, nb * 4 * A); A_(&t->dvt, nb * 2 * A); A_(&t->dlogits, nb * A); A_(&t->dv1, nb * 64); A_(&t->dv2, nb);
  t->col_floats = M * 9 * 128;
  A_(&t->col, t->col_floats);
  A_(&t->W1m, 36 * 32); A_(&t->W2m, 288 * 64); A_(&t->W3m, 576 * 128); A_(&t->W2d, 576 * 32); A_(&t->W3d, 1152 * 64);
  A_(&t->Wh, 768); A_(&t->bh, 8); A_(&t->dWh, 768); A_(&t->dbh, 8);
  // split-K partials: every weight-gradient GEMM is cut into at most ~2*148 CTAs of <=128x128 outputs
  t->partial_floats = (size_t)(SM_TARGET_CTAS + 64) * 128 * 128;
  if (t->partial_floats < 2 * A * 4 * A) t->partial_floats = 2 * A * 4 * A;
  A_(&t->partial, t->partial_floats);
  t->cs_floats = ((M + 1023) / 1024 + 1) * 512;
  A_(&t->cs_partial, t->cs_floats);
  A_(&t->rowloss, 3 * nb);
  if (st != AZP_OK) {
    azp_set_global_error(t->error);
    for (void* p : t->allocs) cudaFree(p);
    delete t;
    return st;
  }
  *out = t;
  return AZP_OK;
}

extern "C" void azp_trainer_destroy(azp_trainer* t) {
  if (!t) return;
  AzpDeviceGuard dev_guard(t->device);
  for (void* p : t->allocs) cudaFree(p);
  delete t;
}

extern "C" const char* azp_trainer_last_error(const azp_trainer* t) {
  if (t && !t->error.empty()) return t->error.c_str();
  return azp_global_error();
}

extern "C" size_t azp_trainer_weight_count(const azp_trainer* t) { return t ? t->n_params : 0; }

extern "C" int azp_train_loss_grad(azp_trainer* t, int n, const float* params, const float* planes, const float* pi,
                                   const float* z, float* grads, float* losses, void* stream) {
  if (!t) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(t->device);
  if (n < 1 || n > t->max_batch || !params || !planes || !pi || !z || !losses) {
    t->set_error("azp_train_loss_grad: n must be 1..max_batch and params/planes/pi/z/losses non-null");
    return AZP_ERR_INVALID_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int B = t->board, A = t->A, M = n * A;
  const float* P = params;
  int st;

  // ---- operand packing (weights change every step) ----
  pack_conv_weights_kernel<<<(32 * 4 * 9 + 255) / 256, 256, 0, s>>>(P + t->o_c1w, 4, 32, t->W1m, nullptr);
  pack_conv_weights_kernel<<<(64 * 32 * 9 + 255) / 256, 256, 0, s>>>(P + t->o_c2w, 32, 64, t->W2m, t->W2d);
  pack_conv_weights_kernel<<<(128 * 64 * 9 + 255) / 256, 256, 0, s>>>(P + t->o_c3w, 64, 128, t->W3m, t->W3d);
  pack_heads_kernel<<<3, 256, 0, s>>>(P + t->o_pcw, P + t->o_pcb, P + t->o_vcw, P + t->o_vcb, t->Wh, t->bh);
  azp_count_launch(4);

  // ---- forward (net.py:55-72) ----
  im2col(s, planes, t->col, n, B, 4, true, false);
  gemm(s, true, true, t->col, 36, 1, t->W1m, 32, 1, t->a1, M, 32, 36, 2, P + t->o_c1b);
  im2col(s, t->a1, t->col, n, B, 32, false, false);
  gemm(s, true, true, t->col, 288, 1, t->W2m, 64, 1, t->a2, M, 64, 288, 2, P + t->o_c2b);
  im2col(s, t->a2, t->col, n, B, 64, false, false);
  gemm(s, true, true, t->col, 576, 1, t->W3m, 128, 1, t->a3, M, 128, 576, 2, P + t->o_c3b);
  gemm(s, true, false, t->a3, 128, 1, t->Wh, 1, 128, t->hh, M, 6, 128, 2, t->bh);
  {
    const long long tot = (long long)M * 6;
    heads_to_nchw_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(t->hh, n, A, t->pt, t->vt);
    azp_count_launch();
  }
  gemm(s, true, false, t->pt, 4 * A, 1, P + t->o_pfw, 1, 4 * A, t->logits, n, A, 4 * A, 1, P + t->o_pfb);
  gemm(s, true, false, t->vt, 2 * A, 1, P + t->o_v1w, 1, 2 * A, t->v1, n, 64, 2 * A, 2, P + t->o_v1b);
  gemm(s, true, false, t->v1, 64, 1, P + t->o_v2w, 1, 64, t->v2, n, 1, 64, 1, P + t->o_v2b);

  // ---- loss (controller.py:117-123) ----
  loss_grad_kernel<<<(n + 3) / 4, 128, 0, s>>>(t->logits, pi, t->v2, z, n, A, grads ? t->dlogits : nullptr,
                                                grads ? t->dv2 : nullptr, t->rowloss);
  loss_reduce_kernel<<<1, 256, 0, s>>>(t->rowloss, n, losses);
  azp_count_launch(2);
  if (!grads) return azp_check_launch();

  // ---- backward ----
  float* G = grads;
  // value_fc2
  if ((st = gemm_splitk(t, s, false, true, t->dv2, 1, 1, t->v1, 64, 1, G + t->o_v2w, 1, 64, n))) return st;   // dW = dv2^T v1
  colsum(t, s, t->dv2, n, 1, G + t->o_v2b);
  gemm(s, true, true, t->dv2, 1, 1, P + t->o_v2w, 64, 1, t->dv1, n, 64, 1, 3, nullptr, t->v1);     // dv1 = dv2 W2 .* (v1>0)
  // value_fc1
  if ((st = gemm_splitk(t, s, false, true, t->dv1, 1, 64, t->vt, 2 * A, 1, G + t->o_v1w, 64, 2 * A, n))) return st;
  colsum(t, s, t->dv1, n, 64, G + t->o_v1b);
  gemm(s, true, true, t->dv1, 64, 1, P + t->o_v1w, 2 * A, 1, t->dvt, n, 2 * A, 64);
  // policy_fc
  if ((st = gemm_splitk(t, s, false, true, t->dlogits, 1, A, t->pt, 4 * A, 1, G + t->o_pfw, A, 4 * A, n))) return st;
  colsum(t, s, t->dlogits, n, A, G + t->o_pfb);
  gemm(s, true, true, t->dlogits, A, 1, P + t->o_pfw, 4 * A, 1, t->dpt, n, 4 * A, A);
  // 1x1 head convolutions
  {
    const long long tot = (long long)M * 6;
    heads_grad_from_nchw_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(t->dpt, t->dvt, t->hh, n, A, t->dhh);
    azp_count_launch();
  }
  if ((st = gemm_splitk(t, s, false, true, t->dhh, 1, 6, t->a3, 128, 1, t->dWh, 6, 128, M))) return st;
  colsum(t, s, t->dhh, M, 6, t->dbh);
  if (cudaMemcpyAsync(G + t->o_pcw, t->dWh, 512 * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(G + t->o_vcw, t->dWh + 512, 256 * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(G + t->o_pcb, t->dbh, 4 * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(G + t->o_vcb, t->dbh + 4, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
    t->set_error("cudaMemcpyAsync of the head gradients failed");
    return AZP_ERR_CUDA;
  }
  gemm(s, true, true, t->dhh, 6, 1, t->Wh, 128, 1, t->dz3, M, 128, 6, 3, nullptr, t->a3);          // d(act3) .* (a3>0)
  // conv3
  colsum(t, s, t->dz3, M, 128, G + t->o_c3b);
  im2col(s, t->a2, t->col, n, B, 64, false, false);
  if ((st = gemm_splitk(t, s, false, true, t->col, 1, 576, t->dz3, 128, 1, G + t->o_c3w, 576, 128, M, 64, 128))) return st;
  im2col(s, t->dz3, t->col, n, B, 128, false, true);
  gemm(s, true, true, t->col, 1152, 1, t->W3d, 64, 1, t->dz2, M, 64, 1152, 3, nullptr, t->a2);
  // conv2
  colsum(t, s, t->dz2, M, 64, G + t->o_c2b);
  im2col(s, t->a1, t->col, n, B, 32, false, false);
  if ((st = gemm_splitk(t, s, false, true, t->col, 1, 288, t->dz2, 64, 1, G + t->o_c2w, 288, 64, M, 32, 64))) return st;
  im2col(s, t->dz2, t->col, n, B, 64, false, true);
  gemm(s, true, true, t->col, 576, 1, t->W2d, 32, 1, t->dz1, M, 32, 576, 3, nullptr, t->a1);
  // conv1 (no input gradient)
  colsum(t, s, t->dz1, M, 32, G + t->o_c1b);
  im2col(s, planes, t->col, n, B, 4, true, false);
  if ((st = gemm_splitk(t, s, false, true, t->col, 1, 36, t->dz1, 32, 1, G + t->o_c1w, 36, 32, M, 4, 32))) return st;
  st = azp_check_launch();
  if (st != AZP_OK) t->set_error(azp_global_error());
  return st;
}

extern "C" int azp_adamw_step(size_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int step,
                              double lr, double beta1, double beta2, double eps, double weight_decay, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || step < 1) {
    azp_set_global_error("azp_adamw_step: null buffer or step < 1");
    return AZP_ERR_INVALID_ARG;
  }
  if (n == 0) return AZP_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float decay = (float)(1.0 - lr * weight_decay);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
  const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, params, grads, exp_avg, exp_avg_sq, decay, w1,
                                                                   (float)beta2, w2, step_size, bc2_sqrt, (float)eps);
  azp_count_launch();
  return azp_check_launch();
}
