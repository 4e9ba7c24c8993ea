// Stateless batched rules primitives over bitboards.
#include "azp_internal.cuh"

// ---- pack / unpack -------------------------------------------------------------------
__global__ void pack_boards_kernel(int A, int NW, int n, const int8_t* __restrict__ cells,
                                   uint64_t* __restrict__ x, uint64_t* __restrict__ o) {
  // one warp per 64-cell word: lanes ballot two halves
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n * NW) return;
  int board = gw / NW, word = gw % NW;
  uint64_t wx = 0, wo = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    int cell = word * 64 + h * 32 + lane;
    int v = cell < A ? cells[(size_t)board * A + cell] : 0;
    uint32_t bx = __ballot_sync(AZP_FULL, v == 1), bo = __ballot_sync(AZP_FULL, v == 2);
    wx |= (uint64_t)bx << (32 * h);
    wo |= (uint64_t)bo << (32 * h);
  }
  if (lane == 0) { x[gw] = wx; o[gw] = wo; }
}

__global__ void unpack_boards_kernel(int A, int NW, int n, const uint64_t* __restrict__ x,
                                     const uint64_t* __restrict__ o, int8_t* __restrict__ cells) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * A) return;
  int board = i / A, cell = i % A;
  uint64_t wx = x[(size_t)board * NW + (cell >> 6)], wo = o[(size_t)board * NW + (cell >> 6)];
  int bit = cell & 63;
  cells[i] = ((wx >> bit) & 1) ? 1 : (((wo >> bit) & 1) ? 2 : 0);
}

// ---- legal mask: ~(x|o) & board_mask (games.py:35-47; no terminal guard) ---------------
__global__ void legal_mask_kernel(int A, int NW, int n, const uint64_t* __restrict__ x,
                                  const uint64_t* __restrict__ o, uint8_t* __restrict__ mask) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * A) return;
  int board = i / A, cell = i % A;
  uint64_t occ = x[(size_t)board * NW + (cell >> 6)] | o[(size_t)board * NW + (cell >> 6)];
  mask[i] = ((occ >> (cell & 63)) & 1) ? 0 : 1;
}

// ---- apply move (games.py:64-82) ------------------------------------------------------
__global__ void apply_moves_kernel(int A, int NW, int n, uint64_t* x, uint64_t* o, uint8_t* to_move,
                                   int16_t* last, const int16_t* __restrict__ moves, int32_t* status) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int m = moves[i];
  if (m < 0 || m >= A) { status[i] = AZP_ERR_INVALID_MOVE; return; }
  uint64_t bit = 1ull << (m & 63);
  size_t w = (size_t)i * NW + (m >> 6);
  if ((x[w] | o[w]) & bit) { status[i] = AZP_ERR_INVALID_MOVE; return; }
  if (to_move[i] == 1) x[w] |= bit; else o[w] |= bit;
  to_move[i] = 3 - to_move[i];
  last[i] = (int16_t)m;
  status[i] = AZP_OK;
}

// ---- terminal status: one warp per board ----------------------------------------------
template <int NW>
__global__ void terminal_status_kernel(RulesParams rp, int n, const uint64_t* __restrict__ x,
                                       const uint64_t* __restrict__ o, uint8_t* __restrict__ result) {
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n) return;
  BB<NW> bx = bb_load<NW>(x + (size_t)gw * NW), bo = bb_load<NW>(o + (size_t)gw * NW);
  int r = warp_result<NW>(rp, bx, bo, lane);
  if (lane == 0) result[gw] = (uint8_t)r;
}

// ---- encode (games.py:86-129): planes float32 [n][4][A] -------------------------------
__global__ void encode_kernel(int A, int NW, int n, const uint64_t* __restrict__ x,
                              const uint64_t* __restrict__ o, const uint8_t* __restrict__ to_move,
                              const int16_t* __restrict__ last, float* __restrict__ planes) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * A) return;
  int board = i / A, cell = i % A;
  uint64_t wx = x[(size_t)board * NW + (cell >> 6)], wo = o[(size_t)board * NW + (cell >> 6)];
  int bit = cell & 63;
  bool isx = (wx >> bit) & 1, iso = (wo >> bit) & 1;
  bool xm = to_move[board] == 1;
  float* p = planes + (size_t)board * 4 * A + cell;
  p[0] = (xm ? isx : iso) ? 1.f : 0.f;          // side to move's stones
  p[A] = (xm ? iso : isx) ? 1.f : 0.f;          // opponent's stones: +1.0 (games.py:120)
  p[2 * A] = (last[board] == cell) ? 1.f : 0.f; // one-hot last action
  p[3 * A] = 0.f;
}

// ---- pi from visit counts (mcts.py:143-170): one warp per row ---------------------------
// counts are float32 in the reference; log in float32, then float64 (NumPy >= 2 promotes
// the division by an np.float64 temperature), exp, normalise, cast to float32.
__device__ __forceinline__ void warp_policy_from_visits(int A, const int32_t* visits, const uint8_t* legal,
                                                        double T, float* pi, int lane) {
  if (T <= 1e-7) T = 1e-7;
  double mx = -1.0e300;
  int k = 0;
  for (int a = lane; a < A; a += 32)
    if (legal[a]) { mx = fmax(mx, (double)logf((float)visits[a] + 1e-8f) / T); k++; }
  mx = warp_max_d(mx);
  k = warp_sum_i(k);
  double sum = 0.0;
  for (int a = lane; a < A; a += 32)
    if (legal[a]) sum += exp((double)logf((float)visits[a] + 1e-8f) / T - mx);
  sum = warp_sum_d(sum);
  bool uniform = !(sum >= 1e-8);            // also catches NaN (mcts.py:167-168)
  for (int a = lane; a < A; a += 32) {
    float p = 0.f;
    if (legal[a]) p = uniform ? (float)(1.0 / k) : (float)(exp((double)logf((float)visits[a] + 1e-8f) / T - mx) / sum);
    pi[a] = p;
  }
}

__global__ void extract_policy_kernel(int A, int n, const int32_t* __restrict__ visits,
                                      const uint8_t* __restrict__ legal, const double* __restrict__ temperature,
                                      float* __restrict__ pi) {
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n) return;
  warp_policy_from_visits(A, visits + (size_t)gw * A, legal + (size_t)gw * A, temperature[gw],
                          pi + (size_t)gw * A, lane);
}

// ---- augmentation + example writer ------------------------------------------------------
// Symmetry s of a BxB grid, expressed as "output cell (i,j) reads input cell src(i,j)".
// k counter-clockwise quarter turns (torch.rot90 / np.rot90 convention, out[i,j] = in[j, B-1-i]
// for k=1), optionally followed by a left-right flip.
__device__ __forceinline__ int sym_source(int B, int k, int flip, int i, int j) {
  if (flip) j = B - 1 - j;
  int r = i, c = j;
  for (int t = 0; t < k; t++) { int nr = c, nc = B - 1 - r; r = nr; c = nc; }
  return r * B + c;
}

// grid: (position, symmetry); block: threads over the 5*A output floats
__global__ void augment_write_kernel(int B, int mode, int n, const float* __restrict__ planes,
                                     const float* __restrict__ pi, const int32_t* __restrict__ z,
                                     float* __restrict__ out_planes, float* __restrict__ out_pi,
                                     int32_t* __restrict__ out_z, long long out_offset,
                                     const long long* __restrict__ out_offset_dev, long long out_capacity,
                                     const int* __restrict__ n_dev) {
  const int A = B * B;
  const int mult = mode == AZP_AUGMENT_D4 ? 8 : (mode == AZP_AUGMENT_REFERENCE ? 4 : 1);
  if (n_dev) n = *n_dev;
  if (out_offset_dev) out_offset += *out_offset_dev;
  // more examples than the ring holds: only the newest out_capacity survive (deque(maxlen) semantics), the others are
  // not written at all, so no two blocks of this launch ever target the same slot
  const long long skip = (long long)n * mult > out_capacity ? (long long)n * mult - out_capacity : 0;
  int pos = blockIdx.x / mult, s = blockIdx.x % mult;
  for (; pos < n; pos += gridDim.x / mult) {
    if ((long long)pos * mult + s < skip) continue;
    // state symmetry: k quarter turns (+flip for s>=4 in D4); pi symmetry: the SAME element in
    // D4/NONE mode, but ALWAYS one quarter turn in reference mode (self_play.py:104-105 as written)
    int ks = s & 3, fs = (mode == AZP_AUGMENT_D4) ? (s >> 2) : 0;
    int kp = (mode == AZP_AUGMENT_REFERENCE) ? 1 : ks, fp = fs;
    long long slot = (out_offset + (long long)pos * mult + s) % out_capacity;
    const float* ip = planes + (size_t)pos * 4 * A;
    float* op = out_planes + (size_t)slot * 4 * A;
    for (int t = threadIdx.x; t < 4 * A; t += blockDim.x) {
      int ch = t / A, cell = t % A;
      op[t] = ip[ch * A + sym_source(B, ks, fs, cell / B, cell % B)];
    }
    const float* ipi = pi + (size_t)pos * A;
    float* opi = out_pi + (size_t)slot * A;
    for (int t = threadIdx.x; t < A; t += blockDim.x) opi[t] = ipi[sym_source(B, kp, fp, t / B, t % B)];
    if (threadIdx.x == 0) out_z[slot] = z[pos];
  }
}

// ---- host wrappers ------------------------------------------------------------------------
static inline int blocks_for(size_t n, int per) { return (int)((n + per - 1) / per); }

extern "C" int azp_pack_boards(int board, int n, const int8_t* cells, uint64_t* x, uint64_t* o, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  int A = board * board, NW = AZP_WORDS(board);
  pack_boards_kernel<<<blocks_for((size_t)n * NW * 32, 256), 256, 0, (cudaStream_t)stream>>>(A, NW, n, cells, x, o);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_unpack_boards(int board, int n, const uint64_t* x, const uint64_t* o, int8_t* cells, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  int A = board * board, NW = AZP_WORDS(board);
  unpack_boards_kernel<<<blocks_for((size_t)n * A, 256), 256, 0, (cudaStream_t)stream>>>(A, NW, n, x, o, cells);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_legal_mask(int board, int n, const uint64_t* x, const uint64_t* o, uint8_t* mask, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  int A = board * board, NW = AZP_WORDS(board);
  legal_mask_kernel<<<blocks_for((size_t)n * A, 256), 256, 0, (cudaStream_t)stream>>>(A, NW, n, x, o, mask);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_apply_moves(int board, int n, uint64_t* x, uint64_t* o, uint8_t* to_move, int16_t* last,
                               const int16_t* moves, int32_t* status, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  int A = board * board, NW = AZP_WORDS(board);
  apply_moves_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(A, NW, n, x, o, to_move, last, moves, status);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_terminal_status(int board, int win_length, int n, const uint64_t* x, const uint64_t* o,
                                   uint8_t* result, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || win_length < 1 || win_length > board || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  RulesParams rp = azp_make_rules(board, win_length);
  int grid = blocks_for((size_t)n * 32, 128);
  AZP_DISPATCH_NW(rp.NW, terminal_status_kernel, grid, 128, 0, (cudaStream_t)stream, rp, n, x, o, result);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_encode(int board, int n, const uint64_t* x, const uint64_t* o, const uint8_t* to_move,
                          const int16_t* last, float* planes, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  int A = board * board, NW = AZP_WORDS(board);
  encode_kernel<<<blocks_for((size_t)n * A, 256), 256, 0, (cudaStream_t)stream>>>(A, NW, n, x, o, to_move, last, planes);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_extract_policy(int board, int n, const int32_t* visits, const uint8_t* legal,
                                  const double* temperature, float* pi, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0) return AZP_ERR_INVALID_ARG;
  if (n == 0) return AZP_OK;
  extract_policy_kernel<<<blocks_for((size_t)n * 32, 128), 128, 0, (cudaStream_t)stream>>>(board * board, n, visits, legal, temperature, pi);
  azp_count_launch();
  return azp_check_launch();
}
int azp_augment_launch(int board, int mode, int n_max, const int* n_dev, const float* planes, const float* pi,
                       const int32_t* z, float* out_planes, float* out_pi, int32_t* out_z,
                       long long out_offset, const long long* out_offset_dev, long long out_capacity,
                       cudaStream_t stream) {
  int mult = mode == AZP_AUGMENT_D4 ? 8 : (mode == AZP_AUGMENT_REFERENCE ? 4 : 1);
  if (n_max <= 0) return AZP_OK;
  int grid = (int)(((long long)n_max < 4096 ? n_max : 4096) * mult);
  augment_write_kernel<<<grid, 128, 0, stream>>>(board, mode, n_max, planes, pi, z, out_planes, out_pi, out_z,
                                                 out_offset, out_offset_dev, out_capacity, n_dev);
  azp_count_launch();
  return azp_check_launch();
}
extern "C" int azp_augment_write(int board, int mode, int n, const float* planes, const float* pi, const int32_t* z,
                                 float* out_planes, float* out_pi, int32_t* out_z, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n < 0 || mode < 0 || mode > 2) return AZP_ERR_INVALID_ARG;
  int mult = mode == AZP_AUGMENT_D4 ? 8 : (mode == AZP_AUGMENT_REFERENCE ? 4 : 1);
  return azp_augment_launch(board, mode, n, nullptr, planes, pi, z, out_planes, out_pi, out_z, 0, nullptr,
                            (long long)n * mult + 1, (cudaStream_t)stream);
}

// ---- replay ring writer (ReplayBuffer.extend on the device, replay_buffer.py:27-29; + _augment_symmetries) --------
__global__ void ring_advance_kernel(long long* head, const int* n_dev, int n_host, int mult) {
  *head += (long long)(n_dev ? *n_dev : n_host) * mult;
}
extern "C" int azp_augment_write_ring(int board, int mode, int n_max, const int32_t* n_dev, const float* planes,
                                      const float* pi, const int32_t* z, float* ring_planes, float* ring_pi,
                                      int32_t* ring_z, int64_t* ring_head_dev, int64_t ring_capacity, void* stream) {
  if (board < 1 || board > AZP_MAX_BOARD || n_max < 0 || mode < 0 || mode > 2 || !ring_head_dev || ring_capacity < 1)
    return AZP_ERR_INVALID_ARG;
  if (n_max == 0) return AZP_OK;
  const int mult = mode == AZP_AUGMENT_D4 ? 8 : (mode == AZP_AUGMENT_REFERENCE ? 4 : 1);
  cudaStream_t s = (cudaStream_t)stream;
  int st = azp_augment_launch(board, mode, n_max, n_dev, planes, pi, z, ring_planes, ring_pi, ring_z, 0,
                              reinterpret_cast<const long long*>(ring_head_dev), (long long)ring_capacity, s);
  if (st) return st;
  ring_advance_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(ring_head_dev), n_dev, n_max, mult);
  azp_count_launch();
  return azp_check_launch();
}
