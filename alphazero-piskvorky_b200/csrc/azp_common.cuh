// Shared device helpers: bitboards, warp reductions, Philox counter RNG.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/azp_b200.h"

#define AZP_MAXW AZP_MAX_WORDS
#define AZP_WARP 32
#define AZP_FULL 0xffffffffu

// Geometry of the rules, passed BY VALUE to kernels (lives in the constant bank).
struct RulesParams {
  int B, A, NW, WL;
  int dshift[4];                    // bit distance of one step in (0,1),(1,0),(1,1),(1,-1): 1, B, B+1, B-1
  uint64_t board_mask[AZP_MAXW];    // the A valid bits
  uint64_t valid[4][AZP_MAXW];      // cells from which WL steps in direction d stay on the board
};

template <int NW> struct BB { uint64_t w[NW]; };

template <int NW> __device__ __forceinline__ BB<NW> bb_load(const uint64_t* p) {
  BB<NW> b;
#pragma unroll
  for (int i = 0; i < NW; i++) b.w[i] = p[i];
  return b;
}
template <int NW> __device__ __forceinline__ void bb_store(uint64_t* p, const BB<NW>& b) {
#pragma unroll
  for (int i = 0; i < NW; i++) p[i] = b.w[i];
}
template <int NW> __device__ __forceinline__ bool bb_test(const BB<NW>& b, int i) {
  uint64_t word = 0;
  int wi = i >> 6;
#pragma unroll
  for (int k = 0; k < NW; k++) if (k == wi) word = b.w[k];
  return (word >> (i & 63)) & 1ull;
}
template <int NW> __device__ __forceinline__ void bb_set(BB<NW>& b, int i) {
  int wi = i >> 6;
  uint64_t bit = 1ull << (i & 63);
#pragma unroll
  for (int k = 0; k < NW; k++) if (k == wi) b.w[k] |= bit;
}
template <int NW> __device__ __forceinline__ int bb_popcount(const BB<NW>& b) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < NW; k++) c += __popcll(b.w[k]);
  return c;
}
// logical shift right by s bits (0 <= s < 64*NW), register-only (no dynamic indexing)
template <int NW> __device__ __forceinline__ BB<NW> bb_shr(const BB<NW>& a, int s) {
  BB<NW> r;
  const int ws = s >> 6, bs = s & 63;
#pragma unroll
  for (int i = 0; i < NW; i++) {
    uint64_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
      if (k == i + ws) lo = a.w[k];
      if (k == i + ws + 1) hi = a.w[k];
    }
    r.w[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
  }
  return r;
}
template <int NW> __device__ __forceinline__ int bb_lowest(const BB<NW>& b) {   // -1 if empty
  int idx = -1;
#pragma unroll
  for (int k = NW - 1; k >= 0; k--) if (b.w[k]) idx = k * 64 + __ffsll((long long)b.w[k]) - 1;
  return idx;
}

__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(AZP_FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(AZP_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(AZP_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(AZP_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(AZP_FULL, v, o));
  return v;
}
// softmax over ALL A cells of one row of raw logits (controller.py:49), evaluated by one warp:
// p[a] = __expf(row[a] - m) * inv.  Every consumer of the bf16 net's raw logits (root expansion, leaf expansion,
// azp_net_forward's policy output) goes through this one function with the same lane assignment (a = lane + 32k),
// so they all see bit-identical probabilities.
__device__ __forceinline__ void warp_softmax_stats(const float* __restrict__ row, int A, int lane, float& m, float& inv) {
  float mx = -INFINITY;
  for (int a = lane; a < A; a += 32) mx = fmaxf(mx, row[a]);
  mx = warp_max_f(mx);
  float z = 0.f;
  for (int a = lane; a < A; a += 32) z += __expf(row[a] - mx);
  z = warp_sum_f(z);
  m = mx;
  inv = 1.f / z;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(AZP_FULL, v, o));
  return v;
}

// Game result of a whole board, evaluated by one warp (all lanes hold x and o).
// Same answer as the reference's row-major scan (games.py:133-166, 212-227): per colour
// and direction the set of line-START cells is  valid_d & b & (b>>d) & ... & (b>>(WL-1)d)
// (shift-AND chain), the winner is the colour owning the lowest start index; a full board
// without a line is a draw.  Lanes 0..7 = (colour, direction).
template <int NW>
__device__ __forceinline__ int warp_result(const RulesParams& rp, const BB<NW>& x, const BB<NW>& o, int lane) {
  int key = 0x7fffffff;
  if (lane < 8) {
    const int colour = lane >> 2, d = lane & 3;
    BB<NW> b;
#pragma unroll
    for (int k = 0; k < NW; k++) b.w[k] = colour ? o.w[k] : x.w[k];
    BB<NW> m = b;
    for (int i = 1; i < rp.WL; i++) {
      BB<NW> s = bb_shr<NW>(b, i * rp.dshift[d]);
#pragma unroll
      for (int k = 0; k < NW; k++) m.w[k] &= s.w[k];
    }
#pragma unroll
    for (int k = 0; k < NW; k++) m.w[k] &= rp.valid[d][k];
    int idx = bb_lowest<NW>(m);
    if (idx >= 0) key = idx * 2 + colour;
  }
  key = warp_min_i(key);
  if (key != 0x7fffffff) return (key & 1) ? 2 : 1;
  int stones = 0;
#pragma unroll
  for (int k = 0; k < NW; k++) stones += __popcll(x.w[k] | o.w[k]);
  return stones == rp.A ? 3 : 0;
}

// Result of a position reached by playing cell `a` from a NON-TERMINAL position: only the mover can own a line and
// every such line passes through `a`, so instead of the full-board scan lane (d, k) tests the one WL-run that starts
// k steps before `a` along direction d (32 lanes = 4 directions x up to 8 offsets; WL <= 8).  Same answer as
// warp_result on every position reachable in play (games.py:133-166).  `mine` = the mover's stones including `a`.
template <int NW>
__device__ __forceinline__ bool warp_line_through(const RulesParams& rp, const BB<NW>& mine, int a, int lane) {
  const int WL = rp.WL, B = rp.B;
  const int d = lane >> 3, k = lane & 7;
  bool hit = false;
  if (k < WL) {
    const int r = a / B, c = a - r * B;
    const int dr = d == 0 ? 0 : 1, dc = d == 0 ? 1 : (d == 1 ? 0 : (d == 2 ? 1 : -1));
    const int r0 = r - k * dr, c0 = c - k * dc;
    const int r1 = r0 + (WL - 1) * dr, c1 = c0 + (WL - 1) * dc;
    if (r0 >= 0 && r1 < B && c0 >= 0 && c0 < B && c1 >= 0 && c1 < B) {
      hit = true;
      int idx = r0 * B + c0;
      const int step = dr * B + dc;
      for (int j = 0; j < WL; j++, idx += step) hit = hit && bb_test<NW>(mine, idx);
    }
  }
  return __ballot_sync(AZP_FULL, hit) != 0u;
}

// ---- Philox4x32-10 counter-based RNG ------------------------------------------------
struct Philox {
  uint32_t key[2];
  __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int i = 0; i < 10; i++) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};
// key = (engine seed) mixed with the GLOBAL game id, so streams do not depend on rank or slot
__device__ __forceinline__ Philox philox_for_game(uint64_t seed, int64_t game_id) {
  uint64_t k = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(game_id + 1));
  k ^= k >> 31; k *= 0xBF58476D1CE4E5B9ull; k ^= k >> 29;
  Philox p; p.key[0] = (uint32_t)k; p.key[1] = (uint32_t)(k >> 32);
  return p;
}
__device__ __forceinline__ double u01_open(uint32_t a, uint32_t b) {   // (0,1), 53 bits
  uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
  v &= (1ull << 53) - 1;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

#define AZP_RNG_DIRICHLET 1u
#define AZP_RNG_ACTION 2u

// Gamma(alpha, 1) for alpha < 1 (Marsaglia-Tsang on alpha+1, boosted by U^(1/alpha)).
// Counter = (cell, ply, purpose | attempt<<8, 0): independent of how cells map to lanes.
__device__ inline double gamma_variate(const Philox& rng, double alpha, uint32_t cell, uint32_t ply) {
  const double d = alpha + 1.0 - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  for (uint32_t attempt = 0; attempt < 64; attempt++) {
    uint32_t r[4], r2[4];
    rng(cell, ply, AZP_RNG_DIRICHLET | (attempt << 8), 0u, r);
    rng(cell, ply, AZP_RNG_DIRICHLET | (attempt << 8), 1u, r2);
    double u1 = u01_open(r[0], r[1]), u2 = u01_open(r[2], r[3]);
    double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);          // Box-Muller
    double v = 1.0 + c * z;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = u01_open(r2[0], r2[1]);
    if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) {
      double boost = u01_open(r2[2], r2[3]);
      return d * v * pow(boost, 1.0 / alpha);
    }
  }
  return alpha;   // unreachable in practice (acceptance > 95 % per attempt)
}
