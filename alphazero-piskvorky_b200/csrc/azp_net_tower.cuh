// Residual tower of BASELINE config 4 (10 blocks x 128 channels, BN folded) on the bf16 tcgen05 path.
// Included into azp_net_umma.cu (same translation unit: shares its PTX wrappers, ConvGeom, make_map, FC kernel).
//
//   tower_first_{bits,planes}_kernel   4 -> C first conv (+folded BN, ReLU), SIMT, padded NHWC bf16 out
//   conv_tower_kernel                  C -> C 3x3 conv: weights (288 KB per layer) do NOT fit in shared memory,
//                                      so they are STREAMED by TMA through a ring and every weight chunk is
//                                      used for TWO 128-row tiles (two TMEM accumulators, two issuer threads);
//                                      epilogue = bias (+ residual) + ReLU + halo zero
//   tower_heads_kernel                 both 1x1 head convs (+folded BN, ReLU) -> bf16 feature rows for fc_umma_kernel
#pragma once

constexpr int TW_C = 128;             // channels (the kernel is specialised for 128)
constexpr int TW_WST = 6;             // weight ring stages, 16 KB each ([128 cout][64 cin] bf16, 128B swizzle)
constexpr int TW_THREADS = 384;       // warp 0 A-TMA, warps 1/10 MMA issuers (tile 0/1 of a pair), 2..9 epilogue, 11 W-TMA

// ---- first conv from bitboards: one block per slot, 8 output channels per work item -------------------------
template <int NWT>
__global__ void __launch_bounds__(256) tower_first_bits_kernel(ConvGeom g, int C, const uint64_t* __restrict__ lx,
                                                               const uint64_t* __restrict__ lo,
                                                               const uint8_t* __restrict__ ltm,
                                                               const int16_t* __restrict__ llast,
                                                               const float* __restrict__ w /*[27][C]*/,
                                                               const float* __restrict__ bias,
                                                               __nv_bfloat16* __restrict__ act,
                                                               const int32_t* __restrict__ count_dev, int count_host,
                                                               int back) {
  const int count = count_dev ? *count_dev : count_host;
  extern __shared__ float tw_sm[];
  float* sw = tw_sm;                       // [27][C]
  float* sb = sw + 27 * C;                 // [C]
  __shared__ int8_t cells[AZP_MAX_BOARD * AZP_MAX_BOARD + 3];
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias[i];
  const int groups = C / 8;
  for (int bi = blockIdx.x; bi < count; bi += gridDim.x) {
    const int slot = back ? g.G - 1 - bi : bi;
    __syncthreads();
    const bool xm = ltm[slot] == 1;
    const int last = llast[slot];
    for (int a = threadIdx.x; a < g.A; a += blockDim.x) {
      uint64_t wx = lx[(size_t)slot * NWT + (a >> 6)], wo = lo[(size_t)slot * NWT + (a >> 6)];
      bool isx = (wx >> (a & 63)) & 1, iso = (wo >> (a & 63)) & 1;
      cells[a] = (int8_t)(((xm ? isx : iso) ? 1 : 0) | ((xm ? iso : isx) ? 2 : 0) | (a == last ? 4 : 0));
    }
    __syncthreads();
    for (int item = threadIdx.x; item < g.stride * groups; item += blockDim.x) {
      const int pos = item / groups, cg = item % groups;
      const int y = pos / g.P, x = pos % g.P;
      uint4* dst = reinterpret_cast<uint4*>(act + ((size_t)slot * g.stride + pos) * C + cg * 8);
      if (y < 1 || x < 1 || x > g.B) { *dst = make_uint4(0, 0, 0, 0); continue; }
      const int r = y - 1, c = x - 1;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = sb[cg * 8 + j];
#pragma unroll
      for (int tap = 0; tap < 9; tap++) {
        const int rr = r + tap / 3 - 1, cc = c + tap % 3 - 1;
        if (rr < 0 || rr >= g.B || cc < 0 || cc >= g.B) continue;
        const int v = cells[rr * g.B + cc];
        if (!v) continue;
#pragma unroll
        for (int pl = 0; pl < 3; pl++) {
          if (!(v & (1 << pl))) continue;
          const float* wr = sw + (tap * 3 + pl) * C + cg * 8;
#pragma unroll
          for (int j = 0; j < 8; j++) acc[j] += wr[j];
        }
      }
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
        packed[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// general float32 planes [n][4][A] -> first conv output; same (tap, plane) summation order as the bits kernel
__global__ void __launch_bounds__(256) tower_first_planes_kernel(ConvGeom g, int C, int n, const float* __restrict__ planes,
                                                                 const float* __restrict__ w4 /*[36][C]*/,
                                                                 const float* __restrict__ bias,
                                                                 __nv_bfloat16* __restrict__ act) {
  extern __shared__ float tw_sm[];
  float* sw = tw_sm;                       // [36][C]
  float* sb = sw + 36 * C;                 // [C]
  float* sp = sb + C;                      // [4][A]
  for (int i = threadIdx.x; i < 36 * C; i += blockDim.x) sw[i] = w4[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias[i];
  const int groups = C / 8;
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * g.A; i += blockDim.x) sp[i] = planes[(size_t)slot * 4 * g.A + i];
    __syncthreads();
    for (int item = threadIdx.x; item < g.stride * groups; item += blockDim.x) {
      const int pos = item / groups, cg = item % groups;
      const int y = pos / g.P, x = pos % g.P;
      uint4* dst = reinterpret_cast<uint4*>(act + ((size_t)slot * g.stride + pos) * C + cg * 8);
      if (y < 1 || x < 1 || x > g.B) { *dst = make_uint4(0, 0, 0, 0); continue; }
      const int r = y - 1, c = x - 1;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = sb[cg * 8 + j];
      for (int tap = 0; tap < 9; tap++) {
        const int rr = r + tap / 3 - 1, cc = c + tap % 3 - 1;
        if (rr < 0 || rr >= g.B || cc < 0 || cc >= g.B) continue;
        for (int pl = 0; pl < 4; pl++) {
          const float v = sp[pl * g.A + rr * g.B + cc];
          const float* wr = sw + (tap * 4 + pl) * C + cg * 8;
#pragma unroll
          for (int j = 0; j < 8; j++) acc[j] = fmaf(v, wr[j], acc[j]);
        }
      }
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
        packed[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// ---- C -> C 3x3 convolution with streamed weights --------------------------------------------------------------
struct TowerArgs {
  ConvGeom g;
  const float* bias;                  // [C] (BN folded)
  const __nv_bfloat16* residual;      // block input to add before the ReLU, or null
  __nv_bfloat16* out;
  const int32_t* count_dev; int count_host; int back;
  int a_rows;                         // 128 + 2P + 2 rows per A box
  int a_bytes;                        // bytes reserved per A buffer (rows rounded to 8) * 128
};

__global__ void __launch_bounds__(TW_THREADS, 1) conv_tower_kernel(const __grid_constant__ CUtensorMap map_act,
                                                                   const __grid_constant__ CUtensorMap map_w,
                                                                   const TowerArgs args) {
  constexpr int C = TW_C;
  constexpr int W_BYTES = C * 128;                          // one (tap, k-chunk) weight tile: [128 cout][64 cin]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;                                      // [2 kc][2 tile][a_bytes]
  uint8_t* s_w = s_a + 4 * (size_t)args.a_bytes;            // [TW_WST][W_BYTES]
  uint8_t* s_tail = s_w + TW_WST * W_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_tail);   // [2 kc][2 tile]
  uint64_t* a_empty = a_full + 4;                           // [2 kc]
  uint64_t* w_full = a_empty + 2;                           // [TW_WST]
  uint64_t* w_empty = w_full + TW_WST;                      // [TW_WST]
  uint64_t* tfull = w_empty + TW_WST;                       // [2 accumulator sets]
  uint64_t* tempty = tfull + 2;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);  // [C]

  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const ConvGeom g = args.g;
  const int count = warp_uniform(args.count_dev ? *args.count_dev : args.count_host);
  const int slot0 = args.back ? g.G - count : 0;
  const long long row_begin = (long long)slot0 * g.stride;
  const long long rows_total = (long long)count * g.stride;
  const int n_tiles = (int)((rows_total + TILE_M - 1) / TILE_M);
  const int n_pairs = (n_tiles + 1) / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; i++) mbar_init(&a_full[i], 1);
    for (int i = 0; i < 2; i++) { mbar_init(&a_empty[i], 2); mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < TW_WST; i++) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_bias[i] = args.bias[i];
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== activation producer: per pair and k-chunk (64 channels) one box per tile =====
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)args.a_rows * 128u;
      int pl = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, pl++) {
        for (int kc = 0; kc < 2; kc++) {
          mbar_wait(&a_empty[kc], (uint32_t)((pl & 1) ^ 1));
          for (int j = 0; j < 2; j++) {
            mbar_expect_tx(&a_full[kc * 2 + j], bytes);
            const long long row = row_begin + (long long)(2 * p + j) * TILE_M - g.P - 1;
            tma_load_2d(s_a + (size_t)(kc * 2 + j) * args.a_bytes, &map_act, &a_full[kc * 2 + j], kc * 64, (int)row);
          }
        }
      }
    }
  } else if (warp == 11) {
    // ===== weight producer: the layer's 18 (k-chunk, tap) tiles are streamed once per pair =====
    if (lane == 0) {
      int ws = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int kc = 0; kc < 2; kc++) {
          for (int tap = 0; tap < 9; tap++, ws++) {
            const int st = ws % TW_WST;
            mbar_wait(&w_empty[st], (uint32_t)(((ws / TW_WST) & 1) ^ 1));
            mbar_expect_tx(&w_full[st], W_BYTES);
            tma_load_2d(s_w + (size_t)st * W_BYTES, &map_w, &w_full[st], 0, (tap * 2 + kc) * C);
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===== MMA issuers: warp j accumulates tile j of every pair; the whole warp runs the loop, an elected lane issues
    //       (uniform control flow keeps the descriptors in uniform registers: see elect_one in azp_tc.cuh / azp_net_umma.cu) =====
    {
      const int j = warp == 1 ? 0 : 1;
      const uint32_t idesc = umma_idesc(C);
      const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
      int ws = 0, pl = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, pl++) {
        const int aset = pl & 1;
        mbar_wait(&tempty[aset], (uint32_t)(((pl >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)((aset * 2 + j) * C);
        uint32_t accumulate = 0;
        for (int kc = 0; kc < 2; kc++) {
          mbar_wait(&a_full[kc * 2 + j], (uint32_t)(pl & 1));
          tc_fence_after();
          const uint32_t a_tile = a_base + (uint32_t)(kc * 2 + j) * (uint32_t)args.a_bytes;
          for (int tap = 0; tap < 9; tap++, ws++) {
            const int st = ws % TW_WST;
            mbar_wait(&w_full[st], (uint32_t)((ws / TW_WST) & 1));
            tc_fence_after();
            const uint32_t a_tap = a_tile + (uint32_t)(((tap / 3) * g.P + (tap % 3)) * 128);   // row offset dy*P + dx
            const uint32_t w_tap = w_base + (uint32_t)st * W_BYTES;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              umma_bf16_e(d_tmem, umma_desc(a_tap + k * 32, 128), umma_desc(w_tap + k * 32, 128), idesc, accumulate);
              accumulate = 1;
            }
            umma_commit_e(&w_empty[st]);
          }
          umma_commit_e(&a_empty[kc]);
        }
        umma_commit_e(&tfull[aset]);
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ===== epilogue: warp (q, h) owns TMEM lanes 32q..32q+31 and channel half h of both tiles of a pair =====
    const int q = warp & 3, h = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int pl = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, pl++) {
      const int aset = pl & 1;
      mbar_wait(&tfull[aset], (uint32_t)((pl >> 1) & 1));
      tc_fence_after();
      for (int j = 0; j < 2; j++) {
        const unsigned R = (unsigned)row_begin + (unsigned)(2 * p + j) * TILE_M + q * 32 + lane;
        const int slot = (int)(R / (unsigned)g.stride), pos = (int)(R % (unsigned)g.stride);
        const int y = pos / g.P, x = pos % g.P;
        const bool in_range = slot < g.G;
        const bool interior = y >= 1 && x >= 1 && x <= g.B && in_range;
        const uint32_t taddr = tmem_base + lane_addr + (uint32_t)((aset * 2 + j) * C + h * 64);
        uint4* dst = reinterpret_cast<uint4*>(args.out + (size_t)R * C + h * 64);
        const uint4* res = args.residual ? reinterpret_cast<const uint4*>(args.residual + (size_t)R * C + h * 64) : nullptr;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          uint32_t packed[16];
#pragma unroll
          for (int v = 0; v < 4; v++) {
            uint4 rv = make_uint4(0, 0, 0, 0);
            if (res && interior) rv = res[c0 / 8 + v];
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int jj = v * 4 + u;                      // bf16 pair index inside this 32-channel chunk
              const __nv_bfloat162 rb = *reinterpret_cast<const __nv_bfloat162*>(&rw[u]);
              float v0 = __uint_as_float(r[2 * jj]) + s_bias[h * 64 + c0 + 2 * jj] + __low2float(rb);
              float v1 = __uint_as_float(r[2 * jj + 1]) + s_bias[h * 64 + c0 + 2 * jj + 1] + __high2float(rb);
              v0 = interior ? fmaxf(v0, 0.f) : 0.f;
              v1 = interior ? fmaxf(v1, 0.f) : 0.f;
              __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
              packed[jj] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
          if (in_range) {
#pragma unroll
            for (int v = 0; v < 4; v++)
              dst[c0 / 8 + v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[aset]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- 1x1 head convs (+ folded BN, ReLU) -> bf16 feature rows -------------------------------------------------------
__global__ void __launch_bounds__(256) tower_heads_kernel(ConvGeom g, int C, int NJ, int PC, int PKp, int FK,
                                                          const __nv_bfloat16* __restrict__ act,
                                                          const float* __restrict__ hw /*[NJ][C]*/,
                                                          const float* __restrict__ hb, __nv_bfloat16* __restrict__ feat,
                                                          const int32_t* __restrict__ count_dev, int count_host, int back) {
  const int count = count_dev ? *count_dev : count_host;
  extern __shared__ float tw_sm[];
  for (int i = threadIdx.x; i < NJ * C + NJ; i += blockDim.x) tw_sm[i] = i < NJ * C ? hw[i] : hb[i - NJ * C];
  __syncthreads();
  const long long total = (long long)count * g.A;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int bi = (int)(i / g.A), pix = (int)(i % g.A);
    const int slot = back ? g.G - count + bi : bi;
    const size_t row = (size_t)slot * g.stride + (size_t)(pix / g.B + 1) * g.P + (pix % g.B + 1);
    const uint4* a = reinterpret_cast<const uint4*>(act + row * C);
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c8 = 0; c8 < C / 8; c8++) {
      const uint4 v = a[c8];
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&wv[u]);
        const float x0 = __low2float(b2), x1 = __high2float(b2);
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (j < NJ) dot[j] = fmaf(x1, tw_sm[j * C + c8 * 8 + 2 * u + 1], fmaf(x0, tw_sm[j * C + c8 * 8 + 2 * u], dot[j]));
      }
    }
    __nv_bfloat16* f = feat + (size_t)slot * FK + pix;
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < NJ) f[j < PC ? j * g.A : PKp + (j - PC) * g.A] = __float2bfloat16(fmaxf(dot[j] + tw_sm[NJ * C + j], 0.f));
  }
}
