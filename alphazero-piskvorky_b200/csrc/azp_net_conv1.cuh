// conv1 (4 -> 32, 3x3, alphazero/net.py:37, 57) as a NINE-TAP implicit GEMM on 16-byte input rows.
// Included into azp_net_umma.cu (same translation unit: shares its PTX wrappers, ConvGeom and launch helpers).
//
// conv1_umma_kernel expands the 27 neighbourhood bits of every position into a 32-wide bf16 A row with table lookups:
// ~1.5 k warp instructions per 128-row tile, issue-bound at 18-21 us per 2048-leaf launch for 0.3 % of the net's FLOPs.
// Here the gathering of the neighbourhood is left to the tensor core, exactly as in conv2 / conv3: the producers write
// ONE 16-byte row per padded position ("act0": 8 bf16 channels = side to move's stone, opponent's stone, last move,
// a constant one that carries the bias, plane 3 (float-planes path only), three zeros), and a 3x3 tap is the constant
// row offset ky*P + kx into that array.  The rows are 16 bytes, so the operand uses the NO-SWIZZLE K-major canonical
// layout ((8,m),(T,2)):((1T,SBO),(1,LBO)) with SBO = 128 B: rows uniformly 16 B apart, any row may start a
// descriptor.  A K = 16 UMMA consumes two 8-channel K blocks whose distance is the descriptor's LBO -- so ONE
// instruction covers TWO taps (LBO = the row distance between them): five UMMAs (M128 N64 K16) per tile, the tenth
// "tap" has zero weights.  (`pairs = 0`: nine UMMAs whose second K block meets zero weights; kept as a cross-check.)
// W is split into bf16 hi + lo halves of one N = 64 operand as before (the epilogue adds the two 32-column halves), so
// conv1 stays fp32-accurate; inputs are exact in bf16 on the leaf path.  The float-planes variant (azp_net_forward)
// runs the SAME instruction sequence on rows that also carry plane 3, so on 0/1 planes with an empty plane 3 both
// variants agree bit for bit (the search sees the numbers net_forward reports).
// Unlike the bit-expansion kernel, a halo row's neighbourhood is not empty: the epilogue zeroes halo rows itself.
#pragma once

constexpr int C1X_NPG = 2;                                   // producer groups (tile tl is built by group tl % 2); two CTAs per SM
constexpr int C1X_THREADS = 32 * (1 + 4 * C1X_NPG + 8);      // warp 0 issuer, 1..8 producers, 9..16 two epilogue groups
constexpr int C1X_STAGES = 4;                                // input stages == TMEM accumulators (64 columns each)
constexpr int C1X_STAGE_ROWS = 192;                          // >= 128 + 2P + 3 rows for P <= 24
constexpr int C1X_STAGE_BYTES = C1X_STAGE_ROWS * 16;
constexpr int conv1x_smem_bytes() { return 1024 + C1X_W_BYTES + C1X_STAGES * C1X_STAGE_BYTES + 3 * C1X_STAGES * 8 + 64; }

struct Conv1XArgs {
  ConvGeom g;
  const __nv_bfloat16* w;          // [10 taps][64 = hi|lo x 32 cout][8 k] bf16; k: planes 0..2, bias (centre tap), plane 3, zeros
  const int32_t* count_dev; int count_host; int back;
  const uint64_t* lx; const uint64_t* lo; const uint8_t* ltm; const int16_t* llast; int NW;   // leaf batch (bitboards)
  const float* planes;             // PLANES: [count][4][A] float32
  __nv_bfloat16* out;              // act1 [rows][32]
  int pairs;                       // 1: five two-tap UMMAs per tile, 0: nine
};


template <bool PLANES>
__global__ void __launch_bounds__(C1X_THREADS, 2) conv1_taps_kernel(const Conv1XArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                                         // 10 x [64][8] bf16
  uint8_t* s_a = s_w + C1X_W_BYTES;                            // C1X_STAGES x [192 rows][8] bf16
  uint64_t* full = reinterpret_cast<uint64_t*>(s_a + C1X_STAGES * C1X_STAGE_BYTES);
  uint64_t* done = full + C1X_STAGES;                          // UMMAs of the stage retired: stage free AND accumulator ready
  uint64_t* tempty = done + C1X_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C1X_STAGES);
  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const ConvGeom g = args.g;

  const unsigned tl_idx = threadIdx.x == 0 ? AZP_TL_ENTRY(AZP_TL_CONV1, gridDim.x) : 0u;
  // ---- prologue (independent of the previous kernel: overlaps its tail under programmatic dependent launch) ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < C1X_STAGES; i++) { mbar_init(&full[i], 4); mbar_init(&done[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C1X_W_BYTES / 16; i += C1X_THREADS)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(args.w)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(tmem_slot, 64 * C1X_STAGES);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();                      // leaf batch / leaf count are written by the descent kernel before us
  pdl_launch_dependents();
  if (threadIdx.x == 0) AZP_TL_START(AZP_TL_CONV1, tl_idx);
  const int count = warp_uniform(args.count_dev ? *args.count_dev : args.count_host);
  const int slot0 = args.back ? g.G - count : 0;
  const long long row_begin = (long long)slot0 * g.stride;
  const long long rows_total = (long long)count * g.stride;
  const long long row_end = row_begin + rows_total;
  const int n_tiles = (int)((rows_total + TILE_M - 1) / TILE_M);
  const unsigned stride = (unsigned)g.stride, stride_rcp = (unsigned)(0x100000000ull / (unsigned)g.stride);
  const unsigned p_rcp = (65536u + (unsigned)g.P - 1u) / (unsigned)g.P;     // pos / P == (pos * p_rcp) >> 16 for pos < 2730

  if (warp == 0) {
    // ===== MMA issuer: the whole warp runs the loop, an elected lane issues (see elect_one in azp_net_umma.cu) =====
    {
      const uint32_t idesc = umma_idesc(64);
      const uint32_t w_base = smem_u32(s_w), a_base = smem_u32(s_a);
      int tl = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
        const int st = tl % C1X_STAGES;
        const uint32_t ph = (tl / C1X_STAGES) & 1;
        mbar_wait_sleep(&tempty[st], ph ^ 1);
        mbar_wait_sleep(&full[st], ph);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(st * 64);
        const uint32_t a = a_base + (uint32_t)(st * C1X_STAGE_BYTES);
        if (args.pairs) {
#pragma unroll
          for (int p = 0; p < 5; p++) {
            const int t0 = 2 * p, t1 = t0 + 1;
            const int off0 = (t0 / 3) * g.P + t0 % 3;
            const int off1 = t1 < 9 ? (t1 / 3) * g.P + t1 % 3 : off0 + 1;      // the tenth block has zero weights
            umma_bf16_e(d, umma_desc_rows16(a + (uint32_t)off0 * 16u, (uint32_t)(off1 - off0) * 16u),
                      umma_desc_rows16(w_base + (uint32_t)t0 * C1X_TAP_BYTES, C1X_TAP_BYTES), idesc, p ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int t0 = 0; t0 < 9; t0++) {
            const int off0 = (t0 / 3) * g.P + t0 % 3;
            umma_bf16_e(d, umma_desc_rows16(a + (uint32_t)off0 * 16u, 16u),
                      umma_desc_rows16(w_base + (uint32_t)t0 * C1X_TAP_BYTES, (uint32_t)(9 - t0) * C1X_TAP_BYTES), idesc,
                      t0 ? 1u : 0u);
          }
        }
        umma_commit_e(&done[st]);
      }
    }
  } else if (warp <= 4 * C1X_NPG) {
    // ===== input-row producers: group pg builds the tiles with tl % 2 == pg; stage row j = padded row R0 - (P+1) + j =====
    // The raw inputs of a group's NEXT tile are requested (plain loads into registers) before the current tile is
    // built, so the L2 latency of the leaf records overlaps a whole tile period instead of stalling every tile.
    const int pg = (warp - 1) >> 2;
    const int pt = threadIdx.x - 32 - pg * 128;              // 0..127
    const int NR = TILE_M + 2 * g.P + 3;
    struct Raw { uint32_t a, b, c, d; int meta; int cell; };  // bitboards: wx (a,b), wo (c,d), meta = to-move | last << 8; planes: 4 floats
    auto fetch = [&](int t, Raw (&raw)[2]) {
      const long long Rs = row_begin + (long long)t * TILE_M - (g.P + 1);
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
        raw[pass].cell = -1;
        const int j = pt + pass * 128;
        const long long Rj = Rs + j;
        if (t < n_tiles && j < NR && Rj >= row_begin && Rj < row_end) {
          const unsigned R = (unsigned)Rj;
          unsigned slot = __umulhi(R, stride_rcp);            // floor(2^32 / stride) underestimates by < 1
          if (R - slot * stride >= stride) slot++;
          const unsigned pos = R - slot * stride;
          const unsigned y = (pos * p_rcp) >> 16, x = pos - y * (unsigned)g.P;
          if (y >= 1u && x >= 1u && x <= (unsigned)g.B) {
            const unsigned cell = (y - 1u) * (unsigned)g.B + (x - 1u);
            raw[pass].cell = (int)cell;
            if (!PLANES) {
              const unsigned wi = cell >> 6;
              const uint2 wx = __ldg(reinterpret_cast<const uint2*>(args.lx + (size_t)slot * args.NW + wi));
              const uint2 wo = __ldg(reinterpret_cast<const uint2*>(args.lo + (size_t)slot * args.NW + wi));
              raw[pass].a = wx.x; raw[pass].b = wx.y; raw[pass].c = wo.x; raw[pass].d = wo.y;
              raw[pass].meta = (int)__ldg(args.ltm + slot) | ((int)__ldg(args.llast + slot) * 256);
            } else {
              const float* pb = args.planes + (size_t)((int)slot - slot0) * 4 * g.A + cell;
              raw[pass].a = __float_as_uint(pb[0]); raw[pass].b = __float_as_uint(pb[g.A]);
              raw[pass].c = __float_as_uint(pb[2 * g.A]); raw[pass].d = __float_as_uint(pb[3 * g.A]);
            }
          }
        }
      }
    };
    Raw cur[2], nxt[2];
    fetch((int)blockIdx.x + pg * (int)gridDim.x, cur);
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
      if ((tl % C1X_NPG) != pg) continue;
      const int st = tl % C1X_STAGES;
      const uint32_t ph = (tl / C1X_STAGES) & 1;
      fetch(t + C1X_NPG * (int)gridDim.x, nxt);
      uint4 v[2];
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
        v[pass] = make_uint4(0u, 0u, 0u, 0u);
        if (cur[pass].cell >= 0) {
          if (!PLANES) {
            const unsigned sh = (unsigned)cur[pass].cell & 63u;
            const unsigned bx = ((sh < 32u ? cur[pass].a : cur[pass].b) >> (sh & 31u)) & 1u;
            const unsigned bo = ((sh < 32u ? cur[pass].c : cur[pass].d) >> (sh & 31u)) & 1u;
            const bool xm = (cur[pass].meta & 0xFF) == 1;
            const int last = cur[pass].meta >> 8;             // arithmetic shift: -1 stays -1
            const unsigned mine = xm ? bx : bo, theirs = xm ? bo : bx;
            v[pass].x = (mine ? 0x3F80u : 0u) | (theirs ? 0x3F800000u : 0u);
            v[pass].y = (last == cur[pass].cell ? 0x3F80u : 0u) | 0x3F800000u;
          } else {
            __nv_bfloat162 h01 = __floats2bfloat162_rn(__uint_as_float(cur[pass].a), __uint_as_float(cur[pass].b));
            __nv_bfloat162 h23 = __floats2bfloat162_rn(__uint_as_float(cur[pass].c), 1.f);
            __nv_bfloat162 h45 = __floats2bfloat162_rn(__uint_as_float(cur[pass].d), 0.f);
            v[pass].x = *reinterpret_cast<uint32_t*>(&h01);
            v[pass].y = *reinterpret_cast<uint32_t*>(&h23);
            v[pass].z = *reinterpret_cast<uint32_t*>(&h45);
          }
        }
      }
      mbar_wait(&done[st], ph ^ 1);                          // the UMMAs that read this stage four tiles ago have retired
      uint8_t* sp = s_a + (size_t)st * C1X_STAGE_BYTES;
      *reinterpret_cast<uint4*>(sp + pt * 16) = v[0];
      if (pt + 128 < C1X_STAGE_ROWS) *reinterpret_cast<uint4*>(sp + (pt + 128) * 16) = v[1];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
#pragma unroll
      for (int pass = 0; pass < 2; pass++) cur[pass] = nxt[pass];
    }
  } else {
    // ===== epilogue: group eg drains the tiles with (tl & 1) == eg; warp q owns TMEM lanes 32q..32q+31 =====
    const int eg = (warp - 1 - 4 * C1X_NPG) >> 2;
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int m = q * 32 + lane;
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
      if ((tl & 1) != eg) continue;
      const int st = tl % C1X_STAGES;
      const uint32_t ph = (tl / C1X_STAGES) & 1;
      const unsigned R = (unsigned)(row_begin + (long long)t * TILE_M + m);
      unsigned slot = __umulhi(R, stride_rcp);
      if (R - slot * stride >= stride) slot++;
      const unsigned pos = R - slot * stride;
      const unsigned y = (pos * p_rcp) >> 16, x = pos - y * (unsigned)g.P;
      const bool interior = y >= 1u && x >= 1u && x <= (unsigned)g.B && (long long)R < row_end;
      mbar_wait_sleep(&done[st], ph);
      tc_fence_after();
      uint32_t packed[16];
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {                         // 16 channels at a time: hi-part + lo-part products
        uint32_t r[16], r2[16];
        tmem_ld16_nowait(tmem_base + lane_addr + (uint32_t)(st * 64 + hf * 16), r);
        tmem_ld16_nowait(tmem_base + lane_addr + (uint32_t)(st * 64 + 32 + hf * 16), r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float v0 = interior ? fmaxf(__uint_as_float(r[2 * j]) + __uint_as_float(r2[2 * j]), 0.f) : 0.f;
          const float v1 = interior ? fmaxf(__uint_as_float(r[2 * j + 1]) + __uint_as_float(r2[2 * j + 1]), 0.f) : 0.f;
          __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
          packed[hf * 8 + j] = *reinterpret_cast<uint32_t*>(&hh);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[st]);
      // each lane owns the 64 bytes of its output row; lane pairs exchange halves so that every store instruction writes whole
      // 64-byte rows (store_rows64_paired)
      uint8_t* gp = reinterpret_cast<uint8_t*>(args.out) + (size_t)R * 64;
      store_rows64_paired(packed, gp, 64, true, true, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) AZP_TL_END(AZP_TL_CONV1, tl_idx, gridDim.x);
  if (warp == 0) tmem_dealloc(tmem_base, 64 * C1X_STAGES);
}
