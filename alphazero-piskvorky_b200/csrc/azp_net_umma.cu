// Policy/value network, bf16 tcgen05 path (AZP_DTYPE_BF16) for the conv-stem GomokuNet
// (alphazero/net.py:28-72).  Hand-written sm_100a: TMA-staged activations, weights resident in
// shared memory, tcgen05.mma (UMMA) implicit GEMM with fp32 accumulators in TMEM, fused epilogues.
//
// Activation layout in HBM ("padded NHWC", bf16): one row of C channels per padded board position;
// a board occupies (B+1) row groups of P positions (P = B+1 for B >= 7 -- one halo column per row group and nothing else;
// 16 for the smallest boards, so that a 128-row tile never spans more than three boards):
//     row(slot, r, c) = slot*(B+1)*P + (r+1)*P + (c+1)
// Row group 0 and column 0 of every row group (and columns > B) are ZERO and double as the bottom /
// right halo of the previous row group / board, so a 3x3 tap (dy,dx) is the constant row offset
// dy*P + dx and the implicit-GEMM A operand of a tap is a plain [128 rows x C] window of that matrix.
// UMMA shared-memory descriptors may start at ANY row of a swizzled tile (the swizzle is a function of absolute
// address bits), so ONE TMA box of 128+2P+2 rows per 128-row output tile feeds all nine taps.
//
// Pipeline per step (leaf batch of n positions):
//   conv1               bitboards -> act1 [rows][32]      tcgen05: nine-tap implicit GEMM on 16-byte input rows, two taps per UMMA
//                                                         (azp_net_conv1.cuh; in the exact mode with a machine-full of games it
//                                                         runs as the TAIL of the descent kernel instead: azp_conv1_tail.cuh)
//   conv_umma<32,64>    act1 -> act2 [rows][64]           tcgen05, bias+ReLU+pad-zero epilogue
//   conv_umma<64,128>   act2 -> head features [n][6][A]   tcgen05, epilogue = bias+ReLU + both 1x1 head
//                                                         convs (act3 never leaves the SM)
//   fc_umma_kernel      features -> logits / value        tcgen05 GEMMs (policy_fc, value_fc1), thread-per-row tails
// Every issuer WARP runs its issue loop whole and elects one lane per tcgen05 instruction (azp_tc.cuh: elect_one).
#include "azp_net.cuh"
#include "azp_tc.cuh"
#include <cuda.h>
#include <stdlib.h>

// Bring-up instrumentation is COMPILED OUT of the default build: the per-tile clock stamps (AZP_CLK) and the
// timing-experiment switches (AZP_DEBUG_CONV, which skip work and therefore produce wrong results by design) exist only
// in a `make BRINGUP=1` build (-DAZP_BRINGUP); otherwise the pointers / bits below are compile-time constants and every
// stamp store and debug branch is dead code.  A stray environment variable cannot touch the product kernels.
#ifdef AZP_BRINGUP
#define AZP_CLKPTR(p) (p)
#define AZP_DBGBITS(x) (x)
#else
#define AZP_CLKPTR(p) (static_cast<long long*>(nullptr))
#define AZP_DBGBITS(x) 0
#endif

namespace {


__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_row) : "memory");
}
// AZP_DISCARD=1 (experiment, off by default): dead activations leave L2 WITHOUT a write-back.  act1 / act2 are written once
// and read once per evaluation (100 MB per 2048 leaves at 15x15); the kernel that follows the last reader (it is past its
// griddepcontrol.wait: the reader has completed) invalidates the lines, the next evaluation rewrites every byte of them before
// anything reads them.  Measured: no gain (conv1 / conv2 are bound by their store pattern into L2, not by HBM write-backs)
// and +3 us in the FC kernel that issues 524 k discards.  `bytes` is a multiple of 128, all threads of the grid take part.
__device__ __forceinline__ void discard_l2(const void* base, long long bytes) {
  const long long nthreads = (long long)gridDim.x * gridDim.y * blockDim.x;
  const long long tid = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
  const char* p = static_cast<const char*>(base);
  for (long long o = tid * 128; o < bytes; o += nthreads * 128)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + o) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand with a hardware swizzle whose span is one row
// (64 B rows -> SWIZZLE_64B, 128 B rows -> SWIZZLE_128B): 8-row atoms, stride between atoms = 8 rows.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t row_bytes, int base_mode = 0) {
  const uint64_t layout = row_bytes == 128 ? 2ull : 4ull;        // SWIZZLE_128B : SWIZZLE_64B
  const uint64_t sbo = (8u * row_bytes) >> 4;
  // "matrix base offset" (bits 49-51): position of the start address inside the repeating swizzle pattern
  const uint64_t base = base_mode == 2 ? (uint64_t)((smem_addr >> 7) & (row_bytes == 128 ? 7u : 3u)) : 0ull;
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (base << 49) | (layout << 61);
}


// ---- conv1 from bitboards (SIMT) ------------------------------------------------------------------------
// One block per slot; work item = (padded position, 8 output channels).  Inputs are 0/1 planes
// (games.py:86-129; plane 3 is all zero), so conv1 needs no multiplies: for every input plane and kernel row
// the three column taps form a 3-bit pattern that indexes a table of pre-summed weight rows
// (tab: [plane(3)][ky][pattern(8)][32] fp32).  Same association order as conv1_planes_kernel, so the leaf
// path and the float-planes path agree bit for bit.
template <int NWT>
__global__ void __launch_bounds__(256) conv1_bits_kernel(ConvGeom g, const uint64_t* __restrict__ lx,
                                                         const uint64_t* __restrict__ lo,
                                                         const uint8_t* __restrict__ ltm,
                                                         const int16_t* __restrict__ llast,
                                                         const float* __restrict__ tab, const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ act1,
                                                         const int32_t* __restrict__ count_dev, int count_host,
                                                         int back) {
  const int count = count_dev ? *count_dev : count_host;
  __shared__ __align__(16) float st[72 * 32];
  __shared__ float sb[32];
  __shared__ int16_t codes[AZP_MAX_BOARD * AZP_MAX_BOARD + 3];
  for (int i = threadIdx.x; i < 72 * 32; i += blockDim.x) st[i] = tab[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  for (int bi = blockIdx.x; bi < count; bi += gridDim.x) {
    const int slot = back ? g.G - 1 - bi : bi;
    __syncthreads();
    const bool xm = ltm[slot] == 1;
    const int last = llast[slot];
    const uint64_t* own = (xm ? lx : lo) + (size_t)slot * NWT;
    const uint64_t* opp = (xm ? lo : lx) + (size_t)slot * NWT;
    for (int a = threadIdx.x; a < g.A; a += blockDim.x) {
      const int c = a % g.B;
      int code = 0;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int cc = c + kx - 1, aa = a + kx - 1;
        if (cc < 0 || cc >= g.B) continue;
        const int b0 = (int)((own[aa >> 6] >> (aa & 63)) & 1), b1 = (int)((opp[aa >> 6] >> (aa & 63)) & 1);
        code |= (b0 << kx) | (b1 << (3 + kx)) | ((aa == last ? 1 : 0) << (6 + kx));
      }
      codes[a] = (int16_t)code;
    }
    __syncthreads();
    // 4 threads per padded position, 8 output channels each: one 16-byte store per thread
    for (int item = threadIdx.x; item < g.stride * 4; item += blockDim.x) {
      const int pos = item >> 2, cg = item & 3;
      const int y = pos / g.P, x = pos % g.P;
      uint4* dst = reinterpret_cast<uint4*>(act1 + ((size_t)slot * g.stride + pos) * 32 + cg * 8);
      if (y < 1 || x < 1 || x > g.B) {                 // halo rows/columns stay zero
        *dst = make_uint4(0, 0, 0, 0);
        continue;
      }
      const int r = y - 1, c = x - 1;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = sb[cg * 8 + j];
      int code3[3];
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int rr = r + ky - 1;
        code3[ky] = (rr >= 0 && rr < g.B) ? codes[rr * g.B + c] : 0;
      }
#pragma unroll
      for (int pl = 0; pl < 3; pl++) {
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const int pat = (code3[ky] >> (3 * pl)) & 7;
          if (!pat) continue;       // empty pattern = all-zero table row (x + 0.0f == x): skip its shared-memory read,
                                    // the table traffic is what bounds this kernel and most patterns are empty
          const float4* tr = reinterpret_cast<const float4*>(st + ((pl * 3 + ky) * 8 + pat) * 32 + cg * 8);
          const float4 w0 = tr[0], w1 = tr[1];
          acc[0] += w0.x; acc[1] += w0.y; acc[2] += w0.z; acc[3] += w0.w;
          acc[4] += w1.x; acc[5] += w1.y; acc[6] += w1.z; acc[7] += w1.w;
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

// general float32 planes [n][4][A] -> act1 (used by azp_net_forward on arbitrary inputs)
__global__ void __launch_bounds__(256) conv1_planes_kernel(ConvGeom g, int n, const float* __restrict__ planes,
                                                           const float* __restrict__ w4 /*[tap][4][32]*/,
                                                           const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ act1) {
  __shared__ float sw[36 * 32];
  __shared__ float sb[32];
  extern __shared__ float sp[];          // [4][A]
  for (int i = threadIdx.x; i < 36 * 32; i += blockDim.x) sw[i] = w4[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * g.A; i += blockDim.x) sp[i] = planes[(size_t)slot * 4 * g.A + i];
    __syncthreads();
    for (int pos = threadIdx.x; pos < g.stride; pos += blockDim.x) {
      const int y = pos / g.P, x = pos % g.P;
      uint4* dst = reinterpret_cast<uint4*>(act1 + ((size_t)slot * g.stride + pos) * 32);
      if (y < 1 || x < 1 || x > g.B) {
#pragma unroll
        for (int q = 0; q < 4; q++) dst[q] = make_uint4(0, 0, 0, 0);
        continue;
      }
      const int r = y - 1, c = x - 1;
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; j++) acc[j] = sb[j];
      // same association order as the table-driven producer of the fused kernel: for every (plane, kernel
      // row) the three column taps are summed first, then added to the accumulator -- bit-identical to it
      // on 0/1 planes
      for (int pl = 0; pl < 4; pl++) {
        for (int ky = 0; ky < 3; ky++) {
          const int rr = r + ky - 1;
          if (rr < 0 || rr >= g.B) continue;
          float part[32];
#pragma unroll
          for (int j = 0; j < 32; j++) part[j] = 0.f;
          for (int kx = 0; kx < 3; kx++) {
            const int cc = c + kx - 1;
            if (cc < 0 || cc >= g.B) continue;
            const float v = sp[pl * g.A + rr * g.B + cc];
            const float* wr = sw + ((ky * 3 + kx) * 4 + pl) * 32;
#pragma unroll
            for (int j = 0; j < 32; j++) part[j] = fmaf(v, wr[j], part[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; j++) acc[j] += part[j];
        }
      }
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
        packed[j] = *reinterpret_cast<uint32_t*>(&h);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
  }
}


// ---- conv1 from bitboards, merged-table SIMT kernel (AZP_CONV1=table) ---------------------------------------------------------
// Inputs are 0/1 planes (games.py:86-129) and plane 3 is all zero, so for a position (r, c) and a kernel row ky the
// contribution of BOTH stone planes is a function of six bits -- the side to move's and the opponent's stones in columns
// c-1, c, c+1 of board row r+ky-1 -- and is read from a table:
//     tab2[ky][own3 | opp3 << 3][32]  =  (sum of w[co][0][ky][kx] over the set own bits, kx ascending)
//                                      + (the same for plane 1)                                   (24 KB of fp32)
//     act1[r, c][co] = relu( ((b[co] + tab2[0][.]) + tab2[1][.]) + tab2[2][.]  (+ w[co][2][ky][kx] of the last move, if it
//                            lies in the 3x3 window: at most nine positions of a board) )
// Three 128-byte table rows per position instead of the nine pattern rows of the first SIMT kernel (which was bound by
// exactly that shared-memory traffic), no multiplies, fp32 exact up to the summation order; conv1_table_planes_kernel adds
// the same terms in the same order, so the leaf path and azp_net_forward agree bit for bit on 0/1 planes.
// Work item = (interior position, 8 output channels): four threads per position.  Two neighbouring positions read
// different 64-byte halves of their table rows in each of the two 128-bit loads (thread (p, cg) takes 16-byte chunk
// cg + 4*((p+k)&1) in load k), so the eight threads of a quarter warp always cover all 32 banks; each thread ends up with
// channels [4cg, 4cg+4) and [16+4cg, 16+4cg+4) and writes two 8-byte halves of the position's 64-byte bf16 row.
// Halo rows are never written: they were zeroed when the buffer was allocated and nothing else touches them.
constexpr int C1T_TAB_FLOATS = 3 * 64 * 32;

constexpr int C1T_WARPS = 8;
constexpr int C1T_CODES = (AZP_MAX_BOARD + 2) * AZP_MAX_BOARD;      // per warp: [row r+1][c], rows 0 and B+1 stay zero

// One WARP per board (a CTA-per-board version spent its time in the per-board chain global load -> barrier -> codes ->
// barrier -> items with only four boards in flight per SM; here every warp runs that chain on its own, 32 boards per SM).
template <int NWT>
__global__ void __launch_bounds__(32 * C1T_WARPS, 4) conv1_table_kernel(ConvGeom g, const uint64_t* __restrict__ lx,
                                                          const uint64_t* __restrict__ lo,
                                                          const uint8_t* __restrict__ ltm,
                                                          const int16_t* __restrict__ llast,
                                                          const float* __restrict__ tab2 /*[3][64][32]*/,
                                                          const float* __restrict__ lm /*[ky][kx][32]: plane 2*/,
                                                          const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ act1,
                                                          const int32_t* __restrict__ count_dev, int count_host,
                                                          int back) {
  __shared__ __align__(16) float st[C1T_TAB_FLOATS];
  __shared__ __align__(16) float slm[9 * 32];
  __shared__ __align__(16) float sb[32];
  __shared__ int16_t codes_all[C1T_WARPS][C1T_CODES];
  __shared__ int16_t out_row[AZP_MAX_BOARD * AZP_MAX_BOARD];        // cell -> padded row (r+1)*P + (c+1): no division per item
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int a = threadIdx.x; a < g.A; a += blockDim.x) out_row[a] = (int16_t)((a / g.B + 1) * g.P + (a % g.B + 1));
  for (int i = threadIdx.x; i < C1T_TAB_FLOATS / 4; i += blockDim.x)
    reinterpret_cast<float4*>(st)[i] = reinterpret_cast<const float4*>(tab2)[i];
  for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) slm[i] = lm[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  int16_t* codes = codes_all[warp];
  for (int i = lane; i < g.B; i += 32) { codes[i] = 0; codes[(g.B + 1) * g.B + i] = 0; }
  __syncthreads();
  pdl_wait();                      // leaf batch / leaf count are written by the descent kernel before us
  pdl_launch_dependents();
  const int count = count_dev ? *count_dev : count_host;
  const int cg = lane & 3;
  // TWO warps per board (each builds the codes, each writes half of the positions): 32 resident warps per SM hide the
  // dependent shared-memory loads of an item.  Board bi -> CTA bi % grid, warp pair (bi / grid) % 4.
  const int half = warp & 1;
  for (int bi = (warp >> 1) * gridDim.x + blockIdx.x; bi < count; bi += gridDim.x * (C1T_WARPS / 2)) {
    const int slot = back ? g.G - 1 - bi : bi;
    const bool xm = ltm[slot] == 1;
    const int last = llast[slot];
    // lanes 0..NW-1 hold the side to move's words, lanes NW..2NW-1 the opponent's: one coalesced load each
    uint64_t word = 0;
    if (lane < NWT) word = (xm ? lx : lo)[(size_t)slot * NWT + lane];
    else if (lane < 2 * NWT) word = (xm ? lo : lx)[(size_t)slot * NWT + lane - NWT];
    __syncwarp();                                            // the previous board's codes have been consumed
    for (int a0 = 0; a0 < g.A; a0 += 32) {
      const int a = a0 + lane;
      const int c = a % g.B;
      int code = 0;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int aa = min(max(a + kx - 1, 0), g.A - 1);     // clamped: every lane takes part in the shuffles
        const uint64_t wo = __shfl_sync(AZP_FULL, word, aa >> 6), wp = __shfl_sync(AZP_FULL, word, NWT + (aa >> 6));
        const int cc = c + kx - 1;
        if (a < g.A && cc >= 0 && cc < g.B) {
          const int b0 = (int)((wo >> (aa & 63)) & 1), b1 = (int)((wp >> (aa & 63)) & 1);
          code |= (b0 << kx) | (b1 << (3 + kx)) | ((aa == last ? 1 : 0) << (6 + kx));
        }
      }
      if (a < g.A) codes[g.B + a] = (int16_t)code;
    }
    __syncwarp();
    __nv_bfloat16* base = act1 + (size_t)slot * g.stride * 32;
    const int a_half = (g.A + 1) / 2 & ~7;                   // warp 0 of the pair: cells [0, a_half), warp 1: the rest
    const int a_lo = half ? a_half : 0, a_hi = half ? g.A : a_half;
#pragma unroll 2
    for (int item = a_lo * 4 + lane; item < a_hi * 4; item += 32) {
      const int a = item >> 2;
      const int p = a & 1;                                   // neighbouring positions take opposite chunk halves per load
      const int ch0 = cg + 4 * p, ch1 = cg + 4 * (p ^ 1);    // 16-byte chunks = channels [4 ch, 4 ch + 4)
      const int16_t* cs = codes + a;                         // codes of board rows r-1, r, r+1 (zero border rows)
      const int c0 = cs[0], c1 = cs[g.B], c2 = cs[2 * g.B];
      float4 u = reinterpret_cast<const float4*>(sb)[ch0], v = reinterpret_cast<const float4*>(sb)[ch1];
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int code = ky == 0 ? c0 : (ky == 1 ? c1 : c2);
        const float4* row = reinterpret_cast<const float4*>(st + (ky * 64 + (code & 63)) * 32);
        const float4 t0 = row[ch0], t1 = row[ch1];
        u.x += t0.x; u.y += t0.y; u.z += t0.z; u.w += t0.w;
        v.x += t1.x; v.y += t1.y; v.z += t1.z; v.w += t1.w;
      }
      if ((c0 | c1 | c2) >> 6) {                             // the last move lies in this position's 3x3 window
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const int pat = (ky == 0 ? c0 : (ky == 1 ? c1 : c2)) >> 6;
          if (!pat) continue;
          const int kx = pat == 1 ? 0 : (pat == 2 ? 1 : 2);
          const float4* row = reinterpret_cast<const float4*>(slm + (ky * 3 + kx) * 32);
          const float4 t0 = row[ch0], t1 = row[ch1];
          u.x += t0.x; u.y += t0.y; u.z += t0.z; u.w += t0.w;
          v.x += t1.x; v.y += t1.y; v.z += t1.z; v.w += t1.w;
        }
      }
      __nv_bfloat162 a0 = __floats2bfloat162_rn(fmaxf(u.x, 0.f), fmaxf(u.y, 0.f)), a1 = __floats2bfloat162_rn(fmaxf(u.z, 0.f), fmaxf(u.w, 0.f));
      __nv_bfloat162 b0 = __floats2bfloat162_rn(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)), b1 = __floats2bfloat162_rn(fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
      uint8_t* rowp = reinterpret_cast<uint8_t*>(base + (size_t)out_row[a] * 32);
      *reinterpret_cast<uint2*>(rowp + ch0 * 8) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
      *reinterpret_cast<uint2*>(rowp + ch1 * 8) = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
    }
  }
}

// general float32 planes [n][4][A] -> act1, SAME summation order as conv1_table_kernel: per kernel row the two stone
// planes' column sums are added first (= the table entry), then the rows to the bias; the last-move plane and plane 3
// follow.  On 0/1 planes with an empty plane 3 every extra term is an exact zero, so both kernels agree bit for bit.
__global__ void __launch_bounds__(256) conv1_table_planes_kernel(ConvGeom g, int n, const float* __restrict__ planes,
                                                                 const float* __restrict__ w4 /*[tap][4][32]*/,
                                                                 const float* __restrict__ bias,
                                                                 __nv_bfloat16* __restrict__ act1) {
  __shared__ float sw[36 * 32];
  __shared__ float sb[32];
  extern __shared__ float sp[];          // [4][A]
  for (int i = threadIdx.x; i < 36 * 32; i += blockDim.x) sw[i] = w4[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * g.A; i += blockDim.x) sp[i] = planes[(size_t)slot * 4 * g.A + i];
    __syncthreads();
    for (int item = threadIdx.x; item < g.A * 4; item += blockDim.x) {
      const int a = item >> 2, cgp = item & 3;               // 8 channels per thread: [8 cgp, 8 cgp + 8)
      const int r = a / g.B, c = a - r * g.B;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = sb[cgp * 8 + j];
      auto column_sum = [&](int pl, int ky, float (&part)[8]) {
#pragma unroll
        for (int j = 0; j < 8; j++) part[j] = 0.f;
        const int rr = r + ky - 1;
        if (rr < 0 || rr >= g.B) return;
        for (int kx = 0; kx < 3; kx++) {
          const int cc = c + kx - 1;
          if (cc < 0 || cc >= g.B) continue;
          const float v = sp[pl * g.A + rr * g.B + cc];
          const float* wr = sw + ((ky * 3 + kx) * 4 + pl) * 32 + cgp * 8;
#pragma unroll
          for (int j = 0; j < 8; j++) part[j] = fmaf(v, wr[j], part[j]);
        }
      };
      float p0[8], p1[8];
      for (int ky = 0; ky < 3; ky++) {                       // stone planes: (own + opp) per kernel row, rows in order
        column_sum(0, ky, p0);
        column_sum(1, ky, p1);
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += p0[j] + p1[j];
      }
      for (int pl = 2; pl < 4; pl++)                         // last-move plane, then plane 3
        for (int ky = 0; ky < 3; ky++) {
          column_sum(pl, ky, p0);
#pragma unroll
          for (int j = 0; j < 8; j++) acc[j] += p0[j];
        }
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
        packed[j] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(act1 + ((size_t)slot * g.stride + (size_t)(r + 1) * g.P + (c + 1)) * 32 + cgp * 8) =
          make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// ---- tcgen05 implicit-GEMM 3x3 convolution ---------------------------------------------------------------
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 =
// epilogue (warp w owns TMEM lanes 32*(w%4)..+31 == tile rows).
// Per 128-row tile: three (dx = -1,0,+1) TMA loads of (128+2P) rows; each feeds 3 taps (dy) x CIN/16
// UMMAs (M=128, N=COUT, K=16).  Weights [tap][COUT][CIN] stay in shared memory for the kernel's lifetime.
template <int CIN, int COUT, int STAGES, bool HEADS>
struct ConvSmem {
  static constexpr int ROWB = CIN * 2;
  static constexpr int W_TAP_BYTES = COUT * ROWB;
  static constexpr int W_BYTES = 9 * W_TAP_BYTES;
};

constexpr int SCHED_GROUPS = 16;
struct ConvArgs {
  ConvGeom g;
  const float* bias;              // [COUT]
  __nv_bfloat16* out;             // !HEADS: activations [rows][COUT]
  const float* head_b;            // HEADS: [8] bias of the fused 1x1 head convs (policy rows first)
  __nv_bfloat16* feat;            // HEADS: [G][FK] bf16: policy features at ch*A+pix, value features at PKp+ch*A+pix
  int NJ, PC, PKp, FK;
  const int32_t* count_dev; int count_host; int back;
  int stage_rows;                 // 128 + 2P
  // BITS: the A operand is produced in-kernel from the leaf batch's bitboards (conv1 fused in)
  const uint64_t* lx; const uint64_t* lo; const uint8_t* ltm; const int16_t* llast; int NW;
  const float* w1;                // [plane(3)][ky][pattern(8)][32] fp32 partial sums of conv1 weights
  const float* b1;                // [32]
  int debug;                      // bring-up switches (AZP_DEBUG_CONV): 1 skip epilogue math, 2 skip conv MMAs, 4 skip A loads
  int single;                     // 0: three dx-shifted TMA loads per tile; 1/2: ONE load of 128+2P+2 rows, the nine taps are
                                  // row-offset descriptors into it (2 = descriptors carry the swizzle base offset)
  int stage_bytes;                // bytes of one ring stage
  // Plain conv (conv2): DYNAMIC tile scheduling.  Per launch a few CTAs lose 3-9 us to stalls no SM owns (per-CTA durations
  // average out to +-4 % over a thousand launches, the slowest CTA of a single launch takes 35 % longer than the median:
  // tools/step_timeline.py), so a static tile list makes the launch as long as its unluckiest CTA.  The CTAs form
  // SCHED_GROUPS groups (cta % SCHED_GROUPS) that each own the tiles their members would have had statically and hand them
  // out through their own counter sched[group] -- ONE chip-wide counter serialises 4096 same-address atomics at ~27 cycles
  // each (measured: the median CTA went from 23.6 to 27.9 us).  A CTA's first tile is its own index; sched[SCHED_GROUPS] =
  // CTAs that have left: the last one out resets everything for the next launch.  null = static list t = cta + k * grid.
  int* sched;
  const void* dead; int dead_row_bytes;     // activations of the layer BEFORE this kernel's input (dead once we run), or null
  int paired_stores;              // plain conv epilogue: AZP_PAIRED_STORES=1 lane pairs write whole 64-byte row segments (default: one row half per lane)
  long long* clk;                 // bring-up: per-tile clock64 stamps of CTA clk_cta ([tile][16]) or null
  int clk_cta;                    // bring-up: which CTA stamps (AZP_CLK_CTA, default 0)
  // HEADS only: the FC layers (policy_fc, value_fc1 + ReLU + value_fc2 + tanh; net.py:64-70) run INSIDE this kernel as
  // extra items of the accumulator rotation -- job = (board tile of 128 boards, 64-column slice of the policy outputs or
  // the value head) -- once the head features of the board tile have been written by every CTA's epilogue
  int fc_fused;                   // 0: the FC layers are a separate launch (fc_umma_kernel)
  int fc_slices, fc_pk_chunks, fc_vk_chunks, fc_v_off, fc_A;
  const float* fc_pbias; const float* fc_vbias; const float* fc_vw2; const float* fc_vb2;
  float* fc_logits; float* fc_value;
  unsigned int* feat_done;        // [board tiles] -> FC producer: one count per conv tile touching the board tile, features stored
  unsigned int* fc_done;          // [board tiles] FC epilogue -> consumers of the logits / values: one count per (job, lane quarter)
};

constexpr int FC_FUSED_STAGE_BYTES = TILE_M * 128 + 64 * 128;      // A: 128 boards x 64 K, B: 64 outputs x 64 K (bf16)
constexpr int FC_MAX_BOARD_TILES = 512;                            // 65535 leaf rows / 128

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A is read from tensor memory (lane = row, two bf16 per column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) umma_bf16_ts(tmem_d, tmem_a, desc_b, idesc, accumulate);
}

constexpr int CONV_THREADS = 384;     // warp 0 TMA, warps 1 and 11 MMA issuers (even / odd tiles), warps 2..9 epilogue
                                      // (2 per TMEM lane quarter), warp 10 head-GEMM issuer
constexpr int CONV_THREADS_BITS = 640; // + warps 12..19: conv1-from-bitboards producers of the A operand
constexpr int CONV_THREADS_EPI2 = 640; // plain conv (no heads): warps 12..19 = a second epilogue group (odd tiles)
constexpr int CODES = 256;            // >= 128 + 4P + 2 codes per tile (P <= 24)
constexpr int HEAD_N = 16;            // fused 1x1 head convs: 6 (or 3) outputs padded to the UMMA minimum N

// HEADS = true additionally fuses both 1x1 head convolutions (net.py:47,51) as a SECOND tensor-core
// GEMM per tile: the epilogue writes relu(acc + bias) as bf16 straight back into tensor memory, and a
// TS-mode UMMA (A from TMEM) multiplies it with the [16 x COUT] head weights; the wide activation never
// leaves the SM and the epilogue warps only touch each accumulator element once.
// BITS = true (conv2 only, CIN == 32) fuses conv1: eight producer warps compute relu(conv1) of the leaf
// batch straight from its bitboards (inputs are 0/1, so <= 27 weight-row adds per position) and write the
// three dx-shifted, 64B-swizzled A stages into shared memory themselves -- no act1 round trip through
// HBM/L2 and no TMA loads for the activations.
template <int CIN, int COUT, int STAGES, bool HEADS, bool BITS>
__global__ void __launch_bounds__(BITS ? CONV_THREADS_BITS : (HEADS ? CONV_THREADS : CONV_THREADS_EPI2), 1) conv_umma_kernel(const __grid_constant__ CUtensorMap map_act,
                                                                    const __grid_constant__ CUtensorMap map_w,
                                                                    const __grid_constant__ CUtensorMap map_hw,
                                                                    const __grid_constant__ CUtensorMap map_feat,
                                                                    const __grid_constant__ CUtensorMap map_wp,
                                                                    const __grid_constant__ CUtensorMap map_wv,
                                                                    const ConvArgs args) {
  using S = ConvSmem<CIN, COUT, STAGES, HEADS>;
  constexpr int ROWB = S::ROWB;
  constexpr int KSTEPS = CIN / 16;
  constexpr int HW_BYTES = HEADS ? (COUT / 64) * HEAD_N * 128 : 0;     // head weights: K chunks of 64 x 16 rows
  // TMEM columns: [0, 2*COUT) two conv accumulators; HEADS adds two bf16 copies of the activated tile
  // (COUT/2 columns each) and two head accumulators (HEAD_N columns each), all double-buffered so the
  // epilogue of tile t never waits for the head GEMM of tile t-1
  // The head GEMM's K loop is split over HEAD_CHAINS independent accumulators (summed by the epilogue):
  // back-to-back UMMAs into ONE narrow accumulator serialise on its ~64-cycle dependent-issue latency
  constexpr int HEAD_CHAINS = 4;
  // The plain conv (conv2) is epilogue-latency bound with one epilogue group: it gets FOUR accumulators and TWO
  // epilogue groups (even / odd tiles) so two tiles are drained concurrently.
  // HEADS: FOUR accumulators of COUT columns and nothing else -- the bf16 copy of the activated tile (A operand of the
  // head GEMM) and the head GEMM's partial accumulators ALIAS the accumulator being drained: channel half h keeps its
  // 32 packed columns at [h*64, h*64+32) (written over fp32 columns that are already in registers), the four head
  // chains live in the freed columns [32,64) and [96,128).  An accumulator is released only after its head outputs
  // were read, but with four in rotation the issuers never wait for an epilogue (they did with two: 9 us per launch).
  constexpr int NACC = BITS ? 2 : 4;
  constexpr int EGROUPS = (HEADS || BITS) ? 1 : 2;
  static_assert(!HEADS || (COUT == 128 && HEAD_CHAINS * HEAD_N == 64), "aliased head layout assumes 128 output channels");
  constexpr uint32_t NEED = NACC * COUT;
  constexpr uint32_t TMEM_COLS = NEED <= 32 ? 32 : (NEED <= 64 ? 64 : (NEED <= 128 ? 128 : (NEED <= 256 ? 256 : 512)));
  constexpr int HALF = COUT / 2;                            // channels per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, still a shared pointer
  const int stage_bytes = args.stage_bytes;                 // multiple of the swizzle pattern size
  uint8_t* s_w = smem;                                      // 9 taps x [COUT][CIN]
  uint8_t* s_a = s_w + S::W_BYTES;                          // STAGES x [stage_rows][CIN]
  uint8_t* s_hw = s_a + (size_t)STAGES * stage_bytes;       // HEADS: [COUT/64][16][64] bf16, 128B swizzle
  uint8_t* s_out = s_hw + HW_BYTES;                         // plain conv: EGROUPS x [128 rows][64 ch] bf16 output staging
  uint8_t* s_tail = s_out + (HEADS ? 0 : EGROUPS * TILE_M * 128);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_tail);     // [STAGES]
  uint64_t* empty = full + STAGES;                          // [STAGES]
  uint64_t* tfull = empty + STAGES;                         // [NACC <= 4]
  uint64_t* tempty = tfull + 4;                             // [NACC <= 4]
  uint64_t* wbar = tempty + 4;                              // [1]
  uint64_t* act_full = wbar + 1;                            // [4] epilogue -> MMA: bf16 tile of accumulator a is in TMEM
  uint64_t* head_full = act_full + 4;                       // [4] MMA -> epilogue: head outputs of accumulator a ready
  uint64_t* hw_done = head_full + 4;                        // [4] epilogue -> signaller: head features of the item are stored
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hw_done + 4);
  // dynamic scheduling: tile of local item tl at [tl & 15] (-1 = no more work).  NOT a static __shared__ array: with one in
  // the kernel the dynamic window no longer starts at offset 0 while the compiler still folds the manual 1024-byte alignment
  // above away on the strength of __align__(1024) -- every swizzled TMA / UMMA address is then misaligned
  int* s_tile = reinterpret_cast<int*>(tmem_slot + 2);      // [16]
  float* s_bias = reinterpret_cast<float*>(s_tile + 16);    // [COUT] (+ [8] head bias)
  float* s_w1 = s_bias + COUT + 8;                          // BITS: [3 planes][3 ky][8 patterns][32] conv1 tables + [32] bias
  int16_t* s_cells = reinterpret_cast<int16_t*>(s_w1 + 72 * 32 + 32); // BITS: [2][CODES] column-pattern codes

  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const ConvGeom g = args.g;
  auto wall_ns = [] { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); return (long long)gt; };
  if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && threadIdx.x == 0) AZP_CLKPTR(args.clk)[0 * 16 + 14] = wall_ns();      // kernel entry
  [[maybe_unused]] constexpr int TL_KID = HEADS ? AZP_TL_CONV3 : AZP_TL_CONV2;
  const unsigned tl_idx = threadIdx.x == 0 ? AZP_TL_ENTRY(TL_KID, gridDim.x) : 0u;

  // ---- prologue: touches nothing the previous kernel of the stream writes, so under programmatic dependent
  //      launch it overlaps that kernel's tail (barriers, TMEM, bias and the resident weights) ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], BITS ? 8 : 1); mbar_init(&empty[i], 1); }
    // HEADS: an accumulator is released by the four warps that read its head outputs; otherwise by all eight
    for (int i = 0; i < NACC; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], HEADS ? 4 : 8); }
    mbar_init(wbar, 1);
    for (int i = 0; i < 4; i++) { mbar_init(&act_full[i], 8); mbar_init(&head_full[i], 1); mbar_init(&hw_done[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(wbar, S::W_BYTES + HW_BYTES);
    for (int tap = 0; tap < 9; tap++) tma_load_2d(s_w + tap * S::W_TAP_BYTES, &map_w, wbar, 0, tap * COUT);
    if (HEADS)
      for (int kc = 0; kc < COUT / 64; kc++) tma_load_2d(s_hw + kc * HEAD_N * 128, &map_hw, wbar, kc * 64, 0);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_bias[i] = args.bias[i];
  if (HEADS && threadIdx.x < 8) s_bias[COUT + threadIdx.x] = args.head_b[threadIdx.x];
  if (threadIdx.x < 16) s_tile[threadIdx.x] = -1;           // no item's tag (see tile_of)
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && threadIdx.x == 0) AZP_CLKPTR(args.clk)[1 * 16 + 14] = wall_ns();      // barriers + TMEM ready
  pdl_wait();                      // the leaf count and the input activations are the previous kernels' outputs
  pdl_launch_dependents();
  if (threadIdx.x == 0) AZP_TL_START(TL_KID, tl_idx);
  if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && threadIdx.x == 0) AZP_CLKPTR(args.clk)[2 * 16 + 14] = wall_ns();      // predecessor finished
  const int count = warp_uniform(args.count_dev ? *args.count_dev : args.count_host);
  const int slot0 = args.back ? g.G - count : 0;
  const long long row_begin = (long long)slot0 * g.stride;
  const long long rows_total = (long long)count * g.stride;
  const int n_tiles = (int)((rows_total + TILE_M - 1) / TILE_M);
  if (args.dead) discard_l2(static_cast<const char*>(args.dead) + row_begin * args.dead_row_bytes, (long long)n_tiles * TILE_M * args.dead_row_bytes);
  // ---- work list of this CTA: its conv tiles (t = blockIdx.x + k * gridDim.x) and then, with the FC layers fused in, its
  //      share of the FC jobs j = (board tile, slice); the jobs go to the CTAs with the FEWEST conv tiles first
  const int n_cta = (int)gridDim.x, cta = (int)blockIdx.x;
  const int n_conv = cta < n_tiles ? (n_tiles - 1 - cta) / n_cta + 1 : 0;
  const bool fc_on = HEADS && !BITS && args.fc_fused != 0;
  const int fc_per_bt = args.fc_slices + 1;                       // policy slices + the value job
  const int fc_jobs = fc_on ? ((count + TILE_M - 1) / TILE_M) * fc_per_bt : 0;
  const int fc_first = n_cta - 1 - cta;
  const int n_fc = fc_first < fc_jobs ? (fc_jobs - 1 - fc_first) / n_cta + 1 : 0;
  const int n_items = n_conv + n_fc;
  const bool dyn = !HEADS && !BITS && args.sched != nullptr && args.single != 0;      // never together with the fused FC jobs
  // Dynamic scheduling: ring entry of local item tl = (tl & 0x7FF) << 20 | (tile + 1), tile + 1 == 0: no more work.  The
  // producer publishes an entry when it requests the tile's data, several items before anyone needs it, so the consumers
  // POLL the tag instead of learning the tile behind a barrier: the epilogue has its row arithmetic done before the
  // accumulator is ready, as with the static list (behind the barrier it cost every CTA +18 %).
  auto tile_of = [&](int tl) -> int {
    const volatile int* e = s_tile + (tl & 15);
    const int tag = tl & 0x7FF;
    int v = *e;
    if (((unsigned)v >> 20) != (unsigned)tag) {
      const long long t0 = clock64();
      while ((((unsigned)(v = *e)) >> 20) != (unsigned)tag)
        if (clock64() - t0 > 4000000000ll) __trap();
    }
    return (v & 0xFFFFF) - 1;
  };
  if (!BITS && !HEADS && blockIdx.x == 0 && args.feat_done) {
    // conv2 runs between two conv3 launches: it resets the counters of the fused FC jobs (every consumer of the previous
    // evaluation has finished -- this kernel is past its griddepcontrol.wait)
    for (int i = threadIdx.x; i < FC_MAX_BOARD_TILES; i += blockDim.x) { args.feat_done[i] = 0u; args.fc_done[i] = 0u; }
  }
  // item tl of the work list: conv tile (returns true, sets t) or FC job (false, sets bt / sl; sl == fc_slices: value head)
  auto item_is_conv = [&](int tl, int& t, int& bt, int& sl) -> bool {
    if (tl < n_conv) { t = cta + tl * n_cta; return true; }
    const int j = fc_first + (tl - n_conv) * n_cta;
    bt = j / fc_per_bt; sl = j - bt * fc_per_bt; t = 0;
    return false;
  };
  auto fc_chunks = [&](int sl) -> int { return sl == args.fc_slices ? args.fc_vk_chunks : args.fc_pk_chunks; };
  // bring-up: wall-clock stamps of the first FC job of the LAST CTA (row 48 of the stamp buffer)
  auto fc_stamp = [&](int k) {
    if (AZP_CLKPTR(args.clk) && blockIdx.x == gridDim.x - 1) AZP_CLKPTR(args.clk)[48 * 16 + k] = wall_ns();
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      if (!BITS && args.single) {
        // one box of 128+2P+2 rows per tile, starting one row above/left of the dy=-1,dx=-1 window
        const uint32_t bytes = (uint32_t)(args.stage_rows + 2) * ROWB;
        if (dyn) {
          // every tile index is requested ONE item ahead, so the atomic's round trip hides behind the stage wait; two
          // consecutive end markers, one for each issuer / epilogue-group parity
          // group gi = cta % SCHED_GROUPS has m members; its k-th tile is round k / m of member k % m
          const int gi = cta % SCHED_GROUPS, m = (n_cta - gi + SCHED_GROUPS - 1) / SCHED_GROUPS;
          auto next_tile = [&]() {
            const int k = m + atomicAdd(args.sched + gi, 1);
            const int r = k / m;
            return r * n_cta + gi + (k - r * m) * SCHED_GROUPS;
          };
          int t = cta, t_next = next_tile();
          for (int tl = 0, ends = 0; ends < 2; tl++) {
            const int st = tl % STAGES;
            mbar_wait(&empty[st], ((tl / STAGES) & 1) ^ 1);
            *reinterpret_cast<volatile int*>(s_tile + (tl & 15)) = ((tl & 0x7FF) << 20) | (t < n_tiles ? t + 1 : 0);
            if (t < n_tiles) {
              mbar_expect_tx(&full[st], bytes);
              const long long row = row_begin + (long long)t * TILE_M - g.P - 1;
              tma_load_2d(s_a + (size_t)st * stage_bytes, &map_act, &full[st], 0, (int)row);
            } else {
              mbar_arrive(&full[st]);
              ends++;
            }
            t = t_next;
            if (t < n_tiles) t_next = next_tile();
          }
        }
        for (int tl = 0; !dyn && tl < n_items; tl++) {
          int t, bt, sl;
          if (item_is_conv(tl, t, bt, sl)) {
            const int st = it % STAGES;
            if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta) AZP_CLKPTR(args.clk)[it * 16 + 0] = clock64();
            mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1);
            if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta) AZP_CLKPTR(args.clk)[it * 16 + 1] = clock64();
            if (AZP_DBGBITS(args.debug) & 4) { mbar_arrive(&full[st]); it++; continue; }
            mbar_expect_tx(&full[st], bytes);
            const long long row = row_begin + (long long)t * TILE_M - g.P - 1;
            tma_load_2d(s_a + (size_t)st * stage_bytes, &map_act, &full[st], 0, (int)row);
            it++;
            continue;
          }
          // FC job: its A operand = the head features of boards [128 bt, 128 bt + 128), written by the epilogues of every
          // CTA whose conv tiles touch those boards -- one count per such tile
          const int nb = min(TILE_M, count - bt * TILE_M);
          const long long lo = (long long)bt * TILE_M * g.stride, hi = ((long long)bt * TILE_M + nb) * g.stride - 1;
          const int t_lo = (int)(lo / TILE_M), t_hi = min(n_tiles - 1, (int)(hi / TILE_M));
          const unsigned target = (unsigned)(t_hi - t_lo + 1);
          if (tl == n_conv) fc_stamp(0);
          {
            const long long t0 = clock64();
            while (ld_acquire_u32(args.feat_done + bt) < target) {
              __nanosleep(64);
              if (clock64() - t0 > 4000000000ll) __trap();
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");         // generic-proxy stores (acquired above) -> TMA reads
          if (tl == n_conv) fc_stamp(1);
          const bool is_value = sl == args.fc_slices;
          const CUtensorMap* mw = is_value ? &map_wv : &map_wp;
          const int k_off = is_value ? args.fc_v_off : 0, kcs = fc_chunks(sl);
          for (int kc = 0; kc < kcs; kc++, it++) {
            const int st = it % STAGES;
            mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1);
            mbar_expect_tx(&full[st], (uint32_t)FC_FUSED_STAGE_BYTES);
            uint8_t* sa = s_a + (size_t)st * stage_bytes;
            tma_load_2d(sa, &map_feat, &full[st], k_off + kc * 64, bt * TILE_M);
            tma_load_2d(sa + TILE_M * 128, mw, &full[st], kc * 64, is_value ? 0 : sl * 64);     // rows past the matrix: zero fill
          }
          if (tl == n_conv) fc_stamp(2);
        }
      }
      for (int t = blockIdx.x; !BITS && !args.single && t < n_tiles; t += gridDim.x) {
        for (int dxi = 0; dxi < 3; dxi++, it++) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          if (AZP_DBGBITS(args.debug) & 4) { mbar_arrive(&full[st]); continue; }
          mbar_expect_tx(&full[st], (uint32_t)stage_bytes);
          const long long row = row_begin + (long long)t * TILE_M - g.P + (dxi - 1);
          tma_load_2d(s_a + (size_t)st * stage_bytes, &map_act, &full[st], 0, (int)row);
        }
      }
    }
  } else if (BITS && warp >= 12) {
    // ===== A-operand producers: conv1 (4->32, 3x3) + ReLU from bitboards, written as ONE swizzled stage of
    //       128+2P+2 rows per tile (the nine taps are row-offset descriptors into it) =====
    const int pt = threadIdx.x - CONV_THREADS;               // 0..255
    for (int i = pt; i < 72 * 32 + 32; i += 256) s_w1[i] = i < 72 * 32 ? args.w1[i] : args.b1[i - 72 * 32];
    int it = 0, tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++, it++) {
      const long long R0 = row_begin + (long long)t * TILE_M;
      const long long rlo = R0 - g.P - 1;                    // first padded row of the stage
      const int rows_needed = args.stage_rows + 2;
      int16_t* cb = s_cells + (tl & 1) * CODES;
      // One code per padded position in [rlo - P, rlo + rows_needed + P): per input plane (0 = side to
      // move's stones, 1 = opponent's, 2 = last move; games.py:86-129) the 3-bit pattern of columns
      // c-1, c, c+1, plus bit 9 = "interior cell".  Halo positions get 0, so vertical neighbours are simply
      // the codes at +-P and no bounds checks remain in the convolution below.
      for (int i = pt; i < rows_needed + 2 * g.P; i += 256) {
        const long long qrow = rlo - g.P + i;
        int code = 0;
        if (qrow >= 0) {
          const int slot = (int)(qrow / g.stride), pos = (int)(qrow % g.stride);
          const int y = pos / g.P, x = pos % g.P;
          if (y >= 1 && x >= 1 && x <= g.B && slot < g.G) {
            const int c = x - 1, a = (y - 1) * g.B + c;
            const bool xm = args.ltm[slot] == 1;
            const uint64_t* own = (xm ? args.lx : args.lo) + (size_t)slot * args.NW;
            const uint64_t* opp = (xm ? args.lo : args.lx) + (size_t)slot * args.NW;
            const int last = args.llast[slot];
            code = 0x200;
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
              const int cc = c + kx - 1, aa = a + kx - 1;
              if (cc < 0 || cc >= g.B) continue;
              const int b0 = (int)((own[aa >> 6] >> (aa & 63)) & 1), b1 = (int)((opp[aa >> 6] >> (aa & 63)) & 1);
              code |= (b0 << kx) | (b1 << (3 + kx)) | ((aa == last ? 1 : 0) << (6 + kx));
            }
          }
        }
        cb[i] = (int16_t)code;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");         // codes (and, first time, the conv1 tables) visible
      const int st = it % STAGES;
      if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && pt == 0) AZP_CLKPTR(args.clk)[tl * 16 + 0] = clock64();
      mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1);
      if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && pt == 0) AZP_CLKPTR(args.clk)[tl * 16 + 1] = clock64();
      for (int idx = pt; idx < rows_needed; idx += 256) {
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; j++) packed[j] = 0u;
        const int16_t* cs = cb + idx;                        // cs[ky * P] = code of the cell one row above/at/below
        if (cs[g.P] & 0x200) {
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; j++) acc[j] = s_w1[72 * 32 + j];
          // conv1 = bias + sum over (plane, kernel row) of a table row indexed by the 3-bit column pattern
          // The two stone planes are added branch-free (pattern 0 selects an all-zero table row, and
          // x + 0.0f == x), so all table loads of a position are in flight together; the last-move plane is
          // non-zero for at most nine positions of a board and keeps its branch.
          const int c0 = cs[0], c1 = cs[g.P], c2 = cs[2 * g.P];
#pragma unroll
          for (int pl = 0; pl < 3; pl++) {
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
              const int pat = ((ky == 0 ? c0 : (ky == 1 ? c1 : c2)) >> (3 * pl)) & 7;
              if (pl == 2 && !pat) continue;
              const float4* tr = reinterpret_cast<const float4*>(s_w1 + ((pl * 3 + ky) * 8 + pat) * 32);
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const float4 w4 = tr[j];
                acc[4 * j] += w4.x; acc[4 * j + 1] += w4.y; acc[4 * j + 2] += w4.z; acc[4 * j + 3] += w4.w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; j++) {
            __nv_bfloat162 hh = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
            packed[j] = *reinterpret_cast<uint32_t*>(&hh);
          }
        }
        uint8_t* rowp = s_a + (size_t)st * stage_bytes + idx * 64;
        const int sw = (idx >> 1) & 3;                       // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,8]
#pragma unroll
        for (int c = 0; c < 4; c++)
          *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) =
              make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
      if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && pt == 0) AZP_CLKPTR(args.clk)[tl * 16 + 6] = clock64();
    }
  } else if (warp == 1 || warp == 11) {
    // ===== MMA issuers: two threads, even / odd tiles.  A single thread alternates "wait for barriers" and
    //       "issue" and leaves the tensor pipe idle ~10-30 % of a tile; with two, one waits while the other
    //       issues (each owns one of the two TMEM accumulators).  The whole warp runs the loop, an elected lane issues
    //       (see elect_one) =====
    {
      const uint32_t idesc = umma_idesc(COUT);
      const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
      const int my_parity = warp == 1 ? 0 : 1;
      mbar_wait(wbar, 0);
      if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && warp == 1 && lane == 0) AZP_CLKPTR(args.clk)[3 * 16 + 14] = wall_ns();           // weights resident
      tc_fence_after();
      int it = 0;
      for (int tl = 0; dyn || tl < n_items; tl++) {
        int t = 0, bt = 0, sl = 0;
        const bool is_conv = dyn ? true : item_is_conv(tl, t, bt, sl);
        if ((tl & 1) != my_parity) { it += is_conv ? (args.single ? 1 : 3) : fc_chunks(sl); continue; }
        const int acc = tl % NACC;
        const uint32_t acc_ph = (tl / NACC) & 1;
        if (dyn) {
          // the stage's tile index arrives with its data; an end marker is handed on to this parity's epilogue group
          mbar_wait(&tempty[acc], acc_ph ^ 1);
          mbar_wait(&full[it % STAGES], (it / STAGES) & 1);
          if (warp_uniform(tile_of(tl)) < 0) {
            if (elect_one()) mbar_arrive(&tfull[acc]);
            break;
          }
        }
        if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && lane == 0) AZP_CLKPTR(args.clk)[tl * 16 + 2] = clock64();
        mbar_wait(&tempty[acc], acc_ph ^ 1);              // epilogue has drained this accumulator
        if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && lane == 0) AZP_CLKPTR(args.clk)[tl * 16 + 3] = clock64();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * COUT);
        uint32_t accumulate = 0;
        if (!is_conv) {
          // FC job: [128 boards x 64 K] x [64 outputs x 64 K]^T per chunk into the first 64 columns of the accumulator --
          // the instruction sequence of fc_umma_kernel, so both give bit-identical logits / values
          const uint32_t idesc_fc = umma_idesc(64);
          const int kcs = fc_chunks(sl);
          for (int kc = 0; kc < kcs; kc++, it++) {
            const int st = it % STAGES;
            mbar_wait(&full[st], (it / STAGES) & 1);
            if (tl == n_conv && kc == 0 && lane == 0) fc_stamp(3);
            tc_fence_after();
            const uint32_t a_addr = a_base + (uint32_t)st * (uint32_t)stage_bytes, b_addr = a_addr + TILE_M * 128;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              umma_bf16_e(d_tmem, umma_desc(a_addr + k * 32, 128), umma_desc(b_addr + k * 32, 128), idesc_fc, accumulate);
              accumulate = 1;
            }
            umma_commit_e(&empty[st]);
          }
          umma_commit_e(&tfull[acc]);
          if (tl == n_conv && lane == 0) fc_stamp(4);
          continue;
        }
        if (args.single) {
          const int st = it % STAGES;
          mbar_wait(&full[st], (it / STAGES) & 1);
          if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && lane == 0) AZP_CLKPTR(args.clk)[tl * 16 + 4] = clock64();
          tc_fence_after();
          const uint32_t a_stage = a_base + (uint32_t)st * (uint32_t)stage_bytes;
#pragma unroll
          for (int dxi = 0; dxi < 3; dxi++) {
#pragma unroll
            for (int dyi = 0; dyi < 3; dyi++) {
              const uint32_t a_tap = a_stage + (uint32_t)((dyi * g.P + dxi) * ROWB);   // row offset dy*P + dx
              const uint32_t w_tap = w_base + (uint32_t)((dyi * 3 + dxi) * S::W_TAP_BYTES);
#pragma unroll
              for (int k = 0; k < KSTEPS; k++) {
                if ((AZP_DBGBITS(args.debug) & 2) && (dxi | dyi | k)) continue;       // bring-up: one UMMA per tile instead of 9 * KSTEPS
                umma_bf16_e(d_tmem, umma_desc(a_tap + k * 32, ROWB, args.single), umma_desc(w_tap + k * 32, ROWB), idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          umma_commit_e(&empty[st]);
          it++;
        }
        for (int dxi = 0; !args.single && dxi < 3; dxi++, it++) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t a_stage = a_base + (uint32_t)st * (uint32_t)stage_bytes;
#pragma unroll
          for (int dyi = 0; dyi < 3; dyi++) {
            const uint32_t a_tap = a_stage + (uint32_t)(dyi * g.P * ROWB);
            const uint32_t w_tap = w_base + (uint32_t)((dyi * 3 + dxi) * S::W_TAP_BYTES);
#pragma unroll
            for (int k = 0; k < KSTEPS; k++) {
              if (AZP_DBGBITS(args.debug) & 2) break;
              umma_bf16_e(d_tmem, umma_desc(a_tap + k * 32, ROWB), umma_desc(w_tap + k * 32, ROWB), idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_e(&empty[st]);                        // stage reusable once these MMAs retire
        }
        umma_commit_e(&tfull[acc]);                         // accumulator complete
        if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && lane == 0) {
          AZP_CLKPTR(args.clk)[tl * 16 + 5] = clock64();
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          AZP_CLKPTR(args.clk)[tl * 16 + 13] = (long long)gt;          // wall-clock ns next to the cycle stamp: SM clock under load
        }
      }
    }
  } else if (HEADS && warp == 10) {
    // ===== head-GEMM issuer: its own thread, so the conv issue loop never waits for the epilogue =====
    {
      const uint32_t idesc_head = umma_idesc(HEAD_N);
      const uint32_t hw_base = smem_u32(s_hw);
      // 1x1 head GEMM of local tile j: [128 x COUT] (TMEM, bf16) x [COUT x 16] (smem) -> TMEM cols COL_HEAD..
      auto issue_head = [&](int j) {
        const int b = j % NACC;
        mbar_wait(&act_full[b], (uint32_t)((j / NACC) & 1));
        tc_fence_after();
        const uint32_t acc_base = tmem_base + (uint32_t)(b * COUT);
#pragma unroll
        for (int k = 0; k < COUT / 16; k++) {
          const int c = k % HEAD_CHAINS;                    // chain c accumulates K steps c, c+4
          umma_bf16_ts_e(acc_base + (c < 2 ? 32 + c * HEAD_N : 96 + (c - 2) * HEAD_N),
                       acc_base + (k / 4) * 64 + (k % 4) * 8,
                       umma_desc(hw_base + (k / 4) * HEAD_N * 128 + (k % 4) * 32, 128), idesc_head,
                       k >= HEAD_CHAINS ? 1u : 0u);
        }
        umma_commit_e(&head_full[b]);
      };
      // This thread is also the SIGNALLER of the fused FC jobs: the four epilogue warps that store an item's head features
      // arrive on hw_done (release / acquire at CTA scope), then this otherwise idle thread makes the stores visible
      // device-wide with ONE fence and counts the tile in for every board tile it touches (the FC producers acquire the
      // sum).  A fence in the epilogue warps themselves cost 17 % of conv3: they are the busiest threads of the kernel.
      auto signal_item = [&](int j) {
        if (lane == 0) {
          mbar_wait(&hw_done[j % NACC], (uint32_t)((j / NACC) & 1));
          if (fc_on && j < n_conv) {
            __threadfence();
            const unsigned R0 = (unsigned)row_begin + (unsigned)(cta + j * n_cta) * TILE_M;
            const int s_first = (int)(R0 / (unsigned)g.stride);
            const int s_last = min(count - 1, (int)((R0 + TILE_M - 1) / (unsigned)g.stride));
            atomicAdd(args.feat_done + (s_first >> 7), 1u);
            if ((s_last >> 7) != (s_first >> 7)) atomicAdd(args.feat_done + (s_last >> 7), 1u);
          }
        }
        __syncwarp();
      };
      mbar_wait(wbar, 0);
      tc_fence_after();
      for (int tl = 0; tl < n_items; tl++) {
        if (tl < n_conv) {
          issue_head(tl);
          if (tl > 0) signal_item(tl - 1);          // the epilogue writes the heads of tile tl-1 right after the math of tile tl
        } else {
          // FC job: no head GEMM, but the barrier pair keeps its one-phase-per-item rhythm (the epilogue has read the
          // job's outputs out of the accumulator when its eight warps have arrived).  The previous item's features go
          // out FIRST: this very job may be waiting for them.
          if (tl > 0) signal_item(tl - 1);
          const int b = tl % NACC;
          mbar_wait(&act_full[b], (uint32_t)((tl / NACC) & 1));
          if (lane == 0) mbar_arrive(&head_full[b]);
        }
        if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && lane == 0) AZP_CLKPTR(args.clk)[tl * 16 + 6] = clock64();
      }
      if (n_items > 0) signal_item(n_items - 1);
    }
  } else if ((warp >= 2 && warp < 10) || (EGROUPS == 2 && warp >= 12 && warp < 20)) {
    // ===== epilogue: warp (q, h) owns TMEM lanes 32q..32q+31 (tile rows) and channel half h; with two groups,
    //       group 0 (warps 2..9) drains even tiles and group 1 (warps 12..19) odd tiles =====
    const int egroup = warp >= 12 ? 1 : 0;
    const int q = warp & 3;
    const int h = (warp - (egroup ? 12 : 2)) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // padded row -> (slot, y, x) without integer divisions (they cost ~100 instructions per tile in warps whose dependent
    // instruction chain IS the tile period): slot by a 32-bit reciprocal with one correction step, y = pos / P exact for
    // pos < 2730 (pos < 25 * 24)
    const unsigned e_stride = (unsigned)g.stride, e_stride_rcp = (unsigned)(0x100000000ull / (unsigned)g.stride);
    const unsigned e_p_rcp = (65536u + (unsigned)g.P - 1u) / (unsigned)g.P;
    auto row_split = [&](unsigned R, int& slot, int& y, int& x) {
      unsigned sq = __umulhi(R, e_stride_rcp);
      if (R - sq * e_stride >= e_stride) sq++;
      const unsigned pos = R - sq * e_stride;
      const unsigned yy = (pos * e_p_rcp) >> 16;
      slot = (int)sq; y = (int)yy; x = (int)(pos - yy * (unsigned)g.P);
    };
    float bias_r[HEADS ? 1 : HALF];         // plain conv: this warp's 32 output-channel biases live in registers
    if (!HEADS) {
#pragma unroll
      for (int j = 0; j < (HEADS ? 1 : HALF); j++) bias_r[j] = s_bias[h * HALF + j];
    }
    // head outputs of item j of the work list (conv tile tj): +bias, ReLU, scatter to feat[slot][j][pixel]; an FC job has
    // no head outputs -- its accumulator is simply handed back
    auto write_heads = [&](int j) {
      const int b = j % NACC;
      mbar_wait(&head_full[b], (uint32_t)((j / NACC) & 1));
      tc_fence_after();
      if (j >= n_conv) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&tempty[b]); mbar_arrive(&hw_done[b]); }
        return;
      }
      const int tj = cta + j * n_cta;
      const uint32_t acc_base = tmem_base + lane_addr + (uint32_t)(b * COUT);
      uint32_t r[8];
      tmem_ld8(acc_base + 32, r);
#pragma unroll
      for (int c = 1; c < HEAD_CHAINS; c++) {               // sum the partial accumulators of the K split
        uint32_t r2[8];
        tmem_ld8(acc_base + (c < 2 ? 32 + c * HEAD_N : 96 + (c - 2) * HEAD_N), r2);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);               // head outputs are in registers: the accumulator is free
      const unsigned R = (unsigned)row_begin + (unsigned)tj * TILE_M + q * 32 + lane;
      int slot, y, x;
      row_split(R, slot, y, x);
      if (y >= 1 && x >= 1 && x <= g.B && slot < g.G) {
        __nv_bfloat16* f = args.feat + (size_t)slot * args.FK + (y - 1) * g.B + (x - 1);
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j < args.NJ)
            f[j < args.PC ? j * g.A : args.PKp + (j - args.PC) * g.A] =
                __float2bfloat16(fmaxf(__uint_as_float(r[j]) + s_bias[COUT + j], 0.f));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&hw_done[b]);              // -> signaller (the head issuer's thread): fence + count
    };
    bool heads_pending = false;           // write_heads(tl - 1) still to be done when item tl starts
    for (int tl = 0; dyn || tl < n_items; tl++) {
      if (EGROUPS == 2 && (tl & 1) != egroup) continue;
      int t = 0, fbt = 0, fsl = 0;
      const bool is_conv = dyn ? true : item_is_conv(tl, t, fbt, fsl);
      const int acc = tl % NACC;
      const uint32_t acc_ph = (tl / NACC) & 1;
      if (dyn) {
        t = tile_of(tl);
        if (t < 0) break;
      }
      const bool stamp = AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && (warp == 2 || warp == 12) && lane == 0;
      if (HEADS && !is_conv) {
        // ---- FC job.  The features it waits for may be the ones THIS warp still has to write (the previous item's
        //      heads), so they go out first -- waiting for the job's accumulator before that could never end.
        const bool fstamp = tl == n_conv && warp == 2 && lane == 0;
        if (fstamp) fc_stamp(5);
        if (heads_pending && h == ((tl - 1) & 1)) write_heads(tl - 1);
        heads_pending = false;
        if (fstamp) fc_stamp(6);
        mbar_wait(&tfull[acc], acc_ph);
        if (fstamp) fc_stamp(7);
        tc_fence_after();
        if (h == 0) {
          const bool is_value = fsl == args.fc_slices;
          const int trow = q * 32 + lane;
          const int row = fbt * TILE_M + trow;                     // board = row of the evaluated batch
          const uint32_t fc_addr = tmem_base + lane_addr + (uint32_t)(acc * COUT);
          // every conv item of this CTA precedes its FC jobs and this warp group has drained them all: the resident conv
          // weights are dead, their shared memory stages the logits tile for a coalesced write-out
          float* stage = reinterpret_cast<float*>(s_w);
          constexpr int PITCH = 65;                                // odd pitch: conflict-free row-per-lane writes
          float sp[2] = {0.f, 0.f};
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            uint32_t r[32];
            tmem_ld32(fc_addr + hf * 32, r);
            if (!is_value) {
#pragma unroll
              for (int i = 0; i < 32; i++) stage[trow * PITCH + hf * 32 + i] = __uint_as_float(r[i]);
            } else {
              // value_fc1 + ReLU, value_fc2: the same two 32-column partial sums, in the same order, as fc_umma_kernel
#pragma unroll
              for (int i = 0; i < 32; i++)
                sp[hf] = fmaf(fmaxf(__uint_as_float(r[i]) + __ldg(args.fc_vbias + hf * 32 + i), 0.f),
                              __ldg(args.fc_vw2 + hf * 32 + i), sp[hf]);
            }
          }
          if (is_value) {
            if (row < count) args.fc_value[row] = tanhf((sp[0] + sp[1]) + __ldg(args.fc_vb2));
          } else {
            // raw logits (+bias); the softmax over all cells happens in the consumer (controller.py:49).  Warp w of the
            // group writes rows w, w+4, ..: 64 consecutive floats per row, two per lane
            asm volatile("bar.sync 5, 128;" ::: "memory");
            const int nrows = min(TILE_M, count - fbt * TILE_M);
            const int c0 = fsl * 64 + lane, c1 = c0 + 32;
            const float b0 = c0 < args.fc_A ? __ldg(args.fc_pbias + c0) : 0.f, b1 = c1 < args.fc_A ? __ldg(args.fc_pbias + c1) : 0.f;
            for (int rr = q; rr < nrows; rr += 4) {
              float* out = args.fc_logits + (size_t)(fbt * TILE_M + rr) * args.fc_A;
              if (c0 < args.fc_A) out[c0] = stage[rr * PITCH + lane] + b0;
              if (c1 < args.fc_A) out[c1] = stage[rr * PITCH + 32 + lane] + b1;
            }
            asm volatile("bar.sync 5, 128;" ::: "memory");         // the staging tile may be refilled by the next job
          }
        }
        if (fstamp) fc_stamp(8);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&act_full[acc]);
        if (h == 0) {
          __threadfence();                                         // release: this lane quarter's outputs of the job
          __syncwarp();
          if (lane == 0) atomicAdd(args.fc_done + fbt, 1u);
        }
        heads_pending = true;
        continue;
      }
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 8] = clock64();
      mbar_wait(&tfull[acc], acc_ph);
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 9] = clock64();
      tc_fence_after();
      const unsigned R = (unsigned)row_begin + (unsigned)t * TILE_M + q * 32 + lane;   // global padded row (< 2^31)
      bool interior = true;
      if (!HEADS) {
        int slot, y, x;
        row_split(R, slot, y, x);
        interior = y >= 1 && x >= 1 && x <= g.B && slot < g.G;
      }
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(acc * COUT + h * HALF);
      if (AZP_DBGBITS(args.debug) & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (HEADS) mbar_arrive(&act_full[acc]); else mbar_arrive(&tempty[acc]); }
        if (HEADS && heads_pending && h == ((tl - 1) & 1)) write_heads(tl - 1);
        heads_pending = true;
      } else if (!HEADS) {
        // bias + ReLU + halo zero in registers, the accumulator is released at once.  Each lane then owns 64 contiguous
        // bytes of its output row and writes them as two 256-bit stores (st.global.v8: full 32-byte sectors).  16-byte
        // row stores (one half-used sector per lane and instruction) bound this kernel once, and the shared-memory
        // staging + TMA store that replaced them costs 32 KB of shared-memory port time per tile in a kernel that is
        // bound by exactly that port
        static_assert(HEADS || HALF == 32, "plain-conv epilogue expects 64 output channels");
        uint32_t r[32];
        if (AZP_DBGBITS(args.debug) & 64) {
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = (uint32_t)(tl + j);        // bring-up: no TMEM load
        } else {
          tmem_ld32(taddr, r);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float v0 = fmaxf(__uint_as_float(r[2 * j]) + bias_r[HEADS ? 0 : 2 * j], 0.f);
          const float v1 = fmaxf(__uint_as_float(r[2 * j + 1]) + bias_r[HEADS ? 0 : 2 * j + 1], 0.f);
          __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
          packed[j] = interior ? *reinterpret_cast<uint32_t*>(&hh) : 0u;      // halo rows stay zero
        }
        if (!(AZP_DBGBITS(args.debug) & 16)) {
          // direct path (default; AZP_DEBUG_CONV=16 selects the staged TMA-store path, 0.9 % slower end to end): each lane owns 64 contiguous bytes of its output row -> two 256-bit stores (full 32 B sectors),
          // no shared-memory round trip (the staged path costs 32 KB of shared-memory port time per tile)
          uint8_t* gp = reinterpret_cast<uint8_t*>(args.out) + (size_t)R * 128 + h * 64;
          if ((AZP_DBGBITS(args.debug) & 32) && packed[0] != 0x12345678u) continue;      // bring-up: no global stores
          if (args.paired_stores) {
            store_rows64_paired(packed, gp, 128, true, true, lane);
          } else {
#pragma unroll
            for (int v = 0; v < 2; v++)
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                           ::"l"(gp + v * 32), "r"(packed[8 * v]), "r"(packed[8 * v + 1]), "r"(packed[8 * v + 2]),
                             "r"(packed[8 * v + 3]), "r"(packed[8 * v + 4]), "r"(packed[8 * v + 5]), "r"(packed[8 * v + 6]),
                             "r"(packed[8 * v + 7]) : "memory");
          }
          if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 11] = clock64();
          continue;
        }
        uint8_t* obuf = s_out + egroup * (TILE_M * 128);
        const int m = q * 32 + lane;
        // previous TMA store of this group has finished reading the buffer (its issuer waited before this barrier)
        asm volatile("bar.sync %0, 256;" ::"r"(3 + egroup) : "memory");
#pragma unroll
        for (int v = 0; v < 4; v++)
          *reinterpret_cast<uint4*>(obuf + m * 128 + (((h * 4 + v) ^ (m & 7)) << 4)) =
              make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 256;" ::"r"(3 + egroup) : "memory");
        if (h == 0 && q == 0 && lane == 0) {
          const unsigned R0 = (unsigned)row_begin + (unsigned)t * TILE_M;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"((uint64_t)&map_hw), "r"(smem_u32(obuf)), "r"(0), "r"((int)R0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 11] = clock64();
      } else {
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          uint32_t packed[16];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + h * HALF + c0);     // 16-byte aligned (see s_tail layout)
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 bb = b4[j];
            __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(__uint_as_float(r[4 * j]) + bb.x, 0.f),
                                                      fmaxf(__uint_as_float(r[4 * j + 1]) + bb.y, 0.f));
            __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(__uint_as_float(r[4 * j + 2]) + bb.z, 0.f),
                                                      fmaxf(__uint_as_float(r[4 * j + 3]) + bb.w, 0.f));
            packed[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
            packed[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
          // in place: the 16 packed columns of this chunk overwrite fp32 columns this warp already holds in registers
          tmem_st16(tmem_base + lane_addr + (uint32_t)(acc * COUT + h * HALF + c0 / 2), packed);
        }
        if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 10] = clock64();
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&act_full[acc]);
        if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 11] = clock64();
        // the PREVIOUS tile's head GEMM was queued behind this tile's conv MMAs: it has retired by now
        if (heads_pending && h == ((tl - 1) & 1)) write_heads(tl - 1);
        heads_pending = true;
        if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 12] = clock64();
      }
    }
    if (!HEADS && h == 0 && q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (HEADS && heads_pending && h == ((n_items - 1) & 1)) write_heads(n_items - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (AZP_CLKPTR(args.clk) && (int)blockIdx.x == args.clk_cta && threadIdx.x == 0) AZP_CLKPTR(args.clk)[4 * 16 + 14] = wall_ns();      // all roles done
  if (threadIdx.x == 0) fc_stamp(9);
  if (threadIdx.x == 0) AZP_TL_END(TL_KID, tl_idx, gridDim.x);
  if (dyn && threadIdx.x == 0 && atomicInc(reinterpret_cast<unsigned*>(args.sched) + SCHED_GROUPS, gridDim.x - 1) == gridDim.x - 1)
    for (int i = 0; i < SCHED_GROUPS; i++) args.sched[i] = 0;
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- conv3 + 1x1 heads on a CTA PAIR (tcgen05 cta_group::2) ------------------------------------------------------
// The single-CTA conv3 above is bound by the shared-memory port: an SS-mode M128 N128 K16 UMMA reads 4 KB of A and
// 4 KB of B per 64 tensor cycles = the whole 128 B/cycle.  Here two CTAs of a cluster (the two SMs of a TPC) execute
// ONE M256 N128 K16 UMMA: each SM supplies the A rows of ITS 128-row tile and only HALF of B (64 of the 128 output
// channels' weight rows), so a tile costs 4 + 2 KB per 64 cycles (96 B/cycle) and the resident weights shrink to
// 72 KB per SM (two more activation stages).  Protocol (CUTLASS's 2-SM scheme):
//   * both CTAs run a TMA producer for their own tile; the loads complete on the LEADER's full barrier (mapa), the
//     leader's producer posts the expected bytes of both;
//   * the leader's two issuer threads (even / odd tile pairs) issue every UMMA; tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" / "head accumulator ready" arrivals to the same barrier in both CTAs;
//   * the epilogue warps of BOTH CTAs arrive on the LEADER's "accumulator drained" / "bf16 tile written" barriers
//     (mbarrier.arrive on the mapa address), since only the leader issues;
//   * the head GEMM is a TS-mode cta_group::2 UMMA: A = each CTA's activated tile in ITS tensor memory, B = 8 of the 16
//     (padded) head-weight rows per CTA.
constexpr int C3P_STAGES = 5;
constexpr int C3P_CIN = 64, C3P_COUT = 128;
constexpr int C3P_W_TAP_BYTES = (C3P_COUT / 2) * C3P_CIN * 2;          // 64 rows x 128 B
constexpr int C3P_W_BYTES = 9 * C3P_W_TAP_BYTES;
constexpr int C3P_HW_BYTES = (C3P_COUT / 64) * (HEAD_N / 2) * 128;     // 2 K chunks x 8 rows x 128 B

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared variable of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c_inner, int c_row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(leader_bar), "r"(c_inner), "r"(c_row) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
// arrive (once the issuing thread's earlier UMMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor of the pair UMMA: D=f32, A=B=bf16, K-major, M=256 (128 rows per CTA), N
__device__ __forceinline__ uint32_t umma_idesc_pair(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__device__ __forceinline__ bool tp_limit_ok(int n_pairs, int ncl) { return (n_pairs + ncl - 1) / ncl <= 64; }

__global__ void __launch_bounds__(CONV_THREADS, 1) conv3_pair_kernel(const __grid_constant__ CUtensorMap map_act,
                                                                     const __grid_constant__ CUtensorMap map_w,
                                                                     const __grid_constant__ CUtensorMap map_hw,
                                                                     const ConvArgs args) {
  constexpr int CIN = C3P_CIN, COUT = C3P_COUT, STAGES = C3P_STAGES;
  constexpr int ROWB = CIN * 2;
  constexpr int KSTEPS = CIN / 16;
  constexpr uint32_t COL_ACT = 2 * COUT;
  constexpr uint32_t COL_HEAD = COL_ACT + COUT;
  constexpr int HEAD_CHAINS = 4;
  constexpr uint32_t TMEM_COLS = 512;                       // 256 (two accumulators) + 128 (two bf16 tiles) + 128 (head partials)
  constexpr int HALF = COUT / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = args.stage_bytes;
  uint8_t* s_w = smem;                                      // 9 taps x [64 of the 128 couts][CIN]
  uint8_t* s_a = s_w + C3P_W_BYTES;                         // STAGES x [stage_rows + 2][CIN]
  uint8_t* s_hw = s_a + (size_t)STAGES * stage_bytes;       // [2 K chunks][8 of the 16 head rows][64] bf16, 128B swizzle
  uint8_t* s_tail = s_hw + C3P_HW_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_tail);     // [STAGES]  (the leader's are used)
  uint64_t* empty = full + STAGES;                          // [STAGES]  per CTA, multicast commit
  uint64_t* tfull = empty + STAGES;                         // [2]       per CTA, multicast commit
  uint64_t* tempty = tfull + 2;                             // [2]       leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* wbar = tempty + 2;                              // [1]       leader's: weights of both CTAs resident
  uint64_t* act_full = wbar + 1;                            // [2]       leader's: 16 arrivals
  uint64_t* head_full = act_full + 2;                       // [2]       per CTA, multicast commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(head_full + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);  // [COUT] + [8] head bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom g = args.g;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  // ---- prologue (independent of the previous kernel's outputs) ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16);
      mbar_init(&act_full[i], 16); mbar_init(&head_full[i], 1);
    }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_bias[i] = args.bias[i];
  if (threadIdx.x < 8) s_bias[COUT + threadIdx.x] = args.head_b[threadIdx.x];
  __syncthreads();
  cluster_sync_all();                                       // both CTAs' barriers exist before anything signals across
  if (threadIdx.x == 0) {
    if (leader) mbar_expect_tx(wbar, 2 * (C3P_W_BYTES + C3P_HW_BYTES));
    const uint32_t lw = mapa_u32(wbar, 0);
    for (int tap = 0; tap < 9; tap++)
      tma_load_2d_pair(s_w + tap * C3P_W_TAP_BYTES, &map_w, lw, 0, tap * COUT + (int)rank * (COUT / 2));
    for (int kc = 0; kc < COUT / 64; kc++)
      tma_load_2d_pair(s_hw + kc * (HEAD_N / 2) * 128, &map_hw, lw, kc * 64, (int)rank * (HEAD_N / 2));
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();
  pdl_launch_dependents();
  const int count = args.count_dev ? *args.count_dev : args.count_host;
  const int slot0 = args.back ? g.G - count : 0;
  const long long row_begin = (long long)slot0 * g.stride;
  const long long rows_total = (long long)count * g.stride;
  const int n_tiles = (int)((rows_total + TILE_M - 1) / TILE_M);
  const int n_pairs = (n_tiles + 1) / 2;
  const int cid = (int)cluster_id_x(), ncl = (int)cluster_count_x();
  const bool stampc = AZP_CLKPTR(args.clk) && cid == 0 && leader && tp_limit_ok(n_pairs, ncl);
  // this CTA's tile of pair tp is 2*tp + rank (a phantom tile past the end loads zero-filled rows and is discarded)

  if (warp == 0) {
    // ===== TMA producer (both CTAs): one box of 128+2P+2 rows per tile into the own stage =====
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(args.stage_rows + 2) * ROWB;
      int it = 0;
      for (int tp = cid; tp < n_pairs; tp += ncl, it++) {
        const int st = it % STAGES;
        if (stampc) AZP_CLKPTR(args.clk)[it * 16 + 0] = clock64();
        mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1);
        if (stampc) AZP_CLKPTR(args.clk)[it * 16 + 1] = clock64();
        if (leader) mbar_expect_tx(&full[st], 2 * bytes);
        const long long row = row_begin + (long long)(2 * tp + (int)rank) * TILE_M - g.P - 1;
        tma_load_2d_pair(s_a + (size_t)st * stage_bytes, &map_act, mapa_u32(&full[st], 0), 0, (int)row);
      }
    }
  } else if (warp == 1 || warp == 11) {
    // ===== UMMA issuers (leader CTA only): even / odd tile pairs, one TMEM accumulator each =====
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_pair(COUT);
      const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
      const int my_parity = warp == 1 ? 0 : 1;
      mbar_wait(wbar, 0);
      tc_fence_after();
      int tl = 0;
      for (int tp = cid; tp < n_pairs; tp += ncl, tl++) {
        if ((tl & 1) != my_parity) continue;
        const int acc = tl & 1;
        if (stampc) AZP_CLKPTR(args.clk)[tl * 16 + 2] = clock64();
        mbar_wait(&tempty[acc], ((tl >> 1) & 1) ^ 1);          // both CTAs' epilogues have drained this accumulator
        if (stampc) AZP_CLKPTR(args.clk)[tl * 16 + 3] = clock64();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * COUT);
        const int st = tl % STAGES;
        mbar_wait(&full[st], (tl / STAGES) & 1);               // both CTAs' boxes have landed
        if (stampc) AZP_CLKPTR(args.clk)[tl * 16 + 4] = clock64();
        tc_fence_after();
        const uint32_t a_stage = a_base + (uint32_t)st * (uint32_t)stage_bytes;
        uint32_t accumulate = 0;
#pragma unroll
        for (int dxi = 0; dxi < 3; dxi++) {
#pragma unroll
          for (int dyi = 0; dyi < 3; dyi++) {
            const uint32_t a_tap = a_stage + (uint32_t)((dyi * g.P + dxi) * ROWB);
            const uint32_t w_tap = w_base + (uint32_t)((dyi * 3 + dxi) * C3P_W_TAP_BYTES);
#pragma unroll
            for (int k = 0; k < KSTEPS; k++) {
              umma_bf16_pair(d_tmem, umma_desc(a_tap + k * 32, ROWB, args.single), umma_desc(w_tap + k * 32, ROWB), idesc, accumulate);
              accumulate = 1;
            }
          }
        }
        umma_commit_pair(&empty[st]);
        umma_commit_pair(&tfull[acc]);
        if (stampc) AZP_CLKPTR(args.clk)[tl * 16 + 5] = clock64();
      }
    }
  } else if (warp == 10) {
    // ===== head-GEMM issuer (leader CTA only) =====
    if (leader && lane == 0) {
      const uint32_t idesc_head = umma_idesc_pair(HEAD_N);
      const uint32_t hw_base = smem_u32(s_hw);
      mbar_wait(wbar, 0);
      tc_fence_after();
      int tl = 0;
      for (int tp = cid; tp < n_pairs; tp += ncl, tl++) {
        const int b = tl & 1;
        mbar_wait(&act_full[b], (uint32_t)((tl >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < COUT / 16; k++)
          umma_bf16_ts_pair(tmem_base + COL_HEAD + (b * HEAD_CHAINS + (k % HEAD_CHAINS)) * HEAD_N,
                            tmem_base + COL_ACT + b * (COUT / 2) + k * 8,
                            umma_desc(hw_base + (k / 4) * (HEAD_N / 2) * 128 + (k % 4) * 32, 128), idesc_head,
                            k >= HEAD_CHAINS ? 1u : 0u);
        umma_commit_pair(&head_full[b]);
        if (stampc) AZP_CLKPTR(args.clk)[tl * 16 + 6] = clock64();
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ===== epilogue (both CTAs, own tile): warp (q, h) owns TMEM lanes 32q..32q+31 and channel half h =====
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    auto write_heads = [&](int tj, int j) {
      mbar_wait(&head_full[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t r[8];
      tmem_ld8(tmem_base + lane_addr + COL_HEAD + (j & 1) * HEAD_CHAINS * HEAD_N, r);
#pragma unroll
      for (int c = 1; c < HEAD_CHAINS; c++) {
        uint32_t r2[8];
        tmem_ld8(tmem_base + lane_addr + COL_HEAD + ((j & 1) * HEAD_CHAINS + c) * HEAD_N, r2);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
      }
      if (tj >= n_tiles) return;                            // phantom tile of an odd tile count
      const unsigned R = (unsigned)row_begin + (unsigned)tj * TILE_M + q * 32 + lane;
      const int slot = (int)(R / (unsigned)g.stride), pos = (int)(R % (unsigned)g.stride);
      const int y = pos / g.P, x = pos % g.P;
      if (y >= 1 && x >= 1 && x <= g.B && slot < g.G) {
        __nv_bfloat16* f = args.feat + (size_t)slot * args.FK + (y - 1) * g.B + (x - 1);
#pragma unroll
        for (int j2 = 0; j2 < 8; j2++)
          if (j2 < args.NJ)
            f[j2 < args.PC ? j2 * g.A : args.PKp + (j2 - args.PC) * g.A] =
                __float2bfloat16(fmaxf(__uint_as_float(r[j2]) + s_bias[COUT + j2], 0.f));
      }
    };
    const uint32_t l_tempty0 = mapa_u32(&tempty[0], 0), l_tempty1 = mapa_u32(&tempty[1], 0);
    const uint32_t l_act0 = mapa_u32(&act_full[0], 0), l_act1 = mapa_u32(&act_full[1], 0);
    int tl = 0, t_prev = 0;
    for (int tp = cid; tp < n_pairs; tp += ncl, tl++) {
      const int t = 2 * tp + (int)rank;
      const int acc = tl & 1;
      const bool stamp = stampc && warp == 2 && lane == 0;
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 8] = clock64();
      mbar_wait(&tfull[acc], (tl >> 1) & 1);
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 9] = clock64();
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(acc * COUT + h * HALF);
      if (tl >= 2) {             // head GEMM of tile tl-2 (same TMEM buffers) has retired
        mbar_wait(&head_full[tl & 1], (uint32_t)(((tl - 2) >> 1) & 1));
        tc_fence_after();
      }
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          float v0 = fmaxf(__uint_as_float(r[2 * j]) + s_bias[h * HALF + c0 + 2 * j], 0.f);
          float v1 = fmaxf(__uint_as_float(r[2 * j + 1]) + s_bias[h * HALF + c0 + 2 * j + 1], 0.f);
          __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
          packed[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        tmem_st16(tmem_base + lane_addr + COL_ACT + (uint32_t)((tl & 1) * (COUT / 2) + (h * HALF + c0) / 2), packed);
      }
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 10] = clock64();
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(acc ? l_tempty1 : l_tempty0);
        mbar_arrive_cluster((tl & 1) ? l_act1 : l_act0);
      }
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 11] = clock64();
      if (tl > 0 && h == 0) write_heads(t_prev, tl - 1);
      if (stamp) AZP_CLKPTR(args.clk)[tl * 16 + 12] = clock64();
      t_prev = t;
    }
    if (tl > 0 && h == 0) write_heads(t_prev, tl - 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the leader's UMMAs read the peer's shared and tensor memory: leave together
  if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

// ---- conv1 (4 -> 32, 3x3) on the tensor core ------------------------------------------------------------------
// The first conv sees 0/1 planes (games.py:86-129), so its implicit-GEMM A operand is just the 27 neighbourhood BITS
// of a position (3 kernel rows x 3 stone/last-move planes x 3 columns; plane 3 is all zero) expanded to bf16, plus a
// constant-one column that carries the bias:  act1[row][co] = relu( sum_k A[row][k] * W[co][k] ),
//     k = ky*9 + plane*3 + kx  (k < 27),  k = 27: bias,  k = 28..31: zero.
// Halo rows get an all-zero A row, so their output is exactly 0 and the epilogue is a bare ReLU + pack.
// W is split into bf16 hi + lo parts (two UMMA passes over the same A) so the products with the exact 0/1 inputs
// reproduce the fp32 weights to ~2^-17 relative: conv1 stays fp32-accurate like the SIMT kernel it replaces.
// PLANES = true reads arbitrary float planes [n][4][A] instead (azp_net_forward): same K layout for planes 0..2 and
// a second 16-wide K chunk for plane 3, issued AFTER the shared chunks -- on 0/1 inputs with an empty plane 3 it adds
// exact zeros, so both variants agree bit for bit and the search sees the numbers net_forward reports.
// Persistent, warp-specialised: warp 0 = TMEM owner + MMA issuer, warps 1..8 = two producer groups (even / odd
// tiles) that build the swizzled A tile in shared memory, warps 9..16 = two epilogue groups (TMEM -> ReLU -> bf16 ->
// swizzled staging tile -> one TMA store).
constexpr int C1_NPG = 2;                    // producer groups per CTA (tile tl is built by group tl % C1_NPG); TWO CTAs per SM
constexpr int C1_THREADS = 32 * (1 + 4 * C1_NPG + 8);
constexpr int C1_STAGES = 4;                 // A stages == TMEM accumulators (tile tl uses stage/accumulator tl % 4)
constexpr int C1_A_BYTES = TILE_M * 64;      // one A tile: 128 rows x 32 bf16
constexpr int C1_W_BYTES = 32 * 64;          // one weight matrix: 32 cout x 32 k bf16
constexpr int C1_TAB = 4 * 3 * 24;            // per tile: <= 4 boards x 3 planes x (B+2 <= 24) row-bit words (uint32)

struct Conv1Args {
  ConvGeom g;
  const __nv_bfloat16* w;          // [4][32][32] bf16: hi, lo (planes 0..2 + bias), plane-3 hi, plane-3 lo
  const int32_t* count_dev; int count_host; int back;
  const uint64_t* lx; const uint64_t* lo; const uint8_t* ltm; const int16_t* llast; int NW;   // leaf batch (bitboards)
  const float* planes;             // PLANES: [count][4][A] float32
  __nv_bfloat16* out;              // act1 [rows][32] (direct-store epilogue)
  int staged;                      // 1: swizzled smem staging + TMA store (AZP_CONV1_STORE=tma), 0: 256-bit stores from registers
  long long* clk;                  // bring-up (AZP_CLK=3): per-tile clock64 stamps of CTA 0 ([tile][16]) or null
};
#define C1_STAMP(cond, k) do { if (AZP_CLKPTR(args.clk) && blockIdx.x == 0 && tl < 64 && (cond)) AZP_CLKPTR(args.clk)[tl * 16 + (k)] = clock64(); } while (0)

constexpr int C1_BB_ITEMS = 32;              // staged leaf inputs of the <= 2 boards a tile touches: 2 x (2*NW words + tm + last)
constexpr int conv1_smem_bytes(bool planes) {
  return 1024 + 4 * C1_W_BYTES + (planes ? 2 : 1) * C1_STAGES * C1_A_BYTES + 4 * C1_A_BYTES + 256 * 16 + C1_NPG * 2 * C1_TAB * 4 +
         C1_NPG * 2 * C1_BB_ITEMS * 8 + 4 * C1_STAGES * 8 + 64;
}

template <bool PLANES>
__global__ void __launch_bounds__(C1_THREADS, 2) conv1_umma_kernel(const __grid_constant__ CUtensorMap map_out,
                                                                   const Conv1Args args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                                         // 4 x [32][32] bf16, 64B-swizzled rows
  uint8_t* s_a = s_w + 4 * C1_W_BYTES;                         // C1_STAGES x [128][32] (planes 0..2 + bias)
  uint8_t* s_a3 = s_a + C1_STAGES * C1_A_BYTES;                // C1_STAGES x [128][32] (plane 3; PLANES only)
  uint8_t* s_out = s_a3 + (PLANES ? C1_STAGES * C1_A_BYTES : 0);   // 2 epilogue groups x 2 x [128][32] bf16 staging
  uint4* s_lut = reinterpret_cast<uint4*>(s_out + 4 * C1_A_BYTES);            // [256]: 8 input bits -> 8 bf16 (0.0 / 1.0)
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_lut + 256);                 // [C1_NPG groups][2][C1_TAB] row-bit tables
  uint64_t* s_bb = reinterpret_cast<uint64_t*>(s_tab + C1_NPG * 2 * C1_TAB);  // [C1_NPG groups][2][C1_BB_ITEMS]
  uint64_t* full = s_bb + C1_NPG * 2 * C1_BB_ITEMS;
  uint64_t* done = full + C1_STAGES;                           // MMAs of the stage retired: A stage free AND accumulator ready
  uint64_t* tempty = done + C1_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C1_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom g = args.g;

  // ---- prologue (independent of the previous kernel: overlaps its tail under programmatic dependent launch) ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < C1_STAGES; i++) {
      mbar_init(&full[i], 4); mbar_init(&done[i], 1); mbar_init(&tempty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 512) {                                     // weights -> shared memory in the UMMA 64B-swizzle layout
    const int mat = threadIdx.x >> 7, co = (threadIdx.x >> 2) & 31, c = threadIdx.x & 3;
    const uint4 v = reinterpret_cast<const uint4*>(args.w)[threadIdx.x];
    *reinterpret_cast<uint4*>(s_w + mat * C1_W_BYTES + co * 64 + ((c ^ ((co >> 1) & 3)) << 4)) = v;
  }
  if (threadIdx.x < 256) {
    const uint32_t e8 = threadIdx.x;
    uint32_t wd[4];
#pragma unroll
    for (int j = 0; j < 4; j++) wd[j] = (((e8 >> (2 * j)) & 1u) ? 0x3F80u : 0u) | (((e8 >> (2 * j + 1)) & 1u) ? 0x3F800000u : 0u);
    s_lut[e8] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(tmem_slot, 64 * C1_STAGES);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();                      // leaf batch / leaf count are written by the descent kernel before us
  pdl_launch_dependents();
  const int count = args.count_dev ? *args.count_dev : args.count_host;
  const int slot0 = args.back ? g.G - count : 0;
  const long long row_begin = (long long)slot0 * g.stride;
  const long long rows_total = (long long)count * g.stride;
  const int n_tiles = (int)((rows_total + TILE_M - 1) / TILE_M);

  if (warp == 0) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // hi and lo weight parts are the two halves of ONE N = 64 B operand (rows 0..31 / 32..63 of the same swizzled
      // array): two UMMAs per tile instead of four; the epilogue adds the two 32-column halves
      const uint32_t idesc = umma_idesc(64);
      const uint32_t w_base = smem_u32(s_w), a_base = smem_u32(s_a), a3_base = smem_u32(s_a3);
      int tl = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
        const int st = tl % C1_STAGES;
        const uint32_t ph = (tl / C1_STAGES) & 1;
        mbar_wait_sleep(&tempty[st], ph ^ 1);
        C1_STAMP(true, 5);
        mbar_wait_sleep(&full[st], ph);
        C1_STAMP(true, 6);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(st * 64);
        const uint32_t a = a_base + (uint32_t)(st * C1_A_BYTES);
        umma_bf16(d, umma_desc(a, 64), umma_desc(w_base, 64), idesc, 0u);
        umma_bf16(d, umma_desc(a + 32, 64), umma_desc(w_base + 32, 64), idesc, 1u);
        if (PLANES) {
          const uint32_t a3 = a3_base + (uint32_t)(st * C1_A_BYTES);
          umma_bf16(d, umma_desc(a3, 64), umma_desc(w_base + 2 * C1_W_BYTES, 64), idesc, 1u);
        }
        umma_commit(&done[st]);
        C1_STAMP(true, 7);
      }
    }
  } else if (warp <= 4 * C1_NPG) {
    // ===== A-operand producers: group pg builds the tiles with tl % C1_NPG == pg, one row per thread =====
    const int pg = (warp - 1) >> 2;
    const int pt = threadIdx.x - 32 - pg * 128;              // 0..127
    // the 128 rows of a tile touch at most `nb` boards: 2 for B >= 9, up to 4 for the smallest boards
    const int nb = min((TILE_M + g.stride - 2) / g.stride + 1, 4);
    const int per_board = 2 * args.NW + 2;
    const int n_items = min(nb * per_board, C1_BB_ITEMS);
    // Leaf inputs of tile t, item = board * per_board + j (j < NW: x word, j < 2NW: o word, then the aligned 32-bit
    // words holding the board's to-move byte / last-move int16).  They are fetched with cp.async one tile ahead:
    // plain loads would be drained by the group's next bar.sync (a barrier waits for the warp's outstanding loads),
    // exposing the full global latency on every tile.
    const unsigned stride_rcp = (unsigned)(0x100000000ull / (unsigned)g.stride);
    const int ib = (pt >= per_board) + (pt >= 2 * per_board) + (pt >= 3 * per_board), ij = pt - ib * per_board;
    auto request = [&](int t, int buf) {
      if (pt < n_items && t < n_tiles) {
        const unsigned R0t = (unsigned)(row_begin + (long long)t * TILE_M);
        unsigned q = __umulhi(R0t, stride_rcp);               // floor(2^32 / stride) underestimates by < 1
        if (R0t - q * (unsigned)g.stride >= (unsigned)g.stride) q++;
        const int slot = (int)q + ib;
        if (slot < g.G) {
          const uint32_t dst = smem_u32(s_bb + (pg * 2 + buf) * C1_BB_ITEMS + pt);
          if (ij < 2 * args.NW) {
            const uint64_t* src = (ij < args.NW ? args.lx : args.lo) + (size_t)slot * args.NW + (ij < args.NW ? ij : ij - args.NW);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
          } else {
            const char* src = ij == 2 * args.NW ? reinterpret_cast<const char*>(args.ltm) + (slot & ~3)
                                                : reinterpret_cast<const char*>(args.llast) + ((2 * slot) & ~3);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (!PLANES) {
      request((int)blockIdx.x + pg * (int)gridDim.x, 0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const int TR = g.B + 2;                                  // table rows per (board, plane): rr = -1 .. B
    const unsigned tr_rcp = (65536u + (unsigned)TR - 1u) / (unsigned)TR;      // e / TR == (e * tr_rcp) >> 16 for e < 128
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
      if ((tl % C1_NPG) != pg) continue;
      const int st = tl % C1_STAGES;
      const uint32_t ph = (tl / C1_STAGES) & 1;
      const long long R0 = row_begin + (long long)t * TILE_M;
      uint32_t words[16], words3[16];
#pragma unroll
      for (int j = 0; j < 16; j++) { words[j] = 0u; words3[j] = 0u; }
      if (!PLANES) {
        const int buf = (tl / C1_NPG) & 1;
        const uint64_t* bb = s_bb + (pg * 2 + buf) * C1_BB_ITEMS;
        C1_STAMP(pt == 0, 0);
        request(t + C1_NPG * (int)gridDim.x, buf ^ 1);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + pg) : "memory");        // bb[buf] (landed last iteration) is visible
        C1_STAMP(pt == 0, 1);
        // Row-bit table: tab[board][plane][rr + 1] = (the B cells of board row rr in that plane) << 1, zero for
        // rr = -1 / B.  The 3-bit column pattern (c-1, c, c+1) of cell (rr, c) is then (tab >> c) & 7 with the
        // board edges handled by the zero bits at both ends.  Planes: side to move's stones, opponent's, last move.
        uint32_t* tab = s_tab + (pg * 2 + buf) * C1_TAB;
        unsigned sA = __umulhi((unsigned)R0, stride_rcp);
        if ((unsigned)R0 - sA * (unsigned)g.stride >= (unsigned)g.stride) sA++;
        for (int e = pt; e < nb * 3 * TR; e += 128) {        // one pass for every supported board (<= 126 entries)
          const int combo = (int)(((unsigned)e * tr_rcp) >> 16), ri = e - combo * TR, rr = ri - 1;
          const int b = combo / 3, pl = combo - 3 * b;
          const int slot = (int)sA + b;
          uint32_t v = 0;
          if (rr >= 0 && rr < g.B && slot < g.G) {
            const uint64_t* sx = bb + b * per_board;
            if (pl == 2) {
              const int last = (int)(int16_t)((uint32_t)sx[2 * args.NW + 1] >> (16 * (slot & 1)));     // -1 if none
              const int rel = last - rr * g.B;
              if (last >= 0 && rel >= 0 && rel < g.B) v = 2u << rel;
            } else {
              const bool xm = (((uint32_t)sx[2 * args.NW] >> (8 * (slot & 3))) & 0xFFu) == 1u;
              const uint64_t* wsrc = ((pl == 0) == xm) ? sx : sx + args.NW;
              const int bit = rr * g.B, wi = bit >> 6, sh = bit & 63;
              uint64_t w = wsrc[wi] >> sh;
              if (sh && wi + 1 < args.NW) w |= wsrc[wi + 1] << (64 - sh);
              v = ((uint32_t)w & ((1u << g.B) - 1u)) << 1;
            }
          }
          tab[e] = v;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + pg) : "memory");
        C1_STAMP(pt == 0, 2);
        unsigned pos = (unsigned)R0 + (unsigned)pt - sA * (unsigned)g.stride;
        int b = 0;
        while (pos >= (unsigned)g.stride) { pos -= (unsigned)g.stride; b++; }
        int y, x;
        if (g.P == 16) { y = (int)(pos >> 4); x = (int)(pos & 15u); } else { y = (int)(pos / (unsigned)g.P); x = (int)pos - y * g.P; }
        if (y >= 1 && x >= 1 && x <= g.B && (int)sA + b < g.G) {
          const uint32_t* tb = tab + b * 3 * TR + (y - 1);
          uint32_t bits = 1u << 27;
#pragma unroll
          for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int pl = 0; pl < 3; pl++) bits |= ((tb[pl * TR + ky] >> (x - 1)) & 7u) << (ky * 9 + pl * 3);
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const uint4 v = s_lut[(bits >> (8 * c)) & 0xFFu];
            words[4 * c] = v.x; words[4 * c + 1] = v.y; words[4 * c + 2] = v.z; words[4 * c + 3] = v.w;
          }
        }
      } else {
        const long long R = R0 + pt;
        const int slot = (int)(R / g.stride), pos = (int)(R % g.stride);
        const int y = pos / g.P, x = pos % g.P;
        if (y >= 1 && x >= 1 && x <= g.B && slot - slot0 < count && slot < g.G) {
          const float* pb = args.planes + (size_t)(slot - slot0) * 4 * g.A;
          __nv_bfloat16 vals[32], vals3[16];
#pragma unroll
          for (int k = 0; k < 32; k++) vals[k] = __float2bfloat16(k == 27 ? 1.f : 0.f);
#pragma unroll
          for (int k = 0; k < 16; k++) vals3[k] = __float2bfloat16(0.f);
#pragma unroll
          for (int ky = 0; ky < 3; ky++) {
            const int rr = y - 1 + ky - 1;
            if (rr < 0 || rr >= g.B) continue;
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
              const int cc = x - 1 + kx - 1;
              if (cc < 0 || cc >= g.B) continue;
#pragma unroll
              for (int pl = 0; pl < 3; pl++) vals[ky * 9 + pl * 3 + kx] = __float2bfloat16(pb[pl * g.A + rr * g.B + cc]);
              vals3[ky * 3 + kx] = __float2bfloat16(pb[3 * g.A + rr * g.B + cc]);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; j++) {
            __nv_bfloat162 h2; h2.x = vals[2 * j]; h2.y = vals[2 * j + 1];
            words[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            __nv_bfloat162 h2; h2.x = vals3[2 * j]; h2.y = vals3[2 * j + 1];
            words3[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
      }
      C1_STAMP(pt == 0, 3);
      mbar_wait(&done[st], ph ^ 1);
      C1_STAMP(pt == 0, 4);
      const int sw = (pt >> 1) & 3;                          // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,8]
      uint8_t* rowp = s_a + (size_t)st * C1_A_BYTES + pt * 64;
#pragma unroll
      for (int c = 0; c < 4; c++)
        *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = make_uint4(words[4 * c], words[4 * c + 1], words[4 * c + 2], words[4 * c + 3]);
      if (PLANES) {
        uint8_t* rowp3 = s_a3 + (size_t)st * C1_A_BYTES + pt * 64;
#pragma unroll
        for (int c = 0; c < 4; c++)
          *reinterpret_cast<uint4*>(rowp3 + ((c ^ sw) << 4)) = make_uint4(words3[4 * c], words3[4 * c + 1], words3[4 * c + 2], words3[4 * c + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
      if (!PLANES) asm volatile("cp.async.wait_group 0;" ::: "memory");   // next tile's leaf inputs have landed
    }
  } else {
    // ===== epilogue: group eg drains the tiles with (tl & 1) == eg; warp q owns TMEM lanes 32q..32q+31 =====
    const int eg = (warp - 1 - 4 * C1_NPG) >> 2;
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int m = q * 32 + lane;
    const bool issuer = ((warp - 1 - 4 * C1_NPG) & 3) == 0 && lane == 0;
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tl++) {
      if ((tl & 1) != eg) continue;
      const int st = tl % C1_STAGES;
      const uint32_t ph = (tl / C1_STAGES) & 1;
      C1_STAMP(issuer, 8);
      mbar_wait_sleep(&done[st], ph);
      C1_STAMP(issuer, 9);
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
          __nv_bfloat162 hh = __floats2bfloat162_rn(fmaxf(__uint_as_float(r[2 * j]) + __uint_as_float(r2[2 * j]), 0.f),
                                                    fmaxf(__uint_as_float(r[2 * j + 1]) + __uint_as_float(r2[2 * j + 1]), 0.f));
          packed[hf * 8 + j] = *reinterpret_cast<uint32_t*>(&hh);
        }
      }
      C1_STAMP(issuer, 10);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[st]);
      if (!args.staged) {
        // each lane owns the 64 contiguous bytes of its output row: two 256-bit stores (full 32-byte sectors), no
        // shared-memory round trip and no group barriers
        uint8_t* gp = reinterpret_cast<uint8_t*>(args.out) + ((size_t)row_begin + (size_t)t * TILE_M + m) * 64;
#pragma unroll
        for (int v = 0; v < 2; v++)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"l"(gp + v * 32), "r"(packed[8 * v]), "r"(packed[8 * v + 1]), "r"(packed[8 * v + 2]),
                         "r"(packed[8 * v + 3]), "r"(packed[8 * v + 4]), "r"(packed[8 * v + 5]), "r"(packed[8 * v + 6]),
                         "r"(packed[8 * v + 7]) : "memory");
        C1_STAMP(issuer, 13);
        continue;
      }
      // two staging buffers per group: the TMA store issued two tiles ago (same buffer) has finished reading it --
      // its issuer waited for that (wait_group.read 1) before arriving at this barrier
      uint8_t* obuf = s_out + (eg * 2 + ((tl >> 1) & 1)) * C1_A_BYTES;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + C1_NPG + eg) : "memory");
      C1_STAMP(issuer, 11);
#pragma unroll
      for (int c = 0; c < 4; c++)
        *reinterpret_cast<uint4*>(obuf + m * 64 + ((c ^ ((m >> 1) & 3)) << 4)) =
            make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + C1_NPG + eg) : "memory");
      C1_STAMP(issuer, 12);
      if (issuer) {
        const unsigned R0 = (unsigned)row_begin + (unsigned)t * TILE_M;
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"((uint64_t)&map_out), "r"(smem_u32(obuf)), "r"(0), "r"((int)R0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      C1_STAMP(issuer, 13);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64 * C1_STAGES);
}

// ---- FC heads as tensor-core GEMMs -------------------------------------------------------------------------
// blockIdx.y <  n_slices: policy_fc, one 64-column slice of the outputs:  logits[b][a] = feat[b][0:PK] . Wp[a][:] + bp[a]
// blockIdx.y == n_slices: value_fc1 (+ReLU), value_fc2, tanh                                   (net.py:64-70)
// One CTA = 128 boards (UMMA M) x 64 outputs (UMMA N).  A = bf16 feature rows, B = bf16 weight rows (both K-major,
// TMA, 128B swizzle, K chunks of 64), fp32 accumulators in TMEM.  The kernel is bound by its operand loads
// (latency, few CTAs), so the outputs are split over 4-6 CTAs per board tile and the policy job stops at the RAW
// LOGITS: the softmax over all cells (controller.py:49) is done by the consumer -- one warp per board in the
// expansion kernels / softmax_rows_kernel -- where it costs two warp reductions; no cross-CTA reduction here.
struct FcArgs {
  int A, G, stages;
  int n_slices;                   // policy output slices of 64 columns
  int pk_chunks, vk_chunks;       // ceil(K / 64) of the policy / value job
  int v_off;                      // first feature column of the value job
  const float* pbias;             // [A]
  const float* vbias;             // [64]
  const float* vfc2_w; const float* vfc2_b;
  float* logits; float* value;
  const int32_t* count_dev; int count_host; int back; int out_by_index;
  const void* dead; int dead_row_bytes, dead_stride;      // act2 (dead once conv3 has completed), see discard_l2; or null
  long long* clk;                 // bring-up (AZP_CLK=4): wall-clock ns stamps of CTA (0,0) or null
};
#define FC_STAMP(cond, k) do { if (AZP_CLKPTR(args.clk) && blockIdx.x == 0 && blockIdx.y == 0 && (cond)) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); AZP_CLKPTR(args.clk)[k] = (long long)gt_; } } while (0)

constexpr int FC_N = 64;
constexpr int FC_EPI_WARPS = 8;                        // warp (q, h): TMEM lanes 32q..32q+31 (boards), columns 32h..32h+31
constexpr int FC_THREADS = 64 + 32 * FC_EPI_WARPS;
constexpr int FC_STAGE_BYTES = TILE_M * 128 + FC_N * 128;
constexpr int FC_PITCH = FC_N + 1;                     // odd pitch: conflict-free row-per-thread staging writes

__global__ void __launch_bounds__(FC_THREADS, 1) fc_umma_kernel(const __grid_constant__ CUtensorMap map_feat,
                                                                const __grid_constant__ CUtensorMap map_wp,
                                                                const __grid_constant__ CUtensorMap map_wv,
                                                                const FcArgs args) {
  const int tile = blockIdx.x;
  const bool is_value = (int)blockIdx.y == args.n_slices;
  const int slice = is_value ? 0 : (int)blockIdx.y;
  const int k_chunks = is_value ? args.vk_chunks : args.pk_chunks;
  const int k_off = is_value ? args.v_off : 0;
  const CUtensorMap* map_w = is_value ? &map_wv : &map_wp;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_tail = smem + (size_t)args.stages * FC_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_tail);
  uint64_t* empty = full + 8;
  uint64_t* tfull = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);          // [64]
  float* s_w2 = s_bias + 64;                                        // [64] + [1] (+pad)
  float* s_part = s_w2 + 72;                                        // [2][128] value partial sums
  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  FC_STAMP(threadIdx.x == 0, 0);
  const unsigned tl_idx = threadIdx.x == 0 ? AZP_TL_ENTRY(AZP_TL_FC, gridDim.x * gridDim.y) : 0u;

  // ---- prologue: constants only (overlaps the previous kernel's tail under programmatic dependent launch) ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < args.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64) {
    const int c = slice * FC_N + threadIdx.x;
    s_bias[threadIdx.x] = is_value ? args.vbias[threadIdx.x] : (c < args.A ? args.pbias[c] : 0.f);
  }
  if (threadIdx.x < 65) s_w2[threadIdx.x] = threadIdx.x < 64 ? args.vfc2_w[threadIdx.x] : args.vfc2_b[0];
  if (warp == 1) tmem_alloc(tmem_slot, FC_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  FC_STAMP(threadIdx.x == 0, 1);
  pdl_wait();                      // leaf count + head features come from the previous kernels of the stream
  pdl_launch_dependents();
  FC_STAMP(threadIdx.x == 0, 2);
  if (threadIdx.x == 0) AZP_TL_START(AZP_TL_FC, tl_idx);
  const int count = args.count_dev ? *args.count_dev : args.count_host;
  const int row0 = (args.back ? args.G - count : 0) + tile * TILE_M;
  if (args.dead) {
    const long long rb = (long long)(args.back ? args.G - count : 0) * args.dead_stride;
    const long long nt = ((long long)count * args.dead_stride + TILE_M - 1) / TILE_M;
    discard_l2(static_cast<const char*>(args.dead) + rb * args.dead_row_bytes, nt * TILE_M * args.dead_row_bytes);
  }
  if (tile * TILE_M >= count) {    // nothing to do for this tile: release TMEM and leave (uniform per CTA)
    if (threadIdx.x == 0) AZP_TL_END(AZP_TL_FC, tl_idx, gridDim.x * gridDim.y);
    if (warp == 1) tmem_dealloc(tmem_base, FC_N);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < k_chunks; kc++) {
        const int st = kc % args.stages;
        const uint32_t ph = (kc / args.stages) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)FC_STAGE_BYTES);
        uint8_t* sa = smem + (size_t)st * FC_STAGE_BYTES;
        tma_load_2d(sa, &map_feat, &full[st], k_off + kc * 64, row0);
        tma_load_2d(sa + TILE_M * 128, map_w, &full[st], kc * 64, slice * FC_N);      // rows past the matrix: zero fill
      }
    }
  } else if (warp == 1) {
    {                                  // whole warp runs the loop, an elected lane issues (see elect_one)
      const uint32_t idesc = umma_idesc(FC_N);
      uint32_t accumulate = 0;
      for (int kc = 0; kc < k_chunks; kc++) {
        const int st = kc % args.stages;
        const uint32_t ph = (kc / args.stages) & 1;
        mbar_wait(&full[st], ph);
        if (kc == 0) FC_STAMP(lane == 0, 3);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)st * FC_STAGE_BYTES);
        const uint32_t b_addr = a_addr + TILE_M * 128;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          umma_bf16_e(tmem_base, umma_desc(a_addr + k * 32, 128), umma_desc(b_addr + k * 32, 128), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit_e(&empty[st]);
      }
      umma_commit_e(tfull);
      FC_STAMP(lane == 0, 4);
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int ew = warp - 2;                                   // 0..7
    mbar_wait_sleep(tfull, 0);
    FC_STAMP(warp == 2 && lane == 0, 5);
    tc_fence_after();
    const int trow = q * 32 + lane;
    const int bi = tile * TILE_M + trow;                       // index within the evaluated batch
    const bool valid = bi < count;
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), r);
    if (!is_value) {
      // logits + bias staged in the (now idle) pipeline buffers, then written out coalesced: 256 contiguous bytes
      // per board row and slice
      float* stage = reinterpret_cast<float*>(smem);
#pragma unroll
      for (int i = 0; i < 32; i++) stage[trow * FC_PITCH + h * 32 + i] = __uint_as_float(r[i]) + s_bias[h * 32 + i];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (args.logits) {
        const int A = args.A;
        const int nrows = min(TILE_M, count - tile * TILE_M);
        const int orow0 = args.out_by_index ? tile * TILE_M : (args.back ? args.G - count + tile * TILE_M : tile * TILE_M);
        float* out = args.logits + (size_t)orow0 * A + slice * FC_N;
        const int ncols = min(FC_N, A - slice * FC_N);
        for (int rr = ew; rr < nrows; rr += FC_EPI_WARPS) {
          const float* src = stage + rr * FC_PITCH;
          for (int c = lane; c < ncols; c += 32) out[(size_t)rr * A + c] = src[c];
        }
      }
    } else {
      float sp = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i++) sp = fmaf(fmaxf(__uint_as_float(r[i]) + s_bias[h * 32 + i], 0.f), s_w2[h * 32 + i], sp);
      s_part[h * 128 + trow] = sp;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (h == 0 && valid && args.value) {
        const int row = args.back ? args.G - count + bi : bi;
        args.value[args.out_by_index ? bi : row] = tanhf((s_part[trow] + s_part[128 + trow]) + s_w2[64]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  FC_STAMP(threadIdx.x == 0, 6);
  if (threadIdx.x == 0) AZP_TL_END(AZP_TL_FC, tl_idx, gridDim.x * gridDim.y);
  if (warp == 1) tmem_dealloc(tmem_base, FC_N);
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor [rows][cols] row-major; box = [box_rows][cols]; swizzle span == one row
int make_map(azp_engine* e, CUtensorMap* map, const void* base, uint64_t rows, int cols, int box_rows,
             int box_cols = 0) {
  if (box_cols == 0) box_cols = cols;
  EncodeTiledFn enc = get_encode();
  if (!enc) { e->set_error("cuTensorMapEncodeTiled unavailable"); return AZP_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  static int promo = -1;
  if (promo < 0) { const char* v = getenv("AZP_TMA_PROMO"); promo = v ? atoi(v) : 2; }
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                  : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { e->set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return AZP_ERR_CUDA; }
  return AZP_OK;
}

#include "azp_net_tower.cuh"
#include "azp_net_conv1.cuh"

struct UmmaWorkspace {
  ConvGeom g;
  bool tower = false;                 // residual tower (config 4) instead of the conv stem
  __nv_bfloat16* tact[3] = {nullptr, nullptr, nullptr};   // tower: rotating activation buffers [rows][128]
  CUtensorMap map_tact[3];
  int t_a_rows = 0, t_a_bytes = 0, smem_tower = 0;
  __nv_bfloat16* act1 = nullptr;      // [rows][32]
  __nv_bfloat16* act2 = nullptr;      // [rows][64]
  __nv_bfloat16* feat = nullptr;      // [G + 128][FK] bf16 head features (rows past G stay zero)
  int PKp = 0, VKp = 0, FK = 0;       // padded feature widths (multiples of 8 elements = 16 bytes)
  int fc_stages = 0, fc_b_rows = 0, fc_smem = 0;
  CUtensorMap map_act1, map_act2, map_feat;
  CUtensorMap map_act1_s, map_act2_s;  // single-load boxes (128+2P+2 rows)
  CUtensorMap map_act2_out;            // conv2 epilogue: TMA store of one [128 rows][64 ch] tile
  CUtensorMap map_act1_out;            // conv1 (tensor core) epilogue: TMA store of one [128 rows][32 ch] tile
  int conv1_mode = 3;                  // 3: conv1_taps_kernel, 1: conv1_umma_kernel, 2: conv1_table_kernel, 0: conv1_bits_kernel (see azp_umma_create)
  int conv1_pairs = 1;                 // conv1_taps_kernel: two taps per UMMA (AZP_CONV1_PAIRS=0: one)
  int single = 0;
  long long* clk = nullptr;            // AZP_CLK=1: per-tile stamps of conv3's CTA 0
  size_t rows = 0;
  int sms = 148;
  int smem2 = 0, smem3 = 0;
  int pair = 0, smem3p = 0, pair_clusters = 0;   // conv3 on CTA pairs (cta_group::2): AZP_CONV3_PAIR=1 (default 0)
  int fuse_fc = 0;                     // AZP_FUSE_FC=1: FC layers inside the conv3 kernel instead of the fc_umma_kernel launch
  unsigned int* feat_done = nullptr;   // [FC_MAX_BOARD_TILES] see ConvArgs
  unsigned int* fc_done = nullptr;     // [FC_MAX_BOARD_TILES]
  int fuse_min_games = -1;             // AZP_FUSE_DESCENT_MIN_GAMES; -1: 8 x SMs
  int fuse_descent = 1;                // AZP_FUSE_DESCENT=0 (read when the engine is created): descent kernels without the conv1 tail
  int discard = 0;                     // AZP_DISCARD=1: dead activations are invalidated in L2 (measured: no gain, see discard_l2)
  // AZP_L2_PERSIST=1: act1 | act2 (ONE allocation) get an L2 access-policy window (persisting hits, streaming misses).
  // Measured on B200 (79 MB carve-out of the 126 MB L2, hit ratio 0.82 over the 96 MB window): conv1 / conv2 unchanged --
  // their stores are bound by the row-per-lane write pattern at 3.6 TB/s into L2 (tools/bench_cu/store_rate.cu), not by
  // HBM write-backs -- and the tree kernel loses 3.5 us to the smaller unreserved L2: off by default
  int l2_persist = 0;
  void* window_base = nullptr; size_t window_bytes = 0; float window_hit = 1.f;
  std::vector<cudaStream_t> policy_streams;
  int* sched2 = nullptr;               // [SCHED_GROUPS + 1] conv2's dynamic tile scheduler (see ConvArgs::sched); AZP_DYN_TILES=0: static lists
};
struct UmmaWeights {
  // tower
  std::vector<__nv_bfloat16*> tw;     // per 128->128 conv: [tap][k-chunk][cout][64]
  std::vector<float*> tb;             // per conv: folded bias [128]
  std::vector<CUtensorMap> map_tw;
  std::vector<int> t_residual;        // per conv: add the block input before the ReLU
  float* tf_bits = nullptr; float* tf_planes = nullptr; float* tf_b = nullptr;   // first conv [27|36][C], [C]
  float* th_w = nullptr; float* th_b = nullptr;                                   // 1x1 heads [NJ][C], [NJ]
  __nv_bfloat16* w2 = nullptr;        // [9][64][32]
  __nv_bfloat16* w3 = nullptr;        // [9][128][64]
  float* b1 = nullptr; float* b2 = nullptr; float* b3 = nullptr;
  float* w1_bits = nullptr;           // [tap][3][32]
  float* w1_tab = nullptr;            // [plane(3)][ky][pattern(8)][32]: sum over set pattern bits kx of w1[co][plane][ky][kx]
  float* w1_planes = nullptr;         // [tap][4][32]
  float* w1_tab2 = nullptr;           // [ky][own3 | opp3 << 3][32]: both stone planes' column sums of one kernel row (conv1_table_kernel)
  float* w1_lm = nullptr;             // [ky][kx][32]: the last-move plane's weights
  __nv_bfloat16* w1_umma = nullptr;   // [4][32][32]: hi, lo (k = ky*9+plane*3+kx, k=27 bias), plane-3 hi, lo (k = ky*3+kx)
  __nv_bfloat16* w1_taps = nullptr;   // [10][64][8]: per tap (the tenth is zero) hi | lo rows x (planes 0..2, bias at the centre tap, plane 3, zeros)
  __nv_bfloat16* head_w16 = nullptr;  // [16][128] 1x1 head conv weights (policy rows first, zero padded)
  float* head_b8 = nullptr;           // [8]
  __nv_bfloat16* wp16 = nullptr;      // [A][PKp] policy_fc weights, K contiguous, zero padded
  __nv_bfloat16* wv16 = nullptr;      // [64][VKp] value_fc1 weights
  CUtensorMap map_w2, map_w3, map_hw, map_wp, map_wv;
  CUtensorMap map_w3h, map_hwh;       // pair kernel: boxes of HALF the output channels / head rows
};

constexpr int STAGES2 = 6, STAGES3 = 3;

// bring-up builds only (make BRINGUP=1): which kernel stamps its tiles (AZP_CLK=1 conv3, 2 conv2, 3 conv1, 4 FC) and the
// AZP_DEBUG_CONV experiment bits; constants in the default build
inline long long* bringup_clk(const UmmaWorkspace* ws, int which) {
#ifdef AZP_BRINGUP
  const char* v = getenv("AZP_CLK");
  return (v && atoi(v) == which) ? ws->clk : nullptr;
#else
  (void)ws; (void)which;
  return nullptr;
#endif
}
inline int bringup_debug_bits() {
#ifdef AZP_BRINGUP
  static int dbg = -1;
  if (dbg < 0) { const char* v = getenv("AZP_DEBUG_CONV"); dbg = v ? atoi(v) : 0; }
  return dbg;
#else
  return 0;
#endif
}

// bytes of one activation stage: the TMA box of a conv tile, or (conv3 with the FC layers fused in) an FC operand pair
inline int conv_stage_bytes(int stage_rows, int rowb, bool fc_fused) {
  const int conv = (stage_rows + 2 + 7) / 8 * 8 * rowb;
  return fc_fused && conv < FC_FUSED_STAGE_BYTES ? FC_FUSED_STAGE_BYTES : conv;
}

template <int CIN, int COUT, int STAGES, bool HEADS, bool BITS>
int conv_smem_bytes(int stage_rows, bool fc_fused = false) {
  using S = ConvSmem<CIN, COUT, STAGES, HEADS>;
  return 1024 + S::W_BYTES + STAGES * conv_stage_bytes(stage_rows, S::ROWB, fc_fused) + (HEADS ? (COUT / 64) * HEAD_N * 128 : 0) +
         (HEADS ? 0 : 2 * TILE_M * 128) + (2 * STAGES + 21) * 8 + 16 + 64 + (COUT + 8) * 4 + (BITS ? (72 * 32 + 32) * 4 + 2 * CODES * 2 : 0) + 64;
}

}  // namespace

int azp_umma_create(azp_engine* e) {
  const bool tower = e->cfg.net_kind == AZP_NET_TOWER;
  if (tower && e->cfg.tower_channels != TW_C) {
    e->set_error("the bf16 tcgen05 tower kernel is specialised for 128 channels (other widths: fp32 path)");
    return AZP_ERR_UNSUPPORTED;
  }
  if (e->cfg.board > 23) { e->set_error("board too large for the padded layout"); return AZP_ERR_UNSUPPORTED; }
  UmmaWorkspace* ws = new UmmaWorkspace();
  e->ws->umma = ws;
  ws->tower = tower;
  ConvGeom& g = ws->g;
  // Row pitch: the halo column x = 0 of a row group is also the right-hand halo of the row group before it, so B+1
  // positions per row group are enough: (B+1)^2 rows per board (100 at 9x9, 256 at 15x15, 400 at 19x19) instead of the
  // 16 / 24-wide groups of the first layout (160 / 256 / 480): the tensor work wasted on halo rows falls from 49 % to
  // 19 % at 9x9 and from 25 % to 10 % at 19x19.  UMMA descriptors and TMA boxes take any row offset, so nothing else
  // depends on the pitch.  Boards below 7x7 keep a pitch of 16 (the conv1 producers stage at most 4 boards per tile).
  g.B = e->cfg.board; g.A = e->A; g.P = g.B >= 7 ? g.B + 1 : 16; g.stride = (g.B + 1) * g.P; g.G = e->slots;
  ws->rows = (size_t)g.G * g.stride + 2 * g.P + TILE_M;        // tail halo + a partial last tile
  int st;
  const int PCh = tower ? 2 : 4, VCh = tower ? 1 : 2;          // head channels (policy, value)
  if (!tower) {
    if ((st = e->alloc(&ws->act1, ws->rows * (32 + 64)))) return st;         // act1 | act2 (one L2 access-policy window)
    ws->act2 = ws->act1 + ws->rows * 32;
    { const char* v = getenv("AZP_L2_PERSIST"); ws->l2_persist = v ? atoi(v) : 0; }
    if (ws->l2_persist) {
      cudaDeviceProp pp;
      AZP_CUDA(e, cudaGetDeviceProperties(&pp, e->cfg.device));
      const size_t want = ws->rows * (32 + 64) * sizeof(__nv_bfloat16);
      size_t carve = (size_t)pp.persistingL2CacheMaxSize < want ? (size_t)pp.persistingL2CacheMaxSize : want;
      if (carve == 0 || pp.accessPolicyMaxWindowSize <= 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) {
        cudaGetLastError();
        ws->l2_persist = 0;
      } else {
        ws->window_base = ws->act1;
        ws->window_bytes = want < (size_t)pp.accessPolicyMaxWindowSize ? want : (size_t)pp.accessPolicyMaxWindowSize;
        ws->window_hit = carve >= ws->window_bytes ? 1.f : (float)((double)carve / (double)ws->window_bytes);
        if (getenv("AZP_VERBOSE")) fprintf(stderr, "[azp] L2 window %zu MB, persisting carve-out %zu MB (max %d MB), hit ratio %.2f\n", ws->window_bytes >> 20, carve >> 20, pp.persistingL2CacheMaxSize >> 20, ws->window_hit);
      }
    }
  } else {
    ws->t_a_rows = TILE_M + 2 * g.P + 2;
    ws->t_a_bytes = (ws->t_a_rows + 7) / 8 * 8 * 128;
    for (int i = 0; i < 3; i++) {
      if ((st = e->alloc(&ws->tact[i], ws->rows * TW_C))) return st;
      AZP_CUDA(e, cudaMemset(ws->tact[i], 0, ws->rows * TW_C * sizeof(__nv_bfloat16)));
      if ((st = make_map(e, &ws->map_tact[i], ws->tact[i], ws->rows, TW_C, ws->t_a_rows, 64))) return st;
    }
    ws->smem_tower = 1024 + 4 * ws->t_a_bytes + TW_WST * TW_C * 128 + 24 * 8 + 16 + TW_C * 4 + 64;
    AZP_CUDA(e, cudaFuncSetAttribute(conv_tower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->smem_tower));
    AZP_CUDA(e, cudaFuncSetAttribute(tower_first_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((37 * TW_C + 4 * g.A) * sizeof(float))));
  }
  ws->PKp = (PCh * g.A + 7) / 8 * 8; ws->VKp = (VCh * g.A + 7) / 8 * 8; ws->FK = ws->PKp + ws->VKp;
  if ((st = e->alloc(&ws->feat, ((size_t)g.G + TILE_M) * ws->FK))) return st;
  AZP_CUDA(e, cudaMemset(ws->feat, 0, ((size_t)g.G + TILE_M) * ws->FK * sizeof(__nv_bfloat16)));
  if ((st = make_map(e, &ws->map_feat, ws->feat, (uint64_t)g.G + TILE_M, ws->FK, TILE_M, 64))) return st;
  {
    { const char* v = getenv("AZP_FC_STAGES"); ws->fc_stages = v ? atoi(v) : 6; }   // x 24 KB operand stages; the first two double as the logits staging tile
    ws->fc_smem = 1024 + ws->fc_stages * FC_STAGE_BYTES + 17 * 8 + 16 + (64 + 72 + 256) * 4 + 64;
  }
  cudaDeviceProp prop0;
  AZP_CUDA(e, cudaGetDeviceProperties(&prop0, e->cfg.device));
  ws->sms = prop0.multiProcessorCount;
  AZP_CUDA(e, cudaFuncSetAttribute(fc_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->fc_smem));
  if (tower) return AZP_OK;
  AZP_CUDA(e, cudaMemset(ws->act1, 0, ws->rows * (32 + 64) * sizeof(__nv_bfloat16)));
  const int stage_rows = TILE_M + 2 * g.P;
  if ((st = make_map(e, &ws->map_act1, ws->act1, ws->rows, 32, stage_rows))) return st;
  if ((st = make_map(e, &ws->map_act2, ws->act2, ws->rows, 64, stage_rows))) return st;
  if ((st = make_map(e, &ws->map_act1_s, ws->act1, ws->rows, 32, stage_rows + 2))) return st;
  if ((st = make_map(e, &ws->map_act2_s, ws->act2, ws->rows, 64, stage_rows + 2))) return st;
  if ((st = make_map(e, &ws->map_act2_out, ws->act2, ws->rows, 64, TILE_M))) return st;
  if ((st = make_map(e, &ws->map_act1_out, ws->act1, ws->rows, 32, TILE_M))) return st;
  // conv1 of the leaf batch: 1 = tensor-core kernel (default), 2 = merged-table SIMT kernel (AZP_CONV1=table: three table rows
  // per position; parity-green, but 11.3 M warp instructions per launch make it issue-bound at 23.9 us against the tensor
  // kernel's 18-20 us, profiles/r2c_ncu_conv1_table_kernel.txt), 0 = the first SIMT kernel with nine pattern rows per
  // position (AZP_CONV1=simt); each has a float-planes twin that agrees with it bit for bit
  // 3 (default) = nine-tap tensor-core kernel on 16-byte input rows (azp_net_conv1.cuh), AZP_CONV1=umma selects 1
  { const char* v = getenv("AZP_CONV1"); ws->conv1_mode = !v ? 3 : (!strcmp(v, "simt") ? 0 : (!strcmp(v, "table") ? 2 : (!strcmp(v, "umma") ? 1 : 3))); }
  { const char* v = getenv("AZP_CONV1_PAIRS"); ws->conv1_pairs = v ? atoi(v) : 1; }
  AZP_CUDA(e, cudaFuncSetAttribute(conv1_taps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, conv1x_smem_bytes()));
  AZP_CUDA(e, cudaFuncSetAttribute(conv1_taps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, conv1x_smem_bytes()));
  AZP_CUDA(e, cudaFuncSetAttribute(conv1_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, conv1_smem_bytes(false)));
  AZP_CUDA(e, cudaFuncSetAttribute(conv1_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, conv1_smem_bytes(true)));
  { const char* v = getenv("AZP_SINGLE_LOAD"); ws->single = v ? atoi(v) : 1; }
  if (getenv("AZP_VERBOSE")) {
    int o1 = 0, o2 = 0, o3 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, conv1_taps_kernel<false>, C1X_THREADS, conv1x_smem_bytes());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, conv1_umma_kernel<false>, C1_THREADS, conv1_smem_bytes(false));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, fc_umma_kernel, FC_THREADS, 1024);
    fprintf(stderr, "[azp] resident CTAs per SM by the occupancy API: conv1_taps %d, conv1_umma %d, fc (1 KB smem) %d\n", o1, o2, o3);
    cudaGetLastError();
  }
#ifdef AZP_BRINGUP
  if (getenv("AZP_CLK")) { if ((st = e->alloc(&ws->clk, (size_t)64 * 16))) return st; AZP_CUDA(e, cudaMemset(ws->clk, 0, 64 * 16 * 8)); }
#endif
  cudaDeviceProp prop;
  AZP_CUDA(e, cudaGetDeviceProperties(&prop, e->cfg.device));
  ws->sms = prop.multiProcessorCount;
  // measured (profiles/r2c_fc_fused_timeline.txt): with the jobs at the tail of every CTA's work list the fused kernel costs
  // what conv3 + the separately launched FC kernel cost under programmatic dependent launch (-6 us bracketed, +1 % end to
  // end): off by default
  { const char* v = getenv("AZP_FUSE_FC"); ws->fuse_fc = v ? atoi(v) : 0; }
  if (ws->single != 1) ws->fuse_fc = 0;
  if ((st = e->alloc(&ws->feat_done, (size_t)FC_MAX_BOARD_TILES))) return st;
  if ((st = e->alloc(&ws->fc_done, (size_t)FC_MAX_BOARD_TILES))) return st;
  AZP_CUDA(e, cudaMemset(ws->feat_done, 0, FC_MAX_BOARD_TILES * sizeof(unsigned int)));
  AZP_CUDA(e, cudaMemset(ws->fc_done, 0, FC_MAX_BOARD_TILES * sizeof(unsigned int)));
  { const char* v = getenv("AZP_DISCARD"); ws->discard = v ? atoi(v) : 0; }
  { const char* v = getenv("AZP_FUSE_DESCENT"); ws->fuse_descent = v ? atoi(v) : 1; }
  { const char* v = getenv("AZP_FUSE_DESCENT_MIN_GAMES"); ws->fuse_min_games = v ? atoi(v) : -1; }
  // measured (profiles/r3_*): balancing the CTAs does not shorten conv2 -- the launch is bound by the chip-wide write stream, the
  // per-CTA spread is unfair arbitration of it -- and the polling costs a little: opt-in (AZP_DYN_TILES=1)
  { const char* v = getenv("AZP_DYN_TILES"); if (v && atoi(v) != 0) { if ((st = e->alloc(&ws->sched2, (size_t)SCHED_GROUPS + 1))) return st; AZP_CUDA(e, cudaMemset(ws->sched2, 0, (SCHED_GROUPS + 1) * sizeof(int))); } }
  ws->smem2 = conv_smem_bytes<32, 64, STAGES2, false, true>(stage_rows);
  ws->smem3 = conv_smem_bytes<64, 128, STAGES3, true, false>(stage_rows, ws->fuse_fc != 0);
  if (ws->smem3 > (int)prop.sharedMemPerBlockOptin) {       // cannot happen for boards <= 23; never launch an oversized kernel
    ws->fuse_fc = 0;
    ws->smem3 = conv_smem_bytes<64, 128, STAGES3, true, false>(stage_rows, false);
  }
  AZP_CUDA(e, cudaFuncSetAttribute(conv_umma_kernel<32, 64, STAGES2, false, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, ws->smem2));
  AZP_CUDA(e, cudaFuncSetAttribute(conv_umma_kernel<32, 64, STAGES2, false, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, ws->smem2));
  AZP_CUDA(e, cudaFuncSetAttribute(conv_umma_kernel<64, 128, STAGES3, true, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, ws->smem3));
  AZP_CUDA(e, cudaFuncSetAttribute(fc_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->fc_smem));
  {
    // conv3 on CTA pairs: persistent grid = 2 x (clusters that can be co-resident; a GPC with an odd SM count strands one SM)
    const char* v = getenv("AZP_CONV3_PAIR");
    ws->pair = v ? atoi(v) : 0;      // measured 4 % SLOWER than the single-CTA kernel: opt-in
    ws->smem3p = 1024 + C3P_W_BYTES + C3P_STAGES * ((stage_rows + 2 + 7) / 8 * 8) * 128 + C3P_HW_BYTES +
                 (2 * C3P_STAGES + 9) * 8 + 16 + (C3P_COUT + 8) * 4 + 64;
    if (ws->pair && ws->single != 1) ws->pair = 0;
    if (ws->pair) {
      AZP_CUDA(e, cudaFuncSetAttribute(conv3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->smem3p));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(ws->sms / 2 * 2); cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = ws->smem3p;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, conv3_pair_kernel, &cfg) != cudaSuccess || ncl < 1) {
        cudaGetLastError();
        ws->pair = 0;
      } else {
        ws->pair_clusters = ncl < ws->sms / 2 ? ncl : ws->sms / 2;
      }
      if (getenv("AZP_VERBOSE")) fprintf(stderr, "[azp] conv3 pair mode %d, co-resident clusters %d (using %d), smem %d\n", ws->pair, ncl, ws->pair_clusters, ws->smem3p);
    }
  }
  return AZP_OK;
}

void azp_umma_destroy(azp_engine* e) {
  if (e->ws && e->ws->umma) {
    UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
    if (ws->l2_persist) {                       // detach the window from the caller's streams, release the persisting lines
      cudaStreamAttrValue v;
      memset(&v, 0, sizeof(v));
      for (cudaStream_t q : ws->policy_streams)
        if (cudaStreamSetAttribute(q, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
      if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
    }
    delete ws; e->ws->umma = nullptr;
  }

}

template <typename T>
static int up(azp_engine* e, T** dst, const std::vector<T>& src) {
  int st = e->alloc(dst, src.size());
  if (st) return st;
  AZP_CUDA(e, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return AZP_OK;
}

static int fc_weights_build(azp_engine* e, const HostNet& net, NetWeights* W, UmmaWeights* U);

static int tower_build(azp_engine* e, const HostNet& net, NetWeights* W) {
  UmmaWeights* U = new UmmaWeights();
  W->umma = U;
  W->umma_free = [](void* q) { delete (UmmaWeights*)q; };
  const int C = TW_C;
  int st;
  if (net.convs.size() < 3 || net.convs[0].cin != 4 || net.convs[0].cout != C) {
    e->set_error("bf16 tower expects a 4->128 first conv and 128->128 residual convs");
    return AZP_ERR_UNSUPPORTED;
  }
  const HostConv& c0 = net.convs[0];
  std::vector<float> wb(27 * C), wp(36 * C);
  for (int co = 0; co < C; co++)
    for (int ci = 0; ci < 4; ci++)
      for (int t = 0; t < 9; t++) {
        const float v = c0.w[((size_t)co * 4 + ci) * 9 + t];
        wp[(t * 4 + ci) * C + co] = v;
        if (ci < 3) wb[(t * 3 + ci) * C + co] = v;
      }
  if ((st = up(e, &U->tf_bits, wb))) return st;
  if ((st = up(e, &U->tf_planes, wp))) return st;
  if ((st = up(e, &U->tf_b, c0.b))) return st;
  const size_t nconv = net.convs.size() - 1;
  U->tw.resize(nconv); U->tb.resize(nconv); U->map_tw.resize(nconv); U->t_residual.resize(nconv);
  for (size_t i = 0; i < nconv; i++) {
    const HostConv& c = net.convs[i + 1];
    if (c.cin != C || c.cout != C) { e->set_error("tower conv is not 128->128"); return AZP_ERR_UNSUPPORTED; }
    std::vector<__nv_bfloat16> w((size_t)9 * 2 * C * 64);
    for (int t = 0; t < 9; t++)
      for (int kc = 0; kc < 2; kc++)
        for (int co = 0; co < C; co++)
          for (int k = 0; k < 64; k++)
            w[(((size_t)t * 2 + kc) * C + co) * 64 + k] = __float2bfloat16(c.w[((size_t)co * C + kc * 64 + k) * 9 + t]);
    if ((st = up(e, &U->tw[i], w))) return st;
    if ((st = up(e, &U->tb[i], c.b))) return st;
    if ((st = make_map(e, &U->map_tw[i], U->tw[i], (uint64_t)9 * 2 * C, 64, C))) return st;
    U->t_residual[i] = c.residual;
  }
  if ((st = up(e, &U->th_w, net.head_w))) return st;
  if ((st = up(e, &U->th_b, net.head_b))) return st;
  return fc_weights_build(e, net, W, U);
}

int azp_umma_build(azp_engine* e, const HostNet& net, NetWeights* W) {
  if (e->cfg.net_kind == AZP_NET_TOWER) return tower_build(e, net, W);
  if (net.convs.size() != 3 || net.convs[0].cout != 32 || net.convs[1].cout != 64 || net.convs[2].cout != 128) {
    e->set_error("bf16 path expects the 4->32->64->128 stem");
    return AZP_ERR_UNSUPPORTED;
  }
  UmmaWeights* U = new UmmaWeights();
  W->umma = U;
  W->umma_free = [](void* q) { delete (UmmaWeights*)q; };
  int st;
  // conv1: fp32 [tap][plane][32]
  const HostConv& c1 = net.convs[0];
  std::vector<float> w1b(27 * 32), w1p(36 * 32);
  for (int co = 0; co < 32; co++)
    for (int ci = 0; ci < 4; ci++)
      for (int t = 0; t < 9; t++) {
        float v = c1.w[((size_t)co * 4 + ci) * 9 + t];
        w1p[(t * 4 + ci) * 32 + co] = v;
        if (ci < 3) w1b[(t * 3 + ci) * 32 + co] = v;
      }
  if ((st = up(e, &U->w1_bits, w1b))) return st;
  {
    std::vector<float> tab(72 * 32, 0.f);
    for (int pl = 0; pl < 3; pl++)
      for (int ky = 0; ky < 3; ky++)
        for (int pat = 0; pat < 8; pat++)
          for (int co = 0; co < 32; co++) {
            float t = 0.f;                                  // fixed order kx = 0,1,2 (matched by conv1_planes_kernel)
            for (int kx = 0; kx < 3; kx++)
              if (pat & (1 << kx)) t += c1.w[((size_t)co * 4 + pl) * 9 + ky * 3 + kx];
            tab[((pl * 3 + ky) * 8 + pat) * 32 + co] = t;
          }
    if ((st = up(e, &U->w1_tab, tab))) return st;
  }
  if ((st = up(e, &U->w1_planes, w1p))) return st;
  {
    std::vector<float> tab2((size_t)C1T_TAB_FLOATS, 0.f), lmw(9 * 32, 0.f);
    for (int ky = 0; ky < 3; ky++)
      for (int pat = 0; pat < 64; pat++)
        for (int co = 0; co < 32; co++) {
          float p0 = 0.f, p1 = 0.f;                        // fixed order kx = 0,1,2 per plane, then own + opp
          for (int kx = 0; kx < 3; kx++) {
            if (pat & (1 << kx)) p0 += c1.w[((size_t)co * 4 + 0) * 9 + ky * 3 + kx];
            if (pat & (8 << kx)) p1 += c1.w[((size_t)co * 4 + 1) * 9 + ky * 3 + kx];
          }
          tab2[((size_t)ky * 64 + pat) * 32 + co] = p0 + p1;
        }
    for (int t = 0; t < 9; t++)
      for (int co = 0; co < 32; co++) lmw[t * 32 + co] = c1.w[((size_t)co * 4 + 2) * 9 + t];
    if ((st = up(e, &U->w1_tab2, tab2))) return st;
    if ((st = up(e, &U->w1_lm, lmw))) return st;
  }
  if ((st = up(e, &U->b1, c1.b))) return st;
  {
    // tensor-core conv1: B operands [32 cout][32 k] bf16, value split v = hi + lo
    std::vector<__nv_bfloat16> wu((size_t)4 * 32 * 32, __float2bfloat16(0.f));
    auto put = [&](int mat, int co, int k, float v) {
      const __nv_bfloat16 hi = __float2bfloat16(v);
      wu[((size_t)mat * 32 + co) * 32 + k] = hi;
      wu[((size_t)(mat + 1) * 32 + co) * 32 + k] = __float2bfloat16(v - __bfloat162float(hi));
    };
    for (int co = 0; co < 32; co++) {
      for (int ky = 0; ky < 3; ky++)
        for (int kx = 0; kx < 3; kx++) {
          for (int pl = 0; pl < 3; pl++) put(0, co, ky * 9 + pl * 3 + kx, c1.w[((size_t)co * 4 + pl) * 9 + ky * 3 + kx]);
          put(2, co, ky * 3 + kx, c1.w[((size_t)co * 4 + 3) * 9 + ky * 3 + kx]);
        }
      put(0, co, 27, c1.b[co]);
    }
    if ((st = up(e, &U->w1_umma, wu))) return st;
  }
  {
    // nine-tap conv1: per tap one K block [64 rows = hi | lo x 32 cout][8 k] bf16 (k: planes 0..2, bias at the centre tap,
    // plane 3, zeros); block 9 is all zero
    std::vector<__nv_bfloat16> wt((size_t)C1X_W_ELEMS, __float2bfloat16(0.f));
    for (int t = 0; t < 9; t++)
      for (int co = 0; co < 32; co++)
        for (int k = 0; k < 5; k++) {
          const float v = k < 3 ? c1.w[((size_t)co * 4 + k) * 9 + t] : (k == 3 ? (t == 4 ? c1.b[co] : 0.f) : c1.w[((size_t)co * 4 + 3) * 9 + t]);
          const __nv_bfloat16 hi = __float2bfloat16(v);
          wt[((size_t)t * 64 + co) * 8 + k] = hi;
          wt[((size_t)t * 64 + 32 + co) * 8 + k] = __float2bfloat16(v - __bfloat162float(hi));
        }
    if ((st = up(e, &U->w1_taps, wt))) return st;
  }
  // conv2/conv3: bf16 [tap][cout][cin] (K-major rows of one tap = UMMA B operand)
  auto pack = [&](const HostConv& c, __nv_bfloat16** dst) {
    std::vector<__nv_bfloat16> w((size_t)9 * c.cout * c.cin);
    for (int co = 0; co < c.cout; co++)
      for (int ci = 0; ci < c.cin; ci++)
        for (int t = 0; t < 9; t++)
          w[((size_t)t * c.cout + co) * c.cin + ci] = __float2bfloat16(c.w[((size_t)co * c.cin + ci) * 9 + t]);
    return up(e, dst, w);
  };
  if ((st = pack(net.convs[1], &U->w2))) return st;
  if ((st = pack(net.convs[2], &U->w3))) return st;
  if ((st = up(e, &U->b2, net.convs[1].b))) return st;
  if ((st = up(e, &U->b3, net.convs[2].b))) return st;
  if ((st = make_map(e, &U->map_w2, U->w2, 9 * 64, 32, 64))) return st;
  if ((st = make_map(e, &U->map_w3, U->w3, 9 * 128, 64, 128))) return st;
  // heads: [16][C] bf16 1x1 conv weights (policy rows first, rows >= NJ zero) = B operand of the head GEMM
  const int NJ = net.PC + net.VC;
  std::vector<__nv_bfloat16> hw((size_t)HEAD_N * net.C, __float2bfloat16(0.f));
  std::vector<float> hb(8, 0.f);
  for (int j = 0; j < NJ; j++) {
    hb[j] = net.head_b[j];
    for (int c = 0; c < net.C; c++) hw[(size_t)j * net.C + c] = __float2bfloat16(net.head_w[(size_t)j * net.C + c]);
  }
  if ((st = up(e, &U->head_w16, hw))) return st;
  if ((st = up(e, &U->head_b8, hb))) return st;
  if ((st = make_map(e, &U->map_hw, U->head_w16, HEAD_N, net.C, HEAD_N, 64))) return st;
  if ((st = make_map(e, &U->map_w3h, U->w3, 9 * 128, 64, 64))) return st;
  if ((st = make_map(e, &U->map_hwh, U->head_w16, HEAD_N, net.C, HEAD_N / 2, 64))) return st;
  return fc_weights_build(e, net, W, U);
}

// ---- device-side repack (azp_load_weights_device on an already loaded bf16 stem) ---------------------------------
// Rebuilds every derived weight buffer of the stem straight from the flat state_dict blob ON THE DEVICE, with the
// arithmetic of the host builders above (same bf16 roundings, same summation order), into the buffers the kernels,
// tensor maps and the captured step graph already point to.  No host round trip, no synchronisation: a promotion
// broadcast or a per-step weight refresh is just one more kernel on the stream.
struct StemRepack {
  const float* blob; int A, PKp, VKp;
  size_t o_c1w, o_c1b, o_c2w, o_c2b, o_c3w, o_c3b, o_pcw, o_pcb, o_pfw, o_pfb, o_vcw, o_vcb, o_v1w, o_v1b, o_v2w, o_v2b;
  float *w1_bits, *w1_tab, *w1_planes, *w1_tab2, *w1_lm, *b1, *b2, *b3, *head_b8, *pfc_b, *vfc1_b, *vfc2_w, *vfc2_b;
  __nv_bfloat16 *w1_umma, *w2, *w3, *head_w16, *wp16, *wv16, *w1_taps;
};

__global__ void stem_repack_kernel(StemRepack r) {
  const float* c1w = r.blob + r.o_c1w; const float* c1b = r.blob + r.o_c1b;
  const int A = r.A, PK = 4 * A, VK = 2 * A;
  // segment sizes, in the order they are walked
  const long long n[15] = {27 * 32, 36 * 32, 72 * 32, 4 * 32 * 32, 9 * 64 * 32, 9 * 128 * 64, HEAD_N * 128,
                           (long long)A * r.PKp, 64LL * r.VKp, 32 + 64 + 128 + 8, A + 64 + 64 + 1, C1T_TAB_FLOATS, 9 * 32,
                           C1X_W_ELEMS, 0};
  long long total = 0;
  for (int i = 0; i < 14; i++) total += n[i];
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long i = idx;
    int seg = 0;
    while (i >= n[seg]) { i -= n[seg]; seg++; }
    const int k = (int)i;
    if (seg == 0) {                                   // w1_bits [tap][3][32]
      const int co = k % 32, ci = (k / 32) % 3, t = k / 96;
      r.w1_bits[k] = c1w[(co * 4 + ci) * 9 + t];
    } else if (seg == 1) {                            // w1_planes [tap][4][32]
      const int co = k % 32, ci = (k / 32) % 4, t = k / 128;
      r.w1_planes[k] = c1w[(co * 4 + ci) * 9 + t];
    } else if (seg == 2) {                            // w1_tab [plane(3)][ky][pattern(8)][32], kx summed in order 0,1,2
      const int co = k % 32, pat = (k / 32) % 8, ky = (k / 256) % 3, pl = k / 768;
      float t = 0.f;
      for (int kx = 0; kx < 3; kx++)
        if (pat & (1 << kx)) t = __fadd_rn(t, c1w[(co * 4 + pl) * 9 + ky * 3 + kx]);
      r.w1_tab[k] = t;
    } else if (seg == 3) {                            // w1_umma [4][32][32]: hi, lo (planes 0..2 + bias), plane-3 hi, lo
      const int kk = k % 32, co = (k / 32) % 32, mat = k / 1024;
      float v = 0.f;
      if (mat < 2) {
        if (kk < 27) { const int ky = kk / 9, rem = kk % 9, pl = rem / 3, kx = rem % 3; v = c1w[(co * 4 + pl) * 9 + ky * 3 + kx]; }
        else if (kk == 27) v = c1b[co];
      } else if (kk < 9) {
        v = c1w[(co * 4 + 3) * 9 + kk];
      }
      const __nv_bfloat16 hi = __float2bfloat16(v);
      r.w1_umma[k] = (mat & 1) ? __float2bfloat16(__fsub_rn(v, __bfloat162float(hi))) : hi;
    } else if (seg == 4) {                            // w2 [tap][64][32]
      const int ci = k % 32, co = (k / 32) % 64, t = k / 2048;
      r.w2[k] = __float2bfloat16(r.blob[r.o_c2w + ((size_t)co * 32 + ci) * 9 + t]);
    } else if (seg == 5) {                            // w3 [tap][128][64]
      const int ci = k % 64, co = (k / 64) % 128, t = k / 8192;
      r.w3[k] = __float2bfloat16(r.blob[r.o_c3w + ((size_t)co * 64 + ci) * 9 + t]);
    } else if (seg == 6) {                            // head_w16 [16][128]: policy rows, value rows, zero padding
      const int c = k % 128, j = k / 128;
      const float v = j < 4 ? r.blob[r.o_pcw + j * 128 + c] : (j < 6 ? r.blob[r.o_vcw + (j - 4) * 128 + c] : 0.f);
      r.head_w16[k] = __float2bfloat16(v);
    } else if (seg == 7) {                            // wp16 [A][PKp]
      const int kk = k % r.PKp, a = k / r.PKp;
      r.wp16[k] = __float2bfloat16(kk < PK ? r.blob[r.o_pfw + (size_t)a * PK + kk] : 0.f);
    } else if (seg == 8) {                            // wv16 [64][VKp]
      const int kk = k % r.VKp, j = k / r.VKp;
      r.wv16[k] = __float2bfloat16(kk < VK ? r.blob[r.o_v1w + (size_t)j * VK + kk] : 0.f);
    } else if (seg == 9) {                            // conv biases + head bias
      if (k < 32) r.b1[k] = c1b[k];
      else if (k < 96) r.b2[k - 32] = r.blob[r.o_c2b + k - 32];
      else if (k < 224) r.b3[k - 96] = r.blob[r.o_c3b + k - 96];
      else { const int j = k - 224; r.head_b8[j] = j < 4 ? r.blob[r.o_pcb + j] : (j < 6 ? r.blob[r.o_vcb + j - 4] : 0.f); }
    } else if (seg == 10) {                           // FC biases and the value tail
      if (k < A) r.pfc_b[k] = r.blob[r.o_pfb + k];
      else if (k < A + 64) r.vfc1_b[k - A] = r.blob[r.o_v1b + k - A];
      else if (k < A + 128) r.vfc2_w[k - A - 64] = r.blob[r.o_v2w + k - A - 64];
      else r.vfc2_b[0] = r.blob[r.o_v2b];
    } else if (seg == 11) {                           // w1_tab2 [ky][own3 | opp3 << 3][32]: kx ascending per plane, then own + opp
      const int co = k % 32, pat = (k / 32) % 64, ky = k / 2048;
      float p0 = 0.f, p1 = 0.f;
      for (int kx = 0; kx < 3; kx++) {
        if (pat & (1 << kx)) p0 = __fadd_rn(p0, c1w[(co * 4 + 0) * 9 + ky * 3 + kx]);
        if (pat & (8 << kx)) p1 = __fadd_rn(p1, c1w[(co * 4 + 1) * 9 + ky * 3 + kx]);
      }
      r.w1_tab2[k] = __fadd_rn(p0, p1);
    } else if (seg == 12) {                           // w1_lm [ky][kx][32]
      const int co = k % 32, t = k / 32;
      r.w1_lm[k] = c1w[(co * 4 + 2) * 9 + t];
    } else {                                          // w1_taps [10][64 = hi | lo x 32][8]
      const int kk = k % 8, nrow = (k / 8) % 64, t = k / 512, co = nrow & 31;
      float v = 0.f;
      if (t < 9) {
        if (kk < 3) v = c1w[(co * 4 + kk) * 9 + t];
        else if (kk == 3) v = t == 4 ? c1b[co] : 0.f;
        else if (kk == 4) v = c1w[(co * 4 + 3) * 9 + t];
      }
      const __nv_bfloat16 hi = __float2bfloat16(v);
      r.w1_taps[k] = (nrow & 32) ? __float2bfloat16(__fsub_rn(v, __bfloat162float(hi))) : hi;
    }
  }
}

// policy_fc / value_fc1 as bf16 [out][K padded to 8] (B operands of fc_umma_kernel) + fp32 biases / tails
static int fc_weights_build(azp_engine* e, const HostNet& net, NetWeights* W, UmmaWeights* U) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  const int A = e->A;
  const int PK = net.PC * A, VK = net.VC * A;
  int st;
  if ((PK + 7) / 8 * 8 != ws->PKp || (VK + 7) / 8 * 8 != ws->VKp) { e->set_error("head channel counts do not match the configured net"); return AZP_ERR_UNSUPPORTED; }
  std::vector<__nv_bfloat16> wp((size_t)A * ws->PKp, __float2bfloat16(0.f)), wv((size_t)64 * ws->VKp, __float2bfloat16(0.f));
  for (int a = 0; a < A; a++) for (int k = 0; k < PK; k++) wp[(size_t)a * ws->PKp + k] = __float2bfloat16(net.pfc_w[(size_t)a * PK + k]);
  for (int j = 0; j < 64; j++) for (int k = 0; k < VK; k++) wv[(size_t)j * ws->VKp + k] = __float2bfloat16(net.vfc1_w[(size_t)j * VK + k]);
  if ((st = up(e, &U->wp16, wp))) return st;
  if ((st = up(e, &U->wv16, wv))) return st;
  if ((st = make_map(e, &U->map_wp, U->wp16, A, ws->PKp, 64, 64))) return st;
  if ((st = make_map(e, &U->map_wv, U->wv16, 64, ws->VKp, 64, 64))) return st;
  if ((st = up(e, &W->pfc_b, net.pfc_b))) return st;
  if ((st = up(e, &W->vfc1_b, net.vfc1_b))) return st;
  if ((st = up(e, &W->vfc2_w, net.vfc2_w))) return st;
  if ((st = up(e, &W->vfc2_b, net.vfc2_b))) return st;
  W->C = net.C; W->PC = net.PC; W->VC = net.VC;
  return AZP_OK;
}

static int launch_fc(azp_engine* e, NetWeights* W, const int32_t* count_dev, int count_host, int back, float* logits,
                     float* policy, float* value, int out_by_index, cudaStream_t s);

// Returns AZP_OK after enqueueing the repack on `s`, AZP_ERR_UNSUPPORTED when this weight set cannot be refreshed in
// place (tower, first load): the caller then takes the host path.
int azp_umma_reload_device(azp_engine* e, NetWeights* W, const float* dev_blob, cudaStream_t s) {
  if (e->cfg.net_kind != AZP_NET_STEM || !W || !W->loaded || !W->umma) return AZP_ERR_UNSUPPORTED;
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  UmmaWeights* U = (UmmaWeights*)W->umma;
  const size_t A = e->A;
  StemRepack r;
  r.blob = dev_blob; r.A = (int)A; r.PKp = ws->PKp; r.VKp = ws->VKp;
  size_t o = 0;
  auto take = [&](size_t& off, size_t cnt) { off = o; o += cnt; };
  take(r.o_c1w, 32 * 4 * 9); take(r.o_c1b, 32); take(r.o_c2w, 64 * 32 * 9); take(r.o_c2b, 64);
  take(r.o_c3w, 128 * 64 * 9); take(r.o_c3b, 128); take(r.o_pcw, 4 * 128); take(r.o_pcb, 4);
  take(r.o_pfw, A * 4 * A); take(r.o_pfb, A); take(r.o_vcw, 2 * 128); take(r.o_vcb, 2);
  take(r.o_v1w, 64 * 2 * A); take(r.o_v1b, 64); take(r.o_v2w, 64); take(r.o_v2b, 1);
  r.w1_bits = U->w1_bits; r.w1_tab = U->w1_tab; r.w1_planes = U->w1_planes; r.b1 = U->b1; r.b2 = U->b2; r.b3 = U->b3;
  r.w1_tab2 = U->w1_tab2; r.w1_lm = U->w1_lm;
  r.head_b8 = U->head_b8; r.pfc_b = W->pfc_b; r.vfc1_b = W->vfc1_b; r.vfc2_w = W->vfc2_w; r.vfc2_b = W->vfc2_b;
  r.w1_taps = U->w1_taps;
  r.w1_umma = U->w1_umma; r.w2 = U->w2; r.w3 = U->w3; r.head_w16 = U->head_w16; r.wp16 = U->wp16; r.wv16 = U->wv16;
  if (!r.w1_tab2 || !r.w1_lm || !r.w1_bits || !r.w1_tab || !r.w1_planes || !r.b1 || !r.b2 || !r.b3 || !r.head_b8 || !r.pfc_b || !r.vfc1_b ||
      !r.vfc2_w || !r.vfc2_b || !r.w1_umma || !r.w1_taps || !r.w2 || !r.w3 || !r.head_w16 || !r.wp16 || !r.wv16)
    return AZP_ERR_UNSUPPORTED;
  stem_repack_kernel<<<148 * 2, 256, 0, s>>>(r);
  azp_count_launch();
  return azp_check_launch();
}

// residual tower: first conv -> 2 convs per block (the second adds the block input) -> 1x1 heads -> FC
static int run_tower(azp_engine* e, NetWeights* W, const int32_t* count_dev, int count_host, int back, int n_planes,
                     const float* planes, float* logits, float* policy, float* value, int out_by_index, cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  UmmaWeights* U = (UmmaWeights*)W->umma;
  const ConvGeom& g = ws->g;
  const int zmax = count_dev ? g.G : count_host;
  if (zmax <= 0) return AZP_OK;
  const int C = TW_C;
  const int grid1 = zmax < 4 * ws->sms ? zmax : 4 * ws->sms;
  if (planes) {
    tower_first_planes_kernel<<<grid1, 256, (37 * C + 4 * g.A) * sizeof(float), s>>>(g, C, n_planes, planes, U->tf_planes,
                                                                                    U->tf_b, ws->tact[0]);
  } else {
    const size_t sm = (size_t)28 * C * sizeof(float);
    AZP_DISPATCH_NW(e->NW, tower_first_bits_kernel, grid1, 256, sm, s, g, C, e->tv.leaf_x, e->tv.leaf_o, e->tv.leaf_tm,
                    e->tv.leaf_last, U->tf_bits, U->tf_b, ws->tact[0], count_dev, count_host, back);
  }
  azp_count_launch();
  const int max_tiles = (int)(((long long)zmax * g.stride + TILE_M - 1) / TILE_M);
  const int max_pairs = (max_tiles + 1) / 2;
  const int grid = max_pairs < ws->sms ? max_pairs : ws->sms;
  int xi = 0;                               // buffer holding the current block input x; y = xi+1, block output = xi+2 (mod 3)
  e->prof_flops_per_board = 2.0 * 9.0 * C * C * g.A;
  for (size_t i = 0; i < U->tw.size(); i++) {
    const bool second = U->t_residual[i] != 0;       // second conv of a block: relu(bn2(conv2(y)) + x)
    TowerArgs ta;
    ta.g = g; ta.bias = U->tb[i]; ta.count_dev = count_dev; ta.count_host = count_host; ta.back = back;
    ta.a_rows = ws->t_a_rows; ta.a_bytes = ws->t_a_bytes;
    const int in_i = second ? (xi + 1) % 3 : xi;
    const int out_i = second ? (xi + 2) % 3 : (xi + 1) % 3;
    ta.residual = second ? ws->tact[xi] : nullptr;
    ta.out = ws->tact[out_i];
    azp_prof_begin(e, AZP_PROF_DOMINANT, count_dev, count_host, s);
    conv_tower_kernel<<<grid, TW_THREADS, ws->smem_tower, s>>>(ws->map_tact[in_i], U->map_tw[i], ta);
    azp_count_launch();
    azp_prof_end(e, AZP_PROF_DOMINANT, s);
    if (second) xi = out_i;
  }
  const int NJ = W->PC + W->VC;
  const long long total = (long long)zmax * g.A;
  const int hgrid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  tower_heads_kernel<<<hgrid, 256, (NJ * C + NJ) * sizeof(float), s>>>(g, C, NJ, W->PC, ws->PKp, ws->FK, ws->tact[xi], U->th_w,
                                                                      U->th_b, ws->feat, count_dev, count_host, back);
  azp_count_launch();
  return launch_fc(e, W, count_dev, count_host, back, logits, policy, value, out_by_index, s);
}

// conv2 + conv3(+heads) + FC on rows already holding act1
static int run_trunk(azp_engine* e, NetWeights* W, const int32_t* count_dev, int count_host, int back,
                     float* logits, float* policy, float* value, int out_by_index, bool from_bits, cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  UmmaWeights* U = (UmmaWeights*)W->umma;
  const ConvGeom& g = ws->g;
  const int zmax = count_dev ? g.G : count_host;
  const int max_tiles = (int)(((long long)zmax * g.stride + TILE_M - 1) / TILE_M);
  const int grid = max_tiles < ws->sms ? max_tiles : ws->sms;
  ConvArgs a2;
  a2.g = g; a2.bias = U->b2; a2.out = ws->act2; a2.head_b = nullptr; a2.feat = nullptr; a2.NJ = 0;
  a2.PC = 0; a2.PKp = 0; a2.FK = 0;
  a2.count_dev = count_dev; a2.count_host = count_host; a2.back = back; a2.stage_rows = TILE_M + 2 * g.P;
  a2.single = from_bits ? 1 : ws->single;
  a2.clk = bringup_clk(ws, 2);
  a2.stage_bytes = (a2.single ? (a2.stage_rows + 2 + 7) / 8 * 8 : a2.stage_rows) * 64;
  a2.lx = e->tv.leaf_x; a2.lo = e->tv.leaf_o; a2.ltm = e->tv.leaf_tm; a2.llast = e->tv.leaf_last; a2.NW = e->NW;
  a2.w1 = U->w1_tab; a2.b1 = U->b1;
  a2.fc_fused = 0; a2.fc_slices = 0; a2.fc_pk_chunks = 0; a2.fc_vk_chunks = 0; a2.fc_v_off = 0; a2.fc_A = g.A;
  a2.fc_pbias = nullptr; a2.fc_vbias = nullptr; a2.fc_vw2 = nullptr; a2.fc_vb2 = nullptr; a2.fc_logits = nullptr; a2.fc_value = nullptr;
  a2.feat_done = ws->feat_done; a2.fc_done = ws->fc_done;       // conv2 resets them for the conv3 launch that follows
  a2.sched = from_bits ? nullptr : ws->sched2;
  a2.dead = nullptr; a2.dead_row_bytes = 0;
  { static int ps = -1; if (ps < 0) { const char* v = getenv("AZP_PAIRED_STORES"); ps = v ? atoi(v) : 0; } a2.paired_stores = ps; }      // measured in the default build, same call: 15.16 M with, 15.22-15.25 M without
  a2.debug = bringup_debug_bits();
  { const char* v = getenv("AZP_CLK_CTA"); a2.clk_cta = v ? atoi(v) : 0; }
  azp_prof_begin(e, AZP_PROF_CONV2, count_dev, count_host, s);
  if (from_bits)
    azp_launch(conv_umma_kernel<32, 64, STAGES2, false, true>, grid, CONV_THREADS_BITS, ws->smem2, s, true, ws->map_act1,
               U->map_w2, ws->map_act2_out, ws->map_feat, U->map_wp, U->map_wv, a2);
  else
    azp_launch(conv_umma_kernel<32, 64, STAGES2, false, false>, grid, CONV_THREADS_EPI2, ws->smem2, s, true,
               a2.single ? ws->map_act1_s : ws->map_act1, U->map_w2, ws->map_act2_out, ws->map_feat, U->map_wp, U->map_wv, a2);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_CONV2, s);
  ConvArgs a3 = a2;
  a3.sched = nullptr;
  if (ws->discard && !from_bits) { a3.dead = ws->act1; a3.dead_row_bytes = 64; }
  a3.bias = U->b3; a3.out = nullptr; a3.head_b = U->head_b8; a3.feat = ws->feat;
  a3.NJ = W->PC + W->VC; a3.PC = W->PC; a3.PKp = ws->PKp; a3.FK = ws->FK;
  a3.single = ws->single;
  a3.clk = bringup_clk(ws, 1);
  if (!a3.clk) a3.clk = bringup_clk(ws, 5);            // 5: stamps of the fused FC items (tools/fc_fused_probe.py)
  // The FC layers ride inside conv3 for the leaf batches the search evaluates from the front (raw logits into the engine's
  // policy buffer); azp_net_forward (softmax wanted, caller's buffers), the arena's back-filled batch and the CTA-pair
  // variant keep the separate FC launch.  Both run the same UMMA sequence: bit-identical outputs.
  const bool fused = ws->fuse_fc && !ws->pair && a3.single == 1 && !out_by_index && !back && policy && value;
  a3.stage_bytes = a3.single ? conv_stage_bytes(a3.stage_rows, 128, ws->fuse_fc != 0) : a3.stage_rows * 128;
  if (fused) {
    a3.fc_fused = 1;
    a3.fc_slices = (g.A + FC_N - 1) / FC_N; a3.fc_pk_chunks = (ws->PKp + 63) / 64; a3.fc_vk_chunks = (ws->VKp + 63) / 64;
    a3.fc_v_off = ws->PKp; a3.fc_A = g.A;
    a3.fc_pbias = W->pfc_b; a3.fc_vbias = W->vfc1_b; a3.fc_vw2 = W->vfc2_w; a3.fc_vb2 = W->vfc2_b;
    a3.fc_logits = policy; a3.fc_value = value;
  }
  e->prof_flops_per_board = 2.0 * 9.0 * 64 * 128 * g.A;
  azp_prof_begin(e, AZP_PROF_DOMINANT, count_dev, count_host, s);
  if (ws->pair) {
    const int max_pairs = (max_tiles + 1) / 2;
    const int ncl = max_pairs < ws->pair_clusters ? max_pairs : ws->pair_clusters;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * ncl); cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = ws->smem3p; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = azp_pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, conv3_pair_kernel, ws->map_act2_s, U->map_w3h, U->map_hwh, a3);
  } else {
    azp_launch(conv_umma_kernel<64, 128, STAGES3, true, false>, grid, CONV_THREADS, ws->smem3, s, true,
               a3.single ? ws->map_act2_s : ws->map_act2, U->map_w3, U->map_hw, ws->map_feat, U->map_wp, U->map_wv, a3);
  }
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_DOMINANT, s);
  if (fused) return azp_check_launch();
  return launch_fc(e, W, count_dev, count_host, back, logits, policy, value, out_by_index, s);
}

static int launch_fc(azp_engine* e, NetWeights* W, const int32_t* count_dev, int count_host, int back, float* logits,
                     float* policy, float* value, int out_by_index, cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  UmmaWeights* U = (UmmaWeights*)W->umma;
  const ConvGeom& g = ws->g;
  const int zmax = count_dev ? g.G : count_host;
  // The kernel writes RAW LOGITS.  Leaf path: `policy` is the engine's d_policy, consumed as logits by the expansion
  // kernels (policy_logits = 1).  azp_net_forward: logits go to the caller's buffer (or the engine's scratch) and a
  // one-warp-per-row softmax with the consumers' arithmetic produces `policy`.
  const bool want_softmax = out_by_index != 0;
  float* logits_out = want_softmax ? (logits ? logits : e->d_logits) : policy;
  FcArgs fa;
  fa.A = g.A; fa.G = g.G; fa.stages = ws->fc_stages;
  fa.n_slices = (g.A + FC_N - 1) / FC_N;
  fa.pk_chunks = (ws->PKp + 63) / 64; fa.vk_chunks = (ws->VKp + 63) / 64; fa.v_off = ws->PKp;
  fa.pbias = W->pfc_b; fa.vbias = W->vfc1_b;
  fa.vfc2_w = W->vfc2_w; fa.vfc2_b = W->vfc2_b; fa.logits = logits_out; fa.value = value;
  fa.count_dev = count_dev; fa.count_host = count_host; fa.back = back; fa.out_by_index = out_by_index;
  fa.clk = bringup_clk(ws, 4);
  fa.dead = (ws->discard && !ws->tower) ? ws->act2 : nullptr; fa.dead_row_bytes = 128; fa.dead_stride = g.stride;
  dim3 fc_grid((zmax + TILE_M - 1) / TILE_M, fa.n_slices + 1);
  azp_prof_begin(e, AZP_PROF_FC, count_dev, count_host, s);
  azp_launch(fc_umma_kernel, fc_grid, FC_THREADS, ws->fc_smem, s, true, ws->map_feat, U->map_wp, U->map_wv, fa);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_FC, s);
  int st = azp_check_launch();
  if (st) return st;
  if (want_softmax && policy) return azp_softmax_rows(count_host, g.A, logits_out, policy, s);
  return AZP_OK;
}

// conv1 on the tensor core: leaf bitboards (planes == null) or float planes -> act1
static int launch_conv1_umma(azp_engine* e, UmmaWeights* U, const int32_t* count_dev, int count_host, int back,
                             const float* planes, cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  const ConvGeom& g = ws->g;
  const int zmax = count_dev ? g.G : count_host;
  const int max_tiles = (int)(((long long)zmax * g.stride + TILE_M - 1) / TILE_M);
  const int grid = max_tiles < 2 * ws->sms ? max_tiles : 2 * ws->sms;       // two CTAs per SM
  if (ws->conv1_mode == 3) {
    Conv1XArgs x;
    x.g = g; x.w = U->w1_taps; x.count_dev = count_dev; x.count_host = count_host; x.back = back;
    x.lx = e->tv.leaf_x; x.lo = e->tv.leaf_o; x.ltm = e->tv.leaf_tm; x.llast = e->tv.leaf_last; x.NW = e->NW;
    x.planes = planes; x.out = ws->act1; x.pairs = ws->conv1_pairs;
    azp_prof_begin(e, AZP_PROF_CONV1, count_dev, count_host, s);
    if (planes) azp_launch(conv1_taps_kernel<true>, grid, C1X_THREADS, conv1x_smem_bytes(), s, true, x);
    else azp_launch(conv1_taps_kernel<false>, grid, C1X_THREADS, conv1x_smem_bytes(), s, true, x);
    azp_count_launch();
    azp_prof_end(e, AZP_PROF_CONV1, s);
    return azp_check_launch();
  }
  Conv1Args a;
  a.g = g; a.w = U->w1_umma; a.count_dev = count_dev; a.count_host = count_host; a.back = back;
  a.lx = e->tv.leaf_x; a.lo = e->tv.leaf_o; a.ltm = e->tv.leaf_tm; a.llast = e->tv.leaf_last; a.NW = e->NW;
  a.planes = planes; a.out = ws->act1;
  { static int staged = -1; if (staged < 0) { const char* v = getenv("AZP_CONV1_STORE"); staged = (v && !strcmp(v, "tma")) ? 1 : 0; } a.staged = staged; }
  a.clk = bringup_clk(ws, 3);
  azp_prof_begin(e, AZP_PROF_CONV1, count_dev, count_host, s);
  if (planes) azp_launch(conv1_umma_kernel<true>, grid, C1_THREADS, conv1_smem_bytes(true), s, true, ws->map_act1_out, a);
  else azp_launch(conv1_umma_kernel<false>, grid, C1_THREADS, conv1_smem_bytes(false), s, true, ws->map_act1_out, a);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_CONV1, s);
  return azp_check_launch();
}

// first evaluation on a stream (outside graph capture): attach the activations' L2 access-policy window to it; kernels
// captured from the stream later inherit it
static void umma_stream_policy(azp_engine* e, UmmaWorkspace* ws, cudaStream_t s) {
  if (!ws->l2_persist || e->capturing) return;
  for (cudaStream_t q : ws->policy_streams) if (q == s) return;
  ws->policy_streams.push_back(s);
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  v.accessPolicyWindow.base_ptr = ws->window_base;
  v.accessPolicyWindow.num_bytes = ws->window_bytes;
  v.accessPolicyWindow.hitRatio = ws->window_hit;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

// Can the descent kernels run conv1 of their leaves as their tail?  Exact mode on the bf16 stem with the nine-tap conv1 only
// (AZP_FUSE_DESCENT=0: off).
int azp_umma_conv1_tail(azp_engine* e, int slot, Conv1TailArgs* out) {
  out->enabled = 0;
  if (e->cfg.net_dtype != AZP_DTYPE_BF16 || e->cfg.net_kind != AZP_NET_STEM || e->K != 1 || !e->ws || !e->ws->umma) return 0;
  NetWeights* W = e->weights[slot];
  if (!W || !W->loaded || !W->umma) return 0;
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  UmmaWeights* U = (UmmaWeights*)W->umma;
  static int fuse = -1;
  if (fuse < 0) { const char* v = getenv("AZP_FUSE_CONV1"); fuse = v ? atoi(v) : 0; }
  if (!ws->fuse_descent || ws->tower || ws->conv1_mode != 3 || !ws->conv1_pairs || fuse || !U->w1_taps) return 0;
  // one CTA per SM (see azp_conv1_tail.cuh): 4, 8 or 16 games per CTA, the fewest that seats every game -- with 16 per CTA the
  // 400-game arena squeezed its conv1 into 25 SMs and lost 20 %; more than 16 * SMs games would make the descent a second wave
  const int g4 = (e->cfg.games + 3) / 4;
  const int groups = g4 <= ws->sms ? 1 : (g4 <= 2 * ws->sms ? 2 : 4);
  if (e->cfg.games > 16 * ws->sms) return 0;
  // Fewer games than half a machine-full: every CTA's boards go through its tail one after the other while the conv1 KERNEL
  // spreads the same tiles over two CTAs per SM -- measured at config 5 (400 games of 19x19, four games per CTA on 100 SMs):
  // 4.88 M simulations/s with the tail, 5.20 M without.  AZP_FUSE_DESCENT_MIN_GAMES (read when the engine is created) moves the
  // threshold; the tests lower it to cover the one- and two-group instantiations.
  if (e->cfg.games < (ws->fuse_min_games >= 0 ? ws->fuse_min_games : 8 * ws->sms)) return 0;
  out->g = ws->g; out->w = U->w1_taps; out->out = ws->act1; out->enabled = groups;
  return 1;
}

int azp_umma_eval_leaves(azp_engine* e, NetWeights* W, const int32_t* count_dev, int count_host, int back,
                         cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  umma_stream_policy(e, ws, s);
  UmmaWeights* U = (UmmaWeights*)W->umma;
  const ConvGeom& g = ws->g;
  const int zmax = count_dev ? g.G : count_host;
  if (zmax <= 0) return AZP_OK;
  if (ws->tower) return run_tower(e, W, count_dev, count_host, back, 0, nullptr, nullptr, e->d_policy, e->d_value, 0, s);
  static int fuse = -1;
  if (fuse < 0) { const char* v = getenv("AZP_FUSE_CONV1"); fuse = v ? atoi(v) : 0; }
  const bool conv1_done = e->conv1_done;       // one-shot: set by the engine after a descent kernel with the conv1 tail
  e->conv1_done = false;
  if (conv1_done) {
    // act1 of this batch is already in place
  } else if (!fuse && (ws->conv1_mode == 1 || ws->conv1_mode == 3)) {
    int st = launch_conv1_umma(e, U, count_dev, count_host, back, nullptr, s);
    if (st) return st;
  } else if (!fuse && ws->conv1_mode == 2) {
    // two warps per board, four boards per CTA, four CTAs per SM (boards are dealt to the CTAs round robin)
    const int grid = zmax < 4 * ws->sms ? zmax : 4 * ws->sms;
    azp_prof_begin(e, AZP_PROF_CONV1, count_dev, count_host, s);
    AZP_DISPATCH_NW_PDL(e->NW, conv1_table_kernel, grid, 256, 0, s, g, e->tv.leaf_x, e->tv.leaf_o, e->tv.leaf_tm,
                        e->tv.leaf_last, U->w1_tab2, U->w1_lm, U->b1, ws->act1, count_dev, count_host, back);
    azp_count_launch();
    azp_prof_end(e, AZP_PROF_CONV1, s);
  } else if (!fuse) {
    const int grid = zmax < 4 * ws->sms ? zmax : 4 * ws->sms;
    AZP_DISPATCH_NW(e->NW, conv1_bits_kernel, grid, 256, 0, s, g, e->tv.leaf_x, e->tv.leaf_o, e->tv.leaf_tm,
                    e->tv.leaf_last, U->w1_tab, U->b1, ws->act1, count_dev, count_host, back);
    azp_count_launch();
  }
  return run_trunk(e, W, count_dev, count_host, back, nullptr, e->d_policy, e->d_value, 0, /*from_bits=*/fuse != 0, s);
}

// diagnostics: intermediate activations of the last evaluation as dense float32
__global__ void unpad_act_kernel(ConvGeom g, int n, int C, const __nv_bfloat16* __restrict__ act, float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * g.A * C) return;
  int c = i % C; size_t t = i / C; int a = t % g.A; int slot = t / g.A;
  size_t row = (size_t)slot * g.stride + (size_t)(a / g.B + 1) * g.P + (a % g.B + 1);
  out[i] = __bfloat162float(act[row * C + c]);
}
__global__ void unpack_feat_kernel(int A, int n, int PKp, int FK, const __nv_bfloat16* __restrict__ feat, float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * 6 * A) return;
  int pix = i % A, j = (i / A) % 6; size_t slot = i / (6 * (size_t)A);
  out[i] = __bfloat162float(feat[slot * FK + (j < 4 ? j * A : PKp + (j - 4) * A) + pix]);
}
int azp_umma_debug(azp_engine* e, int layer, int n, float* out, cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  const ConvGeom& g = ws->g;
  if (layer == 9) {            // clock stamps of the last conv3 launch: [64][16] int64 reinterpreted, as floats of deltas
    if (!ws->clk) return AZP_ERR_INVALID_ARG;
    AZP_CUDA(e, cudaMemcpyAsync(out, ws->clk, 64 * 16 * 8, cudaMemcpyDeviceToDevice, s));
    return AZP_OK;
  }
  if (layer == 1 || layer == 2) {
    int C = layer == 1 ? 32 : 64;
    size_t total = (size_t)n * g.A * C;
    unpad_act_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(g, n, C, layer == 1 ? ws->act1 : ws->act2, out);
    azp_count_launch();
    return azp_check_launch();
  }
  if (layer == 3) {
    size_t total = (size_t)n * 6 * g.A;
    unpack_feat_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(g.A, n, ws->PKp, ws->FK, ws->feat, out);
    azp_count_launch();
    return azp_check_launch();
  }
  return AZP_ERR_INVALID_ARG;
}

int azp_umma_eval_planes(azp_engine* e, NetWeights* W, int n, const float* planes, float* logits, float* policy,
                         float* value, cudaStream_t s) {
  UmmaWorkspace* ws = (UmmaWorkspace*)e->ws->umma;
  UmmaWeights* U = (UmmaWeights*)W->umma;
  const ConvGeom& g = ws->g;
  umma_stream_policy(e, ws, s);
  if (ws->tower) return run_tower(e, W, nullptr, n, 0, n, planes, logits, policy, value, 1, s);
  if (ws->conv1_mode == 1 || ws->conv1_mode == 3) {
    int st = launch_conv1_umma(e, U, nullptr, n, 0, planes, s);
    if (st) return st;
  } else if (ws->conv1_mode == 2) {
    const int grid = n < 4 * ws->sms ? n : 4 * ws->sms;
    conv1_table_planes_kernel<<<grid, 256, 4 * g.A * sizeof(float), s>>>(g, n, planes, U->w1_planes, U->b1, ws->act1);
    azp_count_launch();
  } else {
    const int grid = n < 4 * ws->sms ? n : 4 * ws->sms;
    conv1_planes_kernel<<<grid, 256, 4 * g.A * sizeof(float), s>>>(g, n, planes, U->w1_planes, U->b1, ws->act1);
    azp_count_launch();
  }
  return run_trunk(e, W, nullptr, n, 0, logits, policy, value, 1, /*from_bits=*/false, s);
}
