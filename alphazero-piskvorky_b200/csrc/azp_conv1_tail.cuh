// conv1 of the leaves a descent kernel has just emitted, as the TAIL of that kernel (exact mode, K = 1).
//
// The step chain is advance -> conv1 -> conv2 -> conv3 -> fc.  advance (one warp per game, 4 warps per CTA, 4 CTAs per SM)
// is a chain of dependent memory round trips that leaves the SM's issue slots, shared-memory port and tensor pipe idle;
// conv1 is 0.3 % of the net's FLOPs but a launch of its own: 13 us + 1.3 us of dependent-launch gap per step, bound by its
// 33.5 MB of row stores (profiles/r3_step_timeline.txt, r3_store_rate.txt).  Here every CTA of the descent kernel, once its
// four warps have emitted their leaves, runs the nine-tap conv1 (azp_net_conv1.cuh: 16-byte input rows, two taps per UMMA)
// for exactly those (<= 4) boards: its stores overlap the other CTAs' tree walks, and the launch, its gap and its ramp go.
// A board is self-contained in the padded layout (row group 0 / column 0 are halo), so the tail works per board:
// ceil(stride / 128) tiles starting at the board's first row, rows past the board are computed and not stored.
// Same operands, same UMMA sequence per output row as conv1_taps_kernel -> bit-identical act1.
#pragma once
#include "azp_tc.cuh"

constexpr int C1T_MAX_ROWS = 4 * TILE_M + 2 * 24 + 8;        // <= 4 tiles per board (19x19: 400 rows) + halo rows above / below
// GROUPS = warp groups (4 warps = one team's worth of tensor-memory lanes) per CTA: 4 games per group.  The kernel is instantiated
// for 1, 2 and 4 groups and the launcher takes the smallest that seats all games in one CTA per SM (4 * GROUPS * SMs games): few
// games are then spread over many SMs instead of being squeezed into games / 16 of them.
constexpr int C1T_GROUP_BYTES = C1T_MAX_ROWS * 16;
constexpr int c1t_threads(int groups) { return 128 * groups; }
constexpr int c1t_smem(int groups) { return 1024 + C1X_W_BYTES + groups * C1T_GROUP_BYTES + 128; }

// Why ONE 512-thread CTA per SM instead of the descent kernels' four 128-thread CTAs: a kernel that contains tcgen05.alloc was
// given ONE resident CTA per SM on this driver (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for every tcgen05
// kernel here; the 128-thread version of this kernel ran its 512 CTAs as four waves, 78 us instead of 23).  Sixteen warps per
// SM are what the four small CTAs gave the latency-bound tree walk anyway.

// Prologue part (before griddepcontrol.wait: touches nothing a predecessor writes): conv1 weights -> shared memory, barrier init.
template <int C1T_GROUPS>
static __device__ __forceinline__ void conv1_tail_prologue(const Conv1TailArgs& a, uint8_t* smem) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C1X_W_BYTES + C1T_GROUPS * C1T_GROUP_BYTES);
  if (threadIdx.x == 0) {
    for (int i = 0; i < C1T_GROUPS; i++) mbar_init(bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C1X_W_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(a.w)[i];
}

// After griddepcontrol.wait (the previous kernel of the chain, which held tensor memory on this SM, has completed): 128 columns
// per group -- this CTA is alone on its SM (see above).  Every thread of the CTA calls this; returns the base address.
template <int C1T_GROUPS>
static __device__ __forceinline__ uint32_t conv1_tail_alloc(uint8_t* smem) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C1X_W_BYTES + C1T_GROUPS * C1T_GROUP_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C1T_GROUPS);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // the prologue's weight copy -> visible to UMMA
  if ((threadIdx.x >> 5) == 0) tmem_alloc(tmem_slot, 128 * C1T_GROUPS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *tmem_slot;
}
template <int C1T_GROUPS>
static __device__ __forceinline__ void conv1_tail_free(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem, 128 * C1T_GROUPS);
}

// `slots`: shared array of the CTA's (<= 16) leaf slots, -1 = that warp emitted no leaf.  The four warps of TEAM `wg` -- one of
// each tensor-memory lane quarter (warp % 4), teamed up in the order they finished their descents -- take the team's boards as
// soon as all four are there (named barrier 1 + wg; the teams do not wait for each other, so the early finishers' stores
// overlap the slow warps' tree walks), with their own input rows, accumulators (128 of the CTA's 512 tensor-memory columns)
// and mbarrier.  slots[4 * wg + (warp % 4)] is this warp's entry.
// slots[] entries are (slot | emitting warp << 20); `leaves` = the CTA's leaf records in shared memory, `leaf_words` 64-bit words
// per warp: x words, o words, (to-move | last move << 8).
template <int C1T_GROUPS>
static __device__ __forceinline__ void conv1_tail(const Conv1TailArgs& a, uint8_t* smem, uint32_t tmem, int wg, const int* slots,
                                                  const uint64_t* leaves, int leaf_words, int NW) {
  const ConvGeom g = a.g;
  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const int wq = warp & 3, gt = wq * 32 + lane;                        // TMEM lane quarter, thread within the team
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + C1X_W_BYTES + wg * C1T_GROUP_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C1X_W_BYTES + C1T_GROUPS * C1T_GROUP_BYTES);
  uint64_t* bar = bars + wg;
  asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");          // this group's leaf records and slots[] are written
  const uint32_t tmem_base = tmem + (uint32_t)(wg * 128);
  const int tiles = (g.stride + TILE_M - 1) / TILE_M;
  const int NR = tiles * TILE_M + 2 * g.P + 3;                         // stage row j = board row j - (P + 1)
  const unsigned p_rcp = (65536u + (unsigned)g.P - 1u) / (unsigned)g.P;
  const uint32_t idesc = umma_idesc(64);
  const uint32_t w_base = smem_u32(s_w), a_base = smem_u32(s_a);
  uint32_t phase = 0;
  for (int gi = 0; gi < 4; gi++) {
    const int entry = slots[wg * 4 + gi];
    if (entry < 0) continue;                                           // uniform over the team
    const int slot = entry & 0xFFFFF;
    const uint64_t* lx = leaves + (entry >> 20) * leaf_words;
    const uint64_t* lo = lx + NW;
    // ---- input rows of the board: (mine, theirs, last move, one) as bf16, zero outside the board's cells ----
    const int meta = (int)(uint32_t)lx[2 * NW];
    const bool xm = (meta & 0xFF) == 1;
    const int last = meta >> 8;
    for (int j = gt; j < NR; j += 128) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      const int pos = j - (g.P + 1);
      if (pos >= 0 && pos < g.stride) {
        const unsigned y = ((unsigned)pos * p_rcp) >> 16, x = (unsigned)pos - y * (unsigned)g.P;
        if (y >= 1u && x >= 1u && x <= (unsigned)g.B) {
          const unsigned cell = (y - 1u) * (unsigned)g.B + (x - 1u);
          const uint64_t wx = lx[cell >> 6], wo = lo[cell >> 6];
          const unsigned bx = (unsigned)(wx >> (cell & 63u)) & 1u, bo = (unsigned)(wo >> (cell & 63u)) & 1u;
          const unsigned mine = xm ? bx : bo, theirs = xm ? bo : bx;
          v.x = (mine ? 0x3F80u : 0u) | (theirs ? 0x3F800000u : 0u);
          v.y = (last == (int)cell ? 0x3F80u : 0u) | 0x3F800000u;
        }
      }
      *reinterpret_cast<uint4*>(s_a + j * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
    for (int t0 = 0; t0 < tiles; t0 += 2) {
      const int nt = tiles - t0 < 2 ? tiles - t0 : 2;
      if (wq == 0) {                                                   // whole warp, an elected lane issues (see elect_one)
        tc_fence_after();
        for (int tt = 0; tt < nt; tt++) {
          const uint32_t d = tmem_base + (uint32_t)(tt * 64);
          const uint32_t a0 = a_base + (uint32_t)((t0 + tt) * TILE_M) * 16u;
#pragma unroll
          for (int p = 0; p < 5; p++) {
            const int k0 = 2 * p, k1 = k0 + 1;
            const int off0 = (k0 / 3) * g.P + k0 % 3;
            const int off1 = k1 < 9 ? (k1 / 3) * g.P + k1 % 3 : off0 + 1;
            umma_bf16_e(d, umma_desc_rows16(a0 + (uint32_t)off0 * 16u, (uint32_t)(off1 - off0) * 16u),
                        umma_desc_rows16(w_base + (uint32_t)k0 * C1X_TAP_BYTES, C1X_TAP_BYTES), idesc, p ? 1u : 0u);
          }
        }
        umma_commit_e(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      tc_fence_after();
      for (int tt = 0; tt < nt; tt++) {
        const int pos = (t0 + tt) * TILE_M + wq * 32 + lane;          // row of the board
        const unsigned y = ((unsigned)pos * p_rcp) >> 16, x = (unsigned)pos - y * (unsigned)g.P;
        const bool interior = pos < g.stride && y >= 1u && x >= 1u && x <= (unsigned)g.B;
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tt * 64);
        uint32_t packed[16];
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          uint32_t r[16], r2[16];
          tmem_ld16_nowait(taddr + (uint32_t)(hf * 16), r);
          tmem_ld16_nowait(taddr + (uint32_t)(32 + hf * 16), r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float v0 = interior ? fmaxf(__uint_as_float(r[2 * j]) + __uint_as_float(r2[2 * j]), 0.f) : 0.f;
            const float v1 = interior ? fmaxf(__uint_as_float(r[2 * j + 1]) + __uint_as_float(r2[2 * j + 1]), 0.f) : 0.f;
            __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
            packed[hf * 8 + j] = *reinterpret_cast<uint32_t*>(&hh);
          }
        }
        {                                                              // rows past the board belong to the next slot: not stored
          uint8_t* gp = reinterpret_cast<uint8_t*>(a.out) + ((size_t)slot * g.stride + pos) * 64;
          const int pos_even = pos & ~1;
          store_rows64_paired(packed, gp, 64, pos_even < g.stride, pos_even + 1 < g.stride, lane);
        }
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");      // accumulators and the input rows may be reused
    }
  }
}
