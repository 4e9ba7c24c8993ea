// tcgen05 / mbarrier / TMEM PTX wrappers shared by the translation units that issue tensor-core work (azp_net_umma.cu and, for
// the conv1 tail of the descent kernels, azp_tree.cu).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "azp_internal.cuh"

constexpr int TILE_M = 128;


// ---- PTX wrappers -----------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trapped launch, never a hung GPU.
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// Same, for waits that are expected to be long and sit next to issue-bound warps (conv1 producers): back off with
// nanosleep so the spinning warp does not eat the issue slots of the warps it is waiting for.
static __device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
static __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
static __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
static __device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate; one thread issues for the CTA
static __device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-uniform issue: the WHOLE issuer warp runs the issue loop (uniform control flow, descriptors and loop counters live in
// uniform registers) and one elected lane executes each tcgen05 instruction.  With the loop under `if (lane == 0)` the
// compiler cannot use the uniform datapath: every UTCHMMA is preceded by an ELECT / R2UR.BROADCAST / BRA.U.ANY "waterfall"
// over its descriptor registers, ~15 dependent instructions = 100-130 cycles per UMMA from one thread -- more than the 48
// (N = 64) / 64 (N = 128) cycles the instruction occupies the tensor pipe (tools/bench_cu/umma_rate.cu).
static __device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
static __device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
static __device__ __forceinline__ void umma_bf16_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) umma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);
}
static __device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  if (elect_one()) umma_commit(bar);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N
static __device__ __forceinline__ uint32_t umma_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
static __device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
static __device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- nine-tap conv1 (azp_net_conv1.cuh, azp_conv1_tail.cuh): operand geometry ---------------------------------------
constexpr int C1X_TAP_BYTES = 64 * 16;                       // one K block of B: 64 rows (hi 32 | lo 32 output channels) x 8 bf16
constexpr int C1X_TAPS = 10;                                 // nine taps + one all-zero block
constexpr int C1X_W_BYTES = C1X_TAPS * C1X_TAP_BYTES;
constexpr int C1X_W_ELEMS = C1X_TAPS * 64 * 8;
// no-swizzle K-major operand: rows 16 B apart (8-row core matrices 128 B apart), K blocks `lbo_bytes` apart
static __device__ __forceinline__ uint64_t umma_desc_rows16(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) |
         (1ull << 46);
}

// ---- epilogue row stores --------------------------------------------------------------------------------------------
// tcgen05.ld 32x32b pins lane == tile row, so an epilogue lane owns 64 bytes of ITS row, and two 32-byte stores per lane put
// 32 bytes into each of 32 lines per instruction: 3.6 TB/s into L2 chip-wide.  64 CONTIGUOUS bytes per line and instruction
// reach the coalesced rate, 5.4 TB/s (tools/bench_cu/store_rate.cu, profiles/r3_store_rate.txt).  Lanes 2k / 2k+1 own rows 2k / 2k+1:
// after ONE exchange of eight registers with the partner lane, instruction 0 writes row 2k (even lane bytes 0-31, odd lane
// bytes 32-63) and instruction 1 row 2k+1.  `mine`: this lane's 64-byte segment; `pitch`: bytes between consecutive rows;
// `st_even` / `st_odd`: whether the pair's even / odd row is stored at all.  Every lane of the warp must call this.
static __device__ __forceinline__ void store_rows64_paired(const uint32_t (&p)[16], uint8_t* mine, int pitch, bool st_even, bool st_odd,
                                                           int lane) {
  const bool odd = (lane & 1) != 0;
  uint32_t a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? p[j] : p[8 + j], 1);
    a[j] = odd ? recv : p[j];            // row 2k:   even lane its own first half, odd lane the partner's second half
    b[j] = odd ? p[8 + j] : recv;        // row 2k+1: even lane the partner's first half, odd lane its own second half
  }
  uint8_t* seg_even = mine - (odd ? pitch : 0) + (odd ? 32 : 0);
  uint8_t* seg_odd = mine + (odd ? 0 : pitch) + (odd ? 32 : 0);
  if (st_even)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(seg_even), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
  if (st_odd)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(seg_odd), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]) : "memory");
}
