// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16, M = 128) as a function of N, of the shared-memory
// layout of the A operand (no swizzle 16 B rows / 32 B / 64 B / 128 B rows) and of the number of issuing threads.
// One CTA per SM, operands = zeros, nine tap-like row shifts of the A window as in the conv kernels.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/umma_rate tools/bench_cu/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
struct Args { int reps, N, issuers, fill, side; uint8_t* gbuf; };   // side: 0 none, 1 tcgen05.ld loop, 2 st.global.v8 loop, 3 both
template <int RBA, int RBB>
__global__ void __launch_bounds__(384, 1) rate_kernel(Args a, long long* out) {
  constexpr int KS = RBA >= 64 ? RBA / 32 : 1;          // K slices per row
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot_arr[2];
  uint32_t& tslot = tslot_arr[0];
  if (threadIdx.x == 0) tslot_arr[1] = 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    // fill 1: pseudo-random bf16 pairs in (-2, 2) (sign + 7 random mantissa bits, exponent 126..127): realistic operand toggling
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t lo = 0x3F00u | (h & 0x80FFu), hi = 0x3F00u | ((h >> 16) & 0x80FFu);
    reinterpret_cast<uint32_t*>(smem)[i] = a.fill ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform role index
  if (w < a.issuers) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)a.N >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t la = RBA == 128 ? 2u : RBA == 64 ? 4u : RBA == 32 ? 6u : 0u;
    constexpr uint32_t lb = RBB == 128 ? 2u : RBB == 64 ? 4u : RBB == 32 ? 6u : 0u;
    constexpr uint32_t sbo_a = RBA == 16 ? 128u : 8u * RBA, sbo_b = RBB == 16 ? 128u : 8u * RBB, lbo_b = RBB == 16 ? 1024u : 16u;
    const uint32_t a_base = smem_u32(smem) + (uint32_t)w * 40 * 1024, b_base = smem_u32(smem) + 96 * 1024;
    const uint64_t da0 = desc(a_base, 16u, sbo_a, la), db0 = desc(b_base, lbo_b, sbo_b, lb);
    const uint32_t d = tmem + (uint32_t)w * 256u;
    const long long t0 = clock64();
    for (int r = 0; r < a.reps; r++) {
#pragma unroll
      for (int tap = 0; tap < 9; tap++) {
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          const uint64_t da = da0 + (uint64_t)((tap / 3 * 16 + tap % 3) * RBA / 16 + (RBA >= 64 ? ks * 2 : 0));
          const uint64_t db = db0 + (uint64_t)(RBB >= 64 ? (ks % (RBB / 32)) * 2 : 0);
          if (elect_one())
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        }
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar[w])), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && w == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    if (w == 0) *reinterpret_cast<volatile int*>(&tslot + 1) = 1;
  }
  if (w >= 4 && a.side) {
    // eight "epilogue" warps (two per lane quarter), as in the conv kernels: drain 64 fp32 columns per 128-row tile and / or
    // write 64 bytes per row with two 256-bit stores (row pitch 128 B), at full speed, while the issuers run
    volatile int* flag = reinterpret_cast<volatile int*>(&tslot + 1);
    const uint32_t lane_addr = (uint32_t)((w & 3) * 32) << 16;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; j++) r[j] = threadIdx.x * 33u + j;
    uint8_t* gp = a.gbuf + ((size_t)blockIdx.x * 128 + (w & 3) * 32 + (threadIdx.x & 31)) * 128 + ((w >> 2) & 1) * 64;
    long long n_it = 0;
    while (!*flag) {
      if (a.side & 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tmem + lane_addr + 384u + (uint32_t)((w >> 2) & 1) * 32u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if (a.side & 2) {
        uint8_t* q = gp + (size_t)(n_it & 63) * (148 * 128 * 128);
#pragma unroll
        for (int v = 0; v < 2; v++)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"l"(q + v * 32), "r"(r[8 * v]), "r"(r[8 * v + 1]), "r"(r[8 * v + 2]), "r"(r[8 * v + 3]), "r"(r[8 * v + 4]),
                         "r"(r[8 * v + 5]), "r"(r[8 * v + 6]), "r"(r[8 * v + 7]) : "memory");
      }
      n_it++;
    }
    if (blockIdx.x == 0 && threadIdx.x == 128) out[1] = n_it;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int RBA, int RBB>
void run(long long* d) {
  constexpr int KS = RBA >= 64 ? RBA / 32 : 1;
  cudaFuncSetAttribute(rate_kernel<RBA, RBB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  static uint8_t* gbuf = nullptr;
  if (!gbuf) cudaMalloc(&gbuf, (size_t)64 * 148 * 128 * 128 + 4096);
  for (int N : {64, 128})
    for (int issuers : {2}) for (int side : {0, 1, 2, 3}) {
      const int fill = 1;
      Args a{400, N, issuers, fill, side, gbuf};
      rate_kernel<RBA, RBB><<<148, 384, 200 * 1024>>>(a, d);
      rate_kernel<RBA, RBB><<<148, 384, 200 * 1024>>>(a, d);
      long long cc[2] = {0, 0}; cudaError_t e = cudaMemcpy(cc, d, 16, cudaMemcpyDeviceToHost);
      const long long c = cc[0];
      const double total = (double)a.reps * 9 * KS * issuers;
      printf("A rows %3d B  B rows %3d B  N %3d  issuers %d  side %d (%lld side iterations) : %7.1f cycles per MMA per CTA (floor %d)%s\n", RBA, RBB, N, issuers, side, cc[1],
             (double)c / total, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<128, 128>(d); run<64, 64>(d);
  return 0;
}
