// Microbenchmark: chip-wide global STORE throughput into a buffer that fits L2 (67 MB = act2 of 2048 boards at 15x15),
// for the conv2 epilogue's pattern (each lane two 32-byte stores into its own 128-byte row) and for fully coalesced
// 16-byte-per-lane stores; plus the same with a concurrent read stream of half the size (conv2 reads act1 while writing act2).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/store_rate tools/bench_cu/store_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void store_rows(uint8_t* out, size_t rows, int mode, const uint4* rd, size_t rd_n, uint4* sink) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = threadIdx.x + j;
  if (mode == 0) {            // 32 rows per warp pass, lane = row, 64 bytes of the 128-byte row (two warps per 32 rows: halves)
    for (size_t w = warp; w < rows / 32 * 2; w += nwarps) {
      uint8_t* gp = out + ((w >> 1) * 32 + lane) * 128 + (w & 1) * 64;
#pragma unroll
      for (int q = 0; q < 2; q++)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gp + q * 32), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                     "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
  } else if (mode == 2) {     // lane pairs cover 64 contiguous bytes of a 128-byte row: 16 rows x 64 B per instruction, the other half by the twin warp
    for (size_t w = warp; w < rows / 32 * 2; w += nwarps) {
#pragma unroll
      for (int q = 0; q < 2; q++) {
        uint8_t* gp = out + ((w >> 1) * 32 + q * 16 + (lane >> 1)) * 128 + (w & 1) * 64 + (lane & 1) * 32;
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gp), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                     "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
      }
    }
  } else if (mode == 3) {     // four lanes cover a full 128-byte row: 8 rows x 128 B per instruction (one warp owns whole rows)
    for (size_t w = warp; w < rows / 32; w += nwarps) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint8_t* gp = out + (w * 32 + q * 8 + (lane >> 2)) * 128 + (lane & 3) * 32;
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gp), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                     "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
      }
    }
  } else {                    // coalesced: a warp writes 512 contiguous bytes per instruction
    for (size_t w = warp; w < rows * 128 / 512; w += nwarps)
      *reinterpret_cast<uint4*>(out + w * 512 + lane * 16) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  if (rd) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rd_n; i += (size_t)gridDim.x * blockDim.x) {
      const uint4 x = rd[i]; acc.x ^= x.x; acc.y ^= x.y; acc.z ^= x.z; acc.w ^= x.w;
    }
    if (acc.x == 0x12345678u) *sink = acc;
  }
}
int main() {
  const size_t rows = 2048 * 256;   // x 128 B = 67 MB
  uint8_t* out; uint4* rd; uint4* sink;
  cudaMalloc(&out, rows * 128); cudaMalloc(&rd, rows * 64); cudaMalloc(&sink, 16);
  cudaMemset(out, 0, rows * 128); cudaMemset(rd, 1, rows * 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int with_read = 0; with_read < 2; with_read++)
    for (int mode : {0, 2, 3, 1})
      for (int ctas_per_sm : {2}) {
        for (int it = 0; it < 3; it++) store_rows<<<148 * ctas_per_sm, 512>>>(out, rows, mode, with_read ? rd : nullptr, rows * 64 / 16, sink);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int it = 0; it < reps; it++) store_rows<<<148 * ctas_per_sm, 512>>>(out, rows, mode, with_read ? rd : nullptr, rows * 64 / 16, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / reps;
        printf("%s stores, %d x 512 threads per SM%s: %.1f us per 67 MB -> %.2f TB/s written%s\n", mode == 1 ? "coalesced 16 B/lane" : mode == 2 ? "lane pairs, 64 B per row" : mode == 3 ? "lane quads, 128 B per row" : "row-per-lane 2 x 32 B",
               ctas_per_sm, with_read ? " + 33.5 MB read" : "", us, rows * 128 / us * 1e-6, cudaGetLastError() == cudaSuccess ? "" : " ERR");
      }
  return 0;
}
