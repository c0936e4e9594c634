// Micro-benchmark: throughput of shared-memory atomics on B200 in the access patterns of the scoring kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
// Each variant: grid = 148*CTAS, 512 threads, every thread issues ITERS atomics to a table of R 32-bit slots.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int R = 19684;
constexpr int ITERS = 4096;

__device__ __forceinline__ unsigned lcg(unsigned x) { return x * 1664525u + 1013904223u; }

// mode 0: random slot, result used        1: random slot, result unused
//      2: conflict-free (bank = lane), unused   3: LDS random (no atomic)  4: random, LDS+IADD+STS (non-atomic RMW)
//      5: random slot in 2 distinct... 16-bit packed add (same as 1 but word = slot>>1)
template <int MODE>
__global__ void __launch_bounds__(512, 2) k_bench(unsigned* out, long long* cyc) {
  extern __shared__ unsigned acc[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int s = tid; s < R + 32; s += blockDim.x) acc[s] = 0;
  __syncthreads();
  unsigned x = (blockIdx.x * 977u + tid) * 2654435761u + 12345u;
  unsigned sink = 0;
  // 16 slot numbers per thread, fixed for the whole run (registers): the loop body is the shared-memory access alone
  unsigned js[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    x = lcg(x);
    js[k] = (x >> 8) % R;
    if (MODE == 2) js[k] = ((js[k] >> 5) << 5) + lane;
  }
  const long long t0 = clock64();
#pragma unroll 16
  for (int it = 0; it < ITERS; ++it) {
    unsigned j = js[it & 15];
    if (MODE == 0) {
      const unsigned old = atomicAdd(&acc[j], x | 1u);
      sink = max(sink, old);
    } else if (MODE == 1 || MODE == 2) {
      atomicAdd(&acc[j], x | 1u);
    } else if (MODE == 3) {
      sink += acc[j];
    } else if (MODE == 4) {
      acc[j] = acc[j] + (x | 1u);
    } else if (MODE == 5) {
      atomicAdd(&acc[j >> 1], 1u << ((j & 1) * 16));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  unsigned s = sink;
  for (int k = tid; k < R; k += blockDim.x) s += acc[k];
  if (s == 0x12345678u) out[0] = s;
}

template <int MODE>
static void run(const char* name, int ctas) {
  const int grid = 148 * ctas;
  unsigned* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  const size_t smem = (R + 32) * 4;
  cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_bench<MODE><<<grid, 512, smem>>>(out, cyc);
  cudaEventRecord(e0);
  k_bench<MODE><<<grid, 512, smem>>>(out, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  const double warp_instr_per_sm = (double)ITERS * 16 * ctas;  // 16 warps per CTA
  printf("%-44s ctas/SM=%d  %.3f ms  %.1f cycles/CTA  -> %.2f SM-cycles per warp instruction, %.2f lanes/cycle/SM  err=%s\n", name, ctas, ms, avg,
         avg / warp_instr_per_sm, 32.0 * warp_instr_per_sm / avg, cudaGetErrorString(cudaGetLastError()));
  free(h);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int ctas = 1; ctas <= 2; ++ctas) {
    run<0>("ATOMS.ADD random slot, result used", ctas);
    run<1>("ATOMS.ADD random slot, result unused", ctas);
    run<2>("ATOMS.ADD bank = lane (conflict-free)", ctas);
    run<3>("LDS random slot", ctas);
    run<4>("LDS + IADD + STS random slot (non-atomic)", ctas);
    run<5>("ATOMS.ADD packed 16-bit halves, random", ctas);
  }
  return 0;
}
