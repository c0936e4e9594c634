// Small device-wide primitives shared by the translation units (each TU gets its own copy).
#pragma once
#include <stdint.h>

namespace rpk {
typedef unsigned long long u64;

// out[0] = 0, out[k+1] = in[0] + ... + in[k]; one block.
static __global__ void k_scan_i32_i64(const int* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  if (tid == 0) {
    carry_s = 0;
    out[0] = 0;
  }
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    int64_t k = base + tid;
    int64_t v = k < n ? (int64_t)in[k] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = lane < nw ? wsum[lane] : 0;
      int64_t iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, iw, o);
        if (lane >= o) iw += t;
      }
      wsum[lane] = iw - w;
    }
    __syncthreads();
    int64_t carry = carry_s;
    int64_t mine = carry + wsum[warp] + incl;
    if (k < n) out[k + 1] = mine;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = mine;
    __syncthreads();
  }
}

// Heaviest-first row order (longest-processing-time scheduling): bucket rows by log2(work).
static __global__ void k_bucket_count(const u64* __restrict__ work, int64_t begin, int64_t end, int* __restrict__ bcnt) {
  int64_t r = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < end) {
    u64 w = work[r];
    int b = w ? 64 - __clzll((long long)w) : 0;  // 0..64
    atomicAdd(&bcnt[b], 1);
  }
}
static __global__ void k_bucket_offsets(const int* __restrict__ bcnt, int* __restrict__ boff) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int b = 64; b >= 0; --b) {
      boff[b] = acc;
      acc += bcnt[b];
    }
  }
}
static __global__ void k_bucket_scatter(const u64* __restrict__ work, int64_t begin, int64_t end, int* __restrict__ boff,
                                 int* __restrict__ order) {
  int64_t r = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < end) {
    u64 w = work[r];
    int b = w ? 64 - __clzll((long long)w) : 0;
    int pos = atomicAdd(&boff[b], 1);
    order[pos] = (int)r;
  }
}


}  // namespace rpk
