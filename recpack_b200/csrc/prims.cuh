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

// The same scan for long inputs in three steps: per-block inclusive scans + block totals, a scan of the totals, the
// offsets added.  `tot` must hold ceil(n / 1024) + 1 elements.
static __global__ void __launch_bounds__(1024) k_scan3_local(const int* __restrict__ in, int64_t* __restrict__ out, int64_t n,
                                                             int64_t* __restrict__ tot) {
  __shared__ int64_t wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t k = (int64_t)blockIdx.x * 1024 + tid;
  const int64_t v = k < n ? (int64_t)in[k] : 0;
  int64_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int64_t w = wsum[lane];
    int64_t iw = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, iw, o);
      if (lane >= o) iw += t;
    }
    wsum[lane] = iw - w;
  }
  __syncthreads();
  const int64_t mine = wsum[warp] + incl;
  if (k < n) out[k + 1] = mine;
  if (tid == 1023) tot[blockIdx.x] = mine;
  if (k == 0) out[0] = 0;
}
static __global__ void __launch_bounds__(1024) k_scan3_totals(int64_t* __restrict__ tot, int nb) {
  // exclusive scan of the block totals, one block (nb <= a few thousand)
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int k = base + tid;
    const int64_t v = k < nb ? tot[k] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int64_t w = wsum[lane];
      int64_t iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, iw, o);
        if (lane >= o) iw += t;
      }
      wsum[lane] = iw - w;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t mine = carry + wsum[warp] + incl;
    if (k < nb) tot[k] = mine - v;
    __syncthreads();
    if (tid == 1023) carry_s = mine;
    __syncthreads();
  }
}
static __global__ void __launch_bounds__(1024) k_scan3_add(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ tot) {
  const int64_t k = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (blockIdx.x > 0 && k < n) out[k + 1] += tot[blockIdx.x];
}

// out[0] = 0, out[k+1] = in[0] + ... + in[k] on the context's stream: one block for short inputs, three steps otherwise.
// (common.cuh is included before this header by every user.)
static void scan_i32_i64(rpk_ctx* c, const int* in, int64_t* out, int64_t n) {
  if (n <= 8192) {
    k_scan_i32_i64<<<1, 1024, 0, c->stream>>>(in, out, n);
    RPK_LAUNCH_CHECK(c);
    return;
  }
  const int nb = (int)((n + 1023) / 1024);
  int64_t* tot = c->buf<int64_t>("scan_totals", (size_t)nb + 1);
  k_scan3_local<<<nb, 1024, 0, c->stream>>>(in, out, n, tot);
  RPK_LAUNCH_CHECK(c);
  k_scan3_totals<<<1, 1024, 0, c->stream>>>(tot, nb);
  RPK_LAUNCH_CHECK(c);
  k_scan3_add<<<nb, 1024, 0, c->stream>>>(out, n, tot);
  RPK_LAUNCH_CHECK(c);
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
