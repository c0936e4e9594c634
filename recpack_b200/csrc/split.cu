// Per-user random split of the interactions (the data side of the pipeline).
//
// Replaces (reference, /root/reference):
//   recpack/scenarios/splitters.py:233-263   FractionInteractionSplitter.split: for every user u, shuffle the user's
//       interaction ids with np.random.RandomState(seed + u) and send the first ceil(n * in_frac) to data_in.
// The arithmetic lives in numpy (not vendored): RandomState(int) = MT19937 seeded by init_genrand; shuffle of a 1-d
// array = Fisher-Yates from the back, j = random_interval(i) = 32-bit draws masked to the next power of two minus
// one, rejected while > i (numpy/random/mtrand.pyx _shuffle_raw, src/distributions/distributions.c random_interval,
// src/mt19937/mt19937.c).  Reproduced bit for bit: one thread per user, its 624-word generator state in shared memory
// (word k of thread t at k * blockDim + t: conflict-free), its permutation in a global scratch segment.
#include "common.cuh"
#include "internal.h"

namespace rpk {

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr int SPLIT_NT = 64;  // 64 threads x 624 words x 4 B = 156 KB of shared memory

struct Mt {
  unsigned* s;  // this thread's word 0; stride SPLIT_NT
  int pos;
  __device__ __forceinline__ unsigned& at(int k) { return s[k * SPLIT_NT]; }
  __device__ void seed(unsigned v) {
    at(0) = v;
    for (int k = 1; k < MT_N; ++k) {
      v = 1812433253u * (v ^ (v >> 30)) + (unsigned)k;
      at(k) = v;
    }
    pos = MT_N;
  }
  __device__ void twist() {
    int k = 0;
    for (; k < MT_N - MT_M; ++k) {
      const unsigned y = (at(k) & 0x80000000u) | (at(k + 1) & 0x7fffffffu);
      at(k) = at(k + MT_M) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; k < MT_N - 1; ++k) {
      const unsigned y = (at(k) & 0x80000000u) | (at(k + 1) & 0x7fffffffu);
      at(k) = at(k + (MT_M - MT_N)) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    const unsigned y = (at(MT_N - 1) & 0x80000000u) | (at(0) & 0x7fffffffu);
    at(MT_N - 1) = at(MT_M - 1) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    pos = 0;
  }
  __device__ __forceinline__ unsigned next() {
    if (pos == MT_N) twist();
    unsigned y = at(pos++);
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

__global__ void __launch_bounds__(SPLIT_NT) k_split_fraction(int64_t n_users, const int64_t* __restrict__ uids,
                                                             const int64_t* __restrict__ seg, const int64_t* __restrict__ rows,
                                                             double in_frac, unsigned long long seed, int* __restrict__ perm,
                                                             unsigned char* __restrict__ in_mask) {
  extern __shared__ unsigned mt_words[];
  const int64_t g = (int64_t)blockIdx.x * SPLIT_NT + threadIdx.x;
  if (g >= n_users) return;
  Mt mt;
  mt.s = mt_words + threadIdx.x;
  mt.seed((unsigned)(seed + (unsigned long long)uids[g]));
  const int64_t b = seg[g];
  const int n = (int)(seg[g + 1] - b);
  int* p = perm + b;
  for (int t = 0; t < n; ++t) p[t] = t;
  for (int i = n - 1; i >= 1; --i) {
    unsigned mask = (unsigned)i;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    unsigned j;
    do {
      j = mt.next() & mask;
    } while (j > (unsigned)i);
    const int a = p[i];
    p[i] = p[j];
    p[j] = a;
  }
  const int cut = (int)ceil(__dmul_rn((double)n, in_frac));
  for (int t = 0; t < n; ++t) in_mask[rows[b + p[t]]] = t < cut ? 1 : 0;
}

}  // namespace

void run_split_fraction(rpk_ctx* c, int64_t n_users, const int64_t* uids_u, const int64_t* seg_u, const int64_t* rows_u,
                        int64_t n_rows, double in_frac, uint64_t seed, uint8_t* out_in_mask_u) {
  RPK_REQUIRE(n_users >= 0 && n_rows >= 0, "negative dimension");
  RPK_REQUIRE(in_frac >= 0.0 && in_frac <= 1.0, "in_frac must be in [0, 1]");
  RPK_REQUIRE(n_rows < ((int64_t)1 << 31), "more than 2^31 interactions are not supported");
  RPK_REQUIRE(out_in_mask_u, "out_in_mask must not be null");
  cudaStream_t st = c->stream;
  const int64_t* uids = stage_in(c, uids_u, (size_t)n_users, "sp_uids");
  const int64_t* seg = stage_in(c, seg_u, (size_t)n_users + 1, "sp_seg");
  const int64_t* rows = stage_in(c, rows_u, (size_t)n_rows, "sp_rows");
  Out<uint8_t> o;
  o.init(c, out_in_mask_u, (size_t)n_rows, "sp_mask");
  if (n_rows > 0) RPK_CUDA(cudaMemsetAsync(o.dev, 0, (size_t)n_rows, st));
  if (n_users > 0 && n_rows > 0) {
    int* perm = c->buf<int>("sp_perm", (size_t)n_rows);
    const size_t smem = (size_t)MT_N * SPLIT_NT * sizeof(unsigned);
    RPK_CUDA(cudaFuncSetAttribute(k_split_fraction, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_split_fraction<<<ceil_div(n_users, SPLIT_NT), SPLIT_NT, smem, st>>>(n_users, uids, seg, rows, in_frac,
                                                                            (unsigned long long)seed, perm, o.dev);
    RPK_LAUNCH_CHECK(c);
  }
  o.finish(c);
  finish_call(c);
}

}  // namespace rpk
