// Top-K ranking of arbitrary prediction rows and the listwise metrics on rank-ordered lists.
//
// Replaces (reference, /root/reference):
//   recpack/util.py:50-77            get_top_K_ranks            -> k_topk_csr
//   recpack/metrics/dcg.py:21-128    DCGK / NDCGK._calculate    -> k_metrics
//   recpack/metrics/recall.py:21-85  RecallK / CalibratedRecallK
//   recpack/metrics/base.py:106-123  users without true items are not counted
#include <math.h>

#include "common.cuh"
#include "internal.h"
#include "select.cuh"

namespace rpk {

// Order-preserving map double -> u64 (larger double -> larger integer).
__device__ __forceinline__ u64 ordered_bits(double v) {
  u64 b = (u64)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct CsrRowSrc {
  __device__ __forceinline__ bool has_queue() const { return false; }
  template <class F>
  __device__ __forceinline__ bool for_each_queued(F, int*, int, int*) const { return false; }
  const int* idx;
  const double* val;
  int ns;
  __device__ __forceinline__ int nslots() const { return ns; }
  __device__ __forceinline__ u64 margin() const { return 0ull; }
  __device__ __forceinline__ void set_floor(u64) {}
  __device__ __forceinline__ void stats(SelShared* sh) const { generic_stats(*this, sh); }
  __device__ __forceinline__ u64 key_at(int slot) const {
    double v = val[slot];
    if (v == 0.0) v = 0.0;  // -0.0 and +0.0 are the same score
    return ordered_bits(v);
  }
  // every stored entry is ranked, explicit zeros included (util.py:63-73)
  template <class F>
  __device__ __forceinline__ void visit(F f, int stride) const {
    for (int slot = threadIdx.x * stride; slot < ns; slot += blockDim.x * stride) f(slot, key_at(slot));
  }
  template <class F>
  __device__ __forceinline__ void for_each(F f) const { visit(f, 1); }
  template <class F>
  __device__ __forceinline__ void for_each_sampled(F f) const { visit(f, SEL_SAMPLE); }
  __device__ __forceinline__ void entry(int slot, Entry& e) const {
    e.key = key_at(slot);
    e.idx = idx[slot];
    e.aux = 0;
  }
  __device__ __forceinline__ int cmp3(const Entry& a, const Entry& b) const {
    if (a.key != b.key) return a.key > b.key ? 1 : -1;
    return 0;
  }
};

__global__ void __launch_bounds__(256) k_topk_csr(const int64_t* __restrict__ indptr, const int* __restrict__ indices,
                                                  const double* __restrict__ values, int64_t rows, int K, int cap,
                                                  int direct_cap, int* __restrict__ out_idx, int* __restrict__ out_len) {
  extern __shared__ __align__(16) unsigned char smem[];
  Entry* list = reinterpret_cast<Entry*>(smem);
  int* hist = reinterpret_cast<int*>(smem + sel_list_bytes(cap));
  SelShared* sh = reinterpret_cast<SelShared*>(hist + SEL_BINS);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t b = indptr[r];
    const int n = (int)(indptr[r + 1] - b);
    int m = 0;
    if (n > 0) {
      CsrRowSrc src{indices + b, values + b, n};
      m = block_select_topk(src, K, list, cap, direct_cap, hist, sh);
    }
    for (int t = tid; t < K; t += nt) out_idx[r * K + t] = t < m ? list[t].idx : -1;
    if (tid == 0) out_len[r] = m;
    __syncthreads();
  }
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

void run_topk_csr(rpk_ctx* c, int64_t rows, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u,
                  const double* values_u, int K, int32_t* out_idx_u, int32_t* out_len_u) {
  RPK_REQUIRE(rows >= 0 && nnz >= 0, "negative dimension");
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(out_idx_u && out_len_u, "output pointers must not be null");
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)rows + 1, "t_indptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "t_indices");
  const double* values = stage_in(c, values_u, (size_t)nnz, "t_values");
  Out<int32_t> o_idx, o_len;
  o_idx.init(c, out_idx_u, (size_t)rows * K, "t_out_idx");
  o_len.init(c, out_len_u, (size_t)rows, "t_out_len");
  if (rows > 0) {
    const bool tiny = c->flags & DBG_TINY_LIST;
    const int cap = std::max(tiny ? 64 : 1024, next_pow2(2 * K));
    const int direct_cap = tiny ? K : cap;
    const size_t smem = sel_smem_bytes(cap);
    RPK_CUDA(cudaFuncSetAttribute(k_topk_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(rows, (int64_t)c->sm_count * 8);
    k_topk_csr<<<grid, 256, smem, st>>>(indptr, indices, values, rows, K, cap, direct_cap, o_idx.dev, o_len.dev);
    RPK_LAUNCH_CHECK(c);
  }
  o_idx.finish(c);
  o_len.finish(c);
  finish_call(c);
}

// One thread per user; lists are short (N <= a few hundred), true rows are sorted -> binary search.
__global__ void k_metrics(const int* __restrict__ top_idx, const int* __restrict__ top_len, int64_t U, int N,
                          const int64_t* __restrict__ tptr, const int* __restrict__ tidx, int n_metrics,
                          const int* __restrict__ kinds, const int* __restrict__ Ks, const double* __restrict__ discount,
                          const double* __restrict__ idcg, double* __restrict__ per_user) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const int64_t tb = tptr[u];
  const int nt = (int)(tptr[u + 1] - tb);
  if (nt == 0) {
    for (int m = 0; m < n_metrics; ++m) per_user[(int64_t)m * U + u] = nan("");
    return;
  }
  int len = top_len[u];
  if (len > N) len = N;
  for (int m = 0; m < n_metrics; ++m) {
    const int K = Ks[m];
    const int kind = kinds[m];
    const int lim = len < K ? len : K;
    double dcg = 0.0;
    int hits = 0, first_hit = 0;  // first_hit: 1-based rank of the best-ranked hit
    for (int r = 0; r < lim; ++r) {
      const int item = top_idx[u * N + r];
      int lo = 0, hi = nt;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (tidx[tb + mid] < item) lo = mid + 1;
        else hi = mid;
      }
      if (lo < nt && tidx[tb + lo] == item) {
        if (!hits) first_hit = r + 1;
        hits++;
        dcg = __dadd_rn(dcg, discount[r]);
      }
    }
    double v;
    if (kind == RPK_METRIC_NDCG) v = __ddiv_rn(dcg, idcg[nt < K ? nt : K]);
    else if (kind == RPK_METRIC_DCG) v = dcg;
    else if (kind == RPK_METRIC_RECALL) v = __ddiv_rn((double)hits, (double)nt);
    else if (kind == RPK_METRIC_PRECISION) v = __ddiv_rn((double)hits, (double)K);
    else if (kind == RPK_METRIC_RECIPROCAL_RANK) v = first_hit ? __ddiv_rn(1.0, (double)first_hit) : 0.0;
    else if (kind == RPK_METRIC_HITS) v = (double)hits;
    else v = __ddiv_rn((double)hits, (double)(nt < K ? nt : K));
    per_user[(int64_t)m * U + u] = v;
  }
}

// Deterministic reduction: one block per metric, fixed assignment of users to threads, fixed tree.
__global__ void __launch_bounds__(1024) k_metric_sums(const double* __restrict__ per_user, int64_t U, double* __restrict__ sums,
                                                      long long* __restrict__ n_users) {
  __shared__ double s_sum[1024];
  __shared__ long long s_cnt[1024];
  const int m = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  long long cnt = 0;
  for (int64_t u = tid; u < U; u += blockDim.x) {
    double v = per_user[(int64_t)m * U + u];
    if (v == v) {
      acc += v;
      cnt++;
    }
  }
  s_sum[tid] = acc;
  s_cnt[tid] = cnt;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (tid < o) {
      s_sum[tid] += s_sum[tid + o];
      s_cnt[tid] += s_cnt[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    sums[m] = s_sum[0];
    if (m == 0) *n_users = s_cnt[0];
  }
}

void run_metrics_topn(rpk_ctx* c, int64_t U, int N, const int32_t* top_idx_u, const int32_t* top_len_u,
                      const int64_t* true_indptr_u, const int32_t* true_indices_u, int64_t true_nnz, int n_metrics,
                      const int32_t* kinds_u, const int32_t* Ks_u, const double* discount_u, const double* idcg_u, int maxK,
                      double* per_user_u, double* sums_u, int64_t* n_users_u) {
  RPK_REQUIRE(U >= 0 && N >= 1 && n_metrics >= 1 && n_metrics <= 64 && maxK >= 1, "bad metric arguments");
  RPK_REQUIRE(sums_u && n_users_u, "sums / n_users must not be null");
  for (int m = 0; m < n_metrics; ++m) {
    // kinds / Ks are tiny and always host-side in practice; validate when they are
    if (!is_device_ptr(kinds_u)) RPK_REQUIRE(kinds_u[m] >= 0 && kinds_u[m] <= RPK_METRIC_HITS, "unknown metric kind");
    if (!is_device_ptr(Ks_u)) RPK_REQUIRE(Ks_u[m] >= 1 && Ks_u[m] <= maxK, "metric K exceeds the discount table");
  }
  cudaStream_t st = c->stream;
  const int32_t* top_idx = stage_in(c, top_idx_u, (size_t)U * N, "x_top_idx");
  const int32_t* top_len = stage_in(c, top_len_u, (size_t)U, "x_top_len");
  const int64_t* tptr = stage_in(c, true_indptr_u, (size_t)U + 1, "x_tptr");
  const int32_t* tidx = stage_in(c, true_indices_u, (size_t)true_nnz, "x_tidx");
  const int32_t* kinds = stage_in(c, kinds_u, (size_t)n_metrics, "x_kinds");
  const int32_t* Ks = stage_in(c, Ks_u, (size_t)n_metrics, "x_Ks");
  const double* discount = stage_in(c, discount_u, (size_t)maxK, "x_disc");
  const double* idcg = stage_in(c, idcg_u, (size_t)maxK + 1, "x_idcg");
  Out<double> o_pu, o_sums;
  Out<int64_t> o_n;
  double* pu_dev;
  if (per_user_u) {
    o_pu.init(c, per_user_u, (size_t)n_metrics * U, "x_per_user");
    pu_dev = o_pu.dev;
  } else {
    pu_dev = c->buf<double>("x_per_user", (size_t)n_metrics * U);
  }
  o_sums.init(c, sums_u, (size_t)n_metrics, "x_sums");
  o_n.init(c, n_users_u, 1, "x_nusers");
  if (U > 0) {
    k_metrics<<<ceil_div(U, 128), 128, 0, st>>>(top_idx, top_len, U, N, tptr, tidx, n_metrics, kinds, Ks, discount, idcg, pu_dev);
    RPK_LAUNCH_CHECK(c);
  }
  k_metric_sums<<<n_metrics, 1024, 0, st>>>(pu_dev, U, o_sums.dev, reinterpret_cast<long long*>(o_n.dev));
  RPK_LAUNCH_CHECK(c);
  c->mark("metrics");
  o_pu.finish(c);
  o_sums.finish(c);
  o_n.finish(c);
  finish_call(c);
}

// ---- CoverageK (recpack/metrics/coverage.py:13-40): items that appear among the first K places of the list of
// any user with a non-empty y_true row (metrics/base.py:106-123 drops the other users first).
__global__ void k_coverage_mark(const int* __restrict__ top_idx, const int* __restrict__ top_len, int64_t U, int N, int K,
                                const int64_t* __restrict__ tptr, unsigned char* __restrict__ flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= U * K) return;
  const int64_t u = t / K;
  const int r = (int)(t % K);
  if (tptr[u + 1] == tptr[u]) return;
  int len = top_len[u];
  if (len > N) len = N;
  if (r < len) flags[top_idx[u * N + r]] = 1;
}

__global__ void __launch_bounds__(1024) k_coverage_count(const unsigned char* __restrict__ flags, int64_t I,
                                                         long long* __restrict__ out) {
  __shared__ long long s_cnt[32];
  long long c = 0;
  for (int64_t j = threadIdx.x; j < I; j += blockDim.x) c += flags[j] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_cnt[w];
    *out = tot;
  }
}

void run_coverage_topn(rpk_ctx* c, int64_t U, int N, int K, int64_t I, const int32_t* top_idx_u, const int32_t* top_len_u,
                       const int64_t* true_indptr_u, int64_t* out_count_u, uint8_t* out_flags_u) {
  RPK_REQUIRE(U >= 0 && N >= 1 && K >= 1 && K <= N && I >= 0, "bad coverage arguments");
  RPK_REQUIRE(out_count_u, "out_count must not be null");
  cudaStream_t st = c->stream;
  const int32_t* top_idx = stage_in(c, top_idx_u, (size_t)U * N, "x_top_idx");
  const int32_t* top_len = stage_in(c, top_len_u, (size_t)U, "x_top_len");
  const int64_t* tptr = stage_in(c, true_indptr_u, (size_t)U + 1, "x_tptr");
  Out<int64_t> o_n;
  Out<uint8_t> o_f;
  o_n.init(c, out_count_u, 1, "x_cov_n");
  unsigned char* flags;
  if (out_flags_u) {
    o_f.init(c, out_flags_u, (size_t)I, "x_cov_flags");
    flags = o_f.dev;
  } else {
    flags = c->buf<unsigned char>("x_cov_flags", (size_t)I + 1);
  }
  RPK_CUDA(cudaMemsetAsync(flags, 0, (size_t)I, st));
  if (U > 0) {
    k_coverage_mark<<<ceil_div(U * K, 256), 256, 0, st>>>(top_idx, top_len, U, N, K, tptr, flags);
    RPK_LAUNCH_CHECK(c);
  }
  k_coverage_count<<<1, 1024, 0, st>>>(flags, I, reinterpret_cast<long long*>(o_n.dev));
  RPK_LAUNCH_CHECK(c);
  o_n.finish(c);
  o_f.finish(c);
  finish_call(c);
}

}  // namespace rpk
