// EASE on the GPU (recpack/algorithms/ease.py:63-95): the dense co-occurrence Gram on the tensor cores, the closed-form
// item-item model from the inverse, and scoring with a dense, signed model.
//
//   XTX = (X.T @ X).toarray()                         ease.py:79    -> rpk_gram_dense_f64: exact integer counts on tcgen05
//   P = inv(XTX + l2 * I)                              ease.py:80    -> the caller (cuSOLVER potrf / potri through torch)
//   B = I - P @ diag(1 / diag(P)); diag(B) = 0         ease.py:83-84 -> rpk_ease_from_inverse: B_ij = -P_ij / P_jj, B_ii = 0
//   B = B @ diag(1 / n_j^alpha)                        ease.py:86-88 -> the same kernel, column scale w_j
//   scores = X @ B                                     base.py:248   -> rpk_predict_dense_*: sum over the history in
//                                                                       ascending item order, float64, one add per term --
//                                                                       scipy's csr_matvecs order, so equal inputs give
//                                                                       bit-identical scores
#include "common.cuh"
#include "internal.h"
#include "select.cuh"

namespace rpk {

// A[item][u - u0] = 1 for the interactions of users [u0, u1) (one warp per user).
__global__ void k_fill_dense_chunk(const int64_t* __restrict__ indptr, const int* __restrict__ indices, int64_t u0, int64_t u1,
                                   int64_t kd_pad, unsigned char* __restrict__ A) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = u0 + warp; u < u1; u += nwarps)
    for (int64_t k = indptr[u] + lane; k < indptr[u + 1]; k += 32) A[(int64_t)indices[k] * kd_pad + (u - u0)] = 1;
}

// G[i][j] (+)= G16[i][j]
__global__ void k_acc_g16(const unsigned short* __restrict__ g16, int64_t ldg, int64_t I, int first, double* __restrict__ G) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= I * I) return;
  const int64_t i = t / I, j = t - i * I;
  const double v = (double)g16[i * ldg + j];
  G[t] = first ? v : G[t] + v;
}

void run_gram_dense_f64(rpk_ctx* c, int64_t U, int64_t I, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u,
                        double* out_G_u) {
  RPK_REQUIRE(U >= 0 && I >= 1 && nnz >= 0 && out_G_u, "bad arguments");
  RPK_REQUIRE(I < ((int64_t)1 << 24), "item count must be below 2^24");
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)U + 1, "fit_indptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "fit_indices");
  Out<double> o;
  o.init(c, out_G_u, (size_t)I * I, "ease_G");
  const int64_t rows_pad = (I + 255) / 256 * 256;
  const int64_t chunk = 32768;  // users per pass: their counts fit the kernel's 16-bit output
  const int64_t kd_max = std::min<int64_t>(chunk, (U + 127) / 128 * 128);
  unsigned char* A = c->buf<unsigned char>("ease_A", (size_t)rows_pad * std::max<int64_t>(kd_max, 128));
  unsigned short* G16 = c->buf<unsigned short>("ease_G16", (size_t)I * rows_pad);
  const int gblocks = (int)std::min<int64_t>(ceil_div(I * I, 256), 1 << 30);
  if (U == 0) RPK_CUDA(cudaMemsetAsync(o.dev, 0, sizeof(double) * (size_t)I * I, st));
  for (int64_t u0 = 0; u0 < U; u0 += chunk) {
    const int64_t u1 = std::min(U, u0 + chunk);
    const int64_t kd_pad = (u1 - u0 + 127) / 128 * 128;
    RPK_CUDA(cudaMemsetAsync(A, 0, (size_t)rows_pad * kd_pad, st));
    const int wblocks = (int)std::min<int64_t>(((u1 - u0) * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    k_fill_dense_chunk<<<wblocks, 256, 0, st>>>(indptr, indices, u0, u1, kd_pad, A);
    RPK_LAUNCH_CHECK(c);
    run_gram_dense_tc(c, A, rows_pad, kd_pad, 0, I, G16, rows_pad);
    k_acc_g16<<<gblocks, 256, 0, st>>>(G16, rows_pad, I, u0 == 0 ? 1 : 0, o.dev);
    RPK_LAUNCH_CHECK(c);
  }
  o.finish(c);
  finish_call(c);
}

// B_ij = -P_ij / P_jj * w_j, B_ii = 0  (in place when B == P).  The division by the column's diagonal entry is the
// closed form of P @ diag(1 / diag(P)) (ease.py:83); the products are formed in the reference's order:
// fl(fl(P_ij * fl(1 / P_jj)) ...) -> negated -> * w_j.
__global__ void k_ease_finish(const double* P, const double* __restrict__ diag, const double* __restrict__ w, int64_t I, double* B) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= I * I) return;
  const int64_t i = t / I, j = t - i * I;
  double v = 0.0;
  if (i != j) {
    const double inv = __ddiv_rn(1.0, diag[j]);       // np.diag(1.0 / np.diag(P))
    v = -__dmul_rn(P[t], inv);                         // I - P @ diag(.), off the diagonal
    if (w) v = __dmul_rn(v, w[j]);                     // B @ diag(w)
  }
  B[t] = v;
}

void run_ease_from_inverse(rpk_ctx* c, int64_t I, const double* P_u, const double* w_u, double* B_u) {
  RPK_REQUIRE(I >= 1 && P_u && B_u, "bad arguments");
  RPK_REQUIRE(is_device_ptr(P_u) && is_device_ptr(B_u), "rpk_ease_from_inverse works on device matrices");
  const double* w = w_u ? stage_in(c, w_u, (size_t)I, "ease_w") : nullptr;
  // in place is fine except for the diagonal entries, which every column reads and the kernel zeroes: a copy first
  double* dg = c->buf<double>("ease_diag", (size_t)I);
  RPK_CUDA(cudaMemcpy2DAsync(dg, sizeof(double), P_u, sizeof(double) * ((size_t)I + 1), sizeof(double), (size_t)I,
                             cudaMemcpyDeviceToDevice, c->stream));
  k_ease_finish<<<ceil_div(I * I, 256), 256, 0, c->stream>>>(P_u, dg, w, I, B_u);
  RPK_LAUNCH_CHECK(c);
  finish_call(c);
}

// ------------------------------------------------------------------------------------------
// Dense scoring
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 ordered_bits_f64(double v) {
  u64 b = (u64)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(u64 k) {
  const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// Candidates of one (user, column range): the columns with a non-zero score (what csr_matrix(scores) stores).
struct DenseScoreSrc {
  const double* acc;
  int r0, ns;
  __device__ __forceinline__ bool has_queue() const { return false; }
  template <class F>
  __device__ __forceinline__ bool for_each_queued(F, int*, int, int*) const { return false; }
  __device__ __forceinline__ int nslots() const { return ns; }
  __device__ __forceinline__ u64 margin() const { return 0ull; }
  __device__ __forceinline__ void set_floor(u64) {}
  __device__ __forceinline__ void stats(SelShared* sh) const { generic_stats(*this, sh); }
  template <class F>
  __device__ __forceinline__ void visit(F f, int stride) const {
    for (int slot = threadIdx.x * stride; slot < ns; slot += blockDim.x * stride) {
      const double v = acc[slot];
      if (v != 0.0) f(slot, ordered_bits_f64(v));
    }
  }
  template <class F>
  __device__ __forceinline__ void for_each(F f) const { visit(f, 1); }
  template <class F>
  __device__ __forceinline__ void for_each_sampled(F f) const { visit(f, SEL_SAMPLE); }
  __device__ __forceinline__ void entry(int slot, Entry& e) const {
    e.key = ordered_bits_f64(acc[slot]);
    e.idx = r0 + slot;
    e.aux = 0;
  }
  __device__ __forceinline__ int cmp3(const Entry& a, const Entry& b) const {
    if (a.key != b.key) return a.key > b.key ? 1 : -1;
    return 0;
  }
};

struct DenseParams {
  const int64_t* indptr;
  const int* indices;
  const double* B;  // [I x I] row-major
  int64_t I;
  int U, P, R, N, mask, cap, direct_cap;
  int* part_idx;    // [U*P x N]
  u64* part_key;    // ordered bit patterns of the scores
  int* part_len;
  double* full;     // non-null: write every score [U x I] instead of lists
};

// One CTA per (user, column range): the scores of the range are accumulated in shared memory, a thread owns its
// columns and walks the history rows in ascending item order (no atomics: the order of the additions is scipy's).
__global__ void __launch_bounds__(1024, 1) k_predict_dense(DenseParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Entry* list = reinterpret_cast<Entry*>(smem);
  int* hist = reinterpret_cast<int*>(smem + sel_list_bytes(p.cap));
  SelShared* sh = reinterpret_cast<SelShared*>(hist + SEL_BINS);
  double* acc = reinterpret_cast<double*>(smem + sel_smem_bytes(p.cap));
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t total = (int64_t)p.U * p.P;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int u = (int)(w / p.P), pass = (int)(w % p.P);
    const int r0 = pass * p.R;
    const int ns = (int)min((int64_t)p.R, p.I - r0);
    const int64_t xb = p.indptr[u];
    const int d = (int)(p.indptr[u + 1] - xb);
    for (int s = tid; s < ns; s += nt) acc[s] = 0.0;
    // rows in ascending item order (the CSR is canonical); every thread adds the same rows in the same order
    for (int r = 0; r < d; ++r) {
      const double* row = p.B + (int64_t)p.indices[xb + r] * p.I + r0;
      for (int s = tid; s < ns; s += nt) acc[s] = __dadd_rn(acc[s], __ldg(row + s));
    }
    __syncthreads();
    if (p.mask) {
      for (int r = tid; r < d; r += nt) {
        const int j = p.indices[xb + r] - r0;
        if (j >= 0 && j < ns) acc[j] = 0.0;
      }
      __syncthreads();
    }
    if (p.full) {
      for (int s = tid; s < ns; s += nt) p.full[(int64_t)u * p.I + r0 + s] = acc[s];
      __syncthreads();
      continue;
    }
    DenseScoreSrc src{acc, r0, ns};
    const int m = block_select_topk(src, p.N, list, p.cap, p.direct_cap, hist, sh);
    const int64_t slot_out = (int64_t)u * p.P + pass;
    for (int t = tid; t < p.N; t += nt) {
      p.part_idx[slot_out * p.N + t] = t < m ? list[t].idx : -1;
      p.part_key[slot_out * p.N + t] = t < m ? list[t].key : 0ull;
    }
    if (tid == 0) p.part_len[slot_out] = m;
    __syncthreads();
  }
}

// One warp per user: merge the P per-range lists (keys = ordered bit patterns of float64 scores).
__global__ void k_dense_finalize(const int* __restrict__ part_idx, const u64* __restrict__ part_key, const int* __restrict__ part_len,
                                 int64_t U, int P, int N, int* __restrict__ out_idx, double* __restrict__ out_val,
                                 int* __restrict__ out_len) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = warp; u < U; u += nwarps) {
    const int PN = P * N;
    const int* pi = part_idx + u * PN;
    const u64* ps = part_key + u * PN;
    int tot = 0;
    for (int q = 0; q < P; ++q) tot += part_len[u * P + q];
    const int m = min(N, tot);
    for (int e = lane; e < PN; e += 32) {
      const int je = pi[e];
      if (je < 0) continue;
      const u64 se = ps[e];
      int rank = 0;
      for (int f = 0; f < PN; ++f) {
        const int jf = pi[f];
        if (jf < 0) continue;
        const u64 sf = ps[f];
        rank += (sf > se) || (sf == se && jf < je);
      }
      if (rank < N) {
        out_idx[u * N + rank] = je;
        if (out_val) out_val[u * N + rank] = from_ordered_bits(se);
      }
    }
    for (int t = m + lane; t < N; t += 32) {
      out_idx[u * N + t] = -1;
      if (out_val) out_val[u * N + t] = 0.0;
    }
    if (lane == 0) out_len[u] = m;
  }
}

static int next_pow2_(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

void run_predict_dense(rpk_ctx* c, int64_t U, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u, int64_t I,
                       const double* B, int N, int mask_history, int32_t* out_idx_u, double* out_val_u, int32_t* out_len_u,
                       double* out_full_u) {
  RPK_REQUIRE(U >= 0 && nnz >= 0 && I >= 1 && B, "bad arguments");
  RPK_REQUIRE(is_device_ptr(B), "the dense model must live on the device");
  RPK_REQUIRE(U < ((int64_t)1 << 31), "too many users in one call");
  const bool full = out_full_u != nullptr;
  if (!full) {
    RPK_REQUIRE(N >= 1 && N <= 2048, "N must be in [1, 2048]");
    RPK_REQUIRE(out_idx_u && out_len_u, "out_idx / out_len must not be null");
  }
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)U + 1, "p_indptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "p_indices");
  Out<int32_t> o_idx, o_len;
  Out<double> o_val, o_full;
  if (full) {
    o_full.init(c, out_full_u, (size_t)U * I, "d_full");
  } else {
    o_idx.init(c, out_idx_u, (size_t)U * N, "p_out_idx");
    o_val.init(c, out_val_u, (size_t)U * N, "p_out_val");
    o_len.init(c, out_len_u, (size_t)U, "p_out_len");
  }
  if (U > 0) {
    DenseParams p;
    const int Nn = full ? 1 : N;
    p.cap = std::max(256, next_pow2_(2 * Nn));
    p.direct_cap = std::min(p.cap, std::max(64, 2 * Nn));
    const size_t fixed = sel_smem_bytes(p.cap);
    RPK_REQUIRE((size_t)c->smem_max > fixed + 1024 + 8192, "N too large for shared memory");
    const size_t avail = (size_t)c->smem_max - fixed - 1024;
    int P = 1;
    int64_t R = 0;
    for (;; ++P) {
      R = ((I + P - 1) / P + 1) & ~(int64_t)1;
      if ((size_t)R * 8 <= avail) break;
    }
    if ((c->flags & DBG_MULTI_PASS) && P < 2 && I >= 8) {
      P = 2;
      R = ((I + P - 1) / P + 1) & ~(int64_t)1;
    }
    p.indptr = indptr;
    p.indices = indices;
    p.B = B;
    p.I = I;
    p.U = (int)U;
    p.P = P;
    p.R = (int)R;
    p.N = Nn;
    p.mask = mask_history;
    p.full = full ? o_full.dev : nullptr;
    p.part_idx = full ? nullptr : c->buf<int>("d_part_idx", (size_t)U * P * N);
    p.part_key = full ? nullptr : c->buf<u64>("d_part_key", (size_t)U * P * N);
    p.part_len = full ? nullptr : c->buf<int>("d_part_len", (size_t)U * P);
    const size_t smem = fixed + (size_t)R * 8;
    RPK_CUDA(cudaFuncSetAttribute(k_predict_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nt = R >= 8192 ? 1024 : (R >= 2048 ? 512 : 256);
    const int grid = (int)std::min<int64_t>(U * P, (int64_t)c->sm_count * 4);
    c->ev_record(4);
    k_predict_dense<<<grid, nt, smem, st>>>(p);
    RPK_LAUNCH_CHECK(c);
    c->ev_record(5);
    c->ev_valid[2] = true;
    if (!full) {
      const int fgrid = (int)std::min<int64_t>((U * 32 + 255) / 256, (int64_t)c->sm_count * 16);
      k_dense_finalize<<<fgrid, 256, 0, st>>>(p.part_idx, p.part_key, p.part_len, U, P, N, o_idx.dev, o_val.dev, o_len.dev);
      RPK_LAUNCH_CHECK(c);
    }
  }
  o_idx.finish(c);
  o_val.finish(c);
  o_len.finish(c);
  o_full.finish(c);
  finish_call(c);
}

}  // namespace rpk
