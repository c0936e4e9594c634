// Scoring on the GPU: score_uj = sum_{i in hist(u)} S_ij for a K-sparse S, history masking and
// top-N fused in shared memory; optional full CSR output for drop-in predict().
//
// Replaces (reference, /root/reference):
//   recpack/algorithms/base.py:237-255     ItemSimilarityMatrixAlgorithm._predict  (X @ similarity_matrix_)
//   recpack/pipelines/pipeline.py:174-175  history removal
//   recpack/metrics/base.py:189            get_top_K_ranks(y_pred, K) on the prediction rows
//
// Scores are exact integers: every similarity value is stored as q = rint(v * 2^39) | 1 (40 bits), a
// score is the integer sum of q over the history.  Integer sums make the result independent of the order
// in which the atomics land, so the top-N lists are deterministic.  Two kernels:
//   k_predict_a32  top-N: 32-bit approximate sums with one native shared-memory atomic (ATOMS.ADD) per
//                  entry pick the few items that can be in the top N, a second sweep adds their exact q;
//   k_predict      full CSR output, and the exact path for users k_predict_a32 hands over: the sum is kept
//                  in two 32-bit limbs (q & 0xFFFFF, q >> 20), 64-bit compare-and-swap accumulators when a
//                  limb could overflow.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "internal.h"
#include "prims.cuh"
#include "select.cuh"

namespace rpk {

constexpr u64 Q_MASK40 = (((u64)1) << 40) - 1;
constexpr int LIMB_BITS = 20;
constexpr unsigned LIMB_MASK = (1u << LIMB_BITS) - 1;
constexpr int LIMB_CHUNK = 4095;  // rows that can be added before the low limb must be normalised

// ------------------------------------------------------------------------------------------
// Model construction
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool quantize(double v, u64& q) {
  if (!(v >= 0.0) || !(v < 2.0)) return false;
  long long r = __double2ll_rn(v * 549755813888.0);  // 2^39, exact scaling; round half to even like np.rint
  q = ((u64)r) | 1ull;
  return q <= Q_MASK40;
}

// One CTA per row: pack (idx, q), sort by idx, write the row.
__global__ void __launch_bounds__(256) k_model_from_topk(const int* __restrict__ idx, const double* __restrict__ val,
                                                         const int* __restrict__ len, const int64_t* __restrict__ row_src,
                                                         int K, int I, int nrows,
                                                         const int64_t* __restrict__ m_ptr, u64* __restrict__ m_ent,
                                                         unsigned* __restrict__ m_rowmax, int* __restrict__ flag) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64* buf = reinterpret_cast<u64*>(smem);
  __shared__ unsigned s_max;
  const int tid = threadIdx.x, nt = blockDim.x;
  int n2 = 2;
  while (n2 < K) n2 <<= 1;
  for (int i = blockIdx.x; i < nrows; i += gridDim.x) {
    const int64_t src = row_src ? row_src[i] : (int64_t)i;  // where row i lives in the (gathered) input
    int m = len[src];
    if (m > K) m = K;
    if (m < 0) m = 0;
    if (tid == 0) s_max = 0;
    __syncthreads();
    unsigned lmax = 0;
    for (int t = tid; t < n2; t += nt) {
      u64 packed = ~0ull;
      if (t < m) {
        int j = idx[src * K + t];
        u64 q = 1;
        bool ok = quantize(val[src * K + t], q);
        if (!ok || j < 0 || j >= I) atomicOr(flag, 1);
        packed = ((u64)(unsigned)j << 40) | (q & Q_MASK40);
        lmax = max(lmax, (unsigned)(q >> LIMB_BITS) + 1u);
      }
      buf[t] = packed;
    }
    atomicMax(&s_max, lmax);
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < n2; t += nt) {
          int x = t ^ j;
          if (x > t) {
            u64 a = buf[t], b = buf[x];
            bool up = (t & k) == 0;
            if ((a > b) == up) {
              buf[t] = b;
              buf[x] = a;
            }
          }
        }
        __syncthreads();
      }
    if (m_ptr) {
      const int64_t base = m_ptr[i];
      for (int t = tid; t < m; t += nt) m_ent[base + t] = buf[t];
      if (tid == 0) m_rowmax[i] = s_max;
    } else {  // packed rows of K places each, all-ones after the row's entries (they sort last)
      for (int t = tid; t < K; t += nt) m_ent[(int64_t)i * K + t] = buf[t];
    }
    for (int t = tid + 1; t < m; t += nt)
      if ((buf[t] >> 40) == (buf[t - 1] >> 40)) atomicOr(flag, 2);  // duplicate column
    __syncthreads();
  }
}

// One warp per row of a CSR with ascending unique columns.
__global__ void k_model_from_csr(const int64_t* __restrict__ indptr, const int* __restrict__ indices,
                                 const double* __restrict__ values, int64_t I, u64* __restrict__ m_ent,
                                 unsigned* __restrict__ m_rowmax, int* __restrict__ m_len, int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < I; i += nwarps) {
    int64_t b = indptr[i], e = indptr[i + 1];
    unsigned lmax = 0;
    for (int64_t k = b + lane; k < e; k += 32) {
      int j = indices[k];
      u64 q = 1;
      bool ok = quantize(values[k], q);
      if (!ok || j < 0 || j >= I) atomicOr(flag, 1);
      if (k > b && indices[k - 1] >= j) atomicOr(flag, 2);
      m_ent[k] = ((u64)(unsigned)j << 40) | (q & Q_MASK40);
      lmax = max(lmax, (unsigned)(q >> LIMB_BITS) + 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) {
      m_rowmax[i] = lmax;
      m_len[i] = (int)(e - b);
    }
  }
}

// seg[i*(P+1)+p] = offset inside row i of the first entry with column >= p*R
__global__ void k_model_seg(const int64_t* __restrict__ m_ptr, const u64* __restrict__ m_ent, int64_t I, int P, int R,
                            int* __restrict__ seg) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= I * (P + 1)) return;
  int64_t i = t / (P + 1);
  int p = (int)(t % (P + 1));
  int64_t b = m_ptr[i], lo = b, hi = m_ptr[i + 1];
  if (p == P) {
    seg[t] = (int)(hi - b);
    return;
  }
  u64 target = (u64)p * (u64)R;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((m_ent[mid] >> 40) < target) lo = mid + 1;
    else hi = mid;
  }
  seg[t] = (int)(lo - b);
}

// ---- block layout for the 32-bit scoring kernel: every row padded to a multiple of 4 entries (32 bytes, one
// sector) with all-ones entries, whose column 0xFFFFFF lies outside every item range.
__global__ void k_model_pad_len(const int64_t* __restrict__ m_ptr, int64_t I, int* __restrict__ len4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < I) len4[i] = (int)((m_ptr[i + 1] - m_ptr[i] + 3) & ~(int64_t)3);
}

__global__ void k_model_pad(const int64_t* __restrict__ m_ptr, const u64* __restrict__ m_ent,
                            const int64_t* __restrict__ ptr4, int64_t I, u64* __restrict__ ent4) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < I; i += nwarps) {
    const int64_t b = m_ptr[i], n = m_ptr[i + 1] - b, o = ptr4[i], n4 = ptr4[i + 1] - o;
    for (int64_t k = lane; k < n4; k += 32) ent4[o + k] = k < n ? m_ent[b + k] : ~0ull;
  }
}

// blk[i*P + p] = {first 4-entry block, number of blocks} covering the entries of row i with column in
// [p*R, (p+1)*R).  The blocks may also hold neighbours from the adjacent ranges of the same row (and
// padding); the kernel drops those by their column.
__global__ void k_model_blocks(const int64_t* __restrict__ m_ptr, const u64* __restrict__ m_ent,
                               const int64_t* __restrict__ ptr4, int64_t I, int P, int R, int2* __restrict__ blk) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= I * P) return;
  const int64_t i = t / P;
  const int p = (int)(t % P);
  const int64_t b = m_ptr[i], e = m_ptr[i + 1];
  auto lower = [&](u64 target) {
    int64_t lo = b, hi = e;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((m_ent[mid] >> 40) < target) lo = mid + 1;
      else hi = mid;
    }
    return lo - b;
  };
  const int64_t s0 = lower((u64)p * (u64)R), s1 = lower((u64)(p + 1) * (u64)R);
  int2 r = make_int2(0, 0);
  if (s1 > s0) {
    const int64_t first = (ptr4[i] + s0) >> 2, last = (ptr4[i] + s1 + 3) >> 2;
    r = make_int2((int)first, (int)(last - first));
  }
  blk[t] = r;
}

__global__ void k_gather_len(const int* __restrict__ len, const int64_t* __restrict__ row_src, int64_t I, int* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < I) out[i] = len[row_src[i]];
}

static void model_common_begin(rpk_ctx* c, int64_t I) {
  RPK_REQUIRE(I >= 0 && I < ((int64_t)1 << 24), "item count must be below 2^24");
  c->m_I = I;
  c->m_P = 0;  // segment tables must be rebuilt
  c->m_P2 = 0;
  c->m_pad = false;
  int* flag = c->buf<int>("m_flag", 1);
  RPK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
}

static void model_check_flag(rpk_ctx* c) {
  int h = 0;
  RPK_CUDA(cudaMemcpyAsync(&h, c->get<int>("m_flag"), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RPK_CUDA(cudaStreamSynchronize(c->stream));
  if (h & 1) {
    c->m_I = 0;
    throw Error("similarity model: values must lie in [0, 2) and columns in [0, I)");
  }
  if (h & 2) {
    c->m_I = 0;
    throw Error("similarity model: column indices must be unique (and ascending for CSR input) within a row");
  }
}

// rows_in: number of rows of the input arrays (>= I when row_src maps model rows into a larger, e.g.
// all-gathered, array); row_src: int64[I] source row of every model row, or null for the identity.
void run_model_load_topk_rows(rpk_ctx* c, int64_t I, int K, int64_t rows_in, const int32_t* idx_u, const double* val_u,
                              const int32_t* len_u, const int64_t* row_src_u) {
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(rows_in >= I || row_src_u, "fewer input rows than items");
  model_common_begin(c, I);
  cudaStream_t st = c->stream;
  const int32_t* idx = stage_in(c, idx_u, (size_t)rows_in * K, "m_in_idx");
  const double* val = stage_in(c, val_u, (size_t)rows_in * K, "m_in_val");
  const int32_t* len = stage_in(c, len_u, (size_t)rows_in, "m_in_len");
  const int64_t* row_src = row_src_u ? stage_in(c, row_src_u, (size_t)I, "m_in_rowsrc") : nullptr;
  int64_t* m_ptr = c->buf<int64_t>("m_ptr", (size_t)I + 1);
  int* m_len = c->buf<int>("m_len", (size_t)I);
  if (I > 0) {
    if (row_src) {
      k_gather_len<<<ceil_div(I, 256), 256, 0, st>>>(len, row_src, I, m_len);
      RPK_LAUNCH_CHECK(c);
    } else {
      RPK_CUDA(cudaMemcpyAsync(m_len, len, sizeof(int) * (size_t)I, cudaMemcpyDeviceToDevice, st));
    }
  }
  k_scan_i32_i64<<<1, 1024, 0, st>>>(m_len, m_ptr, I);
  RPK_LAUNCH_CHECK(c);
  u64* m_ent = c->buf<u64>("m_ent", (size_t)I * K);
  unsigned* m_rowmax = c->buf<unsigned>("m_rowmax", (size_t)I);
  if (I > 0) {
    int n2 = 2;
    while (n2 < K) n2 <<= 1;
    const int grid = (int)std::min<int64_t>(I, (int64_t)c->sm_count * 16);
    k_model_from_topk<<<grid, 128, (size_t)n2 * sizeof(u64), st>>>(idx, val, len, row_src, K, (int)I, (int)I, m_ptr, m_ent, m_rowmax,
                                                                  c->get<int>("m_flag"));
    RPK_LAUNCH_CHECK(c);
  }
  int64_t total = 0;
  RPK_CUDA(cudaMemcpyAsync(&total, m_ptr + I, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  model_check_flag(c);
  RPK_REQUIRE(total >= 0 && total <= I * (int64_t)K, "similarity model: row lengths exceed K");
  c->m_nnz = total;
  c->m_max_len = K;
}

// One warp per model row: copy its entries out of the packed input rows, check them, record the row maximum.
__global__ void k_model_from_packed(const u64* __restrict__ ent, const int64_t* __restrict__ row_src, int K, int64_t I,
                                    const int64_t* __restrict__ m_ptr, const int* __restrict__ m_len,
                                    u64* __restrict__ m_ent, unsigned* __restrict__ m_rowmax, int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < I; i += nwarps) {
    const int64_t src = row_src ? row_src[i] : i;
    const u64* row = ent + src * K;
    const int64_t base = m_ptr[i];
    const int n = m_len[i];
    unsigned lmax = 0;
    for (int t = lane; t < n; t += 32) {
      const u64 e = row[t];
      const u64 col = e >> 40, q = e & Q_MASK40;
      if (col >= (u64)I || !(q & 1ull)) atomicOr(flag, 1);
      if (t > 0 && (row[t - 1] >> 40) >= col) atomicOr(flag, 2);
      m_ent[base + t] = e;
      lmax = max(lmax, (unsigned)(q >> LIMB_BITS) + 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) m_rowmax[i] = lmax;
  }
}

void run_model_pack_rows(rpk_ctx* c, int64_t I, int K, int64_t rows, const int32_t* idx_u, const double* val_u,
                         const int32_t* len_u, uint64_t* out_u) {
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(I >= 0 && I < ((int64_t)1 << 24), "item count must be below 2^24");
  RPK_REQUIRE(rows >= 0, "negative row count");
  RPK_REQUIRE(out_u != nullptr, "out_ent must not be null");
  cudaStream_t st = c->stream;
  const int32_t* idx = stage_in(c, idx_u, (size_t)rows * K, "pk_in_idx");
  const double* val = stage_in(c, val_u, (size_t)rows * K, "pk_in_val");
  const int32_t* len = stage_in(c, len_u, (size_t)rows, "pk_in_len");
  Out<u64> o;
  o.init(c, reinterpret_cast<u64*>(out_u), (size_t)rows * K, "pk_out");
  int* flag = c->buf<int>("pk_flag", 1);
  RPK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (rows > 0) {
    int n2 = 2;
    while (n2 < K) n2 <<= 1;
    const int grid = (int)std::min<int64_t>(rows, (int64_t)c->sm_count * 16);
    k_model_from_topk<<<grid, 128, (size_t)n2 * sizeof(u64), st>>>(idx, val, len, nullptr, K, (int)I, (int)rows, nullptr, o.dev, nullptr, flag);
    RPK_LAUNCH_CHECK(c);
  }
  int h = 0;
  RPK_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  RPK_CUDA(cudaStreamSynchronize(st));
  RPK_REQUIRE(!(h & 1), "similarity lists: values must lie in [0, 2) and columns in [0, I)");
  RPK_REQUIRE(!(h & 2), "similarity lists: column indices must be unique within a row");
  o.finish(c);
  finish_call(c);
}

void run_model_load_packed_rows(rpk_ctx* c, int64_t I, int K, int64_t rows_in, const uint64_t* ent_u, const int32_t* len_u,
                                const int64_t* row_src_u) {
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(rows_in >= I || row_src_u, "fewer input rows than items");
  model_common_begin(c, I);
  cudaStream_t st = c->stream;
  const u64* ent = stage_in(c, reinterpret_cast<const u64*>(ent_u), (size_t)rows_in * K, "m_in_ent");
  const int32_t* len = stage_in(c, len_u, (size_t)rows_in, "m_in_len");
  const int64_t* row_src = row_src_u ? stage_in(c, row_src_u, (size_t)I, "m_in_rowsrc") : nullptr;
  int64_t* m_ptr = c->buf<int64_t>("m_ptr", (size_t)I + 1);
  int* m_len = c->buf<int>("m_len", (size_t)I);
  if (I > 0) {
    if (row_src) {
      k_gather_len<<<ceil_div(I, 256), 256, 0, st>>>(len, row_src, I, m_len);
      RPK_LAUNCH_CHECK(c);
    } else {
      RPK_CUDA(cudaMemcpyAsync(m_len, len, sizeof(int) * (size_t)I, cudaMemcpyDeviceToDevice, st));
    }
  }
  k_scan_i32_i64<<<1, 1024, 0, st>>>(m_len, m_ptr, I);
  RPK_LAUNCH_CHECK(c);
  u64* m_ent = c->buf<u64>("m_ent", (size_t)I * K);
  unsigned* m_rowmax = c->buf<unsigned>("m_rowmax", (size_t)I);
  if (I > 0) {
    const int grid = (int)std::min<int64_t>((I * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    k_model_from_packed<<<grid, 256, 0, st>>>(ent, row_src, K, I, m_ptr, m_len, m_ent, m_rowmax, c->get<int>("m_flag"));
    RPK_LAUNCH_CHECK(c);
  }
  int64_t total = 0;
  RPK_CUDA(cudaMemcpyAsync(&total, m_ptr + I, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  model_check_flag(c);
  RPK_REQUIRE(total >= 0 && total <= I * (int64_t)K, "similarity model: row lengths exceed K");
  c->m_nnz = total;
  c->m_max_len = K;
}

void run_model_load_topk(rpk_ctx* c, int64_t I, int K, const int32_t* idx_u, const double* val_u, const int32_t* len_u) {
  run_model_load_topk_rows(c, I, K, I, idx_u, val_u, len_u, nullptr);
}

void run_model_load_last_fit(rpk_ctx* c, int64_t token) {
  RPK_REQUIRE(c->lf_idx != nullptr && token == c->lf_token, "no complete fit result of that token is resident on the device");
  run_model_load_topk(c, c->lf_I, c->lf_K, c->lf_idx, c->lf_val, c->lf_len);
}

void run_model_load_csr(rpk_ctx* c, int64_t I, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u,
                        const double* values_u) {
  model_common_begin(c, I);
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)I + 1, "m_in_ptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "m_in_idx");
  const double* values = stage_in(c, values_u, (size_t)nnz, "m_in_val");
  int64_t* m_ptr = c->buf<int64_t>("m_ptr", (size_t)I + 1);
  RPK_CUDA(cudaMemcpyAsync(m_ptr, indptr, sizeof(int64_t) * ((size_t)I + 1), cudaMemcpyDeviceToDevice, st));
  u64* m_ent = c->buf<u64>("m_ent", (size_t)nnz);
  unsigned* m_rowmax = c->buf<unsigned>("m_rowmax", (size_t)I);
  int* m_len = c->buf<int>("m_len", (size_t)I);
  if (I > 0) {
    const int grid = (int)std::min<int64_t>((I * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    k_model_from_csr<<<grid, 256, 0, st>>>(indptr, indices, values, I, m_ent, m_rowmax, m_len, c->get<int>("m_flag"));
    RPK_LAUNCH_CHECK(c);
  }
  int64_t ends[2] = {0, 0};
  RPK_CUDA(cudaMemcpyAsync(&ends[0], m_ptr, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  RPK_CUDA(cudaMemcpyAsync(&ends[1], m_ptr + I, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  model_check_flag(c);
  RPK_REQUIRE(ends[0] == 0 && ends[1] == nnz, "similarity model: indptr does not match nnz");
  c->m_nnz = nnz;
  c->m_max_len = 0;
}

// ------------------------------------------------------------------------------------------
// Scoring kernel
// ------------------------------------------------------------------------------------------
// Candidate sources: the dense accumulator range, or the list of slots touched by this user.
struct ScoreSrc {
  const unsigned* lo;
  const unsigned* hi;
  const u64* wide;     // non-null: 64-bit accumulators
  const int* touched;  // non-null: slot list (sparse mode)
  int r0, ns;
  u64 floor_;
  __device__ __forceinline__ u64 score_at(int j) const {
    return wide ? wide[j] : (((u64)hi[j] << LIMB_BITS) + (u64)lo[j]);
  }
  __device__ __forceinline__ int nslots() const { return ns; }
  __device__ __forceinline__ u64 margin() const { return 0ull; }
  __device__ __forceinline__ void set_floor(u64 thr) { floor_ = thr; }
  // Selection key: the score rounded toward zero to float, as a bit pattern.  The map is monotone
  // (a <= b => key(a) <= key(b)), so margin() = 0 is right: equal keys are told apart by cmp3 on the exact
  // score.  Scores lie in [1, 2^53), which gives constant key bounds -- no pass over the slots is needed.
  __device__ __forceinline__ static u64 key_of(u64 score) { return (u64)__float_as_uint(__ull2float_rz(score)); }
  __device__ __forceinline__ void stats(SelShared* sh) const {
    if (threadIdx.x == 0) {
      sh->count = ns;  // upper bound of the candidate count (slots can be empty or masked)
      sh->kmin = (u64)__float_as_uint(1.0f);
      sh->kmax = (u64)__float_as_uint(9007199254740992.0f);
    }
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void visit(F f, int stride) const {
    for (int slot = threadIdx.x * stride; slot < ns; slot += blockDim.x * stride) {
      const int j = touched ? touched[slot] : slot;
      const u64 sc = score_at(j);
      if (sc != 0) {
        const u64 k = key_of(sc);
        if (k >= floor_) f(slot, k);
      }
    }
  }
  template <class F>
  __device__ __forceinline__ void for_each(F f) const { visit(f, 1); }
  template <class F>
  __device__ __forceinline__ void for_each_sampled(F f) const { visit(f, SEL_SAMPLE); }
  __device__ __forceinline__ void entry(int slot, Entry& e) const {
    const int j = touched ? touched[slot] : slot;
    e.key = score_at(j);
    e.idx = r0 + j;
    e.aux = 0;
  }
  __device__ __forceinline__ int cmp3(const Entry& a, const Entry& b) const {
    if (a.key != b.key) return a.key > b.key ? 1 : -1;
    return 0;
  }
};

enum { PRED_TOPN = 0, PRED_COUNT = 1, PRED_FILL = 2 };

struct PredParams {
  const int64_t* indptr;
  const int* indices;
  const int64_t* m_ptr;
  const u64* m_ent;
  const int* m_seg;
  const unsigned* m_rowmax;
  const int4* work_tab;  // {user, history length, row start lo, hi} in processing order
  const int* n_work;     // non-null: number of work_tab records (device side), replaces U
  int U, P, R, I, N, mask, mode, force_wide;
  int cap, direct_cap, tcap;
  int* queue;
  int* part_idx;
  u64* part_sq;
  int* part_len;
  int64_t* pass_cnt;
  const int64_t* out_indptr;
  int* out_indices;
  double* out_values;
  unsigned long long* prof;  // RPK_PHASE_PROF builds: cycles of thread 0 per phase
};

#ifdef RPK_PHASE_PROF
#define PROF_MARK(k)                                         \
  do {                                                       \
    if (tid == 0) {                                          \
      const long long t_now = clock64();                     \
      atomicAdd(p.prof + (k), (unsigned long long)(t_now - t_prev)); \
      t_prev = t_now;                                        \
    }                                                        \
  } while (0)
#else
#define PROF_MARK(k) do {} while (0)
#endif

// Invariant: the accumulators of a CTA are all zero between work items.  A light user touches few of
// the R slots of a range, so its slots are recorded on first touch (the low limb of a touched slot can
// never be zero again: every q is odd) and only those are ranked and cleared; heavy users, and the
// full-CSR modes, sweep the whole range instead.
__global__ void __launch_bounds__(1024, 1) k_predict(PredParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Entry* list = reinterpret_cast<Entry*>(smem);
  int* hist = reinterpret_cast<int*>(smem + sel_list_bytes(p.cap));
  SelShared* sh = reinterpret_cast<SelShared*>(hist + SEL_BINS);
  u64* acc64 = reinterpret_cast<u64*>(smem + sel_smem_bytes(p.cap));
  unsigned* acc_lo = reinterpret_cast<unsigned*>(acc64);
  unsigned* acc_hi = acc_lo + p.R;
  int* touched = reinterpret_cast<int*>(acc64 + p.R);
  __shared__ int s_work;
  __shared__ int s_next;      // work item fetched ahead (its user row is warmed in L2 on the way)
  __shared__ u64 s_bound;
  __shared__ int s_cnt;
  __shared__ int s_ntouched;

  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int total = (p.n_work ? *p.n_work : p.U) * p.P;
  if (total == 0) return;
  for (int s = tid; s < p.R; s += nt) acc64[s] = 0ull;
  if (tid == 0) s_next = atomicAdd(p.queue, 1);
#ifdef RPK_PHASE_PROF
  long long t_prev = clock64();
#endif
  for (;;) {
    PROF_MARK(7);
    if (tid == 0) {
      s_work = s_next;
      const int nx = atomicAdd(p.queue, 1);
      s_next = nx;
      if (nx < total) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.work_tab + nx / p.P) : "memory");
      s_bound = 0;
      s_cnt = 0;
      s_ntouched = 0;
    }
    __syncthreads();
    const int w = s_work;
    __syncthreads();
    if (w >= total) break;
    const int4 rec = p.work_tab[w / p.P];
    const int u = rec.x;
    const int pass = w % p.P;
    const int r0 = pass * p.R;
    const int ns = min(p.R, p.I - r0);
    const int64_t xb = ((int64_t)(unsigned)rec.z) | ((int64_t)rec.w << 32);
    const int d = rec.y;
    const int64_t slot_out = (int64_t)u * p.P + pass;
    PROF_MARK(0);
    if (d == 0) {  // user without history: empty prediction row (algorithms/base.py:123-127)
      if (p.mode == PRED_TOPN) {
        for (int t = tid; t < p.N; t += nt) {
          p.part_idx[slot_out * p.N + t] = -1;
          p.part_sq[slot_out * p.N + t] = 0;
        }
        if (tid == 0) p.part_len[slot_out] = 0;
      } else if (p.mode == PRED_COUNT) {
        if (tid == 0) p.pass_cnt[slot_out] = 0;
      }
      continue;
    }
    // ---- can the high limb overflow?  sum of per-row maxima bounds every score's high limb
    bool wide = p.force_wide != 0;
    if (!wide && d > LIMB_CHUNK) {
      u64 b = 0;
      for (int r = tid; r < d; r += nt) b += (u64)p.m_rowmax[p.indices[xb + r]];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
      if (lane == 0) atomicAdd(&s_bound, b);
      __syncthreads();
      wide = s_bound >= (((u64)1) << 32);
    }
    PROF_MARK(1);
    // sparse mode: track first touches (needs the limb accumulators and a single chunk)
    const bool track = !wide && d <= LIMB_CHUNK && p.mode == PRED_TOPN && p.tcap > 0;
    // ---- accumulate: history rows are dealt to the warps in chunks (<= 32 rows, one per lane, so that the
    //      segment bounds are fetched in parallel); rows are added in groups of LIMB_CHUNK so that the low
    //      limb (20 bits per term) cannot overflow 32 bits between normalisations
    // the loop is instantiated for the three accumulator modes so that the hot path carries no mode tests
    auto accumulate = [&](auto wide_c, auto track_c) {
      constexpr bool WIDE = decltype(wide_c)::value;
      constexpr bool TRACK = decltype(track_c)::value;
      for (int c0 = 0; c0 < d; c0 += LIMB_CHUNK) {
        const int c1 = min(d, c0 + LIMB_CHUNK);
        int chunk = (c1 - c0 + nwarps - 1) / nwarps;
        chunk = chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk);
        for (int base = c0 + warp * chunk; base < c1; base += nwarps * chunk) {
          const int nvalid = min(chunk, c1 - base);
          int64_t beg = 0;
          int len = 0;
          if (lane < nvalid) {
            const int i = p.indices[xb + base + lane];
            const int* sg = p.m_seg + (int64_t)i * (p.P + 1) + pass;
            const int s0 = sg[0];
            beg = p.m_ptr[i] + s0;
            len = sg[1] - s0;
          }
          // add one entry to the accumulators; in sparse mode record the slot on its first touch
          auto add_entry = [&](u64 ent) {
            bool first = false;
            int j = 0;
            if (ent != 0ull) {
              j = (int)(ent >> 40) - r0;
              const u64 q = ent & Q_MASK40;
              if (WIDE) {
                atomicAdd(&acc64[j], q);
              } else {
                const unsigned old = atomicAdd(&acc_lo[j], (unsigned)q & LIMB_MASK);
                atomicAdd(&acc_hi[j], (unsigned)(q >> LIMB_BITS));
                first = old == 0u;
              }
            }
            if (TRACK) {
              const unsigned m = __ballot_sync(0xffffffffu, first);
              if (m) {
                const int leader = __ffs(m) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(&s_ntouched, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (first) {
                  const int my = pos + __popc(m & ((1u << lane) - 1u));
                  if (my < p.tcap) touched[my] = j;
                }
              }
            }
          };
          // rows are taken four at a time: the first 96 entries of each (a row segment rarely has more) are
          // loaded up front -- 12 independent loads in flight per lane -- and only then added
          for (int l0 = 0; l0 < nvalid; l0 += 4) {
            u64 ent[4][3];
            int64_t bq[4];
            int nq[4];
  #pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int l = l0 + q;  // lanes >= nvalid hold len = 0
              bq[q] = __shfl_sync(0xffffffffu, beg, l & 31);
              nq[q] = l < nvalid ? __shfl_sync(0xffffffffu, len, l & 31) : 0;
  #pragma unroll
              for (int it = 0; it < 3; ++it) {
                const int e = it * 32 + lane;
                ent[q][it] = e < nq[q] ? p.m_ent[bq[q] + e] : 0ull;
              }
            }
  #pragma unroll
            for (int q = 0; q < 4; ++q) {
  #pragma unroll
              for (int it = 0; it < 3; ++it)
                if (it * 32 < nq[q]) add_entry(ent[q][it]);
              for (int e0 = 96; e0 < nq[q]; e0 += 32) {
                const int e = e0 + lane;
                add_entry(e < nq[q] ? p.m_ent[bq[q] + e] : 0ull);
              }
            }
          }
        }
        __syncthreads();
        if (!WIDE && c1 < d) {  // carry the low limb into the high limb before the next chunk
          for (int s = tid; s < ns; s += nt) {
            unsigned l = acc_lo[s];
            acc_hi[s] += l >> LIMB_BITS;
            acc_lo[s] = l & LIMB_MASK;
          }
          __syncthreads();
        }
      }
    };
    if (wide) accumulate(std::true_type{}, std::false_type{});
    else if (track) accumulate(std::false_type{}, std::true_type{});
    else accumulate(std::false_type{}, std::false_type{});
    const int n_touched = s_ntouched;
    const bool sparse = track && n_touched <= p.tcap;
    PROF_MARK(2);
#ifdef RPK_PHASE_PROF
    if (tid == 0) {
      atomicAdd(p.prof + 8, 1ull);
      atomicAdd(p.prof + 9, (unsigned long long)sparse);
      atomicAdd(p.prof + 10, (unsigned long long)n_touched);
    }
#endif
    if (p.mask) {  // pipelines/pipeline.py:174-175 -- before the truncation to N
      for (int r = tid; r < d; r += nt) {
        const int j = p.indices[xb + r] - r0;
        if (j >= 0 && j < ns) {
          if (wide) acc64[j] = 0ull;
          else {
            acc_lo[j] = 0u;
            acc_hi[j] = 0u;
          }
        }
      }
      __syncthreads();
    }
    ScoreSrc src{acc_lo, acc_hi, wide ? acc64 : nullptr, sparse ? touched : nullptr, r0, sparse ? n_touched : ns, 0ull};
    PROF_MARK(3);
    if (p.mode == PRED_TOPN) {
      const int m = block_select_topk(src, p.N, list, p.cap, p.direct_cap, hist, sh);
      PROF_MARK(4);
      for (int t = tid; t < p.N; t += nt) {
        p.part_idx[slot_out * p.N + t] = t < m ? list[t].idx : -1;
        p.part_sq[slot_out * p.N + t] = t < m ? list[t].key : 0ull;
      }
      if (tid == 0) p.part_len[slot_out] = m;
    } else if (p.mode == PRED_COUNT) {
      int cnt = 0;
      for (int s = tid; s < ns; s += nt) cnt += src.score_at(s) != 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
      __syncthreads();
      if (tid == 0) p.pass_cnt[slot_out] = s_cnt;
    } else {
      int64_t out = p.out_indptr[u];
      for (int q = 0; q < pass; ++q) out += p.pass_cnt[(int64_t)u * p.P + q];
      int running = 0;
      for (int base = 0; base < ns; base += nt) {
        const int s = base + tid;
        const u64 sc = s < ns ? src.score_at(s) : 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, sc != 0);
        if (lane == 0) sh->warp_tot[warp] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int q = 0; q < nwarps; ++q) {
          int v = sh->warp_tot[q];
          if (q < warp) off += v;
          tot += v;
        }
        if (sc != 0) {
          const int pos = running + off + __popc(bal & ((1u << lane) - 1u));
          p.out_indices[out + pos] = r0 + s;
          p.out_values[out + pos] = (double)sc * (1.0 / 549755813888.0);
        }
        running += tot;
        __syncthreads();
      }
    }
    __syncthreads();
    PROF_MARK(5);
    // ---- restore the all-zero invariant
    if (sparse) {
      for (int t = tid; t < n_touched; t += nt) {
        const int j = touched[t];
        acc_lo[j] = 0u;
        acc_hi[j] = 0u;
      }
    } else {
      for (int s = tid; s < ns; s += nt) acc64[s] = 0ull;
      if (ns < p.R)
        for (int s = tid; s < ns; s += nt) acc_hi[s] = 0u;
    }
    __syncthreads();
    PROF_MARK(6);
  }
}


// ------------------------------------------------------------------------------------------
// Scoring kernel, top-N mode: 32-bit approximate accumulators + exact scores for the survivors
// ------------------------------------------------------------------------------------------
// The two-limb kernel above pays two shared-memory atomics per similarity entry and 8 bytes per item slot.
// For top-N only a handful of scores per user matter, so this kernel runs two sweeps over the user's rows:
//   1. a_j += (q >> s) | 1 with ONE 32-bit atomic per entry (s = 9 + ceil(log2 d) keeps every sum below
//      2^32).  Each term is within 1 of q / 2^s, so |a_j - score_j / 2^s| <= d: a_j orders two items
//      correctly whenever their a differ by more than 2d.
//   2. the selection keeps every item whose a_j is within 2d of the N-th largest (a few more than N); their
//      slots are marked and the rows are streamed again, adding the exact q of marked slots only.
// The survivors are then ordered by their exact scores.  4 bytes per slot let two CTAs share an SM, so one
// CTA's latency-bound steps (work fetch, selection, output) hide behind the other's atomics.  A user whose
// survivors do not fit the list (huge groups of near-equal scores) is handed to the two-limb kernel.
struct ApproxSrc {
  const unsigned* acc;
  const int* touched;  // non-null: slot list (sparse mode)
  int r0, ns;
  u64 floor_;
  u64 margin_;
  u64 kmax_;
  __device__ __forceinline__ int nslots() const { return ns; }
  __device__ __forceinline__ u64 margin() const { return margin_; }
  __device__ __forceinline__ void set_floor(u64 thr) { floor_ = thr; }
  __device__ __forceinline__ void stats(SelShared* sh) const {
    if (threadIdx.x == 0) {
      sh->count = ns;
      sh->kmin = 1ull;
      sh->kmax = kmax_;
    }
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void visit(F f, int stride) const {
    for (int slot = threadIdx.x * stride; slot < ns; slot += blockDim.x * stride) {
      const int j = touched ? touched[slot] : slot;
      const u64 k = (u64)acc[j];
      if (k != 0 && k >= floor_) f(slot, k);
    }
  }
  template <class F>
  __device__ __forceinline__ void for_each(F f) const { visit(f, 1); }
  template <class F>
  __device__ __forceinline__ void for_each_sampled(F f) const { visit(f, SEL_SAMPLE); }
  __device__ __forceinline__ void entry(int slot, Entry& e) const {
    const int j = touched ? touched[slot] : slot;
    e.key = (u64)acc[j];
    e.idx = r0 + j;
    e.aux = 0;
  }
  __device__ __forceinline__ int cmp3(const Entry& a, const Entry& b) const {  // unused (SURVIVORS_ONLY)
    if (a.key != b.key) return a.key > b.key ? 1 : -1;
    return 0;
  }
};

struct ExactOrder {
  __device__ __forceinline__ int cmp3(const Entry& a, const Entry& b) const {
    if (a.key != b.key) return a.key > b.key ? 1 : -1;
    return 0;
  }
};

struct Pred32Params {
  const int* indices;
  const uint4* ent4;  // model rows in 4-entry blocks (two uint4 each)
  const int2* blk;    // {first block, blocks} per (row, item range)
  const int4* work_tab;
  int U, P, R, I, N, mask;
  int cap, direct_cap, tcap;
  int* queue;
  int* part_idx;
  u64* part_sq;
  int* part_len;
  int* ovf_flag;   // per user: 1 = handed to the two-limb kernel
  int* ovf_count;  // number of such users
  int4* ovf_tab;   // their work records
  unsigned long long* prof;
};

constexpr int A32_BITS = 11;            // selection histogram of the 32-bit kernel: 2048 bins
constexpr int A32_BINS = 1 << A32_BITS;
constexpr int A32_ROWS = 512;           // history rows staged per chunk (>= block size)

// Row table of one chunk of the user's history: first block and number of blocks of every row's segment.
struct RowTab {
  int start[A32_ROWS];
  int nb[A32_ROWS];
};

// Stages rows [c0, c0 + n) of the history, one row per thread.
__device__ __forceinline__ void stage_rows(const Pred32Params& p, int64_t xb, int c0, int n, int pass, RowTab* rt) {
  const int tid = threadIdx.x;
  if (tid < n) {
    const int i = p.indices[xb + c0 + tid];
    const int2 sb = __ldg(p.blk + (int64_t)i * p.P + pass);
    rt->start[tid] = sb.x;
    rt->nb[tid] = sb.y;
  }
  __syncthreads();
}

// Calls f(entry) for every entry of the staged segments (entries of neighbouring ranges and padding
// included -- f drops them by their column).  One warp per row, a lane per entry, so that a load covers
// 256 contiguous bytes; two rows (up to six loads per lane) are in flight before anything is added.
template <class F>
__device__ __forceinline__ void sweep_rows(const Pred32Params& p, const RowTab* rt, int n, F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const u64* ent = reinterpret_cast<const u64*>(p.ent4);
  for (int r = warp; r < n; r += 2 * nwarps) {
    const int r2 = r + nwarps;
    const u64* b1 = ent + (int64_t)rt->start[r] * 4 + lane;
    const int n1 = rt->nb[r] * 4 - lane;  // entries left from this lane's first one
    const u64* b2 = b1;
    int n2 = 0;
    if (r2 < n) {
      b2 = ent + (int64_t)rt->start[r2] * 4 + lane;
      n2 = rt->nb[r2] * 4 - lane;
    }
    u64 e1[3], e2[3];
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      e1[it] = 32 * it < n1 ? __ldg(b1 + 32 * it) : ~0ull;
      e2[it] = 32 * it < n2 ? __ldg(b2 + 32 * it) : ~0ull;
    }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      f(e1[it]);
      f(e2[it]);
    }
    for (int e = 96; e < n1; e += 32) f(__ldg(b1 + e));
    for (int e = 96; e < n2; e += 32) f(__ldg(b2 + e));
  }
}

// Survivors of the approximate scores in one histogram round: 2048 bins over [0, kmax] (kmax = the largest
// sum seen by sweep 1), the bin holding the K-th largest key found with two barriers, then every candidate
// with key >= that bin's lower edge - margin is copied to the list.  Returns their number (unsorted), or -1
// when they do not fit -- the caller then runs the general refinement.  `hist` must be all zero on entry and is
// all zero again on return.
__device__ int a32_select(const unsigned* acc, const int* touched, int n, int r0, unsigned kmax, u64 margin, int K,
                          Entry* list, int cap, int direct_cap, int* hist, SelShared* sh) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  unsigned thr = 1u;
  if (tid == 0) sh->count = 0;
  if (n > direct_cap) {
    const int bits = 32 - __clz(kmax | 1u);
    const int shift = bits > A32_BITS ? bits - A32_BITS : 0;
    for (int slot = tid; slot < n; slot += nt) {
      const unsigned k = acc[touched ? touched[slot] : slot];
      if (k) atomicAdd(&hist[min(k >> shift, (unsigned)(A32_BINS - 1))], 1);
    }
    __syncthreads();
    const int per = (A32_BINS + nt - 1) / nt;
    const int b0 = min(A32_BINS, tid * per), b1 = min(A32_BINS, b0 + per);
    int tsum = 0;
    for (int b = b0; b < b1; ++b) tsum += hist[b];
    int incl = tsum;  // members of my bins and of the higher lanes' bins
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    if (lane == 0) sh->warp_tot[warp] = incl;
    if (tid == 0) sh->bstar = 0;  // fewer than K candidates: everything survives
    __syncthreads();
    int above = (lane > warp && lane < nwarps) ? sh->warp_tot[lane] : 0;  // totals of the higher warps
    above = __reduce_add_sync(0xffffffffu, above) + incl - tsum;
    if (above < K && above + tsum >= K) {
      for (int b = b1 - 1; b >= b0; --b) {
        above += hist[b];
        if (above >= K) {
          sh->bstar = b;
          break;
        }
      }
    }
    __syncthreads();
    for (int b = b0; b < b1; ++b) hist[b] = 0;
    const u64 edge = (u64)sh->bstar << shift;
    thr = edge > margin + 1ull ? (unsigned)(edge - margin) : 1u;
  } else {
    __syncthreads();
  }
  for (int slot = tid; slot < n; slot += nt) {
    const int j = touched ? touched[slot] : slot;
    const unsigned k = acc[j];
    if (k >= thr) {
      const int pos = atomicAdd(&sh->count, 1);
      if (pos < cap) {
        Entry e;
        e.key = (u64)k;
        e.idx = r0 + j;
        e.aux = 0;
        list[pos] = e;
      }
    }
  }
  __syncthreads();
  const int m = sh->count;
  return m > cap ? -1 : m;
}

#ifdef RPK_PHASE_PROF
#define PROF32_MARK(k) PROF_MARK(k)
#else
#define PROF32_MARK(k) do {} while (0)
#endif

__host__ __device__ __forceinline__ size_t a32_fixed_bytes(int cap) {
  return sel_smem_bytes(cap, A32_BINS) + ((sizeof(RowTab) + 15) / 16) * 16;
}

__global__ void __launch_bounds__(512, 2) k_predict_a32(Pred32Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Entry* list = reinterpret_cast<Entry*>(smem);
  int* hist = reinterpret_cast<int*>(smem + sel_list_bytes(p.cap));
  SelShared* sh = reinterpret_cast<SelShared*>(hist + A32_BINS);
  RowTab* rt = reinterpret_cast<RowTab*>(smem + sel_smem_bytes(p.cap, A32_BINS));
  unsigned* acc = reinterpret_cast<unsigned*>(smem + a32_fixed_bytes(p.cap));
  int* touched = reinterpret_cast<int*>(acc + p.R);
  __shared__ int s_work;
  __shared__ int s_next;
  __shared__ int s_ntouched;
  __shared__ unsigned s_kmax;

  const int tid = threadIdx.x, nt = blockDim.x;
  const int total = p.U * p.P;
  for (int s = tid; s < p.R; s += nt) acc[s] = 0u;
  for (int b = tid; b < A32_BINS; b += nt) hist[b] = 0;
  if (tid == 0) s_next = atomicAdd(p.queue, 1);
#ifdef RPK_PHASE_PROF
  long long t_prev = clock64();
#endif
  for (;;) {
    if (tid == 0) {
      s_work = s_next;
      const int nx = atomicAdd(p.queue, 1);
      s_next = nx;
      if (nx < total) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.work_tab + nx / p.P) : "memory");
      s_ntouched = 0;
      s_kmax = 0u;
    }
    __syncthreads();
    const int w = s_work;
    __syncthreads();
    if (w >= total) break;
    const int4 rec = p.work_tab[w / p.P];
    const int u = rec.x;
    const int pass = w % p.P;
    const int r0 = pass * p.R;
    const int ns = min(p.R, p.I - r0);
    const int64_t xb = ((int64_t)(unsigned)rec.z) | ((int64_t)rec.w << 32);
    const int d = rec.y;
    const int64_t slot_out = (int64_t)u * p.P + pass;
    PROF32_MARK(0);
    if (d == 0) {  // user without history: empty prediction row (algorithms/base.py:123-127)
      for (int t = tid; t < p.N; t += nt) {
        p.part_idx[slot_out * p.N + t] = -1;
        p.part_sq[slot_out * p.N + t] = 0;
      }
      if (tid == 0) p.part_len[slot_out] = 0;
      continue;
    }
    const int sft = 9 + (d > 1 ? 32 - __clz(d - 1) : 0);
    const int rows_per_chunk = min(nt, A32_ROWS);
    // ---- sweep 1: approximate scores, slots recorded on their first touch (a sum is never zero again)
    unsigned mx = 0u;  // largest sum this thread produced
    for (int c0 = 0; c0 < d; c0 += rows_per_chunk) {
      const int n = min(rows_per_chunk, d - c0);
      if (c0 > 0) __syncthreads();  // the previous chunk's table is still being read
      stage_rows(p, xb, c0, n, pass, rt);
      sweep_rows(p, rt, n, [&](u64 e) {
        const unsigned j = (unsigned)(e >> 40) - (unsigned)r0;
        if (j < (unsigned)ns) {
          const unsigned a = (unsigned)((e & Q_MASK40) >> sft) | 1u;
          const unsigned old = atomicAdd(&acc[j], a);
          mx = max(mx, old + a);
          if (old == 0u) {
            const int pos = atomicAdd(&s_ntouched, 1);
            if (pos < p.tcap) touched[pos] = (int)j;
          }
        }
      });
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31) == 0 && mx) atomicMax(&s_kmax, mx);
    __syncthreads();
    const int n_touched = s_ntouched;
    const bool sparse = n_touched <= p.tcap;
    PROF32_MARK(1);
#ifdef RPK_PHASE_PROF
    if (tid == 0) {
      atomicAdd(p.prof + 8, 1ull);
      atomicAdd(p.prof + 9, (unsigned long long)sparse);
      atomicAdd(p.prof + 10, (unsigned long long)n_touched);
    }
#endif
    if (p.mask) {  // pipelines/pipeline.py:174-175 -- before the truncation to N
      for (int r = tid; r < d; r += nt) {
        const int j = p.indices[xb + r] - r0;
        if (j >= 0 && j < ns) acc[j] = 0u;
      }
      __syncthreads();
    }
    PROF32_MARK(2);
    ApproxSrc src{acc, sparse ? touched : nullptr, r0, sparse ? n_touched : ns, 0ull, 2ull * (u64)d,
                  (u64)d * ((1ull << (40 - sft)) + 1ull)};
    int m = a32_select(acc, src.touched, src.ns, r0, s_kmax, src.margin_, p.N, list, p.cap, p.direct_cap, hist, sh);
    if (m < 0) {  // a crowded boundary bin: general refinement
      __syncthreads();
      m = block_select_topk<true, A32_BITS>(src, p.N, list, p.cap, p.direct_cap, hist, sh);
      for (int b = tid; b < A32_BINS; b += nt) hist[b] = 0;
    }
    PROF32_MARK(3);
#ifdef RPK_PHASE_PROF
    if (tid == 0) atomicAdd(p.prof + 11, (unsigned long long)(m < 0 ? 0 : m));
#endif
    // ---- restore the all-zero invariant
    if (sparse) {
      for (int t = tid; t < n_touched; t += nt) acc[touched[t]] = 0u;
    } else {
      for (int s = tid; s < ns; s += nt) acc[s] = 0u;
    }
    if (m < 0) {  // survivors do not fit: the exact kernel takes the whole user
      if (tid == 0 && atomicExch(&p.ovf_flag[u], 1) == 0) p.ovf_tab[atomicAdd(p.ovf_count, 1)] = rec;
      __syncthreads();
      continue;
    }
    __syncthreads();
    // ---- sweep 2: exact scores of the survivors (slot -> survivor number + 1, key -> exact sum)
    for (int t = tid; t < m; t += nt) {
      acc[list[t].idx - r0] = (unsigned)t + 1u;
      list[t].key = 0ull;
    }
    __syncthreads();
    PROF32_MARK(4);
    const bool limbs = d <= LIMB_CHUNK;  // two 32-bit adds (20-bit limbs) cannot overflow
    if (m > 0) {
      for (int c0 = 0; c0 < d; c0 += rows_per_chunk) {
        const int n = min(rows_per_chunk, d - c0);
        if (d > rows_per_chunk) {  // a single chunk is still staged from sweep 1
          if (c0 > 0) __syncthreads();
          stage_rows(p, xb, c0, n, pass, rt);
        }
        sweep_rows(p, rt, n, [&](u64 e) {
          const unsigned j = (unsigned)(e >> 40) - (unsigned)r0;
          if (j < (unsigned)ns) {
            const unsigned cn = acc[j];
            if (cn) {
              const u64 q = e & Q_MASK40;
              if (limbs) {
                unsigned* wd = reinterpret_cast<unsigned*>(&list[cn - 1u].key);
                atomicAdd(wd, (unsigned)q & LIMB_MASK);
                atomicAdd(wd + 1, (unsigned)(q >> LIMB_BITS));
              } else {
                atomicAdd(&list[cn - 1u].key, q);
              }
            }
          }
        });
      }
      __syncthreads();
    }
    PROF32_MARK(5);
    for (int t = tid; t < m; t += nt) {
      acc[list[t].idx - r0] = 0u;
      if (limbs) {
        const u64 kv = list[t].key;
        list[t].key = ((kv >> 32) << LIMB_BITS) + (kv & 0xffffffffull);
      }
    }
    __syncthreads();
    ExactOrder ord;
    const int mo = m < p.N ? m : p.N;
    if (m <= SEL_RANK_MAX) {
      // rank every survivor by counting the ones that precede it (nt / 64 threads per survivor) and write
      // it straight to its place
      const int tpe = nt / SEL_RANK_MAX, i = tid / tpe, part = tid % tpe;
      int rank = 0;
      Entry a;
      a.key = 0;
      a.idx = 0;
      if (i < m) {
        a = list[i];
        for (int f = part; f < m; f += tpe) rank += (f != i) && entry_before(ord, list[f], a);
      }
      for (int o = tpe >> 1; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
      if (i < m && part == 0 && rank < p.N) {
        p.part_idx[slot_out * p.N + rank] = a.idx;
        p.part_sq[slot_out * p.N + rank] = a.key;
      }
    } else {
      int n2 = 1;
      while (n2 < m) n2 <<= 1;
      for (int i = m + tid; i < n2; i += nt) {
        Entry s;
        s.key = 0;
        s.idx = SENTINEL_IDX;
        s.aux = 0;
        list[i] = s;
      }
      __syncthreads();
      bitonic_sort_entries(ord, list, n2);
      for (int t = tid; t < mo; t += nt) {
        p.part_idx[slot_out * p.N + t] = list[t].idx;
        p.part_sq[slot_out * p.N + t] = list[t].key;
      }
    }
    for (int t = mo + tid; t < p.N; t += nt) {
      p.part_idx[slot_out * p.N + t] = -1;
      p.part_sq[slot_out * p.N + t] = 0ull;
    }
    if (tid == 0) p.part_len[slot_out] = mo;
    __syncthreads();
    PROF32_MARK(6);
  }
}

// One warp per user: merge the P per-range lists (each best-first) into the final top-N.
// Users flagged in `alt_flag` take their lists from the second set (the exact kernel's, alt_P ranges).
__global__ void k_predict_finalize(const int* __restrict__ part_idx, const u64* __restrict__ part_sq,
                                   const int* __restrict__ part_len, int64_t U, int P, int N, int* __restrict__ out_idx,
                                   double* __restrict__ out_val, int* __restrict__ out_len,
                                   const int* __restrict__ alt_flag, const int* __restrict__ alt_idx,
                                   const u64* __restrict__ alt_sq, const int* __restrict__ alt_len, int alt_P) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = warp; u < U; u += nwarps) {
    const bool alt = alt_flag && alt_flag[u];
    const int Pu = alt ? alt_P : P;
    const int PN = Pu * N;
    const int* pi = (alt ? alt_idx : part_idx) + u * PN;
    const u64* ps = (alt ? alt_sq : part_sq) + u * PN;
    const int* pl = (alt ? alt_len : part_len) + u * Pu;
    int tot = 0;
    for (int q = 0; q < Pu; ++q) tot += pl[q];
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
        if (out_val) out_val[u * N + rank] = (double)se * (1.0 / 549755813888.0);
      }
    }
    for (int t = m + lane; t < N; t += 32) {
      out_idx[u * N + t] = -1;
      if (out_val) out_val[u * N + t] = 0.0;
    }
    if (lane == 0) out_len[u] = m;
  }
}

// Work table in processing order: one 16-byte record per user {user, history length, row start}, so that a
// CTA needs a single load (prefetched one item ahead) to start on a user.
__global__ void k_build_work_tab(const int* __restrict__ order, const int64_t* __restrict__ indptr, int64_t U,
                                 int4* __restrict__ tab) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= U) return;
  const int u = order[k];
  const int64_t xb = indptr[u];
  const int64_t d = indptr[u + 1] - xb;
  tab[k] = make_int4(u, (int)d, (int)(xb & 0xffffffffll), (int)(xb >> 32));
}

__global__ void k_row_lengths(const int64_t* __restrict__ indptr, int64_t U, u64* __restrict__ work) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < U) work[u] = (u64)(indptr[u + 1] - indptr[u]);
}

__global__ void k_sum_passes(const int64_t* __restrict__ pass_cnt, int64_t U, int P, int64_t* __restrict__ row_nnz) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < U) {
    int64_t s = 0;
    for (int q = 0; q < P; ++q) s += pass_cnt[u * P + q];
    row_nnz[u] = s;
  }
}

// ------------------------------------------------------------------------------------------
// Host driver
// ------------------------------------------------------------------------------------------
static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct PredGeom {
  int cap, direct_cap, P, R, nt, tcap;
  size_t fixed, smem;
};

static PredGeom predict_geometry(rpk_ctx* c, int N) {
  PredGeom g;
  const bool tiny = c->flags & DBG_TINY_LIST;
  g.cap = std::max(tiny ? 64 : 256, next_pow2(2 * std::max(N, 1)));
  g.direct_cap = tiny ? std::max(N, 1) : std::min(g.cap, std::max(64, 2 * N));
  g.fixed = sel_smem_bytes(g.cap);
  RPK_REQUIRE((size_t)c->smem_max > g.fixed + 1024 + 8192, "N too large for shared memory");
  const size_t avail = (size_t)c->smem_max - g.fixed - 1024;
  const int64_t I = c->m_I;
  // 8 B of accumulator per item of a range + a list of touched slots (4 B each, up to 8192 of them)
  int P = 1;
  int64_t R = 0, T = 0;
  for (;; ++P) {
    R = ((I + P - 1) / P + 3) & ~(int64_t)3;
    if (R < 4) R = 4;
    T = std::min<int64_t>(R, tiny ? 48 : 8192);
    if ((size_t)(R * 8 + T * 4) <= avail) break;
  }
  if ((c->flags & DBG_MULTI_PASS) && P < 2 && I >= 8) {
    P = 2;
    R = ((I + P - 1) / P + 3) & ~(int64_t)3;
    T = std::min<int64_t>(R, tiny ? 48 : 8192);
  }
  g.P = P;
  g.R = (int)R;
  g.tcap = (int)T;
  g.smem = g.fixed + (size_t)R * sizeof(u64) + (size_t)T * sizeof(int);
  g.nt = g.R >= 8192 ? 1024 : (g.R >= 2048 ? 512 : 256);
  if (const char* e = getenv("RPK_PRED_NT")) {  // tuning hook
    int v = atoi(e);
    if (v >= 64 && v <= 1024 && v % 32 == 0) g.nt = v;
  }
  return g;
}

// Segment table of the two-limb kernel: offset of each item range inside every model row.
static void ensure_segments(rpk_ctx* c, const PredGeom& g) {
  if (c->m_P == g.P && c->m_R == g.R) return;
  const int64_t I = c->m_I;
  int* seg = c->buf<int>("m_seg", (size_t)I * (g.P + 1));
  if (I > 0) {
    k_model_seg<<<ceil_div(I * (g.P + 1), 256), 256, 0, c->stream>>>(c->get<int64_t>("m_ptr"), c->get<u64>("m_ent"), I, g.P,
                                                                      g.R, seg);
    RPK_LAUNCH_CHECK(c);
  }
  c->m_P = g.P;
  c->m_R = g.R;
}

// Padded block layout of the model (once per model) and the block table of a geometry.
static void ensure_blocks(rpk_ctx* c, int P, int R) {
  const int64_t I = c->m_I;
  cudaStream_t st = c->stream;
  const int64_t* m_ptr = c->get<int64_t>("m_ptr");
  const u64* m_ent = c->get<u64>("m_ent");
  if (!c->m_pad) {
    int* len4 = c->buf<int>("m_len4", (size_t)I);
    int64_t* ptr4 = c->buf<int64_t>("m_ptr4", (size_t)I + 1);
    k_model_pad_len<<<ceil_div(I, 256), 256, 0, st>>>(m_ptr, I, len4);
    RPK_LAUNCH_CHECK(c);
    k_scan_i32_i64<<<1, 1024, 0, st>>>(len4, ptr4, I);
    RPK_LAUNCH_CHECK(c);
    // rows are padded by at most 3 entries each
    u64* ent4 = c->buf<u64>("m_ent4", (size_t)c->m_nnz + 3 * (size_t)I + 4);
    k_model_pad<<<(int)std::min<int64_t>(ceil_div(I * 32, 256), (int64_t)c->sm_count * 32), 256, 0, st>>>(m_ptr, m_ent, ptr4, I, ent4);
    RPK_LAUNCH_CHECK(c);
    c->m_pad = true;
    c->m_P2 = 0;
  }
  if (c->m_P2 == P && c->m_R2 == R) return;
  int2* blk = c->buf<int2>("m_blk", (size_t)I * P);
  k_model_blocks<<<ceil_div(I * P, 256), 256, 0, st>>>(m_ptr, m_ent, c->get<int64_t>("m_ptr4"), I, P, R, blk);
  RPK_LAUNCH_CHECK(c);
  c->m_P2 = P;
  c->m_R2 = R;
}

// Geometry of the 32-bit kernel: two CTAs per SM when that costs at most one more item range than one CTA
// per SM would need (or at most three ranges), else one.
struct Pred32Geom {
  int cap, direct_cap, P, R, nt, tcap, ctas;
  size_t smem;
};

static Pred32Geom predict32_geometry(rpk_ctx* c, int N) {
  Pred32Geom g;
  const bool tiny = c->flags & DBG_TINY_LIST;
  g.cap = std::max(tiny ? 64 : 256, next_pow2(2 * std::max(N, 1)));
  g.direct_cap = tiny ? std::max(N, 1) : std::min(g.cap, std::max(64, 2 * N));
  const size_t fixed = a32_fixed_bytes(g.cap);
  const int64_t I = c->m_I;
  auto solve = [&](int ctas, int& P, int64_t& R, int64_t& T) -> bool {
    const size_t per_cta = std::min<size_t>((size_t)c->smem_max, (size_t)c->smem_per_sm / ctas - 1024);
    if (per_cta < fixed + 1024 + 4096) return false;
    const size_t avail = per_cta - fixed - 256;
    for (P = 1;; ++P) {
      R = ((I + P - 1) / P + 3) & ~(int64_t)3;
      if (R < 4) R = 4;
      // touched-slot list: what is left after the accumulators, at least 2048 slots (or all of them)
      const int64_t left = ((int64_t)avail - R * 4) / 4;
      T = std::min<int64_t>(R, tiny ? 48 : std::min<int64_t>(left, 8192));
      if (left >= std::min<int64_t>(R, tiny ? 48 : 2048)) return true;
      if (R <= 4) return false;
    }
  };
  int P1 = 0, P2 = 0;
  int64_t R1 = 0, T1 = 0, R2 = 0, T2 = 0;
  const bool ok1 = solve(1, P1, R1, T1);
  const bool ok2 = solve(2, P2, R2, T2);
  RPK_REQUIRE(ok1, "N too large for shared memory");
  bool two = ok2 && (P2 <= 3 || P2 <= P1 + 1);
  if (const char* e = getenv("RPK_PRED_CTAS")) {  // tuning hook
    const int v = atoi(e);
    if (v == 1) two = false;
    if (v == 2 && ok2) two = true;
  }
  g.ctas = two ? 2 : 1;
  g.P = two ? P2 : P1;
  int64_t R = two ? R2 : R1, T = two ? T2 : T1;
  if ((c->flags & DBG_MULTI_PASS) && g.P < 2 && I >= 8) {
    g.P = 2;
    R = ((I + g.P - 1) / g.P + 3) & ~(int64_t)3;
    T = std::min<int64_t>(R, tiny ? 48 : 4096);
  }
  g.R = (int)R;
  g.tcap = (int)T;
  g.smem = fixed + (size_t)R * 4 + (size_t)T * 4;
  g.nt = two ? 512 : (g.R >= 8192 ? 1024 : (g.R >= 2048 ? 512 : 256));
  if (g.nt > 512) g.nt = 512;  // the kernel is compiled for at most 512 threads
  if (const char* e = getenv("RPK_PRED_NT")) {  // tuning hook
    int v = atoi(e);
    if (v == 64 || v == 128 || v == 256 || v == 512) g.nt = v;
  }
  return g;
}

// Users in processing order (heaviest first) as 16-byte work records, plus the zeroed work counters.
struct PredWork {
  int4* tab;
  int* queue;   // [0]: two-limb kernel, [1]: 32-bit kernel
};

static PredWork prepare_work(rpk_ctx* c, int64_t U, const int64_t* indptr) {
  cudaStream_t st = c->stream;
  u64* work = c->buf<u64>("p_work", (size_t)U);
  int* order = c->buf<int>("p_order", (size_t)U);
  int* bcnt = c->buf<int>("p_bcnt", 65 * 2 + 4);
  int* boff = bcnt + 65;
  int* queue = boff + 65;
  RPK_CUDA(cudaMemsetAsync(bcnt, 0, sizeof(int) * (65 * 2 + 4), st));
  k_row_lengths<<<ceil_div(U, 256), 256, 0, st>>>(indptr, U, work);
  RPK_LAUNCH_CHECK(c);
  k_bucket_count<<<ceil_div(U, 256), 256, 0, st>>>(work, 0, U, bcnt);
  RPK_LAUNCH_CHECK(c);
  k_bucket_offsets<<<1, 32, 0, st>>>(bcnt, boff);
  RPK_LAUNCH_CHECK(c);
  k_bucket_scatter<<<ceil_div(U, 256), 256, 0, st>>>(work, 0, U, boff, order);
  RPK_LAUNCH_CHECK(c);
  int4* tab = c->buf<int4>("p_work_tab", (size_t)U);
  k_build_work_tab<<<ceil_div(U, 256), 256, 0, st>>>(order, indptr, U, tab);
  RPK_LAUNCH_CHECK(c);
  return PredWork{tab, queue};
}

// Launches the two-limb kernel on the records of pp.work_tab (pp.n_work: device-side count, or null = U).
static void launch_predict_kernel(rpk_ctx* c, PredParams& pp, const PredGeom& g, int64_t U) {
  RPK_CUDA(cudaFuncSetAttribute(k_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  int occ = 0;
  RPK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_predict, g.nt, g.smem));
  RPK_REQUIRE(occ >= 1, "predict kernel does not fit on an SM");
  const int64_t total = U * g.P;
  const int grid = (int)std::min<int64_t>(total, (int64_t)c->sm_count * occ);
  pp.prof = nullptr;
#ifdef RPK_PHASE_PROF
  pp.prof = c->buf<unsigned long long>("p_prof", 16);
  RPK_CUDA(cudaMemsetAsync(pp.prof, 0, 16 * sizeof(unsigned long long), c->stream));
#endif
  k_predict<<<grid, g.nt, g.smem, c->stream>>>(pp);
  RPK_LAUNCH_CHECK(c);
}

static void launch_predict(rpk_ctx* c, PredParams& pp, const PredGeom& g, int64_t U, const int64_t* indptr) {
  const PredWork w = prepare_work(c, U, indptr);
  pp.work_tab = w.tab;
  pp.n_work = nullptr;
  pp.queue = w.queue;
  c->ev_record(4);
  launch_predict_kernel(c, pp, g, U);
  c->ev_record(5);
  c->ev_valid[2] = true;
}

static void fill_common(rpk_ctx* c, PredParams& pp, const PredGeom& g, int64_t U, const int64_t* indptr,
                        const int32_t* indices, int N, int mask, int mode) {
  pp.indptr = indptr;
  pp.indices = indices;
  pp.m_ptr = c->get<int64_t>("m_ptr");
  pp.m_ent = c->get<u64>("m_ent");
  pp.m_seg = c->get<int>("m_seg");
  pp.n_work = nullptr;
  pp.m_rowmax = c->get<unsigned>("m_rowmax");
  pp.U = (int)U;
  pp.P = g.P;
  pp.R = g.R;
  pp.I = (int)c->m_I;
  pp.N = N;
  pp.mask = mask;
  pp.mode = mode;
  pp.force_wide = (c->flags & DBG_WIDE_ACC) ? 1 : 0;
  pp.cap = g.cap;
  pp.direct_cap = g.direct_cap;
  pp.tcap = g.tcap;
  pp.part_idx = nullptr;
  pp.part_sq = nullptr;
  pp.part_len = nullptr;
  pp.pass_cnt = nullptr;
  pp.out_indptr = nullptr;
  pp.out_indices = nullptr;
  pp.out_values = nullptr;
}

static void check_predict_args(rpk_ctx* c, int64_t U, int64_t nnz) {
  RPK_REQUIRE(c->m_I > 0 || c->m_nnz == 0, "no similarity model loaded");
  RPK_REQUIRE(c->bufs.count("m_ptr") != 0, "no similarity model loaded (call rpk_model_load_* first)");
  RPK_REQUIRE(U >= 0 && nnz >= 0, "negative dimension");
  RPK_REQUIRE(U < ((int64_t)1 << 27), "too many users in one call; split the batch");
}

void run_predict_topn(rpk_ctx* c, int64_t U, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u, int N,
                      int mask_history, int32_t* out_idx_u, double* out_val_u, int32_t* out_len_u) {
  check_predict_args(c, U, nnz);
  RPK_REQUIRE(N >= 1 && N <= 2048, "N must be in [1, 2048]");
  RPK_REQUIRE(out_idx_u && out_len_u, "out_idx / out_len must not be null");
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)U + 1, "p_indptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "p_indices");
  Out<int32_t> o_idx, o_len;
  Out<double> o_val;
  o_idx.init(c, out_idx_u, (size_t)U * N, "p_out_idx");
  o_val.init(c, out_val_u, (size_t)U * N, "p_out_val");
  o_len.init(c, out_len_u, (size_t)U, "p_out_len");
  if (U > 0) {
    PredGeom g = predict_geometry(c, N);
    ensure_segments(c, g);
    PredParams pp;
    fill_common(c, pp, g, U, indptr, indices, N, mask_history, PRED_TOPN);
    pp.part_idx = c->buf<int>("p_part_idx", (size_t)U * g.P * N);
    pp.part_sq = c->buf<u64>("p_part_sq", (size_t)U * g.P * N);
    pp.part_len = c->buf<int>("p_part_len", (size_t)U * g.P);
    const int fgrid = (int)std::min<int64_t>((U * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    // the block table of the 32-bit kernel indexes 4-entry blocks with 32 bits
    const bool blocks_fit = (c->m_nnz + 3 * c->m_I) / 4 < (int64_t)0x7fffffff;
    if ((c->flags & DBG_WIDE_ACC) || !blocks_fit) {  // debug flag / giant models: every user through the two-limb kernel
      launch_predict(c, pp, g, U, indptr);
      k_predict_finalize<<<fgrid, 256, 0, st>>>(pp.part_idx, pp.part_sq, pp.part_len, U, g.P, N, o_idx.dev, o_val.dev,
                                                o_len.dev, nullptr, nullptr, nullptr, nullptr, 0);
      RPK_LAUNCH_CHECK(c);
    } else {
      const Pred32Geom g2 = predict32_geometry(c, N);
      Pred32Params qp;
      qp.indices = indices;
      ensure_blocks(c, g2.P, g2.R);
      qp.ent4 = reinterpret_cast<const uint4*>(c->get<u64>("m_ent4"));
      qp.blk = c->get<int2>("m_blk");
      qp.U = (int)U;
      qp.P = g2.P;
      qp.R = g2.R;
      qp.I = (int)c->m_I;
      qp.N = N;
      qp.mask = mask_history;
      qp.cap = g2.cap;
      qp.direct_cap = g2.direct_cap;
      qp.tcap = g2.tcap;
      qp.part_idx = c->buf<int>("p_part2_idx", (size_t)U * g2.P * N);
      qp.part_sq = c->buf<u64>("p_part2_sq", (size_t)U * g2.P * N);
      qp.part_len = c->buf<int>("p_part2_len", (size_t)U * g2.P);
      qp.ovf_flag = c->buf<int>("p_ovf_flag", (size_t)U + 1);
      qp.ovf_count = qp.ovf_flag + U;
      qp.ovf_tab = c->buf<int4>("p_ovf_tab", (size_t)U);
      RPK_CUDA(cudaMemsetAsync(qp.ovf_flag, 0, sizeof(int) * ((size_t)U + 1), st));
      const PredWork w = prepare_work(c, U, indptr);
      qp.work_tab = w.tab;
      qp.queue = w.queue + 1;
      qp.prof = nullptr;
#ifdef RPK_PHASE_PROF
      qp.prof = c->buf<unsigned long long>("p_prof32", 16);
      RPK_CUDA(cudaMemsetAsync(qp.prof, 0, 16 * sizeof(unsigned long long), st));
#endif
      RPK_CUDA(cudaFuncSetAttribute(k_predict_a32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g2.smem));
      int occ = 0;
      RPK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_predict_a32, g2.nt, g2.smem));
      RPK_REQUIRE(occ >= 1, "predict kernel does not fit on an SM");
      const int grid = (int)std::min<int64_t>(U * g2.P, (int64_t)c->sm_count * occ);
      c->ev_record(4);
      k_predict_a32<<<grid, g2.nt, g2.smem, st>>>(qp);
      RPK_LAUNCH_CHECK(c);
      // users whose survivors overflowed the list: exact two-limb kernel (normally none; the kernel then exits)
      pp.work_tab = qp.ovf_tab;
      pp.n_work = qp.ovf_count;
      pp.queue = w.queue;
      launch_predict_kernel(c, pp, g, U);
      c->ev_record(5);
      c->ev_valid[2] = true;
#ifdef RPK_PHASE_PROF
      {
        unsigned long long h[16];
        int novf = 0;
        RPK_CUDA(cudaMemcpyAsync(h, qp.prof, sizeof(h), cudaMemcpyDeviceToHost, st));
        RPK_CUDA(cudaMemcpyAsync(&novf, qp.ovf_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        RPK_CUDA(cudaStreamSynchronize(st));
        static const char* nm[7] = {"fetch", "sweep1", "mask", "select", "clean+mark", "sweep2", "sort+out"};
        unsigned long long tot = 0;
        for (int k = 0; k < 7; ++k) tot += h[k];
        fprintf(stderr, "[predict32 phases] grid=%d nt=%d P=%d R=%d tcap=%d smem=%zu occ=%d items=%llu sparse=%llu touched/item=%.0f survivors/item=%.1f overflow users=%d cycles/item=%.0f\n",
                grid, g2.nt, g2.P, g2.R, g2.tcap, g2.smem, occ, h[8], h[9], h[8] ? (double)h[10] / h[8] : 0.0,
                h[8] ? (double)h[11] / h[8] : 0.0, novf, h[8] ? (double)tot / h[8] : 0.0);
        for (int k = 0; k < 7; ++k)
          fprintf(stderr, "  %-10s %5.1f%%  %8.0f cycles/item\n", nm[k], 100.0 * h[k] / (tot ? tot : 1), h[8] ? (double)h[k] / h[8] : 0.0);
      }
#endif
      k_predict_finalize<<<fgrid, 256, 0, st>>>(qp.part_idx, qp.part_sq, qp.part_len, U, g2.P, N, o_idx.dev, o_val.dev,
                                                o_len.dev, qp.ovf_flag, pp.part_idx, pp.part_sq, pp.part_len, g.P);
      RPK_LAUNCH_CHECK(c);
    }
  }
  o_idx.finish(c);
  o_val.finish(c);
  o_len.finish(c);
  finish_call(c);
}

void run_predict_csr_count(rpk_ctx* c, int64_t U, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u,
                           int mask_history, int64_t* out_row_nnz_u) {
  check_predict_args(c, U, nnz);
  RPK_REQUIRE(out_row_nnz_u, "out_row_nnz must not be null");
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)U + 1, "p_indptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "p_indices");
  Out<int64_t> o;
  o.init(c, out_row_nnz_u, (size_t)U, "p_row_nnz");
  if (U > 0) {
    PredGeom g = predict_geometry(c, 1);
    ensure_segments(c, g);
    PredParams pp;
    fill_common(c, pp, g, U, indptr, indices, 1, mask_history, PRED_COUNT);
    pp.pass_cnt = c->buf<int64_t>("p_pass_cnt", (size_t)U * g.P);
    launch_predict(c, pp, g, U, indptr);
    k_sum_passes<<<ceil_div(U, 256), 256, 0, st>>>(pp.pass_cnt, U, g.P, o.dev);
    RPK_LAUNCH_CHECK(c);
    c->pc_U = U;
  }
  o.finish(c);
  finish_call(c);
}

void run_predict_csr_fill(rpk_ctx* c, int64_t U, int64_t nnz, const int64_t* indptr_u, const int32_t* indices_u,
                          int mask_history, const int64_t* out_indptr_u, int32_t* out_indices_u, double* out_values_u) {
  check_predict_args(c, U, nnz);
  RPK_REQUIRE(out_indptr_u && out_indices_u && out_values_u, "output pointers must not be null");
  if (U == 0) return;
  RPK_REQUIRE(c->pc_U == U, "rpk_predict_csr_fill must follow rpk_predict_csr_count on the same input");
  cudaStream_t st = c->stream;
  const int64_t* indptr = stage_in(c, indptr_u, (size_t)U + 1, "p_indptr");
  const int32_t* indices = stage_in(c, indices_u, (size_t)nnz, "p_indices");
  const int64_t* out_indptr = stage_in(c, out_indptr_u, (size_t)U + 1, "p_out_indptr");
  int64_t total = 0;
  if (is_device_ptr(out_indptr_u)) {
    RPK_CUDA(cudaMemcpyAsync(&total, out_indptr + U, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RPK_CUDA(cudaStreamSynchronize(st));
  } else {
    total = out_indptr_u[U];
  }
  Out<int32_t> o_i;
  Out<double> o_v;
  o_i.init(c, out_indices_u, (size_t)total, "p_csr_idx");
  o_v.init(c, out_values_u, (size_t)total, "p_csr_val");
  PredGeom g = predict_geometry(c, 1);
  ensure_segments(c, g);
  PredParams pp;
  fill_common(c, pp, g, U, indptr, indices, 1, mask_history, PRED_FILL);
  pp.pass_cnt = c->get<int64_t>("p_pass_cnt");
  pp.out_indptr = out_indptr;
  pp.out_indices = o_i.dev;
  pp.out_values = o_v.dev;
  launch_predict(c, pp, g, U, indptr);
  o_i.finish(c);
  o_v.finish(c);
  finish_call(c);
}

}  // namespace rpk
