// Scoring on the GPU: score_uj = sum_{i in hist(u)} S_ij for a K-sparse S, history masking and
// top-N fused in shared memory; optional full CSR output for drop-in predict().
//
// Replaces (reference, /root/reference):
//   recpack/algorithms/base.py:237-255     ItemSimilarityMatrixAlgorithm._predict  (X @ similarity_matrix_)
//   recpack/pipelines/pipeline.py:174-175  history removal
//   recpack/metrics/base.py:189            get_top_K_ranks(y_pred, K) on the prediction rows
//
// Scores are exact integers: every similarity value is stored as q = rint(v * 2^39) | 1 (40 bits), a
// score is the integer sum of q over the history.  The sum is accumulated in two 32-bit limbs
// (q & 0xFFFFF, q >> 20) with native shared-memory integer atomics (ATOMS.ADD); a 64-bit
// compare-and-swap accumulator is the fallback when a limb could overflow.  Integer sums make the
// result independent of the order in which the atomics land, so the top-N lists are deterministic.
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
                                                         int K, int I,
                                                         const int64_t* __restrict__ m_ptr, u64* __restrict__ m_ent,
                                                         unsigned* __restrict__ m_rowmax, int* __restrict__ flag) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64* buf = reinterpret_cast<u64*>(smem);
  __shared__ unsigned s_max;
  const int tid = threadIdx.x, nt = blockDim.x;
  int n2 = 2;
  while (n2 < K) n2 <<= 1;
  for (int i = blockIdx.x; i < I; i += gridDim.x) {
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
    const int64_t base = m_ptr[i];
    for (int t = tid; t < m; t += nt) {
      m_ent[base + t] = buf[t];
      if (t > 0 && (buf[t] >> 40) == (buf[t - 1] >> 40)) atomicOr(flag, 2);  // duplicate column
    }
    if (tid == 0) m_rowmax[i] = s_max;
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

__global__ void k_gather_len(const int* __restrict__ len, const int64_t* __restrict__ row_src, int64_t I, int* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < I) out[i] = len[row_src[i]];
}

static void model_common_begin(rpk_ctx* c, int64_t I) {
  RPK_REQUIRE(I >= 0 && I < ((int64_t)1 << 24), "item count must be below 2^24");
  c->m_I = I;
  c->m_P = 0;  // segment table must be rebuilt
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
    k_model_from_topk<<<grid, 128, (size_t)n2 * sizeof(u64), st>>>(idx, val, len, row_src, K, (int)I, m_ptr, m_ent, m_rowmax,
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
};

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
  const int total = p.U * p.P;
  for (int s = tid; s < p.R; s += nt) acc64[s] = 0ull;
  if (tid == 0) s_next = atomicAdd(p.queue, 1);
  for (;;) {
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
    if (p.mode == PRED_TOPN) {
      const int m = block_select_topk(src, p.N, list, p.cap, p.direct_cap, hist, sh);
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
  }
}

// One warp per user: merge the P per-range lists (each best-first) into the final top-N.
__global__ void k_predict_finalize(const int* __restrict__ part_idx, const u64* __restrict__ part_sq,
                                   const int* __restrict__ part_len, int64_t U, int P, int N, int* __restrict__ out_idx,
                                   double* __restrict__ out_val, int* __restrict__ out_len) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int PN = P * N;
  for (int64_t u = warp; u < U; u += nwarps) {
    const int* pi = part_idx + u * PN;
    const u64* ps = part_sq + u * PN;
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

static void launch_predict(rpk_ctx* c, PredParams& pp, const PredGeom& g, int64_t U, const int64_t* indptr) {
  cudaStream_t st = c->stream;
  // heaviest users first
  u64* work = c->buf<u64>("p_work", (size_t)U);
  int* order = c->buf<int>("p_order", (size_t)U);
  int* bcnt = c->buf<int>("p_bcnt", 65 * 2 + 2);
  int* boff = bcnt + 65;
  int* queue = boff + 65;
  RPK_CUDA(cudaMemsetAsync(bcnt, 0, sizeof(int) * (65 * 2 + 2), st));
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
  pp.work_tab = tab;
  pp.queue = queue;
  RPK_CUDA(cudaFuncSetAttribute(k_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  int occ = 0;
  RPK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_predict, g.nt, g.smem));
  RPK_REQUIRE(occ >= 1, "predict kernel does not fit on an SM");
  const int64_t total = U * g.P;
  const int grid = (int)std::min<int64_t>(total, (int64_t)c->sm_count * occ);
  c->ev_record(4);
  k_predict<<<grid, g.nt, g.smem, st>>>(pp);
  RPK_LAUNCH_CHECK(c);
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
    launch_predict(c, pp, g, U, indptr);
    const int grid = (int)std::min<int64_t>((U * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    k_predict_finalize<<<grid, 256, 0, st>>>(pp.part_idx, pp.part_sq, pp.part_len, U, g.P, N, o_idx.dev, o_val.dev, o_len.dev);
    RPK_LAUNCH_CHECK(c);
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
