// Scoring on the GPU: score_uj = sum_{i in hist(u)} S_ij for a K-sparse S, history masking and
// top-N fused in shared memory; optional full CSR output for drop-in predict().
//
// Replaces (reference, /root/reference):
//   recpack/algorithms/base.py:237-255     ItemSimilarityMatrixAlgorithm._predict  (X @ similarity_matrix_)
//   recpack/pipelines/pipeline.py:174-175  history removal
//   recpack/metrics/base.py:189            get_top_K_ranks(y_pred, K) on the prediction rows
//
// Scores are exact integers: every similarity value is stored as q = max(rint(v * 2^e), 1) (40 bits), with the
// power of two 2^e chosen per model so that the largest similarity lands in [2^39, 2^40) -- the resolution is
// relative to the model's largest value (2^-40 of it), whatever its magnitude.  A score is the integer sum of q
// over the history, reported as sum * 2^-e.  Integer sums make the result independent of the order in which the
// atomics land, so the top-N lists are deterministic.  Two kernels:
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
// Scale of the loaded model: q = rint(v * 2^e), scores come back as sum * inv (inv = 2^-e).
struct ModelScale {
  double inv;
  int e;
  int pad;
};

__device__ __forceinline__ bool quantize(double v, int e, u64& q) {
  if (!(v >= 0.0) || !(v <= 1.7e308)) return false;  // negative, NaN or infinite
  const double sv = scalbn(v, e);                   // exact power-of-two scaling
  if (!(sv < 1099511627776.0)) return false;        // does not fit 40 bits (cannot happen with the model's own scale)
  long long r = __double2ll_rn(sv);                 // round half to even like np.rint
  q = r < 1 ? 1ull : (u64)r;                        // a stored entry never vanishes
  if (q > Q_MASK40) q = Q_MASK40;                   // a value within half a step of 2^40 rounds up to it: clamp
  return true;
}

// Largest value of the lists / CSR about to be loaded (bit pattern of a non-negative double orders like an integer).
__global__ void k_model_vmax_lists(const double* __restrict__ val, const int* __restrict__ len, const int64_t* __restrict__ row_src,
                                   int K, int64_t nrows, u64* __restrict__ vmax_bits, int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  u64 best = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nrows * K; t += stride) {
    const int64_t src = row_src ? row_src[t / K] : t / K;
    int m = len[src];
    m = m > K ? K : m;
    if ((int)(t % K) < m) {
      const double v = val[src * K + t % K];
      if (!(v >= 0.0) || !(v <= 1.7e308)) atomicOr(flag, 1);
      else best = max(best, (u64)__double_as_longlong(v));
    }
  }
  best = warp_max_u64(best);
  if ((threadIdx.x & 31) == 0 && best) atomicMax(vmax_bits, best);
}

__global__ void k_model_vmax_flat(const double* __restrict__ val, int64_t n, u64* __restrict__ vmax_bits, int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  u64 best = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const double v = val[t];
    if (!(v >= 0.0) || !(v <= 1.7e308)) atomicOr(flag, 1);
    else best = max(best, (u64)__double_as_longlong(v));
  }
  best = warp_max_u64(best);
  if ((threadIdx.x & 31) == 0 && best) atomicMax(vmax_bits, best);
}

// e = 39 - floor(log2(vmax)): vmax * 2^e lies in [2^39, 2^40).  An empty (or all-zero) model gets e = 39.
// (vmax_bits: bit pattern of a non-negative double, which is also how a device-resident double is passed in.)
__global__ void k_model_scale(const u64* __restrict__ vmax_bits, ModelScale* __restrict__ sc, int forced_e, int use_forced) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int e = 39;
    if (use_forced) {
      e = forced_e;
    } else {
      const double vmax = __longlong_as_double((long long)vmax_bits[0]);
      if (vmax > 0.0) e = 39 - ilogb(vmax);
    }
    sc->e = e;
    sc->inv = scalbn(1.0, -e);
    sc->pad = 0;
  }
}

// One CTA per row: pack (idx, q), sort by idx, write the row.
__global__ void __launch_bounds__(256) k_model_from_topk(const int* __restrict__ idx, const double* __restrict__ val,
                                                         const int* __restrict__ len, const int64_t* __restrict__ row_src,
                                                         int K, int I, int nrows,
                                                         const int64_t* __restrict__ m_ptr, u64* __restrict__ m_ent,
                                                         unsigned* __restrict__ m_rowmax, int* __restrict__ flag,
                                                         const ModelScale* __restrict__ scale) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64* buf = reinterpret_cast<u64*>(smem);
  __shared__ unsigned s_max;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int sc_e = scale->e;
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
        bool ok = quantize(val[src * K + t], sc_e, q);
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
                                 unsigned* __restrict__ m_rowmax, int* __restrict__ m_len, int* __restrict__ flag,
                                 const ModelScale* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int sc_e = scale->e;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < I; i += nwarps) {
    int64_t b = indptr[i], e = indptr[i + 1];
    unsigned lmax = 0;
    for (int64_t k = b + lane; k < e; k += 32) {
      int j = indices[k];
      u64 q = 1;
      bool ok = quantize(values[k], sc_e, q);
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

// ---- block layout for the 32-bit scoring kernel, built per geometry (P item ranges of R items).  The entries of
// every (row, item range) segment are copied out on their own, padded to a multiple of 4 entries (32 bytes, one
// sector), and re-packed as (q << 24) | 4 * slot with slot = column - p*R, the accumulator the kernel adds to (as a
// byte offset): one AND and one shift per entry, no range test -- which bounds R to 2^22 slots, far beyond what
// shared memory holds.  Padding entries have q = 0 and slot = R + (position in the segment & 31):
// they land in 32 scratch accumulators behind the range, a different bank for every lane of the warp that reads them.
// seg[t] = {offset of the segment inside its model row, entries}, t = row * P + range
__global__ void k_model_seg_len(const int64_t* __restrict__ m_ptr, const u64* __restrict__ m_ent, int64_t I, int P, int R,
                                int2* __restrict__ seg, int* __restrict__ len4) {
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
  seg[t] = make_int2((int)s0, (int)(s1 - s0));
  len4[t] = (int)((s1 - s0 + 3) & ~(int64_t)3);
}

// One warp per segment: copy + re-pack + pad; blk[t] = {first 4-entry block, number of blocks, ceil(largest q / 2^20), 0}.
// The third field bounds every term of the segment: the kernel sums it over a user's rows to choose the shift of
// its 32-bit sums.
__global__ void k_model_pad(const int64_t* __restrict__ m_ptr, const u64* __restrict__ m_ent, const int2* __restrict__ seg,
                            const int64_t* __restrict__ ptr4, int64_t I, int P, int R, u64* __restrict__ ent4,
                            int4* __restrict__ blk) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < I * P; t += nwarps) {
    const int64_t i = t / P;
    const int p = (int)(t % P);
    const int2 sg = seg[t];
    const int64_t src = m_ptr[i] + sg.x, o = ptr4[t], n4 = ptr4[t + 1] - o;
    u64 qmax = 0;
    for (int64_t k = lane; k < n4; k += 32) {
      u64 v = (u64)(unsigned)(R + (int)(k & 31)) << 2;  // padding: q = 0, a scratch slot of this lane's own
      if (k < sg.y) {
        const u64 e = m_ent[src + k];
        const u64 q = e & Q_MASK40;
        v = (q << 24) | ((u64)((unsigned)(e >> 40) - (unsigned)p * (unsigned)R) << 2);
        qmax = max(qmax, q);
      }
      ent4[o + k] = v;
    }
    qmax = warp_max_u64(qmax);
    if (lane == 0) blk[t] = make_int4((int)(o >> 2), (int)(n4 >> 2), (int)(unsigned)((qmax + 0xfffffull) >> 20), 0);
  }
}

__global__ void k_gather_len(const int* __restrict__ len, const int64_t* __restrict__ row_src, int64_t I, int* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < I) out[i] = len[row_src[i]];
}

static void model_common_begin(rpk_ctx* c, int64_t I) {
  c->mark("model: begin");
  RPK_REQUIRE(I >= 0 && I < ((int64_t)1 << 24), "item count must be below 2^24");
  c->m_I = I;
  c->m_P = 0;  // segment tables must be rebuilt
  c->m_P2 = 0;
  c->m_pad = false;
  int* flag = c->buf<int>("m_flag", 1);
  RPK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  RPK_CUDA(cudaMemsetAsync(c->buf<u64>("m_vmax", 1), 0, sizeof(u64), c->stream));
  c->buf<ModelScale>("m_scale", 1);
}

// Scale of the model about to be loaded: from the largest value (forced < 0 ... use_forced = false) or as given.
static const ModelScale* model_set_scale(rpk_ctx* c, bool use_forced, int forced_e) {
  ModelScale* sc = c->get<ModelScale>("m_scale");
  k_model_scale<<<1, 32, 0, c->stream>>>(c->get<u64>("m_vmax"), sc, forced_e, use_forced ? 1 : 0);
  RPK_LAUNCH_CHECK(c);
  return sc;
}

static void model_check_flag(rpk_ctx* c) {
  int h = 0, hp = 0;
  ModelScale hs;
  if (c->pack_flag_pending) RPK_CUDA(cudaMemcpyAsync(&hp, c->get<int>("pk_flag"), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RPK_CUDA(cudaMemcpyAsync(&h, c->get<int>("m_flag"), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RPK_CUDA(cudaMemcpyAsync(&hs, c->get<ModelScale>("m_scale"), sizeof(ModelScale), cudaMemcpyDeviceToHost, c->stream));
  RPK_CUDA(cudaStreamSynchronize(c->stream));
  c->m_exp = hs.e;
  c->mark("model: loaded (sync)");
  if (c->pack_flag_pending) {
    c->pack_flag_pending = false;
    h |= hp;  // what the (stream-ordered) packing of this rank's rows found
  }
  if (h & 1) {
    c->m_I = 0;
    throw Error("similarity model: values must be finite and non-negative, columns in [0, I)");
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
  scan_i32_i64(c, m_len, m_ptr, I);
  u64* m_ent = c->buf<u64>("m_ent", (size_t)I * K);
  unsigned* m_rowmax = c->buf<unsigned>("m_rowmax", (size_t)I);
  if (I > 0) {
    k_model_vmax_lists<<<(int)std::min<int64_t>(ceil_div(I * K, 256), (int64_t)c->sm_count * 16), 256, 0, st>>>(
        val, len, row_src, K, I, c->get<u64>("m_vmax"), c->get<int>("m_flag"));
    RPK_LAUNCH_CHECK(c);
  }
  const ModelScale* scale = model_set_scale(c, false, 0);
  if (I > 0) {
    int n2 = 2;
    while (n2 < K) n2 <<= 1;
    const int grid = (int)std::min<int64_t>(I, (int64_t)c->sm_count * 16);
    k_model_from_topk<<<grid, 128, (size_t)n2 * sizeof(u64), st>>>(idx, val, len, row_src, K, (int)I, (int)I, m_ptr, m_ent, m_rowmax,
                                                                  c->get<int>("m_flag"), scale);
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
      if (col >= (u64)I || q == 0ull) atomicOr(flag, 1);
      if (t > 0 && (row[t - 1] >> 40) >= col) atomicOr(flag, 2);
      m_ent[base + t] = e;
      lmax = max(lmax, (unsigned)(q >> LIMB_BITS) + 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) m_rowmax[i] = lmax;
  }
}

// Exponent e of the scale 2^e these lists would get as a model of their own (39 - floor(log2(largest value))).
void run_model_scale_exp(rpk_ctx* c, int K, int64_t rows, const double* val_u, const int32_t* len_u, int32_t* out_exp) {
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(rows >= 0 && out_exp, "bad arguments");
  cudaStream_t st = c->stream;
  const double* val = stage_in(c, val_u, (size_t)rows * K, "pk_in_val");
  const int32_t* len = stage_in(c, len_u, (size_t)rows, "pk_in_len");
  u64* vmax = c->buf<u64>("pk_vmax", 1);
  int* flag = c->buf<int>("pk_flag", 1);
  ModelScale* sc = c->buf<ModelScale>("pk_scale", 1);
  RPK_CUDA(cudaMemsetAsync(vmax, 0, sizeof(u64), st));
  RPK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (rows > 0) {
    k_model_vmax_lists<<<(int)std::min<int64_t>(ceil_div(rows * K, 256), (int64_t)c->sm_count * 16), 256, 0, st>>>(val, len, nullptr, K, rows,
                                                                                                                  vmax, flag);
    RPK_LAUNCH_CHECK(c);
  }
  k_model_scale<<<1, 32, 0, st>>>(vmax, sc, 0, 0);
  RPK_LAUNCH_CHECK(c);
  ModelScale hs;
  int h = 0;
  RPK_CUDA(cudaMemcpyAsync(&hs, sc, sizeof(ModelScale), cudaMemcpyDeviceToHost, st));
  RPK_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  RPK_CUDA(cudaStreamSynchronize(st));
  RPK_REQUIRE(!(h & 1), "similarity lists: values must be finite and non-negative");
  *out_exp = hs.e;
}

// Largest value of the lists as a double in device (or host) memory -- the stream-ordered form of
// rpk_model_scale_exp: no synchronisation when out_vmax is device memory.
void run_model_vmax(rpk_ctx* c, int K, int64_t rows, const double* val_u, const int32_t* len_u, double* out_vmax_u) {
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(rows >= 0 && out_vmax_u, "bad arguments");
  cudaStream_t st = c->stream;
  const double* val = stage_in(c, val_u, (size_t)rows * K, "pk_in_val");
  const int32_t* len = stage_in(c, len_u, (size_t)rows, "pk_in_len");
  Out<double> o;
  o.init(c, out_vmax_u, 1, "pk_vmax_out");
  int* flag = c->buf<int>("pk_flag", 1);
  RPK_CUDA(cudaMemsetAsync(o.dev, 0, sizeof(double), st));
  RPK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (rows > 0) {
    k_model_vmax_lists<<<(int)std::min<int64_t>(ceil_div(rows * K, 256), (int64_t)c->sm_count * 16), 256, 0, st>>>(
        val, len, nullptr, K, rows, reinterpret_cast<u64*>(o.dev), flag);
    RPK_LAUNCH_CHECK(c);
  }
  o.finish(c);
  finish_call(c);
}

void run_model_pack_rows(rpk_ctx* c, int64_t I, int K, int64_t rows, const int32_t* idx_u, const double* val_u,
                         const int32_t* len_u, int scale_exp, const double* vmax_u, uint64_t* out_u) {
  c->mark("pack: begin");
  RPK_REQUIRE(vmax_u || (scale_exp > -1000 && scale_exp < 1100), "scale exponent out of range");
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
  ModelScale* sc = c->buf<ModelScale>("pk_scale", 1);
  const bool deferred = vmax_u != nullptr && is_device_ptr(vmax_u) && !o.host;  // everything stays on the stream
  if (vmax_u) {
    const double* vm = stage_in(c, vmax_u, 1, "pk_vmax_in");
    k_model_scale<<<1, 32, 0, st>>>(reinterpret_cast<const u64*>(vm), sc, 0, 0);
  } else {
    k_model_scale<<<1, 32, 0, st>>>(nullptr, sc, scale_exp, 1);
  }
  RPK_LAUNCH_CHECK(c);
  if (rows > 0) {
    int n2 = 2;
    while (n2 < K) n2 <<= 1;
    const int grid = (int)std::min<int64_t>(rows, (int64_t)c->sm_count * 16);
    k_model_from_topk<<<grid, 128, (size_t)n2 * sizeof(u64), st>>>(idx, val, len, nullptr, K, (int)I, (int)rows, nullptr, o.dev, nullptr, flag, sc);
    RPK_LAUNCH_CHECK(c);
  }
  if (deferred) {
    c->pack_flag_pending = true;  // checked together with the flags of the next model load (no synchronisation here)
    c->mark("pack: rows packed");
    return;
  }
  int h = 0;
  RPK_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  RPK_CUDA(cudaStreamSynchronize(st));
  RPK_REQUIRE(!(h & 1), "similarity lists: values must be finite, non-negative and below 2^(40 - scale_exp); columns in [0, I)");
  RPK_REQUIRE(!(h & 2), "similarity lists: column indices must be unique within a row");
  c->mark("pack: rows packed (sync)");
  o.finish(c);
  finish_call(c);
}

void run_model_load_packed_rows(rpk_ctx* c, int64_t I, int K, int64_t rows_in, const uint64_t* ent_u, const int32_t* len_u,
                                const int64_t* row_src_u, int scale_exp, const double* vmax_u) {
  RPK_REQUIRE(K >= 1 && K <= 4096, "K must be in [1, 4096]");
  RPK_REQUIRE(rows_in >= I || row_src_u, "fewer input rows than items");
  RPK_REQUIRE(vmax_u || (scale_exp > -1000 && scale_exp < 1100), "scale exponent out of range");
  model_common_begin(c, I);
  if (vmax_u) {
    const double* vm = stage_in(c, vmax_u, 1, "pk_vmax_in");
    k_model_scale<<<1, 32, 0, c->stream>>>(reinterpret_cast<const u64*>(vm), c->get<ModelScale>("m_scale"), 0, 0);
    RPK_LAUNCH_CHECK(c);
  } else {
    model_set_scale(c, true, scale_exp);
  }
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
  scan_i32_i64(c, m_len, m_ptr, I);
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
  if (nnz > 0) {
    k_model_vmax_flat<<<(int)std::min<int64_t>(ceil_div(nnz, 256), (int64_t)c->sm_count * 16), 256, 0, st>>>(
        values, nnz, c->get<u64>("m_vmax"), c->get<int>("m_flag"));
    RPK_LAUNCH_CHECK(c);
  }
  const ModelScale* scale = model_set_scale(c, false, 0);
  if (I > 0) {
    const int grid = (int)std::min<int64_t>((I * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    k_model_from_csr<<<grid, 256, 0, st>>>(indptr, indices, values, I, m_ent, m_rowmax, m_len, c->get<int>("m_flag"), scale);
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
  __device__ __forceinline__ bool has_queue() const { return false; }
  template <class F>
  __device__ __forceinline__ bool for_each_queued(F, int*, int, int*) const { return false; }
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
  const ModelScale* scale;   // scores are reported as sum * scale->inv
  const unsigned char* item_ok;  // non-null: item filter, 1 = may be recommended
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
    if (p.item_ok) {  // postprocessing/filters.py:58-101: filtered items are never recommended
      for (int s = tid; s < ns; s += nt)
        if (!p.item_ok[r0 + s]) {
          if (wide) acc64[s] = 0ull;
          else {
            acc_lo[s] = 0u;
            acc_hi[s] = 0u;
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
      const double inv_scale = p.scale->inv;
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
          p.out_values[out + pos] = (double)sc * inv_scale;
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
// Scoring kernel, top-N mode: 32-bit approximate sums, exact scores only where the order needs them
// ------------------------------------------------------------------------------------------
// The two-limb kernel above pays two shared-memory atomics per similarity entry and 8 bytes per item slot.
// For top-N only a handful of scores per user matter, so this kernel works on approximations:
//   1. sweep: a_j += (q >> s) | 1 with ONE fire-and-forget 32-bit atomic per entry.  s is the smallest shift that
//      keeps every sum below 2^32, taken from a bound on the user's terms (sum over the history rows of the largest
//      q of the row's segment).  Each term is within 1 of q / 2^s, so |a_j - score_j / 2^s| <= d (d = history
//      length): a_x - a_y > 2d proves score_x > score_y.
//   2. history items are zeroed, then two dense passes over the accumulators (16-byte loads, untouched vectors are
//      skipped as a whole): a 2048-bin histogram finds the bin of the N-th largest sum, the second pass copies
//      every item within 2d of that bin's lower edge to the survivor list and clears the accumulators.
//   3. when the caller wants the lists only (no scores) and the survivors' sums are all more than 2d apart down to
//      the (N+1)-th, their order is already proven: the sums themselves are written out.  Otherwise a second sweep
//      over the rows adds the exact q of the survivors (two 20-bit limbs, native atomics) and they are ordered by
//      their exact scores.
// k_predict_merge orders the P per-range lists of a user; a user whose lists cannot be ordered with certainty
// (approximate sums of different ranges too close) is scored again with exact sums.  4 bytes per slot let two CTAs
// share an SM, so one CTA's latency-bound steps hide behind the other's atomics.  A user whose survivors do not fit
// the list (huge groups of near-equal scores) is handed to the two-limb kernel.
struct ExactOrder {
  __device__ __forceinline__ int cmp3(const Entry& a, const Entry& b) const {
    if (a.key != b.key) return a.key > b.key ? 1 : -1;
    return 0;
  }
};

struct Pred32Params {
  const int* indices;
  const u64* ent;        // model rows in 4-entry blocks, entries (q << 24) | column
  const int4* blk;       // {first block, blocks, bound of q >> 20, 0} per (row, item range)
  const int4* work_tab;  // {user, history length, row start lo, hi} in processing order
  const int* n_work;     // non-null: number of work_tab records (device side), replaces U
  int U, P, R, I, N, mask;
  int exact;             // 1: exact scores for every list (scores requested / second pass over flagged users)
  int cap;
  int* queue;
  int* part_idx;         // [U*P x N]
  u64* part_key;         // exact sums; approximate sums where part_sft >= 0
  int* part_len;         // [U*P]
  int* part_sft;         // [U*P] -1: exact keys, s >= 0: keys are 32-bit sums of (q >> s) | 1
  const unsigned char* item_ok;  // non-null: item filter (1 = may be recommended), padded to a multiple of 4 items
  int* ovf_flag;         // per user: 1 = handed to the two-limb kernel
  int* ovf_count;        // number of such users
  int4* ovf_tab;         // their work records
  unsigned long long* prof;
};

constexpr int A32_BITS = 10;            // selection histogram: 1024 bins
constexpr int A32_BINS = 1 << A32_BITS;
constexpr int A32_ROWS = 512;           // history rows staged per chunk (= threads per CTA)
constexpr int A32_NT = 512;

// Row table of one chunk of the user's history: {first block, number of blocks} of every row's segment, and the
// history item itself (for the history mask).
struct RowTab {
  int2 seg[A32_ROWS];
  int item[A32_ROWS];
};

// Histogram bin of a 32-bit sum: 32 bins per octave (bit pattern of the sum as a float, top 13 bits), so that bins
// are ~2.2 % wide at every magnitude -- the bin of the N-th largest value holds few items whatever the scale of the
// scores.
__device__ __forceinline__ int a32_bin(unsigned a) {
  const int b = (int)(__float_as_uint((float)a) >> 18) - (127 << 5);
  return b > A32_BINS - 1 ? A32_BINS - 1 : b;
}
// A value that every sum of bin b and above certainly reaches: the lower edge of bin b - 1 (the float conversion
// rounds by at most 2^-24, far less than a bin).
__device__ __forceinline__ unsigned a32_bin_floor(int b) {
  if (b <= 1) return 1u;
  return (unsigned)__uint_as_float((unsigned)(b - 1 + (127 << 5)) << 18);
}

// Loads of two history rows in flight: the first 96 entries of each (a row segment rarely has more).
struct RowPair {
  u64 e1[3], e2[3];
};

__device__ __forceinline__ void pair_load(RowPair& rp, const u64* __restrict__ ent, const RowTab* rt, int r, int n, int lane, int nwarps,
                                          u64 idle) {
  const int r2 = r + nwarps;
  const int2 s1 = rt->seg[r];
  const int2 s2 = r2 < n ? rt->seg[r2] : make_int2(0, 0);
  const u64* b1 = ent + ((int64_t)s1.x << 2) + lane;
  const u64* b2 = ent + ((int64_t)s2.x << 2) + lane;
  const int n1 = (s1.y << 2) - lane, n2 = (s2.y << 2) - lane;  // entries left from this lane's first one
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    rp.e1[it] = 32 * it < n1 ? __ldg(b1 + 32 * it) : idle;
    rp.e2[it] = 32 * it < n2 ? __ldg(b2 + 32 * it) : idle;
  }
}

template <class F>
__device__ __forceinline__ void pair_apply(const RowPair& rp, F f) {
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    f(rp.e1[it]);
    f(rp.e2[it]);
  }
}

// Calls f(entry) for every entry of the staged segments, padding included, and for `idle` in the places of a load
// slot that a short segment leaves empty (idle = a padding entry: q = 0, a scratch slot of the lane's own).  One
// warp per row, a lane per entry, so that a load covers 256 contiguous bytes.  Software pipeline: the loads of the
// next two rows of the warp are issued before the entries of the current two are added, so that a warp always has
// loads in flight.  Segments longer than 96 entries (rare: K / P entries on average) get their tail in a second loop.
template <class F>
__device__ __forceinline__ void sweep_rows(const u64* __restrict__ ent, const RowTab* rt, int n, u64 idle, F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = A32_NT / 32;
  constexpr int step = 2 * nwarps;
  int r = warp;
  if (r >= n) return;
  RowPair A, B;
  pair_load(A, ent, rt, r, n, lane, nwarps, idle);
  for (;;) {
    r += step;
    const bool more_b = r < n;
    if (more_b) pair_load(B, ent, rt, r, n, lane, nwarps, idle);
    pair_apply(A, f);
    if (!more_b) break;
    r += step;
    const bool more_a = r < n;
    if (more_a) pair_load(A, ent, rt, r, n, lane, nwarps, idle);
    pair_apply(B, f);
    if (!more_a) break;
  }
  for (int q = warp; q < n; q += nwarps) {
    const int2 sg = rt->seg[q];
    if (sg.y > 24) {
      const u64* b = ent + ((int64_t)sg.x << 2);
      for (int e = 96 + lane; e < (sg.y << 2); e += 32) f(__ldg(b + e));
    }
  }
}

// m <= 32 entries: warp w ranks entries w, w + nwarps, ... by (key desc, index asc) -- lane f holds opponent f, one
// ballot counts the opponents that come first -- and writes them to their places in `dst`.  Barrier afterwards.
__device__ __forceinline__ void rank_by_warps(const Entry* src, Entry* dst, int m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  u64 kf = 0ull;
  int jf = SENTINEL_IDX;
  if (lane < m) {
    kf = src[lane].key;
    jf = src[lane].idx;
  }
  for (int i = warp; i < m; i += nwarps) {
    const u64 ki = __shfl_sync(0xffffffffu, kf, i);
    const int ji = __shfl_sync(0xffffffffu, jf, i);
    const unsigned first = __ballot_sync(0xffffffffu, lane < m && (kf > ki || (kf == ki && jf < ji)));
    if (lane == 0) {
      Entry e;
      e.key = ki;
      e.idx = ji;
      e.aux = 0;
      dst[__popc(first)] = e;
    }
  }
}

// One warp: `sorted` holds m <= 32 entries best first.  Optionally checks that the first min(m - 1, N) neighbouring
// keys are more than `margin` apart (the approximate sums then prove the order) and -- unless that check fails --
// writes the first N places of the list.  Returns false when the check failed.
__device__ __forceinline__ bool warp_check_write(const Entry* sorted, int m, int N, u64 margin, bool check, int* o_idx, u64* o_key) {
  const int lane = threadIdx.x & 31;
  if (check) {
    const int pairs = (m - 1) < N ? (m - 1) : N;
    bool bad = false;
    for (int k = lane; k < pairs; k += 32) bad |= sorted[k].key - sorted[k + 1].key <= margin;
    if (__any_sync(0xffffffffu, bad)) return false;
  }
  const int mo = m < N ? m : N;
  for (int t = lane; t < N; t += 32) {
    o_idx[t] = t < mo ? sorted[t].idx : -1;
    o_key[t] = t < mo ? sorted[t].key : 0ull;
  }
  return true;
}

// Orders m entries best first by (key desc, index asc).  Up to SEL_RANK_MAX entries are ranked by counting
// (nt / 64 threads per entry) from `src` into `other`; more are sorted in place.  Returns where the result is.
__device__ __forceinline__ Entry* a32_sort(Entry* src, Entry* other, int m) {
  ExactOrder ord;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (m <= SEL_RANK_MAX) {
    const int tpe = nt / SEL_RANK_MAX, i = tid / tpe, part = tid % tpe;
    int rank = 0;
    Entry a;
    a.key = 0;
    a.idx = 0;
    a.aux = 0;
    if (i < m) {
      a = src[i];
      for (int f = part; f < m; f += tpe) rank += (f != i) && entry_before(ord, src[f], a);
    }
    for (int o = tpe >> 1; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (i < m && part == 0) other[rank] = a;
    __syncthreads();
    return other;
  }
  int n2 = 1;
  while (n2 < m) n2 <<= 1;
  for (int i = m + tid; i < n2; i += nt) {
    Entry s;
    s.key = 0;
    s.idx = SENTINEL_IDX;
    s.aux = 0;
    src[i] = s;
  }
  __syncthreads();
  bitonic_sort_entries(ord, src, n2);
  return src;
}

#ifdef RPK_PHASE_PROF
#define PROF32_MARK(k) PROF_MARK(k)
#else
#define PROF32_MARK(k) do {} while (0)
#endif

__host__ __device__ __forceinline__ size_t a32_fixed_bytes(int cap) {
  return sel_smem_bytes(cap, A32_BINS) + 2 * ((sizeof(RowTab) + 15) / 16) * 16;
}

// Work distribution: the items are sorted heaviest first; CTA b starts with item b and takes every further item from a
// shared counter (longest-processing-time scheduling: the few users with five-digit histories then cost their CTAs
// other items instead of coming on top of an equal share of everything -- that tail bounded the kernel once a rank's
// shard became small).  The counter is read TWO items ahead by thread 0 and kept in a register, so the atomic's
// latency hides behind a whole item and the next item can still be staged while the current one is swept.
// Users with histories of at least A32_EXACT_D rows get their exact sums in the same work item: the margin 2 d of
// their approximate sums never proves a list, and a second launch for them would run on a handful of CTAs.
constexpr int A32_EXACT_D = 2048;

__global__ void __launch_bounds__(A32_NT, 2) k_predict_a32(Pred32Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Entry* list = reinterpret_cast<Entry*>(smem);
  Entry* list2 = list + p.cap;  // SEL_RANK_MAX entries
  int* hist = reinterpret_cast<int*>(smem + sel_list_bytes(p.cap));
  SelShared* sh = reinterpret_cast<SelShared*>(hist + A32_BINS);
  RowTab* rtab = reinterpret_cast<RowTab*>(smem + sel_smem_bytes(p.cap, A32_BINS));
  unsigned* acc = reinterpret_cast<unsigned*>(smem + a32_fixed_bytes(p.cap));
  __shared__ int s_flag3[3];  // "exact scores needed" of item k lives in s_flag3[k % 3]; reset two items ahead
  __shared__ int4 s_rec[2];
  __shared__ int s_w[2];
  __shared__ u64 s_bsum[2];       // bound of the sums, in units of 2^20: rows beyond the first chunk ...
  __shared__ unsigned s_bsum32[2];  // ... and the first chunk

  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int total = (p.n_work ? *p.n_work : p.U) * p.P;
  const int G = gridDim.x, bid = blockIdx.x;
  if (bid >= total) return;
  uint4* acc4 = reinterpret_cast<uint4*>(acc);
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc);
  const int nvec = p.R >> 2;  // R is a multiple of 4
  for (int v = tid; v < nvec; v += nt) acc4[v] = make_uint4(0u, 0u, 0u, 0u);
  for (int b = tid; b < A32_BINS; b += nt) hist[b] = 0;
  int pend = total;  // thread 0: the item after the next one (fetched one item ahead of its use)
  if (tid == 0) {
    const int w0 = bid;
    pend = G + atomicAdd(p.queue, 1);
    s_w[0] = w0;
    s_rec[0] = p.work_tab[w0 / p.P];
    s_bsum[0] = 0ull;
    s_bsum[1] = 0ull;
    s_bsum32[0] = 0u;
    s_bsum32[1] = 0u;
    s_flag3[0] = s_flag3[1] = s_flag3[2] = 0;
    sh->count = 0;
    sh->bstar = 0;
  }
  __syncthreads();
  // ---- staging of a work item, in three parts so that its two dependent global loads (history item -> block
  //      table entry) hide behind other work: part 1 issues the load of this thread's history item, part 2 the
  //      load of that row's table entry, part 3 writes the row table and adds up the bound of the sums
  int st_item = -1;
  int4 st_blk = make_int4(0, 0, 0, 0);
  auto stage1 = [&](int buf) {
    st_item = -1;
    if (s_w[buf] < total) {
      const int4 rc = s_rec[buf];
      const int64_t xb2 = ((int64_t)(unsigned)rc.z) | ((int64_t)rc.w << 32);
      if (tid < rc.y) st_item = p.indices[xb2 + tid];
    }
  };
  auto stage2 = [&](int buf) {
    st_blk = make_int4(0, 0, 0, 0);
    if (st_item >= 0) st_blk = __ldg(p.blk + (int64_t)st_item * p.P + (s_w[buf] % p.P));
  };
  auto stage3 = [&](int buf) {
    if (s_w[buf] >= total) return;
    const int4 rc = s_rec[buf];
    const int d2 = rc.y, pass2 = s_w[buf] % p.P;
    const int64_t xb2 = ((int64_t)(unsigned)rc.z) | ((int64_t)rc.w << 32);
    RowTab* rt2 = rtab + buf;
    unsigned b32 = 0;  // first chunk: at most 512 terms of at most 2^20 each
    if (tid < d2) {
      rt2->seg[tid] = make_int2(st_blk.x, st_blk.y);
      rt2->item[tid] = st_item;
      b32 = (unsigned)st_blk.z;
    }
    b32 = __reduce_add_sync(0xffffffffu, b32);
    if (lane == 0 && b32) atomicAdd(&s_bsum32[buf], b32);
    if (d2 > nt) {  // long histories: the rows beyond the first chunk only add to the bound
      u64 bsum = 0;
      for (int r = tid + nt; r < d2; r += nt) bsum += (u64)(unsigned)__ldg(p.blk + (int64_t)p.indices[xb2 + r] * p.P + pass2).z;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
      if (lane == 0 && bsum) atomicAdd(&s_bsum[buf], bsum);
    }
  };
  stage1(0);
  stage2(0);
  stage3(0);
  __syncthreads();
#ifdef RPK_PHASE_PROF
  long long t_prev = clock64();
#endif
  for (int k = 0;; ++k) {
    const int cur = k & 1, nxt = cur ^ 1;
    int& s_flag = s_flag3[k % 3];
    const int w = s_w[cur];
    if (w >= total) break;
    const int4 rec = s_rec[cur];
    const RowTab* rt = rtab + cur;
    // the next work item: its record is copied into shared memory asynchronously while this item is swept; the
    // per-item scratch is reset here too (every thread has left the previous item, nobody reads these before the
    // barrier after the sweep)
    if (tid == 0) {
      int w_n = pend;                    // asked for while the previous item was processed
      pend = G + atomicAdd(p.queue, 1);  // used at the top of the next item
      if (w_n < total) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_rec[nxt])),
                     "l"(p.work_tab + w_n / p.P)
                     : "memory");
      } else {
        w_n = total;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      s_w[nxt] = w_n;
      s_bsum[nxt] = 0ull;
      s_bsum32[nxt] = 0u;
      s_flag3[(k + 1) % 3] = 0;  // the next item's flag; the previous item's may still be read by slow threads
      sh->count = 0;
      sh->bstar = 0;
    }
    const int u = rec.x;
    const int pass = w % p.P;
    const int r0 = pass * p.R;
    const int ns = min(p.R, p.I - r0);
    const int64_t xb = ((int64_t)(unsigned)rec.z) | ((int64_t)rec.w << 32);
    const int d = rec.y;
    const int64_t slot_out = (int64_t)u * p.P + pass;
    int* o_idx = p.part_idx + slot_out * p.N;
    u64* o_key = p.part_key + slot_out * p.N;
    PROF32_MARK(0);
    // ---- shift of the 32-bit sums from the bound staged with the item
    // B = b20 * 2^20 >= sum over the rows of their largest q; the smallest shift with (B >> sft) + d < 2^32
    const u64 b20 = s_bsum[cur] + (u64)s_bsum32[cur];  // < 2^45 (d < 2^24 rows of at most 2^20 each)
    int sft = 0;
    {
      const u64 room = 0xffffffffull - (u64)d;
      auto fits = [&](int sf) { return sf >= 20 ? (b20 >> (sf - 20)) <= room : b20 <= (room >> (20 - sf)); };
      if (!fits(0)) {
        sft = 64 - __clzll((long long)b20) - 12;  // bits(B) - 32
        sft = sft < 1 ? 1 : sft;
        while (!fits(sft)) ++sft;
      }
    }
    const int sh_a = 24 + sft;
    const unsigned margin = 2u * (unsigned)d;
    const u64 idle = (u64)(unsigned)(p.R + lane) << 2;
    // ---- sweep 1: approximate sums, one fire-and-forget atomic per entry
    if (d > 0) {
      RowTab* rtw = rtab + cur;
      for (int c0 = 0; c0 < d; c0 += A32_ROWS) {
        const int n = min(A32_ROWS, d - c0);
        if (c0 > 0) {
          __syncthreads();  // the previous chunk's table is still being read
          for (int r = tid; r < n; r += nt) {
            const int i = p.indices[xb + c0 + r];
            const int4 b = __ldg(p.blk + (int64_t)i * p.P + pass);
            rtw->seg[r] = make_int2(b.x, b.y);
            rtw->item[r] = i;
          }
          __syncthreads();
        }
        // no test per entry: real entries carry their slot, padding lands in the scratch slots behind the range
        sweep_rows(p.ent, rt, n, idle, [&](u64 e) {
          asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(acc_s + ((unsigned)e & 0xffffffu)), "r"((unsigned)(e >> sh_a) | 1u) : "memory");
        });
      }
    }
    if (tid == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");  // the next item's record has landed
    __syncthreads();
    PROF32_MARK(1);
    stage1(nxt);
    if (d == 0) {  // user without history: empty prediction row (algorithms/base.py:123-127)
      for (int t = tid; t < p.N; t += nt) {
        o_idx[t] = -1;
        o_key[t] = 0;
      }
      if (tid == 0) {
        p.part_len[slot_out] = 0;
        p.part_sft[slot_out] = -1;
      }
      stage2(nxt);
      stage3(nxt);
      __syncthreads();
      continue;
    }
    if (p.mask) {  // pipelines/pipeline.py:174-175 -- before the truncation to N
      if (d <= A32_ROWS) {
        if (tid < d) {
          const int j = rt->item[tid] - r0;
          if (j >= 0 && j < ns) acc[j] = 0u;
        }
      } else {
        for (int r = tid; r < d; r += nt) {
          const int j = p.indices[xb + r] - r0;
          if (j >= 0 && j < ns) acc[j] = 0u;
        }
      }
      __syncthreads();
    }
    if (p.item_ok) {  // postprocessing/filters.py:58-101: filtered items are never recommended (4 items per word)
      const unsigned* ok4 = reinterpret_cast<const unsigned*>(p.item_ok + r0);
      for (int v = tid; v < nvec; v += nt) {
        const unsigned okw = 4 * v < ns ? __ldg(ok4 + v) : 0x01010101u;
        if (okw != 0x01010101u) {
          uint4 x = acc4[v];
          if (!(okw & 0xffu)) x.x = 0u;
          if (!(okw & 0xff00u)) x.y = 0u;
          if (!(okw & 0xff0000u)) x.z = 0u;
          if (!(okw & 0xff000000u)) x.w = 0u;
          acc4[v] = x;
        }
      }
      __syncthreads();
    }
    PROF32_MARK(2);
    // ---- survivors.  A dense pass costs a warp its full instruction stream whenever any of its lanes has work, so
    //      the per-item work is kept out of it: every thread takes the maximum of its own slots (no branches, no
    //      atomics), and the N-th largest of these per-thread maxima -- found with one histogram entry per thread,
    //      64 bins per octave -- is a lower bound of the N-th largest sum (N distinct items reach it).  Everything
    //      within the margin below that bin survives.
    unsigned mymax = 0u;
    for (int v = tid; v < nvec; v += nt) {
      const uint4 x = acc4[v];
      mymax = max(mymax, max(max(x.x, x.y), max(x.z, x.w)));
    }
    // N-th largest of the 512 per-thread maxima, in two rounds of a 32-bin histogram (octave, then 32 bins inside
    // the octave).  Every warp scans the 32 bins by itself (one bin per lane), so no warp waits for another one
    // beyond the two barriers.
    const int mybin = mymax ? a32_bin(mymax) : -1;  // (octave << 5) | bin inside the octave
    if (mybin >= 0) atomicAdd(&hist[mybin >> 5], 1);
    __syncthreads();
    PROF32_MARK(12);
    stage2(nxt);
    int bstar = 0;
    {
      const int h1 = hist[lane];
      int incl = h1;  // maxima in my octave and the higher ones
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, incl >= p.N);  // lanes whose octave-and-above hold N maxima
      if (hit) {
        const int oct = 31 - __clz(hit);  // the highest such octave holds the N-th largest
        const int above = __shfl_sync(0xffffffffu, incl - h1, oct);
        if (mybin >= 0 && (mybin >> 5) == oct) atomicAdd(&hist[32 + (mybin & 31)], 1);
        __syncthreads();
        const int h2 = hist[32 + lane];
        int incl2 = h2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_down_sync(0xffffffffu, incl2, o);
          if (lane + o < 32) incl2 += t;
        }
        const unsigned hit2 = __ballot_sync(0xffffffffu, above + incl2 >= p.N);
        bstar = (oct << 5) | (hit2 ? 31 - __clz(hit2) : 0);
      } else {
        __syncthreads();  // fewer than N non-empty threads: everything survives (bstar = 0)
      }
    }
    __syncthreads();  // every warp has read both histograms
    if (mybin >= 0) {
      hist[mybin >> 5] = 0;
      hist[32 + (mybin & 31)] = 0;
    }
    const unsigned edge = a32_bin_floor(bstar);
    const unsigned thr = edge > margin + 1u ? edge - margin : 1u;
    PROF32_MARK(13);
    // copy the survivors, clear the accumulators: one compare per vector, the rest only for the few that pass
    for (int v = tid; v < nvec; v += nt) {
      const uint4 x = acc4[v];
      acc4[v] = make_uint4(0u, 0u, 0u, 0u);
      if (max(max(x.x, x.y), max(x.z, x.w)) < thr) continue;
      const unsigned xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (xs[q] >= thr) {
          const int pos = atomicAdd(&sh->count, 1);
          if (pos < p.cap) {
            Entry e;
            e.key = (u64)xs[q];
            e.idx = r0 + 4 * v + q;
            e.aux = 0;
            list[pos] = e;
          }
        }
    }
    PROF32_MARK(14);
    stage3(nxt);
    __syncthreads();
    const int m = sh->count;
    PROF32_MARK(3);
#ifdef RPK_PHASE_PROF
    if (tid == 0) {
      atomicAdd(p.prof + 8, 1ull);
      atomicAdd(p.prof + 11, (unsigned long long)(m > p.cap ? 0 : m));
    }
#endif
    if (m > p.cap) {  // survivors do not fit: the exact kernel takes the whole user
      if (tid == 0 && atomicExch(&p.ovf_flag[u], 1) == 0) p.ovf_tab[atomicAdd(p.ovf_count, 1)] = rec;
      __syncthreads();
      continue;
    }
    const int mo = m < p.N ? m : p.N;
    bool need_exact = p.exact != 0 || d >= A32_EXACT_D;
    Entry* surv = list;  // where the survivors are
    if (!need_exact) {
      // is the order of the N best (and their separation from the rest) proven by the approximate sums alone?
      if (m <= 32) {
        // the usual case: every warp ranks two survivors (a lane per opponent), then one warp checks the gaps and
        // writes the list
        rank_by_warps(list, list2, m);
        __syncthreads();
        if (warp == 0) {
          const bool sure = warp_check_write(list2, m, p.N, (u64)margin, true, o_idx, o_key);
          if (lane == 0) {
            if (sure) {
              p.part_len[slot_out] = mo;
              p.part_sft[slot_out] = sft;
            } else {
              s_flag = 1;
            }
          }
        }
        __syncthreads();
        need_exact = s_flag != 0;
        if (!need_exact) {
          PROF32_MARK(6);
          continue;
        }
      } else {
        surv = a32_sort(list, list2, m);
        const int pairs = (m - 1) < p.N ? (m - 1) : p.N;
        for (int kk = tid; kk < pairs; kk += nt)
          if (surv[kk].key - surv[kk + 1].key <= (u64)margin) s_flag = 1;
        __syncthreads();
        need_exact = s_flag != 0;
        if (!need_exact) {
          for (int t = tid; t < p.N; t += nt) {
            o_idx[t] = t < mo ? surv[t].idx : -1;
            o_key[t] = t < mo ? surv[t].key : 0ull;
          }
          if (tid == 0) {
            p.part_len[slot_out] = mo;
            p.part_sft[slot_out] = sft;
          }
          __syncthreads();
          PROF32_MARK(6);
          continue;
        }
      }
    }
#ifdef RPK_PHASE_PROF
    if (tid == 0) atomicAdd(p.prof + 9, 1ull);
#endif
    // ---- sweep 2: exact scores of the survivors (slot -> survivor number + 1, key -> exact sum)
    for (int t = tid; t < m; t += nt) {
      acc[surv[t].idx - r0] = (unsigned)t + 1u;
      surv[t].key = 0ull;
    }
    __syncthreads();
    PROF32_MARK(4);
    const bool limbs = d <= LIMB_CHUNK;  // two 32-bit adds (20-bit limbs) cannot overflow
    if (m > 0) {
      RowTab* rtw = rtab + cur;
      for (int c0 = 0; c0 < d; c0 += A32_ROWS) {
        const int n = min(A32_ROWS, d - c0);
        if (d > A32_ROWS) {  // a single chunk is still staged from sweep 1
          if (c0 > 0) __syncthreads();
          for (int r = tid; r < n; r += nt) {
            const int i = p.indices[xb + c0 + r];
            const int4 b = __ldg(p.blk + (int64_t)i * p.P + pass);
            rtw->seg[r] = make_int2(b.x, b.y);
            rtw->item[r] = i;
          }
          __syncthreads();
        }
        sweep_rows(p.ent, rt, n, idle, [&](u64 e) {
          const unsigned j = ((unsigned)e & 0xffffffu) >> 2;
          if (j < (unsigned)ns) {
            const unsigned cn = acc[j];
            if (cn) {
              const u64 q = e >> 24;
              if (limbs) {
                unsigned* wd = reinterpret_cast<unsigned*>(&surv[cn - 1u].key);
                atomicAdd(wd, (unsigned)q & LIMB_MASK);
                atomicAdd(wd + 1, (unsigned)(q >> LIMB_BITS));
              } else {
                atomicAdd(&surv[cn - 1u].key, q);
              }
            }
          }
        });
      }
      __syncthreads();
    }
    PROF32_MARK(5);
    for (int t = tid; t < m; t += nt) {
      acc[surv[t].idx - r0] = 0u;
      if (limbs) {
        const u64 kv = surv[t].key;
        surv[t].key = ((kv >> 32) << LIMB_BITS) + (kv & 0xffffffffull);
      }
    }
    __syncthreads();
    if (m <= 32) {
      Entry* other = surv == list ? list2 : list;
      rank_by_warps(surv, other, m);
      __syncthreads();
      if (warp == 0) warp_check_write(other, m, p.N, 0ull, false, o_idx, o_key);
    } else {
      const Entry* fin = a32_sort(surv, surv == list ? list2 : list, m);
      for (int t = tid; t < p.N; t += nt) {
        o_idx[t] = t < mo ? fin[t].idx : -1;
        o_key[t] = t < mo ? fin[t].key : 0ull;
      }
    }
    if (tid == 0) {
      p.part_len[slot_out] = mo;
      p.part_sft[slot_out] = -1;
    }
    __syncthreads();
    PROF32_MARK(6);
  }
}

// One warp per user: merge the P per-range lists (each best-first, exact keys) into the final top-N.
// Users flagged in `alt_flag` take their lists from the second set (the two-limb kernel's, alt_P ranges).
// only_a / only_b non-null: only users flagged in either are processed (second pass).
__global__ void k_predict_finalize(const int* __restrict__ part_idx, const u64* __restrict__ part_sq,
                                   const int* __restrict__ part_len, int64_t U, int P, int N, int* __restrict__ out_idx,
                                   double* __restrict__ out_val, int* __restrict__ out_len,
                                   const int* __restrict__ alt_flag, const int* __restrict__ alt_idx,
                                   const u64* __restrict__ alt_sq, const int* __restrict__ alt_len, int alt_P,
                                   const int* __restrict__ only_a, const int* __restrict__ only_b,
                                   const ModelScale* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double inv = scale->inv;
  for (int64_t u = warp; u < U; u += nwarps) {
    if ((only_a || only_b) && !((only_a && only_a[u]) || (only_b && only_b[u]))) continue;
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
        if (out_val) out_val[u * N + rank] = (double)se * inv;
      }
    }
    for (int t = m + lane; t < N; t += 32) {
      out_idx[u * N + t] = -1;
      if (out_val) out_val[u * N + t] = 0.0;
    }
    if (lane == 0) out_len[u] = m;
  }
}

// Lists-only mode: the P per-range lists of a user carry exact sums (part_sft < 0) or approximate ones
// (sums of (q >> s) | 1, part_sft = s).  Every key is an interval of the exact score:
//   exact:        [E, E]
//   approximate:  [(a - d) * 2^s, (a + d) * 2^s]        (|a - E / 2^s| <= d, d = history length)
// One warp per user orders the (at most MERGE_MAX) entries by their upper ends and checks that down to the
// (N+1)-th every entry's interval lies strictly above the next one's (equal exact scores: ascending index).  Then
// the order is the exact order and the list is written; otherwise the user is queued for a second, exact pass.
constexpr int MERGE_MAX = 256;
constexpr int MERGE_WARPS = 4;
__global__ void __launch_bounds__(32 * MERGE_WARPS) k_predict_merge(const int* __restrict__ part_idx, const u64* __restrict__ part_key,
                                                                   const int* __restrict__ part_len, const int* __restrict__ part_sft,
                                                                   const int64_t* __restrict__ indptr, int64_t U, int P, int N,
                                                                   int* __restrict__ out_idx, int* __restrict__ out_len,
                                                                   const int* __restrict__ ovf_flag, int* __restrict__ redo_flag,
                                                                   int* __restrict__ redo_count, int4* __restrict__ redo_tab) {
  // per warp: the user's entries as they come (a*) and in order (s*)
  __shared__ u64 a_lo[MERGE_WARPS][MERGE_MAX], a_hi[MERGE_WARPS][MERGE_MAX];
  __shared__ u64 s_lo[MERGE_WARPS][MERGE_MAX], s_hi[MERGE_WARPS][MERGE_MAX];
  __shared__ int a_ix[MERGE_WARPS][MERGE_MAX], s_ix[MERGE_WARPS][MERGE_MAX];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  u64 *alo = a_lo[wib], *ahi = a_hi[wib], *lo = s_lo[wib], *hi = s_hi[wib];
  int *aix = a_ix[wib], *ix = s_ix[wib];
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = warp; u < U; u += nwarps) {
    if (ovf_flag[u]) continue;  // the two-limb kernel produces this user's lists
    const int64_t xb = indptr[u];
    const u64 d = (u64)(indptr[u + 1] - xb);
    // gather the valid entries of the P lists, densely: range q contributes its first part_len places
    int tot = 0;
    __syncwarp();
    for (int q = 0; q < P; ++q) {
      const int n = part_len[u * P + q];
      const int sf = part_sft[u * P + q];
      for (int t = lane; t < n; t += 32) {
        const u64 k = part_key[(u * P + q) * N + t];
        u64 l, h;
        if (sf < 0) {
          l = h = k;
        } else {
          l = (k > d ? k - d : 0ull) << sf;
          h = (k + d) > (~0ull >> sf) ? ~0ull : (k + d) << sf;
        }
        alo[tot + t] = l;
        ahi[tot + t] = h;
        aix[tot + t] = part_idx[(u * P + q) * N + t];
      }
      tot += n;
    }
    __syncwarp();
    // rank every entry against all of them: by upper end, then lower end, then index
    for (int e = lane; e < tot; e += 32) {
      const u64 l = alo[e], h = ahi[e];
      const int je = aix[e];
      int rank = 0;
      for (int f = 0; f < tot; ++f) {
        const u64 hf = ahi[f], lf = alo[f];
        rank += (hf > h) || (hf == h && (lf > l || (lf == l && aix[f] < je)));
      }
      lo[rank] = l;
      hi[rank] = h;
      ix[rank] = je;
    }
    __syncwarp();
    const int m = min(N, tot);
    const int pairs = min(tot - 1, N);
    bool bad = false;
    for (int k = lane; k < pairs; k += 32) {
      const bool exact_pair = lo[k] == hi[k] && lo[k + 1] == hi[k + 1];
      const bool ok = lo[k] > hi[k + 1] || (exact_pair && (hi[k] > hi[k + 1] || ix[k] < ix[k + 1]));
      bad |= !ok;
    }
    if (__any_sync(0xffffffffu, bad)) {
      if (lane == 0) {
        redo_flag[u] = 1;
        redo_tab[atomicAdd(redo_count, 1)] = make_int4((int)u, (int)d, (int)(xb & 0xffffffffll), (int)(xb >> 32));
      }
      continue;
    }
    for (int t = lane; t < N; t += 32) out_idx[u * N + t] = t < m ? ix[t] : -1;
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

// Block layout of the model for one geometry of the 32-bit kernel (rebuilt when the model or the geometry changes).
static void ensure_blocks(rpk_ctx* c, int P, int R) {
  if (c->m_pad && c->m_P2 == P && c->m_R2 == R) return;
  const int64_t I = c->m_I;
  cudaStream_t st = c->stream;
  const int64_t* m_ptr = c->get<int64_t>("m_ptr");
  const u64* m_ent = c->get<u64>("m_ent");
  const int64_t nseg = I * P;
  int2* seg = c->buf<int2>("m_seg2", (size_t)nseg);
  int* len4 = c->buf<int>("m_len4", (size_t)nseg);
  int64_t* ptr4 = c->buf<int64_t>("m_ptr4", (size_t)nseg + 1);
  int4* blk = c->buf<int4>("m_blk", (size_t)nseg);
  // every segment is padded by at most 3 entries
  u64* ent4 = c->buf<u64>("m_ent4", (size_t)c->m_nnz + 3 * (size_t)nseg + 4);
  if (nseg > 0) {
    k_model_seg_len<<<ceil_div(nseg, 256), 256, 0, st>>>(m_ptr, m_ent, I, P, R, seg, len4);
    RPK_LAUNCH_CHECK(c);
  }
  scan_i32_i64(c, len4, ptr4, nseg);
  if (nseg > 0) {
    k_model_pad<<<(int)std::min<int64_t>(ceil_div(nseg * 32, 256), (int64_t)c->sm_count * 32), 256, 0, st>>>(m_ptr, m_ent, seg, ptr4, I, P, R,
                                                                                                          ent4, blk);
    RPK_LAUNCH_CHECK(c);
  }
  c->m_pad = true;
  c->m_P2 = P;
  c->m_R2 = R;
}

// Geometry of the 32-bit kernel: two CTAs per SM when that costs at most one more item range than one CTA
// per SM would need (or at most three ranges), else one.
struct Pred32Geom {
  int cap, P, R, nt, ctas;
  size_t smem;
};

static Pred32Geom predict32_geometry(rpk_ctx* c, int N) {
  Pred32Geom g;
  const bool tiny = c->flags & DBG_TINY_LIST;
  g.cap = std::max(tiny ? 64 : 256, next_pow2(2 * std::max(N, 1)));
  const size_t fixed = a32_fixed_bytes(g.cap);
  const int64_t I = c->m_I;
  auto solve = [&](int ctas, int& P, int64_t& R) -> bool {
    const size_t per_cta = std::min<size_t>((size_t)c->smem_max, (size_t)c->smem_per_sm / ctas - 1024);
    if (per_cta < fixed + 1024 + 4096) return false;
    const size_t avail = per_cta - fixed - 256;
    for (P = 1;; ++P) {
      R = ((I + P - 1) / P + 3) & ~(int64_t)3;
      if (R < 4) R = 4;
      if ((size_t)R * 4 + 128 <= avail) return true;  // + 32 scratch slots for padding entries
      if (R <= 4) return false;
    }
  };
  int P1 = 0, P2 = 0;
  int64_t R1 = 0, R2 = 0;
  const bool ok1 = solve(1, P1, R1);
  const bool ok2 = solve(2, P2, R2);
  RPK_REQUIRE(ok1, "N too large for shared memory");
  bool two = ok2 && (P2 <= 3 || P2 <= P1 + 1);
  if (const char* e = getenv("RPK_PRED_CTAS")) {  // tuning hook
    const int v = atoi(e);
    if (v == 1) two = false;
    if (v == 2 && ok2) two = true;
  }
  g.ctas = two ? 2 : 1;
  g.P = two ? P2 : P1;
  int64_t R = two ? R2 : R1;
  if ((c->flags & DBG_MULTI_PASS) && g.P < 2 && I >= 8) {
    g.P = 2;
    R = ((I + g.P - 1) / g.P + 3) & ~(int64_t)3;
  }
  g.R = (int)R;
  g.smem = fixed + (size_t)R * 4 + 128;
  g.nt = A32_NT;  // the kernel stages one history row per thread
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
  int* queue = boff + 65;  // [0] two-limb kernel, [1] 32-bit kernel, [2] its second (exact) pass
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

static const unsigned char* item_filter_ptr(rpk_ctx* c);

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
  pp.scale = c->get<ModelScale>("m_scale");
  pp.item_ok = item_filter_ptr(c);
}

// Item filter of the predict calls: a copy of the caller's mask, normalised to 0 / 1 bytes and padded (with 1 = allowed)
// so that the kernels can read it four items at a time up to the end of the last item range.
__global__ void k_filter_copy(const unsigned char* __restrict__ in, int64_t I, int64_t padded, unsigned char* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < padded) out[j] = j < I ? (in[j] ? 1 : 0) : 1;
}

void run_predict_item_filter(rpk_ctx* c, const uint8_t* allowed_u, int64_t I) {
  if (!allowed_u) {
    c->filter_I = -1;
    return;
  }
  RPK_REQUIRE(I >= 0, "negative item count");
  const unsigned char* in = stage_in(c, allowed_u, (size_t)I, "p_item_ok_in");
  const int64_t padded = I + 4096;
  unsigned char* ok = c->buf<unsigned char>("p_item_ok", (size_t)padded);
  k_filter_copy<<<ceil_div(padded, 256), 256, 0, c->stream>>>(in, I, padded, ok);
  RPK_LAUNCH_CHECK(c);
  c->filter_I = I;
  if (!is_device_ptr(allowed_u)) RPK_CUDA(cudaStreamSynchronize(c->stream));  // the caller may reuse its buffer
}

static const unsigned char* item_filter_ptr(rpk_ctx* c) {
  if (c->filter_I < 0) return nullptr;
  RPK_REQUIRE(c->filter_I == c->m_I, "the item filter was set for a different number of items than the model has");
  return c->get<unsigned char>("p_item_ok");
}

static void check_predict_args(rpk_ctx* c, int64_t U, int64_t nnz) {
  RPK_REQUIRE(c->m_I > 0 || c->m_nnz == 0, "no similarity model loaded");
  RPK_REQUIRE(c->bufs.count("m_ptr") != 0, "no similarity model loaded (call rpk_model_load_* first)");
  RPK_REQUIRE(U >= 0 && nnz >= 0, "negative dimension");
  RPK_REQUIRE(U < ((int64_t)1 << 27), "too many users in one call; split the batch");
  RPK_REQUIRE(c->bufs.count("m_scale") != 0, "no similarity model loaded (call rpk_model_load_* first)");
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
    c->mark("predict: begin");
    PredGeom g = predict_geometry(c, N);
    RPK_REQUIRE(U * (int64_t)g.P < ((int64_t)1 << 31), "too many (user, item range) work items in one call; split the batch");
    ensure_segments(c, g);
    PredParams pp;
    fill_common(c, pp, g, U, indptr, indices, N, mask_history, PRED_TOPN);
    pp.part_idx = c->buf<int>("p_part_idx", (size_t)U * g.P * N);
    pp.part_sq = c->buf<u64>("p_part_sq", (size_t)U * g.P * N);
    pp.part_len = c->buf<int>("p_part_len", (size_t)U * g.P);
    const ModelScale* scale = c->get<ModelScale>("m_scale");
    const int fgrid = (int)std::min<int64_t>((U * 32 + 255) / 256, (int64_t)c->sm_count * 16);
    // the block table of the 32-bit kernel indexes 4-entry blocks with 32 bits
    const Pred32Geom g2 = predict32_geometry(c, N);
    const bool blocks_fit = (c->m_nnz + 3 * c->m_I * g2.P) / 4 < (int64_t)0x7fffffff;
    if ((c->flags & DBG_WIDE_ACC) || !blocks_fit) {  // debug flag / giant models: every user through the two-limb kernel
      launch_predict(c, pp, g, U, indptr);
      k_predict_finalize<<<fgrid, 256, 0, st>>>(pp.part_idx, pp.part_sq, pp.part_len, U, g.P, N, o_idx.dev, o_val.dev,
                                                o_len.dev, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, scale);
      RPK_LAUNCH_CHECK(c);
    } else {
      RPK_REQUIRE(U * (int64_t)g2.P < ((int64_t)1 << 31), "too many (user, item range) work items in one call; split the batch");
      ensure_blocks(c, g2.P, g2.R);
      // lists only (no scores wanted): the approximate sums settle most lists, the rest are scored again exactly
      const bool lists_only = o_val.dev == nullptr && (int64_t)g2.P * N <= MERGE_MAX && !(c->flags & DBG_EXACT_SCORES);
      Pred32Params qp;
      qp.indices = indices;
      qp.ent = c->get<u64>("m_ent4");
      qp.blk = c->get<int4>("m_blk");
      qp.n_work = nullptr;
      qp.U = (int)U;
      qp.P = g2.P;
      qp.R = g2.R;
      qp.I = (int)c->m_I;
      qp.N = N;
      qp.mask = mask_history;
      qp.exact = lists_only ? 0 : 1;
      qp.item_ok = item_filter_ptr(c);
      qp.cap = g2.cap;
      qp.part_idx = c->buf<int>("p_part2_idx", (size_t)U * g2.P * N);
      qp.part_key = c->buf<u64>("p_part2_sq", (size_t)U * g2.P * N);
      qp.part_len = c->buf<int>("p_part2_len", (size_t)U * g2.P);
      qp.part_sft = c->buf<int>("p_part2_sft", (size_t)U * g2.P);
      // per-user flags: [0, U) handed to the two-limb kernel, [U] their count; [U+1, 2U+1) second exact pass, [2U+1] count
      int* flags = c->buf<int>("p_ovf_flag", 2 * (size_t)U + 2);
      qp.ovf_flag = flags;
      qp.ovf_count = flags + U;
      qp.ovf_tab = c->buf<int4>("p_ovf_tab", (size_t)U);
      int* redo_flag = flags + U + 1;
      int* redo_count = flags + 2 * U + 1;
      int4* redo_tab = c->buf<int4>("p_redo_tab", (size_t)U);
      RPK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (2 * (size_t)U + 2), st));
      const PredWork w = prepare_work(c, U, indptr);
      qp.work_tab = w.tab;
      qp.queue = w.queue + 1;
      qp.prof = nullptr;
#ifdef RPK_PHASE_PROF
      qp.prof = c->buf<unsigned long long>("p_prof32", 16);
      RPK_CUDA(cudaMemsetAsync(qp.prof, 0, 16 * sizeof(unsigned long long), st));  // 0-6 phases, 8-11 counters, 12-14 inside select
#endif
      RPK_CUDA(cudaFuncSetAttribute(k_predict_a32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g2.smem));
      int occ = 0;
      RPK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_predict_a32, g2.nt, g2.smem));
      RPK_REQUIRE(occ >= 1, "predict kernel does not fit on an SM");
      const int grid = (int)std::min<int64_t>(U * g2.P, (int64_t)c->sm_count * occ);
      c->mark("predict: layout + work table");
      c->ev_record(4);
      k_predict_a32<<<grid, g2.nt, g2.smem, st>>>(qp);
      RPK_LAUNCH_CHECK(c);
      c->mark("predict: scoring kernel");
      if (lists_only) {
        // order the per-range lists; users whose order the approximate sums cannot prove are scored again, exactly
        const int mgrid = (int)std::min<int64_t>(ceil_div(U, MERGE_WARPS), (int64_t)c->sm_count * 16);
        k_predict_merge<<<mgrid, 32 * MERGE_WARPS, 0, st>>>(qp.part_idx, qp.part_key, qp.part_len, qp.part_sft, indptr, U, g2.P, N,
                                                           o_idx.dev, o_len.dev, qp.ovf_flag, redo_flag, redo_count, redo_tab);
        RPK_LAUNCH_CHECK(c);
        Pred32Params rp = qp;
        rp.work_tab = redo_tab;
        rp.n_work = redo_count;
        rp.exact = 1;
        rp.queue = w.queue + 2;
        k_predict_a32<<<grid, g2.nt, g2.smem, st>>>(rp);
        RPK_LAUNCH_CHECK(c);
      }
      // users whose survivors overflowed the list: exact two-limb kernel (normally none; the kernel then exits)
      pp.work_tab = qp.ovf_tab;
      pp.n_work = qp.ovf_count;
      pp.queue = w.queue;
      launch_predict_kernel(c, pp, g, U);
      c->ev_record(5);
      c->ev_valid[2] = true;
      c->mark("predict: merge + exact passes");
#ifdef RPK_PHASE_PROF
      {
        unsigned long long h[16];
        int novf = 0, nredo = 0;
        RPK_CUDA(cudaMemcpyAsync(h, qp.prof, sizeof(h), cudaMemcpyDeviceToHost, st));
        RPK_CUDA(cudaMemcpyAsync(&novf, qp.ovf_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        RPK_CUDA(cudaMemcpyAsync(&nredo, redo_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        RPK_CUDA(cudaStreamSynchronize(st));
        static const char* nm[7] = {"fetch", "bound+sweep1", "mask", "select", "tag", "sweep2", "sort+out"};
        unsigned long long tot = 0;
        h[3] += 0;
        for (int k = 0; k < 7; ++k) tot += h[k];
        tot += h[12] + h[13] + h[14];
        fprintf(stderr, "[predict32 phases] grid=%d nt=%d P=%d R=%d smem=%zu occ=%d lists_only=%d items=%llu exact items=%llu survivors/item=%.1f overflow users=%d redo users=%d cycles/item=%.0f\n",
                grid, g2.nt, g2.P, g2.R, g2.smem, occ, (int)lists_only, h[8], h[9], h[8] ? (double)h[11] / h[8] : 0.0, novf, nredo,
                h[8] ? (double)tot / h[8] : 0.0);
        for (int k = 0; k < 7; ++k)
          fprintf(stderr, "  %-12s %5.1f%%  %8.0f cycles/item\n", nm[k], 100.0 * h[k] / (tot ? tot : 1), h[8] ? (double)h[k] / h[8] : 0.0);
        fprintf(stderr, "  select = thread maxima %.0f + boundary bin %.0f + copy/clear %.0f + staging of the next item %.0f cycles/item\n",
                h[8] ? (double)h[12] / h[8] : 0.0, h[8] ? (double)h[13] / h[8] : 0.0, h[8] ? (double)h[14] / h[8] : 0.0,
                h[8] ? (double)h[3] / h[8] : 0.0);
      }
#endif
      if (lists_only) {
        k_predict_finalize<<<fgrid, 256, 0, st>>>(qp.part_idx, qp.part_key, qp.part_len, U, g2.P, N, o_idx.dev, o_val.dev, o_len.dev,
                                                  qp.ovf_flag, pp.part_idx, pp.part_sq, pp.part_len, g.P, qp.ovf_flag, redo_flag, scale);
      } else {
        k_predict_finalize<<<fgrid, 256, 0, st>>>(qp.part_idx, qp.part_key, qp.part_len, U, g2.P, N, o_idx.dev, o_val.dev, o_len.dev,
                                                  qp.ovf_flag, pp.part_idx, pp.part_sq, pp.part_len, g.P, nullptr, nullptr, scale);
      }
      RPK_LAUNCH_CHECK(c);
    }
  }
  c->mark("predict: final lists");
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
