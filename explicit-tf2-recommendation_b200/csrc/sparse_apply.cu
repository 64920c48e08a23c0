// K8: sorted-ID segment reduction of row gradients + Adam apply.
//
// Replaces the IndexedSlices path of the reference train step
// (2.FM/ModelManager.py:176-178): Keras dedups the per-occurrence rows with
// tf.unique + unsorted_segment_sum and then runs Adam._resource_apply_sparse.
// Here: stable radix sort of (id, bag index) -> run heads -> per-run reduction
// in ascending occurrence order (deterministic, no global atomics on hot rows;
// long runs are cut into fixed chunks reduced by whole CTAs and combined in
// chunk order) -> Adam on the unique rows.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "etr_common.cuh"
#include "etr_async.cuh"

namespace etr {

constexpr int kShortRun = 64;     // runs up to this length: one lane group
constexpr int kChunk = 1024;      // longer runs: CTA-reduced chunks of this many rows

// ---------------------------------------------------------------- plan
__global__ void __launch_bounds__(256) make_pairs_kernel(const long long* ids, const int* csr, long long B,
                                                         int F, int L, long long sb, long long sf, long long sl,
                                                         long long pad, int has_pad, long long rows,
                                                         long long n_slots, unsigned* keys, int* bags,
                                                         unsigned long long* err) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_slots;
       t += (long long)gridDim.x * blockDim.x) {
    long long id;
    int bag;
    if (csr) {
      id = ids[t];
      // bag of slot t: largest i with csr[i] <= t  (binary search over B*F+1 offsets)
      int lo = 0, hi = (int)(B * F);
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (csr[mid] <= t) lo = mid; else hi = mid;
      }
      bag = lo;
    } else {
      const long long bf = t / L;
      const int l = (int)(t % L);
      const long long b = bf / F;
      const int f = (int)(bf % F);
      id = ids[b * sb + (long long)f * sf + (long long)l * sl];
      bag = (int)bf;
    }
    unsigned key;
    if (has_pad && id == pad) {
      key = (unsigned)rows;          // sentinel: sorts behind every valid id
    } else if ((unsigned long long)id >= (unsigned long long)rows) {
      flag_bad_id(err, id);
      key = (unsigned)rows;
    } else {
      key = (unsigned)id;
    }
    keys[t] = key;
    bags[t] = bag;
  }
}

struct RunHead {
  const unsigned* keys;
  __host__ __device__ __forceinline__ bool operator()(const int& i) const {
    return i == 0 || keys[i] != keys[i - 1];
  }
};

// n_runs -> n_unique (drop the sentinel run), close seg_start, widen unique ids.
__global__ void __launch_bounds__(256) finish_plan_kernel(const unsigned* keys, int* seg_start, const int* n_runs,
                                                          long long n_slots, unsigned sentinel,
                                                          long long* unique_ids, int* n_unique, int* n_valid) {
  const int runs = *n_runs;
  int uniq = runs;
  if (runs > 0 && keys[seg_start[runs - 1]] == sentinel) uniq = runs - 1;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long u = tid; u < uniq; u += (long long)gridDim.x * blockDim.x)
    unique_ids[u] = (long long)keys[seg_start[u]];
  if (tid == 0) {
    // seg_start[uniq] = first pad slot (or n_slots): both close the last valid run
    const int end = (uniq < runs) ? seg_start[uniq] : (int)n_slots;
    *n_unique = uniq;
    *n_valid = end;
  }
}
__global__ void close_plan_kernel(int* seg_start, const int* n_unique, const int* n_valid) {
  seg_start[*n_unique] = *n_valid;
}

// ------------------------------------------------------- segment reduction
struct LongRun { int u; int base; int nchunks; int pad; };

struct SegParams {
  const int* sorted_bag;
  const int* seg_start;
  const int* n_unique;
  const float* bag_grad;
  int grad_ld;        // floats, multiple of 4
  // src_fields > 0: the per-occurrence rows are NOT materialised -- occurrence bg = b * F + f (single-hot ids in
  // (b, f) order) is the slice [src_col0 + f * grad_ld, + grad_ld) of row b of a [B, src_ld] matrix (the gradient
  // of the Flatten(embeddings) block of a dense layer's input: 3.DCN/CustomLayers.py:1096-1100 backward)
  int src_fields, src_col0; long long src_ld;
  float* unique_grad; // [max_unique, grad_ld]
  // long-run work lists (workspace)
  int* counters;      // [0] #long runs, [1] #chunk items
  LongRun* long_runs;
  int2* items;        // (long run index, chunk index)
  float* partials;    // [items, grad_ld]
  int max_long, max_items;
};

__device__ __forceinline__ const float* grad_row(const SegParams& p, int bg) {
  if (p.src_fields == 0) return p.bag_grad + (long long)bg * p.grad_ld;
  const int b = bg / p.src_fields, f = bg - b * p.src_fields;
  return p.bag_grad + (long long)b * p.src_ld + p.src_col0 + f * p.grad_ld;
}

// Sums gradient rows sorted_bag[i0..i1) into acc[CPL] (lane gl of an LPR-lane
// group owns 16-byte chunks gl, gl+LPR, ...), 4 rows in flight.
template <int LPR, int CPL>
__device__ __forceinline__ void sum_rows(const SegParams& p, int i0, int i1, int gl, int nchunks, float4 (&acc)[CPL]) {
  int i = i0;
  for (; i + 4 <= i1; i += 4) {
    int bg[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bg[q] = __ldg(p.sorted_bag + i + q);
    const float* rp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) rp[q] = grad_row(p, bg[q]);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = gl + j * LPR;
      if (c < nchunks) {
        float4 r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          r[q] = *reinterpret_cast<const float4*>(rp[q] + c * 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[j].x += r[q].x; acc[j].y += r[q].y; acc[j].z += r[q].z; acc[j].w += r[q].w; }
      }
    }
  }
  for (; i < i1; ++i) {
    const float* rp = grad_row(p, __ldg(p.sorted_bag + i));
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = gl + j * LPR;
      if (c < nchunks) {
        const float4 r = *reinterpret_cast<const float4*>(rp + c * 4);
        acc[j].x += r.x; acc[j].y += r.y; acc[j].z += r.z; acc[j].w += r.w;
      }
    }
  }
}

template <int LPR, int CPL>
__global__ void __launch_bounds__(256) seg_reduce_short_kernel(const SegParams p) {
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const int nchunks = p.grad_ld / 4;
  const int n_unique = *p.n_unique;
  const long long group_global = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + g;
  const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
  for (long long u = group_global; u < n_unique; u += ngroups) {
    const int s0 = p.seg_start[u], s1 = p.seg_start[u + 1];
    const int len = s1 - s0;
    if (len > kShortRun) {
      if (gl == 0) {
        const int nch = (len + kChunk - 1) / kChunk;
        const int slot = atomicAdd(&p.counters[0], 1);
        const int base = atomicAdd(&p.counters[1], nch);
        if (slot < p.max_long && base + nch <= p.max_items) {
          p.long_runs[slot] = LongRun{(int)u, base, nch, 0};
          for (int c = 0; c < nch; ++c) p.items[base + c] = make_int2(slot, c);
        }
      }
      continue;
    }
    float4 acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    sum_rows<LPR, CPL>(p, s0, s1, gl, nchunks, acc);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = gl + j * LPR;
      if (c < nchunks) *reinterpret_cast<float4*>(p.unique_grad + u * (long long)p.grad_ld + c * 4) = acc[j];
    }
  }
}

// one CTA per chunk item: 256 threads = (256/LPR) lane groups, each summing a
// contiguous sub-range in order; sub-range partials combined in order by group 0.
template <int LPR, int CPL>
__global__ void __launch_bounds__(256) seg_reduce_chunk_kernel(const SegParams p) {
  constexpr int NG = 256 / LPR;
  extern __shared__ __align__(16) float4 sm4[];      // [NG][nchunks]
  const int gl = threadIdx.x % LPR, g = threadIdx.x / LPR;
  const int nchunks = p.grad_ld / 4;
  int n_items = p.counters[1];
  if (n_items > p.max_items) n_items = p.max_items;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int2 item = p.items[it];
    const LongRun lr = p.long_runs[item.x];
    const int s0 = p.seg_start[lr.u] + item.y * kChunk;
    int s1 = p.seg_start[lr.u + 1];
    if (s1 > s0 + kChunk) s1 = s0 + kChunk;
    const int len = s1 - s0;
    const int per = (len + NG - 1) / NG;
    int a = s0 + g * per, b = a + per;
    if (a > s1) a = s1;
    if (b > s1) b = s1;
    float4 acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    sum_rows<LPR, CPL>(p, a, b, gl, nchunks, acc);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = gl + j * LPR;
      if (c < nchunks) sm4[g * nchunks + c] = acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
      float4 t = sm4[c];
      for (int q = 1; q < NG; ++q) {
        const float4 r = sm4[q * nchunks + c];
        t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
      }
      *reinterpret_cast<float4*>(p.partials + (long long)it * p.grad_ld + c * 4) = t;
    }
    __syncthreads();
  }
}

template <int LPR, int CPL>
__global__ void __launch_bounds__(256) seg_reduce_combine_kernel(const SegParams p) {
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const int nchunks = p.grad_ld / 4;
  int n_long = p.counters[0];
  if (n_long > p.max_long) n_long = p.max_long;
  const long long group_global = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + g;
  const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
  for (long long r = group_global; r < n_long; r += ngroups) {
    const LongRun lr = p.long_runs[r];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = gl + j * LPR;
      if (c >= nchunks) continue;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < lr.nchunks; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(p.partials + (long long)(lr.base + q) * p.grad_ld + c * 4);
        t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
      }
      *reinterpret_cast<float4*>(p.unique_grad + (long long)lr.u * p.grad_ld + c * 4) = t;
    }
  }
}

// ----------------------------------------------------------------- Adam
struct AdamParams {
  char* table; int table_bf16; int stride;   // elements per row
  float* m; float* v;
  const long long* unique_ids; const int* n_unique;
  const float* grad; int grad_ld;
  float lr_t, b1, b2, eps;
  const float* d_lr_t;  // when non-NULL the step size is read from the device (CUDA-graph replay)
  int scatter_only;     // keras_dense: only m += (1-b1) g ; v += (1-b2) g^2
};

__device__ __forceinline__ void adam4(float4& var, float4& m, float4& v, const float4 g, const AdamParams& p,
                                      const float lr_t) {
  float* pv = &var.x; float* pm = &m.x; float* pvv = &v.x; const float* pg = &g.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pm[i] = p.b1 * pm[i] + (1.0f - p.b1) * pg[i];
    pvv[i] = p.b2 * pvv[i] + (1.0f - p.b2) * pg[i] * pg[i];
    pv[i] = pv[i] - lr_t * pm[i] / (sqrtf(pvv[i]) + p.eps);
  }
}

// one thread per (unique row, 4-column chunk); m/v/grad rows share the fp32 stride.
__global__ void __launch_bounds__(256) sparse_adam_kernel(const AdamParams p) {
  const int nch = p.grad_ld / 4;
  const float lr_t = p.d_lr_t ? *p.d_lr_t : p.lr_t;
  const long long total = (long long)(*p.n_unique) * nch;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long u = t / nch;
    const int c = (int)(t % nch);
    const long long row = p.unique_ids[u];
    const float4 g = *reinterpret_cast<const float4*>(p.grad + u * p.grad_ld + c * 4);
    float4* pm = reinterpret_cast<float4*>(p.m + row * p.stride + c * 4);
    float4* pv = reinterpret_cast<float4*>(p.v + row * p.stride + c * 4);
    float4 m = *pm, v = *pv;
    if (p.scatter_only) {
      m.x += (1.0f - p.b1) * g.x; m.y += (1.0f - p.b1) * g.y; m.z += (1.0f - p.b1) * g.z; m.w += (1.0f - p.b1) * g.w;
      v.x += (1.0f - p.b2) * g.x * g.x; v.y += (1.0f - p.b2) * g.y * g.y;
      v.z += (1.0f - p.b2) * g.z * g.z; v.w += (1.0f - p.b2) * g.w * g.w;
      *pm = m; *pv = v;
      continue;
    }
    if (!p.table_bf16) {
      float4* pvar = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.table) + row * p.stride + c * 4);
      float4 var = *pvar;
      adam4(var, m, v, g, p, lr_t);
      *pvar = var; *pm = m; *pv = v;
    } else {
      __nv_bfloat162* pvar = reinterpret_cast<__nv_bfloat162*>(
          reinterpret_cast<__nv_bfloat16*>(p.table) + row * p.stride + c * 4);
      const float2 lo = __bfloat1622float2(pvar[0]), hi = __bfloat1622float2(pvar[1]);
      float4 var = make_float4(lo.x, lo.y, hi.x, hi.y);
      adam4(var, m, v, g, p, lr_t);
      pvar[0] = __floats2bfloat162_rn(var.x, var.y);
      pvar[1] = __floats2bfloat162_rn(var.z, var.w);
      *pm = m; *pv = v;
    }
  }
}

// used-rows L2 (5.DIN/ModelManager.py:175-190): l2 = factor * 0.5 * sum_{u in unique(batch ids)} |row_u|^2 ; grad_u += factor * row_u
__global__ void __launch_bounds__(256) used_rows_l2_kernel(const char* table, int bf16, int stride, int width, const long long* unique_ids,
                                                           const int* n_unique, float factor, float* grad, int grad_ld, float* rowsq) {
  const int n = *n_unique;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
    const long long row = unique_ids[u];
    float s = 0.f;
    for (int c = 0; c < width; ++c) {
      const float x = bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(table)[row * stride + c])
                           : reinterpret_cast<const float*>(table)[row * stride + c];
      s += x * x;
      grad[(long long)u * grad_ld + c] += factor * x;
    }
    rowsq[u] = s;
  }
}
// fixed-order sum (one block, strided partials then a serial tail): deterministic
__global__ void __launch_bounds__(1024) used_rows_l2_finish_kernel(const float* rowsq, const int* n_unique, float factor, float* loss_accum) {
  __shared__ float part[1024];
  const int n = *n_unique;
  float s = 0.f;
  for (int u = threadIdx.x; u < n; u += 1024) s += rowsq[u];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 1024; ++i) t += part[i];
    *loss_accum += 0.5f * factor * t;
  }
}

// keras_dense passes over the whole table: rows x (rc 16-byte chunks of meaningful columns); row stride in
// floats (== 4*rc for a plain table: one flat pass; 64 for a RECORD table whose m / v sit inside the record)
__global__ void __launch_bounds__(256) dense_decay_kernel(float* m, float* v, long long rows, int rc, int stride, float b1,
                                                          float b2) {
  const long long n4 = rows * rc;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const long long off = (t / rc) * stride + (t % rc) * 4;
    float4 a = *reinterpret_cast<float4*>(m + off), b = *reinterpret_cast<float4*>(v + off);
    a.x *= b1; a.y *= b1; a.z *= b1; a.w *= b1;
    b.x *= b2; b.y *= b2; b.z *= b2; b.w *= b2;
    *reinterpret_cast<float4*>(m + off) = a; *reinterpret_cast<float4*>(v + off) = b;
  }
}
// The quotient uses the MUFU forms (sqrt.approx / div.approx, <= 2 ulp each, as the fused row-wise apply does), NOT the
// IEEE ones.  Under keras_dense every row's m decays by 0.9 per step, so ~700 steps after its last touch a row's m is a
// SUBNORMAL float for ~150 steps before it reaches zero, and IEEE division / square root have data-dependent slow
// paths.  Measured (end of round 2): value_keras_dense of the full default bench, i.e. after ~2 500 training steps on
// the table, 27.5 ms per step with the IEEE form, 5.67 ms with this one (5.7 ms on a fresh table either way; an
// all-subnormal m alone costs the IEEE form only +0.5 ms, scripts/dbg_extras.py, so the subnormal window is not the
// whole story -- the branch-free form is what removed the dependence on the table's history).  IEEE = true keeps the
// old form for an A/B (ETR_DENSE_IEEE=1).
template <bool IEEE>
__global__ void __launch_bounds__(256) dense_var_update_kernel(char* table, int bf16, const float* m, const float* v,
                                                               long long rows, int rc, int stride, float lr_host,
                                                               const float* d_lr_t, float eps) {
  const float lr_t = d_lr_t ? *d_lr_t : lr_host;
  const long long n4 = rows * rc;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const long long off = (t / rc) * stride + (t % rc) * 4;
    const float4 a = *reinterpret_cast<const float4*>(m + off), b = *reinterpret_cast<const float4*>(v + off);
    float d[4];
    if (IEEE) {
      d[0] = lr_t * a.x / (sqrtf(b.x) + eps); d[1] = lr_t * a.y / (sqrtf(b.y) + eps);
      d[2] = lr_t * a.z / (sqrtf(b.z) + eps); d[3] = lr_t * a.w / (sqrtf(b.w) + eps);
    } else {
      d[0] = fast_div(lr_t * a.x, fast_sqrt(b.x) + eps); d[1] = fast_div(lr_t * a.y, fast_sqrt(b.y) + eps);
      d[2] = fast_div(lr_t * a.z, fast_sqrt(b.z) + eps); d[3] = fast_div(lr_t * a.w, fast_sqrt(b.w) + eps);
    }
    if (!bf16) {
      float4* px = reinterpret_cast<float4*>(reinterpret_cast<float*>(table) + off);
      float4 x = *px;
      x.x -= d[0]; x.y -= d[1]; x.z -= d[2]; x.w -= d[3];
      *px = x;
    } else {
      __nv_bfloat162* px = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(table) + off);
      float2 lo = __bfloat1622float2(px[0]), hi = __bfloat1622float2(px[1]);
      px[0] = __floats2bfloat162_rn(lo.x - d[0], lo.y - d[1]);
      px[1] = __floats2bfloat162_rn(hi.x - d[2], hi.y - d[3]);
    }
  }
}

__global__ void __launch_bounds__(256) dense_adam_kernel(float* var, float* m, float* v, const float* g, long long n,
                                                         float lr_host, const float* d_lr_t, float b1, float b2,
                                                         float eps) {
  const float lr_t = d_lr_t ? *d_lr_t : lr_host;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float gi = g[t];
    const float mi = b1 * m[t] + (1.0f - b1) * gi;
    const float vi = b2 * v[t] + (1.0f - b2) * gi * gi;
    m[t] = mi; v[t] = vi;
    var[t] = var[t] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// state[0] = t (as float, exact up to 2^24 steps), state[1] = lr_t of step t
__global__ void adam_step_begin_kernel(float* state, float lr, float b1, float b2) {
  const float t = state[0] + 1.0f;
  state[0] = t;
  state[1] = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
}

static inline int lpr_for(int nchunks) {
  int lpr = 1;
  while (lpr < nchunks && lpr < 32) lpr <<= 1;
  return lpr;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace etr

using namespace etr;

extern "C" {

int64_t etr_sparse_plan_slots(const etr_ids* ids, int64_t nnz_if_csr) {
  if (!ids) return -1;
  if (ids->d_csr_offsets) return nnz_if_csr;
  return ids->batch * (int64_t)ids->fields * ids->bag;
}

int etr_sparse_plan(etr_ctx* ctx, const etr_ids* ids, int64_t nnz_if_csr, int64_t table_rows,
                    int32_t* d_sorted_bag, int64_t* d_unique_ids, int32_t* d_seg_start,
                    int32_t* d_n_unique, int32_t* d_n_valid, void* stream) {
  return etr_sparse_plan_keys(ctx, ids, nnz_if_csr, table_rows, d_sorted_bag, d_unique_ids, d_seg_start, d_n_unique,
                              d_n_valid, nullptr, stream);
}

int etr_sparse_plan_keys(etr_ctx* ctx, const etr_ids* ids, int64_t nnz_if_csr, int64_t table_rows,
                         int32_t* d_sorted_bag, int64_t* d_unique_ids, int32_t* d_seg_start,
                         int32_t* d_n_unique, int32_t* d_n_valid, uint32_t* d_sorted_key, void* stream) {
  ETR_CHECK_ARG(ctx && ids && ids->d_ids && d_sorted_bag && d_unique_ids && d_seg_start && d_n_unique && d_n_valid,
                "NULL argument");
  ETR_CHECK_ARG(table_rows > 0 && table_rows < 0xffffffffLL, "table_rows must fit 32 bits");
  const int64_t n = etr_sparse_plan_slots(ids, nnz_if_csr);
  ETR_CHECK_ARG(n >= 0 && n < 0x7fffffffLL, "slot count must fit int32");
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    ETR_CUDA(cudaMemsetAsync(d_n_unique, 0, sizeof(int), s));
    ETR_CUDA(cudaMemsetAsync(d_n_valid, 0, sizeof(int), s));
    ETR_CUDA(cudaMemsetAsync(d_seg_start, 0, sizeof(int), s));
    return ETR_OK;
  }
  int end_bit = 1;
  while (end_bit < 32 && (1ull << end_bit) <= (unsigned long long)table_rows) ++end_bit;

  size_t sort_bytes = 0, sel_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n, 0, end_bit, s);
  cub::CountingInputIterator<int> counting(0);
  RunHead pred{nullptr};
  cub::DeviceSelect::If(nullptr, sel_bytes, counting, (int*)nullptr, (int*)nullptr, (int)n, pred, s);
  const size_t tmp_bytes = align256(sort_bytes > sel_bytes ? sort_bytes : sel_bytes);
  const size_t arr = align256((size_t)n * 4);
  // workspace: keys_in | keys_out | bags_in | n_runs | cub temp
  const size_t need = 3 * arr + 256 + tmp_bytes;
  int st = etr_ws_reserve(ctx, need);
  if (st != ETR_OK) return st;
  char* ws = (char*)ctx->d_ws;
  unsigned* keys_in = (unsigned*)ws;
  unsigned* keys_out = d_sorted_key ? d_sorted_key : (unsigned*)(ws + arr);   // the sorted ids themselves, kept for the caller
  int* bags_in = (int*)(ws + 2 * arr);
  int* n_runs = (int*)(ws + 3 * arr);
  void* tmp = ws + 3 * arr + 256;

  const int grid = grid_for(n, 256, ctx->sm_count, 8);
  make_pairs_kernel<<<grid, 256, 0, s>>>((const long long*)ids->d_ids, ids->d_csr_offsets, ids->batch, ids->fields,
                                         ids->d_csr_offsets ? 1 : ids->bag, ids->stride_b, ids->stride_f,
                                         ids->stride_l, ids->pad_id, ids->has_pad, table_rows, n, keys_in, bags_in,
                                         ctx->d_err);
  ETR_LAUNCH_CHECK(ctx);
  size_t tb = tmp_bytes;
  ETR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys_out, bags_in, d_sorted_bag, (int)n, 0, end_bit, s));
  pred.keys = keys_out;
  tb = tmp_bytes;
  ETR_CUDA(cub::DeviceSelect::If(tmp, tb, counting, d_seg_start, n_runs, (int)n, pred, s));
  finish_plan_kernel<<<grid, 256, 0, s>>>(keys_out, d_seg_start, n_runs, n, (unsigned)table_rows,
                                          (long long*)d_unique_ids, d_n_unique, d_n_valid);
  ETR_LAUNCH_CHECK(ctx);
  close_plan_kernel<<<1, 1, 0, s>>>(d_seg_start, d_n_unique, d_n_valid);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

static int segment_reduce_impl(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                               const int32_t* d_n_unique, int64_t n_slots, const float* d_bag_grad,
                               int32_t grad_ld, int32_t src_fields, int64_t src_ld, int32_t src_col0,
                               float* d_unique_grad, void* stream);

int etr_sparse_segment_reduce(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                              const int32_t* d_n_unique, int64_t n_slots, const float* d_bag_grad,
                              int32_t grad_ld, float* d_unique_grad, void* stream) {
  return segment_reduce_impl(ctx, d_sorted_bag, d_seg_start, d_n_unique, n_slots, d_bag_grad, grad_ld, 0, 0, 0,
                             d_unique_grad, stream);
}

int etr_sparse_segment_reduce_flat(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                                   const int32_t* d_n_unique, int64_t n_slots, const float* d_flat, int64_t flat_ld,
                                   int32_t flat_col0, int32_t fields, int32_t grad_ld, float* d_unique_grad,
                                   void* stream) {
  ETR_CHECK_ARG(fields > 0 && flat_ld % 4 == 0 && flat_col0 % 4 == 0 && flat_col0 >= 0 &&
                    (int64_t)flat_col0 + (int64_t)fields * grad_ld <= flat_ld,
                "flat source: 16-byte aligned slices inside a row");
  return segment_reduce_impl(ctx, d_sorted_bag, d_seg_start, d_n_unique, n_slots, d_flat, grad_ld, fields, flat_ld,
                             flat_col0, d_unique_grad, stream);
}

static int segment_reduce_impl(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                               const int32_t* d_n_unique, int64_t n_slots, const float* d_bag_grad,
                               int32_t grad_ld, int32_t src_fields, int64_t src_ld, int32_t src_col0,
                               float* d_unique_grad, void* stream) {
  ETR_CHECK_ARG(ctx && d_sorted_bag && d_seg_start && d_n_unique && d_bag_grad && d_unique_grad, "NULL argument");
  ETR_CHECK_ARG(grad_ld > 0 && grad_ld % 4 == 0, "grad_ld must be a positive multiple of 4");
  ETR_CHECK_ARG((((uintptr_t)d_bag_grad | (uintptr_t)d_unique_grad) & 15) == 0, "gradients must be 16-byte aligned");
  if (n_slots == 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SegParams p;
  p.sorted_bag = d_sorted_bag; p.seg_start = d_seg_start; p.n_unique = d_n_unique;
  p.bag_grad = d_bag_grad; p.grad_ld = grad_ld; p.unique_grad = d_unique_grad;
  p.src_fields = src_fields; p.src_ld = src_ld; p.src_col0 = src_col0;
  p.max_long = (int)(n_slots / kShortRun + 1);
  p.max_items = (int)(n_slots / kChunk + n_slots / kShortRun + 2);
  const size_t b_cnt = 256, b_runs = align256(sizeof(LongRun) * (size_t)p.max_long),
               b_items = align256(sizeof(int2) * (size_t)p.max_items),
               b_part = align256(sizeof(float) * (size_t)p.max_items * grad_ld);
  // NOTE: the plan's workspace is dead once seg_start / sorted_bag are produced.
  int st = etr_ws_reserve(ctx, b_cnt + b_runs + b_items + b_part);
  if (st != ETR_OK) return st;
  char* ws = (char*)ctx->d_ws;
  p.counters = (int*)ws;
  p.long_runs = (LongRun*)(ws + b_cnt);
  p.items = (int2*)(ws + b_cnt + b_runs);
  p.partials = (float*)(ws + b_cnt + b_runs + b_items);
  ETR_CUDA(cudaMemsetAsync(p.counters, 0, 2 * sizeof(int), s));
  const int nchunks = grad_ld / 4;
  const int lpr = lpr_for(nchunks);
  const int cpl = (nchunks + lpr - 1) / lpr;
  if (cpl > 4) { etr_set_error("segment_reduce: grad rows wider than 512 floats are not covered"); return ETR_EUNSUPPORTED; }
  const int gshort = grid_for(n_slots, 8 * (32 / lpr), ctx->sm_count, 8);
  const int gchunk = ctx->sm_count * 4;
  const size_t csm = (size_t)(256 / lpr) * nchunks * sizeof(float4);
#define ETR_SEG(LPR, CPL)                                                    \
  do {                                                                       \
    seg_reduce_short_kernel<LPR, CPL><<<gshort, 256, 0, s>>>(p);             \
    ETR_LAUNCH_CHECK(ctx);                                                   \
    seg_reduce_chunk_kernel<LPR, CPL><<<gchunk, 256, csm, s>>>(p);           \
    ETR_LAUNCH_CHECK(ctx);                                                   \
    seg_reduce_combine_kernel<LPR, CPL><<<ctx->sm_count, 256, 0, s>>>(p);    \
    ETR_LAUNCH_CHECK(ctx);                                                   \
  } while (0)
  if (cpl == 1) {
    switch (lpr) {
      case 1: ETR_SEG(1, 1); break;
      case 2: ETR_SEG(2, 1); break;
      case 4: ETR_SEG(4, 1); break;
      case 8: ETR_SEG(8, 1); break;
      case 16: ETR_SEG(16, 1); break;
      default: ETR_SEG(32, 1); break;
    }
  } else if (cpl == 2) {
    ETR_SEG(32, 2);
  } else {
    ETR_SEG(32, 4);
  }
#undef ETR_SEG
  return ETR_OK;
}

int etr_sparse_adam_apply(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v,
                          const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t max_unique,
                          const float* d_unique_grad, int32_t grad_ld, float lr_t, const float* d_lr_t,
                          float beta1, float beta2, float eps, int32_t mode, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_m && d_v && d_unique_ids && d_n_unique && d_unique_grad,
                "NULL argument");
  const bool record = table->reserved == ETR_TABLE_RECORD;
  const int row_cols = record ? ETR_RECORD_ROW_FLOATS : table->stride;     // meaningful columns of a row, whole chunks
  ETR_CHECK_ARG(grad_ld == row_cols, "grad_ld must equal the table stride (20 for a RECORD table): gradient rows use the table's column layout");
  ETR_CHECK_ARG(grad_ld % 4 == 0 && table->stride % 4 == 0, "stride must be a multiple of 4");
  ETR_CHECK_ARG(!record || (table->dtype == ETR_F32 && table->stride == 64 && d_m == (float*)table->d_data + 20 &&
                            d_v == (float*)table->d_data + 40), "RECORD table: m / v must be the record's own slots");
  cudaStream_t s = (cudaStream_t)stream;
  AdamParams p;
  p.table = (char*)table->d_data; p.table_bf16 = table->dtype == ETR_BF16; p.stride = table->stride;
  p.m = d_m; p.v = d_v; p.unique_ids = (const long long*)d_unique_ids; p.n_unique = d_n_unique;
  p.grad = d_unique_grad; p.grad_ld = grad_ld; p.lr_t = lr_t; p.d_lr_t = d_lr_t; p.b1 = beta1; p.b2 = beta2; p.eps = eps;
  p.scatter_only = mode == ETR_ADAM_KERAS_DENSE;
  const int rc = row_cols / 4;
  const long long n4 = table->rows * (long long)rc;
  const int gd = grid_for(n4, 256, ctx->sm_count, 8);
  if (mode == ETR_ADAM_KERAS_DENSE) {
    dense_decay_kernel<<<gd, 256, 0, s>>>(d_m, d_v, table->rows, rc, table->stride, beta1, beta2);
    ETR_LAUNCH_CHECK(ctx);
  }
  if (max_unique > 0) {
    const int g = grid_for(max_unique * (grad_ld / 4), 256, ctx->sm_count, 8);
    sparse_adam_kernel<<<g, 256, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
  }
  if (mode == ETR_ADAM_KERAS_DENSE) {
    static int ieee = -1;
    if (ieee < 0) { const char* e = getenv("ETR_DENSE_IEEE"); ieee = (e && atoi(e) == 1) ? 1 : 0; }
    if (ieee) dense_var_update_kernel<true><<<gd, 256, 0, s>>>((char*)table->d_data, p.table_bf16, d_m, d_v, table->rows, rc, table->stride,
                                               lr_t, d_lr_t, eps);
    else dense_var_update_kernel<false><<<gd, 256, 0, s>>>((char*)table->d_data, p.table_bf16, d_m, d_v, table->rows, rc, table->stride,
                                               lr_t, d_lr_t, eps);
    ETR_LAUNCH_CHECK(ctx);
  }
  return ETR_OK;
}

int etr_used_rows_l2(etr_ctx* ctx, const etr_table* table, const int64_t* d_unique_ids, const int32_t* d_n_unique,
                     int64_t max_unique, float factor, float* d_unique_grad, int32_t grad_ld, float* d_loss_accum, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_unique_ids && d_n_unique && d_unique_grad && d_loss_accum, "NULL argument");
  ETR_CHECK_ARG(grad_ld >= table->width, "grad_ld too small");
  if (max_unique <= 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int st = etr_ws_reserve(ctx, (size_t)max_unique * sizeof(float));
  if (st != ETR_OK) return st;
  float* rowsq = (float*)ctx->d_ws;
  used_rows_l2_kernel<<<grid_for(max_unique, 256, ctx->sm_count, 8), 256, 0, s>>>(
      (const char*)table->d_data, table->dtype == ETR_BF16, table->stride, table->width, (const long long*)d_unique_ids, d_n_unique,
      factor, d_unique_grad, grad_ld, rowsq);
  ETR_LAUNCH_CHECK(ctx);
  used_rows_l2_finish_kernel<<<1, 1024, 0, s>>>(rowsq, d_n_unique, factor, d_loss_accum);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_adam_step_begin(etr_ctx* ctx, float* d_state, float lr, float beta1, float beta2, void* stream) {
  ETR_CHECK_ARG(ctx && d_state, "NULL argument");
  adam_step_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_state, lr, beta1, beta2);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_dense_adam_apply(etr_ctx* ctx, float* d_var, float* d_m, float* d_v, const float* d_grad, int64_t n,
                         float lr_t, const float* d_lr_t, float beta1, float beta2, float eps, void* stream) {
  ETR_CHECK_ARG(ctx && d_var && d_m && d_v && d_grad, "NULL argument");
  if (n <= 0) return ETR_OK;
  const int g = grid_for(n, 256, ctx->sm_count, 8);
  dense_adam_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(d_var, d_m, d_v, d_grad, n, lr_t, d_lr_t, beta1, beta2, eps);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
