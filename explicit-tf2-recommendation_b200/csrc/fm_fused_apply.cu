// K2+K8 fused for the FM family (single-hot ids, fp32 table with the fused w column):
// backward of the gather/FM kernel + sorted-ID segment reduction + row-wise Adam in ONE pass
// over the sorted occurrence list, without ever materialising per-occurrence (or per-bag)
// gradient rows.
//
// For a table row r with occurrences (b,f) in its run (SURVEY a', FM / DeepFM):
//   dL/dv_r[c] = sum_(b,f) [ g_b (S_b[c] - v_r[c]) + dflat[b, f*k + c] ]
//              = sum g_b S_b[c]  +  sum dflat[b, f*k+c]  -  v_r[c] * sum g_b
//   dL/dw_r    = sum g_b
// so a run needs, per occurrence, only g_b (4 B), S_b (k floats, saved by the forward kernel)
// and the MLP's input gradient slice -- all L2-resident (B*k*4 = 4 MB, B*F*k*2 = 56 MB at c2) --
// and the row itself once, which the Adam update reads anyway.  Replaces
// tape.gradient + Keras apply_gradients on IndexedSlices (2.FM/ModelManager.py:176-178).
//
// Runs <= 64 occurrences: one lane group (k/4 lanes, one float4 column chunk per lane), fully
// fused.  Longer runs (hot ids): cut into 1024-occurrence chunks, each reduced by a whole CTA
// into a partial row [P_0..P_{k-1}, sum_g] (fixed order), combined in chunk order and then
// updated -- deterministic, no atomics on hot rows.
#include <stdlib.h>

#include <algorithm>

#include "etr_common.cuh"
#include "etr_async.cuh"
#include "fm_fused_tile.cuh"

namespace etr {

constexpr int kFusedShortRun = 64;
constexpr int kFusedChunk = 1024;

struct FusedLong { int u; int base; int nchunks; int pad; };

struct FusedParams {
  float* table; float* m; float* v; int stride; int k;
  int F; unsigned long long magic; int shift;       // b = (bag * magic) >> shift  ==  bag / F
  const int* sorted_bag; const int* seg_start; const long long* unique_ids; const int* n_unique;
  const float* dlogit; const float* sumv;
  const void* dflat; int flat_bf16; long long flat_ld; int flat_col0;
  float lr_t; const float* d_lr_t; float b1, b2, eps;
  int apply;                  // 1: Adam update; 0: only export the gradient rows
  float* unique_grad;         // optional [n_unique, gld]
  int gld;                    // leading dim of exported gradient rows: stride (20 for a RECORD table)
  // peer-sharded export straight into the owners' gradient mailboxes (NVLink peer stores): row u goes to
  // slot_of_u[u] = owner * cap + slot of the request exchange; replaces unique_grad + a push kernel
  const int* slot_of_u; int cap; float* grads_mb[16];
  int* counters; FusedLong* long_runs; int2* items; float* partials; int max_long, max_items;
  const char* shard_base[16]; int world;     // peer-sharded table (export mode only): ids are global
  int sharded;                               // 1: deferred export (the owner finishes dv = P - v * sum_g), also at world 1
};

__device__ __forceinline__ void bag_to_bf(const FusedParams& p, int bag, int& b, int& f) {
  b = (int)(((unsigned long long)(unsigned)bag * p.magic) >> p.shift);
  f = bag - b * p.F;
}

// accumulate occurrences [i0, i1) into acc (this lane's float4 column chunk) and sum_g; UN
// occurrences in flight (2 in the short-run kernel, where 83% of the runs have length 1 and
// registers are better spent on occupancy; 4 in the chunk kernel)
template <int UN>
__device__ __forceinline__ void fused_accumulate(const FusedParams& p, int i0, int i1, int gl, float4& acc, float& sum_g) {
  int i = i0;
  for (; i + UN <= i1; i += UN) {
    int b[UN], f[UN];
    float g[UN];
    float4 s[UN], d[UN];
#pragma unroll
    for (int q = 0; q < UN; ++q) bag_to_bf(p, __ldg(p.sorted_bag + i + q), b[q], f[q]);
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      g[q] = __ldg(p.dlogit + b[q]);
      s[q] = *reinterpret_cast<const float4*>(p.sumv + (long long)b[q] * p.k + gl * 4);
      d[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.dflat) {
        const long long e0 = (long long)b[q] * p.flat_ld + p.flat_col0 + (long long)f[q] * p.k + gl * 4;
        if (p.flat_bf16) {
          const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.dflat) + e0);
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
          d[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          d[q] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dflat) + e0);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      acc.x += g[q] * s[q].x + d[q].x; acc.y += g[q] * s[q].y + d[q].y;
      acc.z += g[q] * s[q].z + d[q].z; acc.w += g[q] * s[q].w + d[q].w;
      sum_g += g[q];
    }
  }
  for (; i < i1; ++i) {
    int b, f;
    bag_to_bf(p, __ldg(p.sorted_bag + i), b, f);
    const float g = __ldg(p.dlogit + b);
    const float4 s = *reinterpret_cast<const float4*>(p.sumv + (long long)b * p.k + gl * 4);
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.dflat) {
      const long long e0 = (long long)b * p.flat_ld + p.flat_col0 + (long long)f * p.k + gl * 4;
      if (p.flat_bf16) {
        const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.dflat) + e0);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
        d = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        d = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dflat) + e0);
      }
    }
    acc.x += g * s.x + d.x; acc.y += g * s.y + d.y; acc.z += g * s.z + d.z; acc.w += g * s.w + d.w;
    sum_g += g;
  }
}

// occurrences i0, i0+step, ... < i1 into acc / sum_g (this lane's float4 column chunk), two in flight
__device__ __forceinline__ void fused_accumulate_strided(const FusedParams& p, int i0, int i1, int step, int gl, float4& acc,
                                                         float& sum_g) {
  const __nv_bfloat16* dfl16 = reinterpret_cast<const __nv_bfloat16*>(p.dflat);
  const float* dfl32 = reinterpret_cast<const float*>(p.dflat);
  auto fetch = [&](int i, float& g, float4& s, float4& d) {
    int b, f;
    bag_to_bf(p, __ldg(p.sorted_bag + i), b, f);
    g = __ldg(p.dlogit + b);
    s = *reinterpret_cast<const float4*>(p.sumv + (long long)b * p.k + gl * 4);
    d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.dflat) {
      const long long e0 = (long long)b * p.flat_ld + p.flat_col0 + (long long)f * p.k + gl * 4;
      if (p.flat_bf16) {
        const uint2 w = *reinterpret_cast<const uint2*>(dfl16 + e0);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
        d = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        d = *reinterpret_cast<const float4*>(dfl32 + e0);
      }
    }
  };
  int i = i0;
  for (; i + step < i1; i += 2 * step) {
    float g0, g1;
    float4 s0, s1, d0, d1;
    fetch(i, g0, s0, d0);
    fetch(i + step, g1, s1, d1);
    acc.x += g0 * s0.x + d0.x; acc.y += g0 * s0.y + d0.y; acc.z += g0 * s0.z + d0.z; acc.w += g0 * s0.w + d0.w;
    sum_g += g0;
    acc.x += g1 * s1.x + d1.x; acc.y += g1 * s1.y + d1.y; acc.z += g1 * s1.z + d1.z; acc.w += g1 * s1.w + d1.w;
    sum_g += g1;
  }
  if (i < i1) {
    float g0;
    float4 s0, d0;
    fetch(i, g0, s0, d0);
    acc.x += g0 * s0.x + d0.x; acc.y += g0 * s0.y + d0.y; acc.z += g0 * s0.z + d0.z; acc.w += g0 * s0.w + d0.w;
    sum_g += g0;
  }
}

__device__ __forceinline__ void adam_update4(float4& var, float4& m, float4& v, const float4 g, float lr_t, float b1,
                                             float b2, float eps) {
  float* pv = &var.x; float* pm = &m.x; float* pvv = &v.x; const float* pg = &g.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pm[i] = b1 * pm[i] + (1.0f - b1) * pg[i];
    pvv[i] = b2 * pvv[i] + (1.0f - b2) * pg[i] * pg[i];
    pv[i] = pv[i] - fast_div(lr_t * pm[i], fast_sqrt(pvv[i]) + eps);        // MUFU sqrt / rcp: <= 2 ulp each
  }
}

// finish one row: grad = P - v_r * sum_g (embedding chunk gl), w chunk by lane 0; Adam and/or export
struct RowState { float4 var, m, v; float wx; };
// issue the row's (var, m, v) loads; called BEFORE the occurrence loop so that their DRAM latency
// overlaps the loop instead of following it.  LPR >= 4: the w column's state is prefetched here too -- lane 1
// of the group reads w, lane 2 its m, lane 3 its v (one scalar each) -- instead of three dependent 16-byte
// loads by lane 0 after the loop (a second exposed DRAM round trip per row in the first version).
template <int LPR>
__device__ __forceinline__ RowState fused_load_row(const FusedParams& p, long long row, int gl) {
  RowState st;
  st.var = st.m = st.v = make_float4(0.f, 0.f, 0.f, 0.f);
  st.wx = 0.f;
  // peer-sharded table (export only): DEFERRED form -- the row is not read here at all (it may live
  // in another GPU's HBM); the exported row is [P, sum_g] and the owner finishes dv = P - v * sum_g
  if (p.sharded) return st;
  st.var = *reinterpret_cast<const float4*>(p.table + row * p.stride + gl * 4);
  if (p.apply) {
    st.m = *reinterpret_cast<const float4*>(p.m + row * p.stride + gl * 4);
    st.v = *reinterpret_cast<const float4*>(p.v + row * p.stride + gl * 4);
    if (LPR >= 4 && gl >= 1 && gl <= 3) {
      const float* src = gl == 1 ? p.table : (gl == 2 ? p.m : p.v);
      st.wx = src[row * p.stride + p.k];
    }
  }
  return st;
}

template <int LPR>
__device__ __forceinline__ void fused_finish_row(const FusedParams& p, long long u, long long row, int gl, int g,
                                                 RowState st, float4 P, float sum_g, float lr_t) {
  float* prow = p.table + row * p.stride;
  float4 var = st.var;
  const float4 gr = make_float4(P.x - var.x * sum_g, P.y - var.y * sum_g, P.z - var.z * sum_g, P.w - var.w * sum_g);
  float* gdst = p.unique_grad ? p.unique_grad + u * p.gld : nullptr;
  if (p.slot_of_u) {
    const int so = p.slot_of_u[u];
    const int owner = so / p.cap;
    gdst = p.grads_mb[owner] + (long long)(so - owner * p.cap) * p.gld;
  }
  if (gdst) *reinterpret_cast<float4*>(gdst + gl * 4) = gr;
  if (p.apply) {
    float4 m = st.m, v = st.v;
    adam_update4(var, m, v, gr, lr_t, p.b1, p.b2, p.eps);
    *reinterpret_cast<float4*>(prow + gl * 4) = var;
    *reinterpret_cast<float4*>(p.m + row * p.stride + gl * 4) = m;
    *reinterpret_cast<float4*>(p.v + row * p.stride + gl * 4) = v;
  }
  if (LPR >= 4) {
    // the w column: state prefetched by lanes 1..3, one scalar Adam by lane 0 (the zero padding columns behind
    // w have zero gradient and zero moments: Adam leaves them untouched, so skipping them is exact)
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
    float w = __shfl_sync(gmask, st.wx, g * LPR + 1);
    float wm = __shfl_sync(gmask, st.wx, g * LPR + 2);
    float wv = __shfl_sync(gmask, st.wx, g * LPR + 3);
    if (gl == 0) {
      if (gdst)
        for (int c = p.k; c < p.gld; c += 4)
          *reinterpret_cast<float4*>(gdst + c) = make_float4(c == p.k ? sum_g : 0.f, 0.f, 0.f, 0.f);
      if (p.apply) {
        wm = p.b1 * wm + (1.0f - p.b1) * sum_g;
        wv = p.b2 * wv + (1.0f - p.b2) * sum_g * sum_g;
        w = w - fast_div(lr_t * wm, fast_sqrt(wv) + p.eps);
        prow[p.k] = w;
        p.m[row * p.stride + p.k] = wm;
        p.v[row * p.stride + p.k] = wv;
      }
    }
  } else if (gl == 0) {
    // chunks behind the embedding: [w, 0, 0, 0] (+ zero padding chunks)
    for (int c = p.k; c < p.gld; c += 4) {
      const float4 gw = make_float4(c == p.k ? sum_g : 0.f, 0.f, 0.f, 0.f);
      if (gdst) *reinterpret_cast<float4*>(gdst + c) = gw;
      if (p.apply && c == p.k) {
        float4 wv = *reinterpret_cast<const float4*>(prow + c);
        float4* pm = reinterpret_cast<float4*>(p.m + row * p.stride + c);
        float4* pv = reinterpret_cast<float4*>(p.v + row * p.stride + c);
        float4 m = *pm, v = *pv;
        adam_update4(wv, m, v, gw, lr_t, p.b1, p.b2, p.eps);
        *reinterpret_cast<float4*>(prow + c) = wv; *pm = m; *pv = v;
      }
    }
  }
}

// runs longer than kFusedShortRun go on the long-run list (cut into chunk items); done BEFORE the short-run
// kernel so that the chunk kernel (side stream) and the short-run kernel are independent and overlap
__global__ void __launch_bounds__(256) fm_fused_classify_kernel(const FusedParams p) {
  const int n_unique = *p.n_unique;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_unique; u += gridDim.x * blockDim.x) {
    const int len = p.seg_start[u + 1] - p.seg_start[u];
    if (len > kFusedShortRun) {
      const int nch = (len + kFusedChunk - 1) / kFusedChunk;
      const int slot = atomicAdd(&p.counters[0], 1);
      const int base = atomicAdd(&p.counters[1], nch);
      if (slot < p.max_long && base + nch <= p.max_items) {
        p.long_runs[slot] = FusedLong{u, base, nch, 0};
        for (int c = 0; c < nch; ++c) p.items[base + c] = make_int2(slot, c);
      }
    }
  }
}

constexpr int kFusedSolo = 8;      // runs up to this length: one lane group; kFusedSolo < len <= kFusedShortRun: the whole warp

template <int LPR>
__global__ void __launch_bounds__(256, 4) fm_fused_short_kernel(const FusedParams p) {
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const int n_unique = *p.n_unique;
  const float lr_t = p.d_lr_t ? *p.d_lr_t : p.lr_t;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  // warp-uniform trip count: the 9..64-occurrence tier below needs every lane of the warp
  for (long long u0 = warp_global * GPW; u0 < n_unique; u0 += nwarps * GPW) {
    const long long u = u0 + g;
    int s0 = 0, s1 = 0;
    long long row = 0;
    RowState st;
    st.var = st.m = st.v = make_float4(0.f, 0.f, 0.f, 0.f);
    st.wx = 0.f;
    if (u < n_unique) {
      s0 = p.seg_start[u];
      s1 = p.seg_start[u + 1];
      if (s1 - s0 <= kFusedShortRun) {               // longer runs are on the long-run list (fm_fused_classify_kernel)
        row = p.unique_ids[u];
        st = fused_load_row<LPR>(p, row, gl);
      }
    }
    const int len = s1 - s0;
    const bool mine = len > 0 && len <= kFusedShortRun;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float sum_g = 0.f;
    if (mine && len <= kFusedSolo) fused_accumulate<2>(p, s0, s1, gl, acc, sum_g);
    // middle tier: a run of 9..64 occurrences would hold the other GPW-1 groups of the warp up for len/2 dependent
    // round trips; instead every group takes each GPW-th occurrence and a fixed shuffle tree combines them
    unsigned coop = __ballot_sync(0xffffffffu, mine && len > kFusedSolo && gl == 0);
    while (coop) {
      const int src = __ffs(coop) - 1;               // lane LPR*gg of the owning group
      coop &= coop - 1;
      const int a = __shfl_sync(0xffffffffu, s0, src), b = __shfl_sync(0xffffffffu, s1, src);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      float ts = 0.f;
      fused_accumulate_strided(p, a + g, b, GPW, gl, t, ts);
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
        t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
      }
      if (src / LPR == g) { acc = t; sum_g = ts; }
    }
    if (mine) fused_finish_row<LPR>(p, u, row, gl, g, st, acc, sum_g, lr_t);
  }
}

// Record layout (round 2): for k = 16 the FM row and its two Adam slots live interleaved in ONE 256-byte record
// [var 0..19 | m 20..39 | v 40..59 | pad] (include/etr.h, ETR_TABLE_RECORD).  The kernels of this file address var, m and
// v through their own pointers with the record stride, so they run unchanged on either layout (the record layout alone
// takes the c2 apply from 172 to 145 us: a row's state is two lines instead of twelve scattered chunks).  The
// occurrence-parallel kernel that streams records through the copy engine is csrc/fm_fused_flat.cu; a row-parallel
// copy-engine pipeline was measured at 200-290 us (profiles/r02_mb_apply.md) and dropped.

// one CTA per chunk item: (256/LPR) lane groups each reduce a contiguous sub-range in order
template <int LPR>
__global__ void __launch_bounds__(256) fm_fused_chunk_kernel(const FusedParams p) {
  constexpr int NG = 256 / LPR;
  __shared__ float4 sm[NG][LPR];
  __shared__ float sg[NG];
  const int gl = threadIdx.x % LPR, g = threadIdx.x / LPR;
  int n_items = p.counters[1];
  if (n_items > p.max_items) n_items = p.max_items;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int2 item = p.items[it];
    const FusedLong lr = p.long_runs[item.x];
    const int s0 = p.seg_start[lr.u] + item.y * kFusedChunk;
    int s1 = p.seg_start[lr.u + 1];
    if (s1 > s0 + kFusedChunk) s1 = s0 + kFusedChunk;
    const int per = (s1 - s0 + NG - 1) / NG;
    int a = s0 + g * per, b = a + per;
    if (a > s1) a = s1;
    if (b > s1) b = s1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float sum_g = 0.f;
    fused_accumulate<4>(p, a, b, gl, acc, sum_g);
    sm[g][gl] = acc;
    if (gl == 0) sg[g] = sum_g;
    __syncthreads();
    if (g == 0) {
      float4 t = sm[0][gl];
      float ts = sg[0];
      for (int q = 1; q < NG; ++q) {
        const float4 r = sm[q][gl];
        t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
        ts += sg[q];
      }
      float* dst = p.partials + (long long)it * p.stride;
      *reinterpret_cast<float4*>(dst + gl * 4) = t;
      if (gl == 0) dst[p.k] = ts;
    }
    __syncthreads();
  }
}

template <int LPR>
__global__ void __launch_bounds__(256) fm_fused_combine_kernel(const FusedParams p) {
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const float lr_t = p.d_lr_t ? *p.d_lr_t : p.lr_t;
  int n_long = p.counters[0];
  if (n_long > p.max_long) n_long = p.max_long;
  const long long group_global = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + g;
  const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
  for (long long r = group_global; r < n_long; r += ngroups) {
    const FusedLong lr = p.long_runs[r];
    const long long row = p.unique_ids[lr.u];
    const RowState st = fused_load_row<LPR>(p, row, gl);
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    float ts = 0.f;
    for (int c = 0; c < lr.nchunks; ++c) {
      const float* src = p.partials + (long long)(lr.base + c) * p.stride;
      const float4 x = *reinterpret_cast<const float4*>(src + gl * 4);
      t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
      ts += src[p.k];
    }
    fused_finish_row<LPR>(p, lr.u, row, gl, g, st, t, ts, lr_t);
  }
}

}  // namespace etr

using namespace etr;



static int fused_impl(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, int32_t k, int32_t fields,
                      int64_t batch, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                      const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t n_slots,
                      const float* d_dlogit, const float* d_sumv, const void* d_dflat, int32_t flat_dtype,
                      int64_t flat_ld, int32_t flat_col0, float lr_t, const float* d_lr_t, float beta1,
                      float beta2, float eps, int32_t apply, float* d_unique_grad, const int32_t* d_slot_of_u, int32_t cap,
                      float* const* h_grads_mb, const void* d_prep, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_sorted_bag && d_seg_start && d_unique_ids && d_n_unique && d_dlogit &&
                    d_sumv, "NULL argument");
  ETR_CHECK_ARG(!apply || (d_m && d_v), "Adam slots missing");
  ETR_CHECK_ARG(apply || d_unique_grad || d_slot_of_u, "nothing to do: apply == 0 and no gradient output");
  if (table->dtype != ETR_F32) { etr_set_error("etr_fm_fused_backward_apply: fp32 tables only"); return ETR_EUNSUPPORTED; }
  const int lpr = k / 4;
  if (k % 4 != 0 || lpr < 1 || lpr > 32 || (lpr & (lpr - 1)) != 0 || table->width != k + 1 || table->stride % 4 != 0 ||
      table->stride < k + 4) {
    etr_set_error("etr_fm_fused_backward_apply: needs k in {4,8,16,32,64,128} and a [V, k+1] table with the w column "
                  "in its own 16-byte chunk (k=%d width=%d stride=%d)", k, table->width, table->stride);
    return ETR_EUNSUPPORTED;
  }
  ETR_CHECK_ARG(fields > 0 && batch >= 0 && (long long)fields * batch == n_slots, "n_slots must equal batch*fields (single-hot)");
  if (d_dflat) {
    const int osz = flat_dtype == ETR_BF16 ? 2 : 4;
    ETR_CHECK_ARG((flat_col0 * osz) % (4 * osz) == 0 && (flat_ld * osz) % (4 * osz) == 0 && (k * osz) % (4 * osz) == 0 &&
                      ((uintptr_t)d_dflat % (4 * osz)) == 0, "dflat must be aligned for 4-element vector loads");
  }
  if (n_slots == 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.table = (float*)table->d_data; p.m = d_m; p.v = d_v; p.stride = table->stride; p.k = k; p.F = fields;
  // exact division of a 31-bit bag index by F: shift = 32 + ceil(log2 F), magic = ceil(2^shift / F)
  int lg = 0;
  while ((1 << lg) < fields) ++lg;
  p.shift = 32 + lg;
  p.magic = ((1ull << p.shift) + (unsigned long long)fields - 1) / (unsigned long long)fields;
  p.sorted_bag = d_sorted_bag; p.seg_start = d_seg_start; p.unique_ids = (const long long*)d_unique_ids; p.n_unique = d_n_unique;
  p.dlogit = d_dlogit; p.sumv = d_sumv; p.dflat = d_dflat; p.flat_bf16 = flat_dtype == ETR_BF16; p.flat_ld = flat_ld;
  p.flat_col0 = flat_col0; p.lr_t = lr_t; p.d_lr_t = d_lr_t; p.b1 = beta1; p.b2 = beta2; p.eps = eps; p.apply = apply;
  p.unique_grad = d_unique_grad;
  // gradient rows are densely packed (20 floats) whenever the table rows are 256-byte records -- a local record
  // table, or a peer-sharded one whose shards hold records (stride 64)
  p.gld = (table->reserved == ETR_TABLE_RECORD || (table->reserved > 0 && table->stride == 64 && table->width <= 20))
              ? ETR_RECORD_ROW_FLOATS : table->stride;
  p.slot_of_u = d_slot_of_u; p.cap = cap;
  p.world = 1;
  if (table->reserved > 0) {
    ETR_CHECK_ARG(table->reserved <= ctx->n_shard_sets, "unknown shard set");
    ETR_CHECK_ARG(!apply, "a peer-sharded table is updated by its owner (push the exported rows, apply there)");
    const EtrShardSet& ss = ctx->shard_sets[table->reserved - 1];
    p.world = ss.world;
    p.sharded = 1;
    for (int g = 0; g < ss.world; ++g) p.shard_base[g] = ss.base[g];
  }
  if (d_slot_of_u) {
    ETR_CHECK_ARG(h_grads_mb && cap > 0 && p.sharded && !apply, "mailbox export needs a peer-sharded table, apply == 0");
    for (int g = 0; g < p.world; ++g) { ETR_CHECK_ARG(h_grads_mb[g] != nullptr, "NULL mailbox pointer"); p.grads_mb[g] = h_grads_mb[g]; }
  }
  // hot configuration (Adam apply on a local RECORD table, k = 16, no / bf16 dflat) -> tiled kernels (csrc/fm_fused_tile.cu);
  // ETR_FUSED_APPLY=rows keeps the row-parallel kernels below
  {
    static int use_tile = -1;
    if (use_tile < 0) { const char* e = getenv("ETR_FUSED_APPLY"); use_tile = !(e && strcmp(e, "rows") == 0); }
    const bool rec_tab = table->reserved == ETR_TABLE_RECORD && table->stride == 64 && d_m == (float*)table->d_data + 20 &&
                         d_v == (float*)table->d_data + 40;
    const bool df_ok = !d_dflat || (flat_dtype == ETR_BF16 && flat_col0 % 4 == 0 && flat_ld % 4 == 0);
    const bool tile_apply = apply && rec_tab && !d_unique_grad && !d_slot_of_u && !p.sharded;
    const bool tile_push = !apply && d_slot_of_u && p.sharded && !d_unique_grad && p.gld == 20;
    if (use_tile && (tile_apply || tile_push) && lpr == 4 && df_ok && batch * 16 < 0x7fffffffLL && n_slots < 0x7fffffffLL) {
      TileArgs a;
      memset(&a, 0, sizeof(a));
      a.table = p.table; a.sorted_bag = p.sorted_bag; a.seg_start = p.seg_start; a.unique_ids = p.unique_ids; a.n_unique = p.n_unique;
      a.dlogit = p.dlogit; a.sumv = p.sumv;
      a.dflat = d_dflat ? reinterpret_cast<const __nv_bfloat16*>(d_dflat) + flat_col0 : nullptr;
      a.flat_ld = flat_ld; a.F = p.F; a.magic = p.magic; a.shift = p.shift;
      a.lr_t = lr_t; a.d_lr_t = d_lr_t; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.n_slots = n_slots;
      if (tile_push) {
        a.slot_of_u = d_slot_of_u; a.cap = cap; a.gld = p.gld;
        for (int g = 0; g < p.world; ++g) a.grads_mb[g] = p.grads_mb[g];
      }
      return fused_tile_launch(ctx, a, d_prep, s);
    }
    if (d_prep && apply) { etr_set_error("etr_fm_fused_backward_apply_prepared: needs an Adam apply on a local RECORD table (k = 16, bf16 dflat)"); return ETR_EUNSUPPORTED; }
  }
  p.max_long = (int)(n_slots / kFusedShortRun + 1);
  p.max_items = (int)(n_slots / kFusedChunk + n_slots / kFusedShortRun + 2);
  auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_cnt = 256, b_runs = a256(sizeof(FusedLong) * (size_t)p.max_long), b_items = a256(sizeof(int2) * (size_t)p.max_items),
               b_part = a256(sizeof(float) * (size_t)p.max_items * p.stride);
  int st = etr_ws_reserve(ctx, b_cnt + b_runs + b_items + b_part);
  if (st != ETR_OK) return st;
  char* ws = (char*)ctx->d_ws;
  p.counters = (int*)ws;
  p.long_runs = (FusedLong*)(ws + b_cnt);
  p.items = (int2*)(ws + b_cnt + b_runs);
  p.partials = (float*)(ws + b_cnt + b_runs + b_items);
  ETR_CUDA(cudaMemsetAsync(p.counters, 0, 2 * sizeof(int), s));
  const int gshort = grid_for(n_slots, 8 * (32 / lpr), ctx->sm_count, 16);
  const int gchunk = ctx->sm_count * 4;
  // classify -> { short-run kernel on the caller's stream || chunk kernel on the ctx side stream } -> combine
  fm_fused_classify_kernel<<<grid_for(n_slots, 256, ctx->sm_count, 4), 256, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  ETR_CUDA(cudaEventRecord(ctx->ev_fork, s));
  ETR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
#define ETR_FUSED(LPR)                                                      \
  do {                                                                      \
    fm_fused_chunk_kernel<LPR><<<gchunk, 256, 0, ctx->side>>>(p);           \
    ETR_LAUNCH_CHECK(ctx);                                                  \
    ETR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));                     \
    fm_fused_short_kernel<LPR><<<gshort, 256, 0, s>>>(p);                   \
    ETR_LAUNCH_CHECK(ctx);                                                  \
    ETR_CUDA(cudaStreamWaitEvent(s, ctx->ev_join, 0));                      \
    fm_fused_combine_kernel<LPR><<<ctx->sm_count, 256, 0, s>>>(p);          \
    ETR_LAUNCH_CHECK(ctx);                                                  \
  } while (0)
  switch (lpr) {
    case 1: ETR_FUSED(1); break;
    case 2: ETR_FUSED(2); break;
    case 4: ETR_FUSED(4); break;
    case 8: ETR_FUSED(8); break;
    case 16: ETR_FUSED(16); break;
    default: ETR_FUSED(32); break;
  }
#undef ETR_FUSED
  return ETR_OK;
}

extern "C" {

int etr_fm_fused_backward_apply(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, int32_t k, int32_t fields,
                                int64_t batch, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                                const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t n_slots,
                                const float* d_dlogit, const float* d_sumv, const void* d_dflat, int32_t flat_dtype,
                                int64_t flat_ld, int32_t flat_col0, float lr_t, const float* d_lr_t, float beta1,
                                float beta2, float eps, int32_t apply, float* d_unique_grad, void* stream) {
  return fused_impl(ctx, table, d_m, d_v, k, fields, batch, d_sorted_bag, d_seg_start, d_unique_ids, d_n_unique, n_slots,
                    d_dlogit, d_sumv, d_dflat, flat_dtype, flat_ld, flat_col0, lr_t, d_lr_t, beta1, beta2, eps, apply,
                    d_unique_grad, nullptr, 0, nullptr, nullptr, stream);
}

int64_t etr_fm_fused_prepare_bytes(int64_t n_slots) { return n_slots > 0 ? (int64_t)fused_tile_prep_bytes(n_slots) : 256; }

int etr_fm_fused_prepare(etr_ctx* ctx, const int32_t* d_seg_start, const int64_t* d_unique_ids, const int32_t* d_n_unique,
                         int64_t n_slots, void* d_prep, int64_t prep_bytes, void* stream) {
  ETR_CHECK_ARG(ctx && d_seg_start && d_unique_ids && d_n_unique && d_prep, "NULL argument");
  ETR_CHECK_ARG(n_slots >= 0 && n_slots < 0x7fffffffLL, "n_slots out of range");
  ETR_CHECK_ARG(prep_bytes >= etr_fm_fused_prepare_bytes(n_slots) && ((uintptr_t)d_prep & 255) == 0,
                "d_prep must hold etr_fm_fused_prepare_bytes(n_slots) bytes, 256-byte aligned");
  if (n_slots == 0) return ETR_OK;
  return fused_tile_prepare(ctx, d_seg_start, (const long long*)d_unique_ids, d_n_unique, n_slots, d_prep, (cudaStream_t)stream);
}

int etr_fm_fused_backward_apply_prepared(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, int32_t k, int32_t fields,
                                         int64_t batch, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                                         const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t n_slots,
                                         const float* d_dlogit, const float* d_sumv, const void* d_dflat, int32_t flat_dtype,
                                         int64_t flat_ld, int32_t flat_col0, float lr_t, const float* d_lr_t, float beta1,
                                         float beta2, float eps, const void* d_prep, int64_t prep_bytes, void* stream) {
  ETR_CHECK_ARG(d_prep && prep_bytes >= etr_fm_fused_prepare_bytes(n_slots), "d_prep missing or too small");
  return fused_impl(ctx, table, d_m, d_v, k, fields, batch, d_sorted_bag, d_seg_start, d_unique_ids, d_n_unique, n_slots,
                    d_dlogit, d_sumv, d_dflat, flat_dtype, flat_ld, flat_col0, lr_t, d_lr_t, beta1, beta2, eps, 1,
                    nullptr, nullptr, 0, nullptr, d_prep, stream);
}

int etr_fm_fused_backward_push(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t fields, int64_t batch,
                               const int32_t* d_sorted_bag, const int32_t* d_seg_start, const int64_t* d_unique_ids,
                               const int32_t* d_n_unique, int64_t n_slots, const float* d_dlogit, const float* d_sumv,
                               const void* d_dflat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                               const int32_t* d_slot_of_u, int32_t cap, float* const* h_grads_mb, const void* d_prep,
                               int64_t prep_bytes, void* stream) {
  ETR_CHECK_ARG(d_slot_of_u && h_grads_mb, "NULL argument");
  ETR_CHECK_ARG(!d_prep || prep_bytes >= etr_fm_fused_prepare_bytes(n_slots), "d_prep too small");
  return fused_impl(ctx, table, nullptr, nullptr, k, fields, batch, d_sorted_bag, d_seg_start, d_unique_ids, d_n_unique,
                    n_slots, d_dlogit, d_sumv, d_dflat, flat_dtype, flat_ld, flat_col0, 0.f, nullptr, 0.f, 0.f, 0.f, 0,
                    nullptr, d_slot_of_u, cap, h_grads_mb, d_prep, stream);
}

}  // extern "C"
