// K0/K1/K2: input assembly, embedding gather + bag pooling + FM terms (forward
// and backward).  HBM-bound: every table row is fetched exactly once per pass
// with 128-bit loads by a group of LPR lanes (LPR*CPL 16-byte chunks cover one
// row), pooled/FM-reduced in registers and never written back unless a GEMM
// operand ("flat") is requested.
//
// Reference arithmetic replaced: 2.FM/CustomLayers.py:138-155 (FM),
// :280-300 (DeepFM gather/Flatten), 3.DCN/CustomLayers.py:240-259 (DCN input).
#include "etr_common.cuh"

#include <stdlib.h>
#include <type_traits>

namespace etr {

struct GatherParams {
  const char* table;
  long long rows;
  int row_bytes;   // stride * esize
  int nchunks;     // 16-byte chunks holding the first `width` columns
  int k;
  int has_w;
  const long long* ids;
  const int* csr;
  long long B;
  int F, L;
  long long sb, sf, sl;
  long long pad;
  int has_pad;
  int mean;
  const float* bias;
  float* logit;
  float* prob;
  float* sumv;
  void* flat;
  int flat_bf16;
  long long flat_ld;
  int flat_col0;
  int flat_vec;    // 1: 16B/8B vector stores are aligned
  // optional dense ("continuous") columns written in front of the flattened
  // embedding: flat[b, j] = 0 for j < col0-cont_n, cont[b, j-(col0-cont_n)] after
  const float* cont;
  long long cont_sb, cont_sc;
  int cont_n;
  int fill_front;  // 1: this launch owns columns [0, flat_col0) of d_flat
  // tiled kernels: the fused w column sits in its own chunk behind a power-of-two number of
  // embedding chunks -> the lane group covers only the embedding and lane 0 fetches w with one
  // extra 4-byte (2-byte for bf16) load, so no lane idles
  int w_extra;
  int esize;
  int l1_alloc;    // 1: row loads allocate in L1 (hot rows of small fields hit there); 0: L1::no_allocate
  // row-sharded table over peer memory (tiled forward kernel only): shard g holds rows g, g+G, ...
  const char* shard_base[16];
  int world;
  // backward only
  const float* dlogit;
  float* bag_grad;
  int grad_ld;
  unsigned long long* err;
};

// rows in flight per lane: keep <= 32 fp32 registers of row data
template <int CPL, int VEC>
struct UnrollFor {
  static constexpr int value = (CPL * VEC >= 16) ? 2 : ((CPL * VEC >= 8) ? 4 : 8);
};

template <typename Elem>
__device__ __forceinline__ void store_flat(const GatherParams& p, long long b, int f, int col,
                                           const float* v) {
  constexpr int VEC = Chunk<Elem>::kElems;
  if (col >= p.k) return;
  const long long e0 = b * p.flat_ld + p.flat_col0 + (long long)f * p.k + col;
  const bool full = (col + VEC <= p.k) && p.flat_vec;
  if (!p.flat_bf16) {
    float* o = reinterpret_cast<float*>(p.flat) + e0;
    if (full) {
#pragma unroll
      for (int i = 0; i < VEC; i += 4) stg_stream16(o + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (col + i < p.k) o[i] = v[i];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.flat) + e0;
    if (full) {
      uint32_t w[VEC / 2];
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      if (VEC == 4) {
        *reinterpret_cast<uint2*>(o) = make_uint2(w[0], w[1]);
      } else {
        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[VEC / 2 - 2], w[VEC / 2 - 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (col + i < p.k) o[i] = __float2bfloat16_rn(v[i]);
    }
  }
}

// Pools bag (b,f) into e[CPL][VEC]; returns the number of valid ids.
template <typename Elem, int LPR, int CPL>
__device__ __forceinline__ int pool_bag(const GatherParams& p, long long b, int f, int gl, bool active,
                                        float (&e)[CPL][Chunk<Elem>::kElems]) {
  constexpr int VEC = Chunk<Elem>::kElems;
  constexpr int kUnroll = UnrollFor<CPL, VEC>::value;
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int i = 0; i < VEC; ++i) e[j][i] = 0.f;
  if (!active) return 0;
  const long long* base;
  long long step;
  int n;
  if (p.csr) {
    const int o0 = p.csr[b * p.F + f];
    n = p.csr[b * p.F + f + 1] - o0;
    base = p.ids + o0;
    step = 1;
  } else {
    base = p.ids + b * p.sb + (long long)f * p.sf;
    step = p.sl;
    n = p.L;
  }
  int cnt = 0;
  for (int l0 = 0; l0 < n; l0 += kUnroll) {
    long long id[kUnroll];
    bool ok[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const bool in = (l0 + u) < n;
      id[u] = in ? __ldg(base + (long long)(l0 + u) * step) : 0;
      ok[u] = in && !(p.has_pad && id[u] == p.pad);
      if (ok[u] && (unsigned long long)id[u] >= (unsigned long long)p.rows) {
        flag_bad_id(p.err, id[u]);
        ok[u] = false;
      }
    }
    Chunk<Elem> r[kUnroll][CPL];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = gl + j * LPR;
        if (ok[u] && c < p.nchunks)
          r[u][j].load(p.table + id[u] * (long long)p.row_bytes + c * 16);
        else
          r[u][j].zero();
      }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      cnt += ok[u] ? 1 : 0;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i) e[j][i] += r[u][j].v[i];
    }
  }
  if (p.mean && cnt > 1) {
    const float inv = 1.0f / (float)cnt;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int i = 0; i < VEC; ++i) e[j][i] *= inv;
  }
  return cnt;
}

// ---------------------------------------------------------------- forward
// BAG=false: single-hot [B,F] fast path -- kUnroll fields in flight per lane.
template <typename Elem, int LPR, int CPL, bool BAG>
__global__ void __launch_bounds__(256) gather_fm_fwd_kernel(const GatherParams p) {
  constexpr int VEC = Chunk<Elem>::kElems;
  constexpr int kUnroll = UnrollFor<CPL, VEC>::value;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const int g = lane / LPR;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const bool want_fm = (p.logit != nullptr) || (p.prob != nullptr) || (p.sumv != nullptr);

  for (long long b0 = warp_global * GPW; b0 < p.B; b0 += nwarps * GPW) {
    const long long b = b0 + g;
    const bool active = b < p.B;
    float S[CPL][VEC], Q[CPL][VEC];
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int i = 0; i < VEC; ++i) S[j][i] = Q[j][i] = 0.f;

    if (p.fill_front && active) {
      const int z = p.flat_col0 - p.cont_n;
      for (int j = gl; j < p.flat_col0; j += LPR) {
        const float v = (j < z) ? 0.f : p.cont[b * p.cont_sb + (long long)(j - z) * p.cont_sc];
        if (p.flat_bf16) reinterpret_cast<__nv_bfloat16*>(p.flat)[b * p.flat_ld + j] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(p.flat)[b * p.flat_ld + j] = v;
      }
    }

    if (!BAG) {
      for (int f0 = 0; f0 < p.F; f0 += kUnroll) {
        long long id[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const bool in = active && (f0 + u) < p.F;
          id[u] = in ? __ldg(p.ids + b * p.sb + (long long)(f0 + u) * p.sf) : 0;
          ok[u] = in && !(p.has_pad && id[u] == p.pad);
          if (ok[u] && (unsigned long long)id[u] >= (unsigned long long)p.rows) {
            flag_bad_id(p.err, id[u]);
            ok[u] = false;
          }
        }
        Chunk<Elem> r[kUnroll][CPL];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const int c = gl + j * LPR;
            if (ok[u] && c < p.nchunks)
              r[u][j].load(p.table + id[u] * (long long)p.row_bytes + c * 16);
            else
              r[u][j].zero();
          }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (f0 + u < p.F) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
#pragma unroll
              for (int i = 0; i < VEC; ++i) {
                const float x = r[u][j].v[i];
                S[j][i] += x;
                Q[j][i] += x * x;
              }
              if (p.flat && active) store_flat<Elem>(p, b, f0 + u, (gl + j * LPR) * VEC, r[u][j].v);
            }
          }
        }
      }
    } else {
      for (int f = 0; f < p.F; ++f) {
        float e[CPL][VEC];
        pool_bag<Elem, LPR, CPL>(p, b, f, gl, active, e);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            S[j][i] += e[j][i];
            Q[j][i] += e[j][i] * e[j][i];
          }
          if (p.flat && active) store_flat<Elem>(p, b, f, (gl + j * LPR) * VEC, e[j]);
        }
      }
    }

    if (want_fm) {
      // second = 0.5 * sum_{c<k} (S_c^2 - Q_c); first = S_k (the fused w column)
      float second = 0.f, first = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const int col = (gl + j * LPR) * VEC + i;
          if (col < p.k) {
            second += S[j][i] * S[j][i] - Q[j][i];
            if (p.sumv && active) p.sumv[b * p.k + col] = S[j][i];
          } else if (col == p.k && p.has_w) {
            first = S[j][i];
          }
        }
      second = group_sum<LPR>(second);
      first = group_sum<LPR>(first);
      if (gl == 0 && active) {
        const float z = ((p.bias ? p.bias[0] : 0.f) + first) + 0.5f * second;
        if (p.logit) p.logit[b] = z;
        if (p.prob) p.prob[b] = sigmoidf_exact(z);
      }
    }
  }
}

// --------------------------------------------------------------- backward
// bag_grad[(b*F+f), c] = dlogit*(S_c - e_fc) (+ dflat)  for c<k ; dlogit at c==k.
template <typename Elem, int CPL>
__device__ __forceinline__ void emit_bag_grad(const GatherParams& p, long long b, int f, int gl, int LPR_,
                                              const float (&S)[CPL][Chunk<Elem>::kElems],
                                              const float (&e)[CPL][Chunk<Elem>::kElems], float dl, float scale,
                                              bool need_rows) {
  constexpr int VEC = Chunk<Elem>::kElems;
  const int grad_chunks = p.grad_ld / 4;
  float* grow = p.bag_grad + (b * p.F + f) * (long long)p.grad_ld;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = gl + j * LPR_;
    float gout[VEC];
    const int col0 = c * VEC;
    float df[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) df[i] = 0.f;
    if (p.flat && col0 < p.k) {
      const long long e0 = b * p.flat_ld + p.flat_col0 + (long long)f * p.k + col0;
      if (p.flat_vec && col0 + VEC <= p.k && !p.flat_bf16) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
          const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.flat) + e0 + i);
          df[i] = t.x; df[i + 1] = t.y; df[i + 2] = t.z; df[i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (col0 + i < p.k)
            df[i] = p.flat_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.flat)[e0 + i])
                                : reinterpret_cast<const float*>(p.flat)[e0 + i];
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = col0 + i;
      float gv = 0.f;
      if (col < p.k) {
        if (need_rows) gv = dl * (S[j][i] - e[j][i]);
        gv += df[i];
      } else if (col == p.k && p.has_w) {
        gv = dl;
      }
      gout[i] = gv * scale;
    }
#pragma unroll
    for (int q = 0; q < VEC / 4; ++q) {
      const int gc = c * (VEC / 4) + q;
      if (gc < grad_chunks && (!p.w_extra || gc * 4 < p.k))
        stg_stream16(grow + gc * 4, make_float4(gout[4 * q], gout[4 * q + 1], gout[4 * q + 2], gout[4 * q + 3]));
    }
  }
  if (p.w_extra && gl == 0) {
    // the w column's chunk (and any padding chunk behind it): [dlogit, 0, 0, 0]
    for (int gc = p.k / 4; gc < grad_chunks; ++gc)
      stg_stream16(grow + gc * 4, make_float4(gc * 4 == p.k ? dl * scale : 0.f, 0.f, 0.f, 0.f));
  }
}

template <typename Elem, int LPR, int CPL, bool BAG>
__global__ void __launch_bounds__(256) gather_fm_bwd_kernel(const GatherParams p) {
  constexpr int VEC = Chunk<Elem>::kElems;
  constexpr int kUnroll = UnrollFor<CPL, VEC>::value;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const int g = lane / LPR;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const bool need_rows = p.dlogit != nullptr;

  for (long long b0 = warp_global * GPW; b0 < p.B; b0 += nwarps * GPW) {
    const long long b = b0 + g;
    const bool active = b < p.B;
    float S[CPL][VEC];
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int i = 0; i < VEC; ++i) S[j][i] = 0.f;
    const float dl = (need_rows && active) ? p.dlogit[b] : 0.f;

    if (!BAG) {
      // ---- single-hot: kUnroll fields in flight per lane in both passes
      if (need_rows) {
        for (int f0 = 0; f0 < p.F; f0 += kUnroll) {
          long long id[kUnroll];
          bool ok[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const bool in = active && (f0 + u) < p.F;
            id[u] = in ? __ldg(p.ids + b * p.sb + (long long)(f0 + u) * p.sf) : 0;
            ok[u] = in && !(p.has_pad && id[u] == p.pad) && (unsigned long long)id[u] < (unsigned long long)p.rows;
          }
          Chunk<Elem> r[kUnroll][CPL];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              const int c = gl + j * LPR;
              if (ok[u] && c < p.nchunks) r[u][j].load(p.table + id[u] * (long long)p.row_bytes + c * 16);
              else r[u][j].zero();
            }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int j = 0; j < CPL; ++j)
#pragma unroll
              for (int i = 0; i < VEC; ++i) S[j][i] += r[u][j].v[i];
        }
      }
      for (int f0 = 0; f0 < p.F; f0 += kUnroll) {
        Chunk<Elem> r[kUnroll][CPL];
        if (need_rows) {
          long long id[kUnroll];
          bool ok[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const bool in = active && (f0 + u) < p.F;
            id[u] = in ? __ldg(p.ids + b * p.sb + (long long)(f0 + u) * p.sf) : 0;
            ok[u] = in && !(p.has_pad && id[u] == p.pad) && (unsigned long long)id[u] < (unsigned long long)p.rows;
          }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              const int c = gl + j * LPR;
              if (ok[u] && c < p.nchunks) r[u][j].load(p.table + id[u] * (long long)p.row_bytes + c * 16);
              else r[u][j].zero();
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (active && f0 + u < p.F) {
            float e[CPL][VEC];
#pragma unroll
            for (int j = 0; j < CPL; ++j)
#pragma unroll
              for (int i = 0; i < VEC; ++i) e[j][i] = need_rows ? r[u][j].v[i] : 0.f;
            emit_bag_grad<Elem, CPL>(p, b, f0 + u, gl, LPR, S, e, dl, 1.0f, need_rows);
          }
        }
      }
    } else {
      if (need_rows) {
        for (int f = 0; f < p.F; ++f) {
          float e[CPL][VEC];
          pool_bag<Elem, LPR, CPL>(p, b, f, gl, active, e);
#pragma unroll
          for (int j = 0; j < CPL; ++j)
#pragma unroll
            for (int i = 0; i < VEC; ++i) S[j][i] += e[j][i];
        }
      }
      for (int f = 0; f < p.F; ++f) {
        float e[CPL][VEC];
        int cnt = 1;
        if (need_rows) {
          cnt = pool_bag<Elem, LPR, CPL>(p, b, f, gl, active, e);   // rows now come from L2
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j)
#pragma unroll
            for (int i = 0; i < VEC; ++i) e[j][i] = 0.f;
          if (p.mean && active) {                                    // only the bag count is needed
            cnt = 0;
            if (p.csr) {
              cnt = p.csr[b * p.F + f + 1] - p.csr[b * p.F + f];
            } else {
              for (int l = 0; l < p.L; ++l) {
                const long long id = __ldg(p.ids + b * p.sb + (long long)f * p.sf + (long long)l * p.sl);
                cnt += (p.has_pad && id == p.pad) ? 0 : 1;
              }
            }
          }
        }
        if (!active) continue;
        const float scale = (p.mean && cnt > 1) ? 1.0f / (float)cnt : 1.0f;
        emit_bag_grad<Elem, CPL>(p, b, f, gl, LPR, S, e, dl, scale, need_rows);
      }
    }
  }
}

// ================================================================ tiled single-hot kernels
// The common case (single-hot ids, row <= 32 chunks).  A CTA works on tiles of
// SPB = 8 warps * (32/LPR) samples.  The ids of the NEXT tile are brought into
// shared memory with cp.async while the rows of the current tile are being
// fetched, so a lane's row loads never wait on an id load (the id -> row
// dependent-load chain is what bounds the simple kernel), and the register
// budget is capped so that 3 CTAs (24 warps) stay resident per SM.
template <bool PEER>
__device__ __forceinline__ float4 ldg_row16_x(const void* q, int l1_alloc) {
  if (PEER || l1_alloc) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(q));
    return r;
  }
  return ldg_row16(q);
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int SPB>
__device__ __forceinline__ void stage_ids_tile(const GatherParams& p, long long tile, long long* dst) {
  const long long b0 = tile * SPB;
  const int total = SPB * p.F;
  const bool field_major = (p.sb == 1);
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int sl, f;
    if (field_major) { sl = e % SPB; f = e / SPB; } else { f = e % p.F; sl = e / p.F; }
    const long long b = b0 + sl;
    if (b < p.B) cp_async8(dst + sl * p.F + f, p.ids + b * p.sb + (long long)f * p.sf);
  }
}

template <typename Elem, int LPR>
__global__ void __launch_bounds__(256, 3) gather_fm_fwd_tile_kernel(const GatherParams p) {
  constexpr int VEC = Chunk<Elem>::kElems;
  constexpr int U = (VEC >= 8) ? 4 : 8;
  constexpr int GPW = 32 / LPR;
  constexpr int SPB = 8 * GPW;
  extern __shared__ __align__(16) long long ids_s[];          // [2][SPB * F]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, g = lane / LPR;
  const int sl = warp * GPW + g;
  const long long ntiles = (p.B + SPB - 1) / SPB;
  const bool want_fm = (p.logit != nullptr) || (p.prob != nullptr) || (p.sumv != nullptr);
  const int c = gl;
  const bool has_chunk = c < p.nchunks;
  int buf = 0;
  long long tile = blockIdx.x;
  if (tile < ntiles) stage_ids_tile<SPB>(p, tile, ids_s);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  for (; tile < ntiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    if (next < ntiles) stage_ids_tile<SPB>(p, next, ids_s + (buf ^ 1) * SPB * p.F);
    cp_async_commit();

    const long long b = tile * SPB + sl;
    const bool active = b < p.B;
    const long long* my = ids_s + buf * SPB * p.F + sl * p.F;
    float S[VEC], Q[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) S[i] = Q[i] = 0.f;
    if (p.fill_front && active) {
      const int z = p.flat_col0 - p.cont_n;
      for (int j = gl; j < p.flat_col0; j += LPR) {
        const float v = (j < z) ? 0.f : p.cont[b * p.cont_sb + (long long)(j - z) * p.cont_sc];
        if (p.flat_bf16) reinterpret_cast<__nv_bfloat16*>(p.flat)[b * p.flat_ld + j] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(p.flat)[b * p.flat_ld + j] = v;
      }
    }
    float wsum = 0.f;
    for (int f0 = 0; f0 < p.F; f0 += U) {
      Chunk<Elem> r[U];
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bool ok = active && (f0 + u) < p.F;
        long long id = ok ? my[f0 + u] : 0;
        ok = ok && !(p.has_pad && id == p.pad);
        if (ok && (unsigned long long)id >= (unsigned long long)p.rows) { flag_bad_id(p.err, id); ok = false; }
        const char* rowp = p.table + id * (long long)p.row_bytes;
        if (p.world > 1) {                       // NVLink peer memory: the owner's shard, local row id / G
          const long long lrow = id / p.world;
          rowp = p.shard_base[(int)(id - lrow * p.world)] + lrow * (long long)p.row_bytes;
        }
        if (ok && has_chunk) {
          if (p.l1_alloc) {
            const float4 t = ldg_row16_x<false>(rowp + c * 16, 1);
            if constexpr (sizeof(Elem) == 4) { r[u].v[0] = t.x; r[u].v[1] = t.y; r[u].v[2] = t.z; r[u].v[3] = t.w; }
            else r[u].load(rowp + c * 16);
          } else {
            r[u].load(rowp + c * 16);
          }
        } else {
          r[u].zero();
        }
        wv[u] = 0.f;
        if (p.w_extra && gl == 0 && ok) {
          if (sizeof(Elem) == 4) wv[u] = __ldg(reinterpret_cast<const float*>(rowp) + p.k);
          else wv[u] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rowp)[p.k]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (f0 + u < p.F) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float x = r[u].v[i];
            S[i] += x;
            Q[i] += x * x;
          }
          wsum += wv[u];
          if (p.flat && active) store_flat<Elem>(p, b, f0 + u, c * VEC, r[u].v);
        }
      }
    }
    if (want_fm) {
      float second = 0.f, first = wsum;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int col = c * VEC + i;
        if (col < p.k) {
          second += S[i] * S[i] - Q[i];
          if (p.sumv && active) p.sumv[b * p.k + col] = S[i];
        } else if (col == p.k && p.has_w) {
          first = S[i];
        }
      }
      second = group_sum<LPR>(second);
      first = group_sum<LPR>(first);
      if (gl == 0 && active) {
        const float z = ((p.bias ? p.bias[0] : 0.f) + first) + 0.5f * second;
        if (p.logit) p.logit[b] = z;
        if (p.prob) p.prob[b] = sigmoidf_exact(z);
      }
    }
    cp_async_wait_all();
    __syncthreads();
    buf ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Lean form of the tiled forward kernel for the FM family's hot configuration (round 2): fp32 rows, k = 16 with the
// fused w column at 16, single-hot ids, no padding, local table.  The generic tiled kernel above re-evaluates every
// option of GatherParams per row (ncu, r01_prof_gather_fwd: 26.4 M warp instructions for 1.70 M rows = 15.5 per row,
// of which IMAD 20.8 %, ISETP 19.2 %, BRA 8.4 % and only LDG 2.8 %, FADD + FFMA 7.7 %: it is ISSUE-bound at 45 % of
// the issue slots with 6 warps per scheduler, not memory-bound).  Here everything that does not depend on the row is
// a template parameter or hoisted out of the field loop: per row a thread executes one shared-memory id read, a range
// check, one 64-bit multiply-add, one 128-bit load (lane 0: + the 4-byte w), 4 FADD + 4 FFMA, and (FLAT) two packs + one
// 8-byte store.  FLAT: 0 no flattened operand, 1 bf16, 2 fp32.
template <int FLAT, int OCC>
__global__ void __launch_bounds__(256, OCC) gather_fm_fwd_lean_kernel(const GatherParams p) {
  constexpr int SPB = 64, U = 8;
  extern __shared__ __align__(16) long long ids_s[];          // [2][SPB * F]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane & 3, g = lane >> 2;
  const int sl = warp * 8 + g;
  const int F = p.F;
  const long long ntiles = (p.B + SPB - 1) / SPB;
  const unsigned long long rows = (unsigned long long)p.rows;
  const long long row_bytes = p.row_bytes;
  const char* const tbase = p.table + c * 16;
  const float bias = p.bias ? p.bias[0] : 0.f;
  int buf = 0;
  long long tile = blockIdx.x;
  if (tile < ntiles) stage_ids_tile<SPB>(p, tile, ids_s);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  for (; tile < ntiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    if (next < ntiles) stage_ids_tile<SPB>(p, next, ids_s + (buf ^ 1) * SPB * F);
    cp_async_commit();

    const long long b = tile * SPB + sl;
    const bool active = b < p.B;
    const long long* my = ids_s + buf * SPB * F + sl * F;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.f;
    if (active) {
      if (p.fill_front) {
        const int z = p.flat_col0 - p.cont_n;
        for (int j = c; j < p.flat_col0; j += 4) {
          const float v = (j < z) ? 0.f : p.cont[b * p.cont_sb + (long long)(j - z) * p.cont_sc];
          if (FLAT == 1) reinterpret_cast<__nv_bfloat16*>(p.flat)[b * p.flat_ld + j] = __float2bfloat16_rn(v);
          else if (FLAT == 2) reinterpret_cast<float*>(p.flat)[b * p.flat_ld + j] = v;
        }
      }
      // this lane's 4 columns of field 0 in the flattened operand; field f is 16 elements further
      char* const obase = FLAT == 1 ? reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(p.flat) + b * p.flat_ld + p.flat_col0 + c * 4)
                                    : reinterpret_cast<char*>(reinterpret_cast<float*>(p.flat) + b * p.flat_ld + p.flat_col0 + c * 4);
      for (int f0 = 0; f0 < F; f0 += U) {
        float4 r[U];
        float wv[U];
        const int nu = F - f0 < U ? F - f0 : U;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          wv[u] = 0.f;
          if (u < nu) {
            const unsigned long long id = (unsigned long long)my[f0 + u];
            if (id < rows) {
              const char* rp = tbase + id * row_bytes;
              r[u] = ldg_row16(rp);
              if (c == 0) wv[u] = __ldg(reinterpret_cast<const float*>(rp) + 16);
            } else {
              flag_bad_id(p.err, (long long)id);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u < nu) {
            S.x += r[u].x; S.y += r[u].y; S.z += r[u].z; S.w += r[u].w;
            Q.x += r[u].x * r[u].x; Q.y += r[u].y * r[u].y; Q.z += r[u].z * r[u].z; Q.w += r[u].w * r[u].w;
            wsum += wv[u];
            if (FLAT == 1) {
              const __nv_bfloat162 lo = __floats2bfloat162_rn(r[u].x, r[u].y), hi = __floats2bfloat162_rn(r[u].z, r[u].w);
              uint2 w2;
              w2.x = *reinterpret_cast<const uint32_t*>(&lo);
              w2.y = *reinterpret_cast<const uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(obase + (f0 + u) * 32) = w2;
            } else if (FLAT == 2) {
              stg_stream16(obase + (f0 + u) * 64, r[u]);
            }
          }
        }
      }
      if (p.sumv) *reinterpret_cast<float4*>(p.sumv + b * 16 + c * 4) = S;
    }
    float second = (S.x * S.x - Q.x) + (S.y * S.y - Q.y) + (S.z * S.z - Q.z) + (S.w * S.w - Q.w);
    second = group_sum<4>(second);
    if (c == 0 && active) {
      const float z = (bias + wsum) + 0.5f * second;
      if (p.logit) p.logit[b] = z;
      if (p.prob) p.prob[b] = sigmoidf_exact(z);
    }
    cp_async_wait_all();
    __syncthreads();
    buf ^= 1;
  }
}

template <typename Elem, int LPR>
__global__ void __launch_bounds__(256, 3) gather_fm_bwd_tile_kernel(const GatherParams p) {
  constexpr int VEC = Chunk<Elem>::kElems;
  constexpr int U = (VEC >= 8) ? 4 : 8;
  constexpr int GPW = 32 / LPR;
  constexpr int SPB = 8 * GPW;
  extern __shared__ __align__(16) long long ids_s[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, g = lane / LPR;
  const int sl = warp * GPW + g;
  const long long ntiles = (p.B + SPB - 1) / SPB;
  const bool need_rows = p.dlogit != nullptr;
  const int c = gl;
  const bool has_chunk = c < p.nchunks;
  int buf = 0;
  long long tile = blockIdx.x;
  if (need_rows) {
    if (tile < ntiles) stage_ids_tile<SPB>(p, tile, ids_s);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
  }
  for (; tile < ntiles; tile += gridDim.x) {
    if (need_rows) {
      const long long next = tile + gridDim.x;
      if (next < ntiles) stage_ids_tile<SPB>(p, next, ids_s + (buf ^ 1) * SPB * p.F);
      cp_async_commit();
    }
    const long long b = tile * SPB + sl;
    const bool active = b < p.B;
    const long long* my = ids_s + buf * SPB * p.F + sl * p.F;
    float S[1][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) S[0][i] = 0.f;
    const float dl = (need_rows && active) ? p.dlogit[b] : 0.f;
    if (need_rows) {
      for (int f0 = 0; f0 < p.F; f0 += U) {
        Chunk<Elem> r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          bool ok = active && (f0 + u) < p.F;
          const long long id = ok ? my[f0 + u] : 0;
          ok = ok && !(p.has_pad && id == p.pad) && (unsigned long long)id < (unsigned long long)p.rows;
          if (ok && has_chunk) r[u].load(p.table + id * (long long)p.row_bytes + c * 16);
          else r[u].zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < VEC; ++i) S[0][i] += r[u].v[i];
      }
    }
    for (int f0 = 0; f0 < p.F; f0 += U) {
      Chunk<Elem> r[U];
      if (need_rows) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          bool ok = active && (f0 + u) < p.F;
          const long long id = ok ? my[f0 + u] : 0;
          ok = ok && !(p.has_pad && id == p.pad) && (unsigned long long)id < (unsigned long long)p.rows;
          if (ok && has_chunk) r[u].load(p.table + id * (long long)p.row_bytes + c * 16);   // L2 hits now
          else r[u].zero();
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (active && f0 + u < p.F) {
          float e[1][VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) e[0][i] = need_rows ? r[u].v[i] : 0.f;
          emit_bag_grad<Elem, 1>(p, b, f0 + u, gl, LPR, S, e, dl, 1.0f, need_rows);
        }
      }
    }
    if (need_rows) {
      cp_async_wait_all();
      __syncthreads();
      buf ^= 1;
    }
  }
}

// ------------------------------------------------------------ row dump
template <typename Elem>
__global__ void __launch_bounds__(256) embedding_gather_kernel(const char* table, long long rows, int row_bytes,
                                                               int width, const long long* ids, long long count,
                                                               float* out, long long out_ld,
                                                               unsigned long long* err) {
  constexpr int VEC = Chunk<Elem>::kElems;
  const int nchunks = (width + VEC - 1) / VEC;
  const long long total = count * nchunks;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long n = t / nchunks;
    const int c = (int)(t % nchunks);
    const long long id = ids[n];
    Chunk<Elem> r;
    if ((unsigned long long)id >= (unsigned long long)rows) {
      flag_bad_id(err, id);
      r.zero();
    } else {
      r.load(table + id * (long long)row_bytes + c * 16);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      if (c * VEC + i < width) out[n * out_ld + c * VEC + i] = r.v[i];
  }
}

// ================================================================ streaming single-hot forward (K1)
// The hot case: single-hot ids without padding and an embedding of a power-of-two number KCH of
// 16-byte chunks; for the FM family the fused w column is the first element of chunk KCH.
// Warps are independent (no shared memory, no CTA barrier): a warp owns tiles of GPW = 32/LPR
// samples, one lane group per sample.  Per batch of U fields
//   * the ids of the NEXT batch are already in flight (each lane loads the ids of "its" fields,
//     checks their range once and keeps them as 32-bit row numbers), so no row load ever waits on
//     an id load;
//   * a lane group broadcasts a row number with one shuffle and issues ONE 128-bit load
//     instruction per row: lanes 0..KCH-1 take the embedding chunks and (HAS_W) lane KCH takes the
//     chunk that holds w.  Measured on B200 (scripts/mb_gather.py, profiles/r01_mb_gather.md): a
//     random row costs the same ~30 ns/1000 whether it is 32, 64 or 128 bytes as long as it is ONE
//     request inside one 128-byte line, while a second request to the same row (the separate
//     4-byte w load of the first version) doubled the kernel time on a DRAM-resident table;
//   * all U loads of the batch are in flight before the first one is consumed.
// About 20 instructions per (sample, field) instead of ~100 in the generic tiled kernel (ncu r01:
// that one was issue-bound, 51 % issue-active at 30 % DRAM).  Out-of-range ids are flagged in the
// error word and read row 0 (the caller raises; TF-CPU semantics).
// peer rows bypass the local L2 (only L1 caches them): let the hot head of a Zipf id space live in L1
template <bool PEER>
__device__ __forceinline__ float4 ldg_row16_p(const void* q) {
  if (!PEER) return ldg_row16(q);
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(q));
  return r;
}

template <typename Elem, int KCH, int U, int FLAT, bool PEER, bool HAS_W>
__global__ void __launch_bounds__(128, 3) gather_fm_fwd_stream_kernel(const GatherParams p, const int wshift) {
  constexpr int VEC = Chunk<Elem>::kElems;
  constexpr int LPR = KCH + (HAS_W ? 1 : 0);        // lanes per row: the embedding chunks (+ the w chunk);
                                                    // 5 for the k=16 fp32 FM row -> 6 rows per instruction
  constexpr int K = KCH * VEC;                      // embedding columns
  constexpr int GPW = 32 / LPR;                     // samples per warp tile (lanes >= GPW*LPR idle)
  constexpr int IPL = (U + LPR - 1) / LPR;          // ids owned per lane per batch
  constexpr int OSZ = FLAT == 2 ? 2 : 4;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const bool in_group = g < GPW;
  const bool is_v = in_group && gl < KCH;           // this lane holds embedding columns
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ntiles = (p.B + GPW - 1) / GPW;
  const int nbatches = (p.F + U - 1) / U;
  const bool want_fm = (p.logit != nullptr) || (p.prob != nullptr) || (p.sumv != nullptr);
  const unsigned row_bytes = (unsigned)p.row_bytes;
  const char* lane_base = p.table + gl * 16;        // (unused for peer tables)
  long long bad_id = 0;                             // an out-of-range id this lane saw
  bool any_bad = false;

  auto chunk_addr = [&](unsigned rid) -> const char* {       // this lane's 16-byte chunk of row `rid`
    if constexpr (PEER) {
      const unsigned owner = rid & (unsigned)(p.world - 1);
      return p.shard_base[owner] + gl * 16 + (unsigned long long)(rid >> wshift) * row_bytes;
    } else {
      return lane_base + (unsigned long long)rid * row_bytes;
    }
  };
  auto load_ids = [&](long long t, int f0, unsigned (&row)[IPL]) {
    const long long b = t * GPW + g;
#pragma unroll
    for (int j = 0; j < IPL; ++j) {
      const int fi = j * LPR + gl;
      const int f = f0 + fi;
      long long id = 0;
      if (fi < U && f < p.F && b < p.B && in_group) id = __ldg(p.ids + b * p.sb + (long long)f * p.sf);
      const bool bad = (unsigned long long)id >= (unsigned long long)p.rows;
      bad_id = bad ? id : bad_id;
      any_bad = any_bad || bad;
      row[j] = bad ? 0u : (unsigned)id;
    }
  };

  unsigned cur[IPL], nxt[IPL];
#pragma unroll
  for (int j = 0; j < IPL; ++j) cur[j] = nxt[j] = 0u;
  if (tile < ntiles) load_ids(tile, 0, cur);
  for (; tile < ntiles; tile += nwarps) {
    const long long b = tile * GPW + g;
    const bool active = b < p.B && in_group;
    float S[VEC], Q[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) S[i] = Q[i] = 0.f;
    if (FLAT && p.fill_front && active) {
      const int z = p.flat_col0 - p.cont_n;
#pragma unroll 1
      for (int j = gl; j < p.flat_col0; j += LPR) {
        const float v = (j < z) ? 0.f : __ldg(p.cont + b * p.cont_sb + (long long)(j - z) * p.cont_sc);
        if (FLAT == 2) reinterpret_cast<__nv_bfloat16*>(p.flat)[b * p.flat_ld + j] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(p.flat)[b * p.flat_ld + j] = v;
      }
    }
    char* frow = nullptr;                             // this lane's chunk of field 0 in the flat operand
    if (FLAT) frow = reinterpret_cast<char*>(p.flat) + (b * p.flat_ld + p.flat_col0 + gl * VEC) * OSZ;

    // FULL: all U fields of the batch exist and every sample of the tile is in range -> no predicates
    auto batch = [&](auto full_tag, const int f0) {
      constexpr bool FULL = decltype(full_tag)::value;
      float4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned rid = __shfl_sync(0xffffffffu, cur[u / LPR], g * LPR + (u % LPR));
        if (in_group) r[u] = ldg_row16_x<PEER>(chunk_addr(rid), p.l1_alloc);   // a missing field carries row 0
      }
      char* fo = FLAT ? frow + (long long)f0 * (K * OSZ) : nullptr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (FULL || f0 + u < p.F) {
          float x[VEC];
          if constexpr (sizeof(Elem) == 4) {
            x[0] = r[u].x; x[1] = r[u].y; x[2] = r[u].z; x[3] = r[u].w;
          } else {
            const uint32_t h[4] = {__float_as_uint(r[u].x), __float_as_uint(r[u].y), __float_as_uint(r[u].z),
                                   __float_as_uint(r[u].w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[2 * i] = __uint_as_float(h[i] << 16);               // low half  = element 2i
              x[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u);   // high half = element 2i+1
            }
          }
          if (want_fm) {                      // lane KCH accumulates sum w in S[0]; its other sums are unused
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              S[i] += x[i];
              Q[i] += x[i] * x[i];
            }
          }
          if (FLAT && (FULL || active) && (!HAS_W || is_v)) {
            char* o = fo + u * (K * OSZ);
            if constexpr (FLAT == 2 && sizeof(Elem) == 2) {
              *reinterpret_cast<float4*>(o) = r[u];                       // bf16 row -> bf16 operand: raw copy
            } else if constexpr (FLAT == 2) {
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]);
              const __nv_bfloat162 h1 = __floats2bfloat162_rn(x[2], x[3]);
              *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0),
                                                        *reinterpret_cast<const uint32_t*>(&h1));
            } else {
#pragma unroll
              for (int i = 0; i < VEC; i += 4)
                stg_stream16(o + i * 4, make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]));
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < IPL; ++j) cur[j] = nxt[j];
    };
    const bool tile_full = (tile + 1) * GPW <= p.B;
#pragma unroll 1
    for (int jb = 0; jb < nbatches; ++jb) {
      const int f0 = jb * U;
      if (jb + 1 < nbatches) load_ids(tile, f0 + U, nxt);
      else if (tile + nwarps < ntiles) load_ids(tile + nwarps, 0, nxt);
      if (tile_full && f0 + U <= p.F) batch(std::true_type{}, f0);
      else batch(std::false_type{}, f0);
    }

    if (want_fm) {
      float second = 0.f;
      if (!HAS_W || is_v) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) second += S[i] * S[i] - Q[i];
        if (p.sumv && active) {
#pragma unroll
          for (int i = 0; i < VEC; i += 4)
            *reinterpret_cast<float4*>(p.sumv + b * K + gl * VEC + i) = make_float4(S[i], S[i + 1], S[i + 2], S[i + 3]);
        }
      }
      // the group is not a power of two wide: fetch the KCH partial sums by lane and add them in the
      // order of a butterfly reduction (bit-identical to the tiled kernel's group_sum)
      float part[KCH];
#pragma unroll
      for (int c = 0; c < KCH; ++c) part[c] = __shfl_sync(0xffffffffu, second, g * LPR + c);
#pragma unroll
      for (int o = KCH / 2; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < o; ++c) part[c] += part[c + o];
      second = part[0];
      const float first = HAS_W ? __shfl_sync(0xffffffffu, S[0], g * LPR + KCH) : 0.f;
      if (gl == 0 && active) {
        const float z = ((p.bias ? p.bias[0] : 0.f) + first) + 0.5f * second;
        if (p.logit) p.logit[b] = z;
        if (p.prob) p.prob[b] = sigmoidf_exact(z);
      }
    }
  }
  if (any_bad) flag_bad_id(p.err, bad_id);
}

// ------------------------------------------------------------ K0 assemble
struct ColPtrs {
  const long long* col[64];
};
__global__ void __launch_bounds__(256) assemble_ids_kernel(const ColPtrs cols, int fields, long long batch,
                                                           long long* out) {
  const long long total = (long long)fields * batch;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(t / batch);
    const long long b = t % batch;
    out[t] = cols.col[f][b];
  }
}

// ------------------------------------------------------------ dispatch
static int fill_params(const char* fn, const etr_table* table, int k, int has_w, const etr_ids* ids,
                       GatherParams* p, etr_ctx* ctx) {
  if (!ctx || !table || !ids) { etr_set_error("%s: NULL argument", fn); return ETR_EINVAL; }
  if (!table->d_data || !ids->d_ids) { etr_set_error("%s: NULL device pointer", fn); return ETR_EINVAL; }
  const int esize = table->dtype == ETR_BF16 ? 2 : 4;
  if (table->dtype != ETR_F32 && table->dtype != ETR_BF16) { etr_set_error("%s: bad table dtype", fn); return ETR_EINVAL; }
  if (k <= 0 || k + (has_w ? 1 : 0) > table->width || table->width > table->stride) {
    etr_set_error("%s: k=%d has_w=%d does not fit table width=%d stride=%d", fn, k, has_w, table->width, table->stride);
    return ETR_EINVAL;
  }
  if ((table->stride * esize) % 16 != 0 || ((uintptr_t)table->d_data & 15)) {
    etr_set_error("%s: table rows must be 16-byte aligned (stride*esize %% 16 == 0)", fn);
    return ETR_EINVAL;
  }
  if (ids->batch < 0 || ids->fields <= 0 || (!ids->d_csr_offsets && ids->bag <= 0)) {
    etr_set_error("%s: bad ids shape", fn);
    return ETR_EINVAL;
  }
  memset(p, 0, sizeof(*p));
  p->table = (const char*)table->d_data;
  p->rows = table->rows;
  p->row_bytes = table->stride * esize;
  p->esize = esize;
  const int used = k + (has_w ? 1 : 0);
  p->nchunks = (used * esize + 15) / 16;
  p->k = k;
  p->has_w = has_w ? 1 : 0;
  p->ids = (const long long*)ids->d_ids;
  p->csr = ids->d_csr_offsets;
  p->B = ids->batch;
  p->F = ids->fields;
  p->L = ids->d_csr_offsets ? 0 : ids->bag;
  p->sb = ids->stride_b; p->sf = ids->stride_f; p->sl = ids->stride_l;
  p->pad = ids->pad_id;
  p->has_pad = ids->has_pad;
  p->mean = ids->pooling == ETR_POOL_MEAN;
  p->err = ctx->d_err;
  p->world = 1;
  if (table->reserved > 0) {
    if (table->reserved > ctx->n_shard_sets) { etr_set_error("%s: unknown shard set %d", fn, table->reserved); return ETR_EINVAL; }
    const EtrShardSet& ss = ctx->shard_sets[table->reserved - 1];
    p->world = ss.world;
    p->rows = ss.rows_global;                 // ids are GLOBAL; owner = id mod G, local row = id div G
    for (int g = 0; g < ss.world; ++g) p->shard_base[g] = ss.base[g];
  }
  return ETR_OK;
}

template <typename Elem, int KCH, int FLAT, bool PEER, bool HAS_W>
static int launch_stream_one(etr_ctx* ctx, const GatherParams& p_in, int wshift, cudaStream_t s) {
  GatherParams p = p_in;
  auto kern = gather_fm_fwd_stream_kernel<Elem, KCH, 13, FLAT, PEER, HAS_W>;
  static int occ = 0;                      // per instantiation; one device architecture (sm_100a)
  if (occ == 0) {
    ETR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0));
    if (occ < 1) occ = 1;
  }
  constexpr int GPW = 32 / (KCH + (HAS_W ? 1 : 0));
  const long long ntiles = (p.B + GPW - 1) / GPW;
  long long grid = (ntiles + 3) / 4;
  if (grid > (long long)ctx->sm_count * occ) grid = (long long)ctx->sm_count * occ;
  p.l1_alloc = 1;                          // measured: no difference on local tables; peer rows want L1
  kern<<<(int)grid, 128, 0, s>>>(p, wshift);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

// returns -1 when the shape is not one the streaming kernel covers
template <typename Elem>
static int launch_stream(etr_ctx* ctx, const GatherParams& p, bool bag, cudaStream_t s) {
  const int emb_bytes = p.k * p.esize;
  const int kch = emb_bytes / 16;
  if (bag || p.has_pad || emb_bytes % 16 != 0 || (kch & (kch - 1))) return -1;
  if (p.rows >= (1ll << 32) || (p.flat && !p.flat_vec)) return -1;
  if (p.has_w ? !(kch == 2 || kch == 4 || kch == 8) : !(kch == 2 || kch == 4 || kch == 8 || kch == 16)) return -1;
  if (p.has_w && p.row_bytes < (kch + 1) * 16) return -1;        // the w chunk must exist in the row
  int wshift = 0;
  if (p.world > 1) {
    if ((p.world & (p.world - 1)) || !p.has_w) return -1;
    while ((1 << wshift) < p.world) ++wshift;
  }
  const int flat = !p.flat ? 0 : (p.flat_bf16 ? 2 : 1);
#define ETR_STREAM_W(KCH, FL)                                                                       \
  return p.world > 1 ? launch_stream_one<Elem, KCH, FL, true, true>(ctx, p, wshift, s)             \
                     : launch_stream_one<Elem, KCH, FL, false, true>(ctx, p, wshift, s)
#define ETR_STREAM_NOW(KCH, FL) return launch_stream_one<Elem, KCH, FL, false, false>(ctx, p, wshift, s)
#define ETR_STREAM_FL(MACRO, KCH)                                                                   \
  do {                                                                                              \
    if (flat == 0) { MACRO(KCH, 0); }                                                               \
    else if (flat == 1) { MACRO(KCH, 1); }                                                          \
    else { MACRO(KCH, 2); }                                                                         \
  } while (0)
  if (p.has_w) {
    switch (kch) {
      case 2: ETR_STREAM_FL(ETR_STREAM_W, 2); break;
      case 4: ETR_STREAM_FL(ETR_STREAM_W, 4); break;
      default: ETR_STREAM_FL(ETR_STREAM_W, 8); break;
    }
  } else {
    switch (kch) {
      case 2: ETR_STREAM_FL(ETR_STREAM_NOW, 2); break;
      case 4: ETR_STREAM_FL(ETR_STREAM_NOW, 4); break;
      case 8: ETR_STREAM_FL(ETR_STREAM_NOW, 8); break;
      default: ETR_STREAM_FL(ETR_STREAM_NOW, 16); break;
    }
  }
#undef ETR_STREAM_W
#undef ETR_STREAM_NOW
#undef ETR_STREAM_FL
  return -1;
}

template <typename Elem, bool BWD>
static int launch_gather(etr_ctx* ctx, const GatherParams& p_in, bool bag, cudaStream_t s) {
  GatherParams p = p_in;
  if (!BWD) {
    // Measured on the c2 workload (scripts/mb_k1.py, profiles/r01_mb_gather.md): every variant is bound by the
    // rows it keeps in flight per SM; with the fused w column the tiled kernel below (24 warps x 8 loads) wins
    // (52 us vs 65 us streaming; a cp.async-staged variant reached 73 us and was dropped), without it the
    // streaming kernel does.  Peer-sharded tables are served by the streaming kernel.  ETR_GATHER=stream|generic
    // overrides the choice (micro-benchmarks).
    const char* which = getenv("ETR_GATHER");
    const bool want_stream = which ? strcmp(which, "stream") == 0 : (!p.has_w || p.world > 1);
    if (want_stream) {
      const int st = launch_stream<Elem>(ctx, p, bag, s);
      if (st >= 0) return st;
    }
  }
  auto lanes_for = [](int nchunks) {
    int l = 1;
    while (l < nchunks && l < 32) l <<= 1;
    return l;
  };
  p.l1_alloc = 0;
  // single-hot + fused w in a chunk of its own behind a power-of-two number of embedding chunks:
  // lane groups cover the embedding only (no idle lanes), lane 0 fetches w separately.  Only the
  // tiled kernels implement this, so it is enabled only when they will be used.
  {
    const int emb_bytes = p.k * p.esize;
    const int emb_chunks = emb_bytes / 16;
    if (!bag && p.has_w && emb_bytes % 16 == 0 && emb_chunks >= 1 && emb_chunks <= 32 &&
        (emb_chunks & (emb_chunks - 1)) == 0) {
      const size_t smem = 2 * (size_t)(8 * (32 / emb_chunks)) * p.F * sizeof(long long);
      if (smem <= 64 * 1024) {
        p.w_extra = 1;
        p.nchunks = emb_chunks;
      }
    }
  }
  // lanes per row: smallest power of two covering the chunks, at most 32; then CPL
  const int lpr = lanes_for(p.nchunks);
  const int cpl = (p.nchunks + lpr - 1) / lpr;
  if (cpl > 4) {
    etr_set_error("gather: row of %d 16-byte chunks is wider than the kernels cover (2 KiB)", p.nchunks);
    return ETR_EUNSUPPORTED;
  }
  const int gpw = 32 / lpr;
  const int threads = 256;
  // tiled single-hot kernels: ids double-buffered in shared memory
  const size_t tile_smem = 2 * (size_t)(8 * gpw) * p.F * sizeof(long long);
  if (p.world > 1 && (BWD || bag || cpl != 1 || tile_smem > 64 * 1024)) {
    etr_set_error("gather: peer-sharded tables are served by the tiled single-hot forward kernel only");
    return ETR_EUNSUPPORTED;
  }
  if (!bag && cpl == 1 && tile_smem <= 64 * 1024) {
    const int tgrid = grid_for(p.B, 8 * gpw, ctx->sm_count, 3);
    // the FM family's hot configuration takes the lean kernel (ETR_GATHER=generic keeps the generic tiled one)
    const char* which_k = getenv("ETR_GATHER");
    const bool flat_ok = !p.flat || (p.flat_vec && p.flat_col0 % 4 == 0 && p.flat_ld % 4 == 0);
    if (!BWD && sizeof(Elem) == 4 && p.k == 16 && p.has_w && p.w_extra && lpr == 4 && !p.has_pad && p.world <= 1 &&
        flat_ok && p.row_bytes >= 68 && !(which_k && strcmp(which_k, "generic") == 0)) {
      static int occ = 0;
      if (!occ) { const char* e = getenv("ETR_K1_OCC"); occ = (e && e[0] == '3') ? 3 : 4; }
      const int lgrid = grid_for(p.B, 64, ctx->sm_count, occ);
#define ETR_LEAN(FL, OC)                                                                                               \
  do {                                                                                                                 \
    if (tile_smem > 48 * 1024)                                                                                         \
      cudaFuncSetAttribute(gather_fm_fwd_lean_kernel<FL, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem); \
    gather_fm_fwd_lean_kernel<FL, OC><<<lgrid, threads, tile_smem, s>>>(p);                                           \
  } while (0)
      const int fl = !p.flat ? 0 : (p.flat_bf16 ? 1 : 2);
      if (occ == 3) { if (fl == 0) ETR_LEAN(0, 3); else if (fl == 1) ETR_LEAN(1, 3); else ETR_LEAN(2, 3); }
      else { if (fl == 0) ETR_LEAN(0, 4); else if (fl == 1) ETR_LEAN(1, 4); else ETR_LEAN(2, 4); }
#undef ETR_LEAN
      ETR_LAUNCH_CHECK(ctx);
      return ETR_OK;
    }
#define ETR_TILE(LPR)                                                                                   \
  do {                                                                                                  \
    if (BWD) {                                                                                          \
      if (tile_smem > 48 * 1024)                                                                        \
        cudaFuncSetAttribute(gather_fm_bwd_tile_kernel<Elem, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem); \
      gather_fm_bwd_tile_kernel<Elem, LPR><<<tgrid, threads, tile_smem, s>>>(p);                        \
    } else {                                                                                            \
      if (tile_smem > 48 * 1024)                                                                        \
        cudaFuncSetAttribute(gather_fm_fwd_tile_kernel<Elem, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem); \
      gather_fm_fwd_tile_kernel<Elem, LPR><<<tgrid, threads, tile_smem, s>>>(p);                        \
    }                                                                                                   \
  } while (0)
    switch (lpr) {
      case 1: ETR_TILE(1); break;
      case 2: ETR_TILE(2); break;
      case 4: ETR_TILE(4); break;
      case 8: ETR_TILE(8); break;
      case 16: ETR_TILE(16); break;
      default: ETR_TILE(32); break;
    }
#undef ETR_TILE
    ETR_LAUNCH_CHECK(ctx);
    return ETR_OK;
  }
  const int grid = grid_for(p.B, (threads / 32) * gpw, ctx->sm_count, 8);
#define ETR_LAUNCH(LPR, CPL)                                                              \
  do {                                                                                    \
    if (BWD && bag)                                                                       \
      gather_fm_bwd_kernel<Elem, LPR, CPL, true><<<grid, threads, 0, s>>>(p);             \
    else if (BWD)                                                                         \
      gather_fm_bwd_kernel<Elem, LPR, CPL, false><<<grid, threads, 0, s>>>(p);            \
    else if (bag)                                                                         \
      gather_fm_fwd_kernel<Elem, LPR, CPL, true><<<grid, threads, 0, s>>>(p);             \
    else                                                                                  \
      gather_fm_fwd_kernel<Elem, LPR, CPL, false><<<grid, threads, 0, s>>>(p);            \
  } while (0)
  if (cpl == 1) {
    switch (lpr) {
      case 1: ETR_LAUNCH(1, 1); break;
      case 2: ETR_LAUNCH(2, 1); break;
      case 4: ETR_LAUNCH(4, 1); break;
      case 8: ETR_LAUNCH(8, 1); break;
      case 16: ETR_LAUNCH(16, 1); break;
      default: ETR_LAUNCH(32, 1); break;
    }
  } else if (cpl == 2) {
    ETR_LAUNCH(32, 2);
  } else {
    ETR_LAUNCH(32, 4);
  }
#undef ETR_LAUNCH
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // namespace etr

using namespace etr;

extern "C" {

int etr_assemble_ids(etr_ctx* ctx, const int64_t* const* h_cols, int32_t fields, int64_t batch,
                     int64_t* d_out, void* stream) {
  ETR_CHECK_ARG(ctx && h_cols && d_out, "NULL argument");
  ETR_CHECK_ARG(fields > 0 && fields <= 64, "fields must be in [1,64] per call");
  ETR_CHECK_ARG(batch >= 0, "negative batch");
  if (batch == 0) return ETR_OK;
  ColPtrs cols;
  for (int f = 0; f < fields; ++f) {
    ETR_CHECK_ARG(h_cols[f] != nullptr, "NULL column pointer");
    cols.col[f] = (const long long*)h_cols[f];
  }
  const int grid = grid_for((long long)fields * batch, 256 * 4, ctx->sm_count, 8);
  assemble_ids_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cols, fields, batch, (long long*)d_out);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_gather_fm_forward(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t has_w,
                          const etr_ids* ids, const float* d_bias, float* d_logit, float* d_prob,
                          float* d_sumv, void* d_flat, int32_t flat_dtype, int64_t flat_ld,
                          int32_t flat_col0, const float* d_cont, int32_t cont_n, int64_t cont_stride_b,
                          int64_t cont_stride_c, void* stream) {
  GatherParams p;
  int st = fill_params(__func__, table, k, has_w, ids, &p, ctx);
  if (st != ETR_OK) return st;
  if (p.B == 0) return ETR_OK;
  p.bias = d_bias; p.logit = d_logit; p.prob = d_prob; p.sumv = d_sumv;
  p.flat = d_flat;
  p.flat_bf16 = flat_dtype == ETR_BF16;
  p.flat_ld = flat_ld;
  p.flat_col0 = flat_col0;
  if (d_flat) {
    ETR_CHECK_ARG(flat_ld >= flat_col0 + (int64_t)p.F * k, "flat_ld too small");
    // one lane stores VEC output elements per chunk: alignment min(16, VEC*osz) bytes
    const int vec = table->dtype == ETR_BF16 ? 8 : 4;
    const int osz = p.flat_bf16 ? 2 : 4;
    const int align = vec * osz > 16 ? 16 : vec * osz;
    p.flat_vec = ((flat_col0 * osz) % align == 0) && ((flat_ld * osz) % align == 0) &&
                 ((k * osz) % align == 0) && (((uintptr_t)d_flat & 15) == 0);
  }
  // front columns [0, flat_col0): cont_n < 0 leaves them to the caller; otherwise the
  // kernel writes zeros to [0, flat_col0-cont_n) and the dense features after them.
  if (d_flat && flat_col0 > 0 && cont_n >= 0) {
    ETR_CHECK_ARG(cont_n <= flat_col0, "cont_n > flat_col0");
    ETR_CHECK_ARG(cont_n == 0 || d_cont != nullptr, "d_cont is NULL");
    p.fill_front = 1;
    p.cont = d_cont; p.cont_n = cont_n; p.cont_sb = cont_stride_b; p.cont_sc = cont_stride_c;
  }
  const bool bag = ids->d_csr_offsets != nullptr || ids->bag != 1;
  if (table->dtype == ETR_BF16)
    return launch_gather<__nv_bfloat16, false>(ctx, p, bag, (cudaStream_t)stream);
  return launch_gather<float, false>(ctx, p, bag, (cudaStream_t)stream);
}

int etr_gather_fm_backward(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t has_w,
                           const etr_ids* ids, const float* d_dlogit, const void* d_dflat,
                           int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0, float* d_bag_grad,
                           int32_t grad_ld, void* stream) {
  GatherParams p;
  int st = fill_params(__func__, table, k, has_w, ids, &p, ctx);
  if (st != ETR_OK) return st;
  if (p.B == 0) return ETR_OK;
  ETR_CHECK_ARG(d_bag_grad != nullptr, "d_bag_grad is NULL");
  ETR_CHECK_ARG(d_dlogit || d_dflat, "need d_dlogit and/or d_dflat");
  ETR_CHECK_ARG(grad_ld % 4 == 0 && grad_ld >= k + (has_w ? 1 : 0), "grad_ld must be a multiple of 4 and >= k+has_w");
  ETR_CHECK_ARG(((uintptr_t)d_bag_grad & 15) == 0, "d_bag_grad must be 16-byte aligned");
  p.dlogit = d_dlogit;
  p.flat = const_cast<void*>(d_dflat);
  p.flat_bf16 = flat_dtype == ETR_BF16;
  p.flat_ld = flat_ld;
  p.flat_col0 = flat_col0;
  p.bag_grad = d_bag_grad;
  p.grad_ld = grad_ld;
  if (d_dflat) {
    ETR_CHECK_ARG(flat_ld >= flat_col0 + (int64_t)p.F * k, "flat_ld too small");
    p.flat_vec = !p.flat_bf16 && (flat_col0 % 4 == 0) && (flat_ld % 4 == 0) && (k % 4 == 0) &&
                 (((uintptr_t)d_dflat & 15) == 0);
  }
  // the gradient row may be wider (in chunks) than the table row for bf16 tables:
  // lanes are assigned by table chunks, each producing VEC fp32 values.
  const bool bag = ids->d_csr_offsets != nullptr || ids->bag != 1;
  if (table->dtype == ETR_BF16)
    return launch_gather<__nv_bfloat16, true>(ctx, p, bag, (cudaStream_t)stream);
  return launch_gather<float, true>(ctx, p, bag, (cudaStream_t)stream);
}

int etr_embedding_gather(etr_ctx* ctx, const etr_table* table, const int64_t* d_ids, int64_t count,
                         float* d_out, int64_t out_ld, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_ids && d_out, "NULL argument");
  ETR_CHECK_ARG(out_ld >= table->width, "out_ld < width");
  const int esize = table->dtype == ETR_BF16 ? 2 : 4;
  ETR_CHECK_ARG((table->stride * esize) % 16 == 0 && ((uintptr_t)table->d_data & 15) == 0,
                "table rows must be 16-byte aligned");
  if (count == 0) return ETR_OK;
  const int vec = 16 / esize;
  const int nchunks = (table->width + vec - 1) / vec;
  const int grid = grid_for(count * nchunks, 256, ctx->sm_count, 8);
  if (table->dtype == ETR_BF16)
    embedding_gather_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const char*)table->d_data, table->rows, table->stride * esize, table->width,
        (const long long*)d_ids, count, d_out, out_ld, ctx->d_err);
  else
    embedding_gather_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const char*)table->d_data, table->rows, table->stride * esize, table->width,
        (const long long*)d_ids, count, d_out, out_ld, ctx->d_err);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
