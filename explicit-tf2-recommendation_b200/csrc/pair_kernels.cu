// K3: field-aware pair interaction (FFM / FwFM / FieldAwareInteractionLayer)
// K4: PNN inner / outer product networks.
//
// K3 replaces 2.FM/CustomLayers.py:438-461 (GatherV2 on [V,F,k] + transpose +
// band-part mask + boolean_mask, ~6 [B,F,F,k] temporaries) and the heads at
// :490-495 / :526-531.  One CTA per sample: every warp pools whole bags of
// F*k(+1)-wide rows straight from HBM with 128-bit loads into a shared-memory
// tile E[F][F*k] (the only copy of the gathered rows), then the P = F(F-1)/2
// pair products are formed from shared memory and reduced in the same kernel.
#include "etr_common.cuh"

namespace etr {

struct PairParams {
  const char* table; long long rows; int row_bytes; int nchunks;   // fp32 table
  int F, k, has_w;
  const long long* ids; const int* csr; long long B; int L; long long sb, sf, sl; long long pad; int has_pad; int mean;
  const float* bias; const float* r; const float* r0;
  float* pairvec; float* pairdot; float* logit; float* prob; float* pooled;
  int es;            // smem row stride in floats ( >= F*k+1, == 4 mod 32 )
  unsigned long long* err;
  // backward
  const float* dlogit; const float* dpairvec; float* bag_grad; int grad_ld;
};

__device__ __forceinline__ void pair_from_index(int p, int F, int& a, int& c) {
  // row-major strict upper triangle: rows a contribute F-1-a pairs
  int rem = p;
  a = 0;
  while (rem >= F - 1 - a) { rem -= F - 1 - a; ++a; }
  c = a + 1 + rem;
}

__device__ __forceinline__ int bag_count(const PairParams& p, long long b, int f) {
  if (p.csr) return p.csr[b * p.F + f + 1] - p.csr[b * p.F + f];
  if (!p.has_pad) return p.L;
  int cnt = 0;
  for (int l = 0; l < p.L; ++l)
    cnt += (__ldg(p.ids + b * p.sb + (long long)f * p.sf + (long long)l * p.sl) == p.pad) ? 0 : 1;
  return cnt;
}

// CPL: 16-byte chunks per lane (row of up to 32*CPL chunks)
template <int CPL>
__global__ void __launch_bounds__(256, 3) field_pair_fwd_kernel(const PairParams p) {
  extern __shared__ __align__(16) float E[];          // [F][es]
  __shared__ float red[8];
  __shared__ int bag_ctr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int P = p.F * (p.F - 1) / 2;
  const int Fk = p.F * p.k;
  if (threadIdx.x == 0) bag_ctr = 0;
  __syncthreads();
  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    // ---- phase 1: pool every field's bag into E[a][:]
    // The ids of a bag are read 64 slots at a time by the whole warp (two per lane), pads / out-of-range ids drop out
    // in a ballot, and the surviving ids are handed out four at a time by shuffles: the row loads of a bag are issued
    // back to back (12 x 128-bit per lane in flight) instead of waiting for each group's ids, and padded slots cost
    // nothing.  The ids of the warp's NEXT bag are fetched before this bag's rows, so that latency is hidden too.
    // Accumulation order = slot order of the valid ids (unchanged: bit-identical sums).
    auto bag_desc = [&](int a, const long long*& base, long long& step, int& n) {
      if (p.csr) { const int o0 = p.csr[b * p.F + a]; n = p.csr[b * p.F + a + 1] - o0; base = p.ids + o0; step = 1; }
      else { base = p.ids + b * p.sb + (long long)a * p.sf; step = p.sl; n = p.L; }
    };
    auto load_ids = [&](const long long* base, long long step, int n, int s0, long long& i0, long long& i1, bool& k0, bool& k1) {
      const int s = s0 + lane;
      i0 = 0; i1 = 0; k0 = false; k1 = false;
      if (s < n) {
        i0 = __ldg(base + (long long)s * step);
        k0 = !(p.has_pad && i0 == p.pad);
        if (k0 && (unsigned long long)i0 >= (unsigned long long)p.rows) { flag_bad_id(p.err, i0); k0 = false; }
      }
      if (s + 32 < n) {
        i1 = __ldg(base + (long long)(s + 32) * step);
        k1 = !(p.has_pad && i1 == p.pad);
        if (k1 && (unsigned long long)i1 >= (unsigned long long)p.rows) { flag_bad_id(p.err, i1); k1 = false; }
      }
    };
    // bags are claimed dynamically (their lengths differ: a static split leaves warps idle at the barrier below)
    auto claim = [&]() {
      int x = 0;
      if (lane == 0) x = atomicAdd(&bag_ctr, 1);
      return __shfl_sync(0xffffffffu, x, 0);
    };
    long long ni0 = 0, ni1 = 0; bool nk0 = false, nk1 = false;
    int a = claim();
    if (a < p.F) {
      const long long* base; long long step; int n;
      bag_desc(a, base, step, n);
      load_ids(base, step, n, 0, ni0, ni1, nk0, nk1);
    }
    while (a < p.F) {
      const int a_next = claim();
      const long long* base; long long step; int n;
      bag_desc(a, base, step, n);
      float4 acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int s0 = 0; s0 < n || s0 == 0; s0 += 64) {
        long long i0, i1; bool k0, k1;
        if (s0 == 0) { i0 = ni0; i1 = ni1; k0 = nk0; k1 = nk1; }
        else load_ids(base, step, n, s0, i0, i1, k0, k1);
        if (s0 == 0 && a_next < p.F) {                 // the next bag's first 64 ids: in flight during this bag's rows
          const long long* nb; long long ns; int nn;
          bag_desc(a_next, nb, ns, nn);
          load_ids(nb, ns, nn, 0, ni0, ni1, nk0, nk1);
        }
        unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
        cnt += __popc(m0) + __popc(m1);
        // two-stage pipeline over pairs of rows: the loads of pair g + 1 are issued before pair g is accumulated
        auto issue = [&](float4 (&r)[2][CPL]) {        // pops up to two valid ids (slot order) and issues their row loads
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const bool ok = (m0 | m1) != 0;
            int src = 0; bool hi = false;
            if (m0) { src = __ffs(m0) - 1; m0 &= m0 - 1; }
            else if (m1) { src = __ffs(m1) - 1; m1 &= m1 - 1; hi = true; }
            const long long x0 = __shfl_sync(0xffffffffu, i0, src), x1 = __shfl_sync(0xffffffffu, i1, src);
            const char* row = p.table + (hi ? x1 : x0) * (long long)p.row_bytes;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              const int c = lane + j * 32;
              r[u][j] = (ok && c < p.nchunks) ? ldg_row16(row + c * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        };
        auto add = [&](const float4 (&r)[2][CPL]) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < CPL; ++j) { acc[j].x += r[u][j].x; acc[j].y += r[u][j].y; acc[j].z += r[u][j].z; acc[j].w += r[u][j].w; }
        };
        if (m0 | m1) {                                 // warp-uniform control flow throughout
          float4 rA[2][CPL], rB[2][CPL];
          issue(rA);
          while (true) {
            if (!(m0 | m1)) { add(rA); break; }
            issue(rB);
            add(rA);
            if (!(m0 | m1)) { add(rB); break; }
            issue(rA);
            add(rB);
          }
        }
      }
      const float inv = (p.mean && cnt > 1) ? 1.0f / (float)cnt : 1.0f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + j * 32;
        if (c < p.nchunks) {
          float4 v = acc[j];
          v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
          *reinterpret_cast<float4*>(&E[a * p.es + c * 4]) = v;
          if (p.pooled) {
            float* o = p.pooled + (b * p.F + a) * (long long)Fk + c * 4;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) if (c * 4 + i < Fk) o[i] = vv[i];
          }
        }
      }
      a = a_next;
    }
    __syncthreads();
    if (threadIdx.x == 0) bag_ctr = 0;                 // (visible to the next sample through the barrier that ends phase 2)
    // ---- phase 2: pairs
    float part = 0.f;
    for (int pi = threadIdx.x; pi < P; pi += blockDim.x) {
      int a, c;
      pair_from_index(pi, p.F, a, c);
      const float* ea = &E[a * p.es + c * p.k];
      const float* ec = &E[c * p.es + a * p.k];
      float dot = 0.f;
      for (int d = 0; d < p.k; ++d) {
        const float v = ea[d] * ec[d];
        dot += v;
        if (p.pairvec) p.pairvec[(b * P + pi) * (long long)p.k + d] = v;
      }
      if (p.pairdot) p.pairdot[b * P + pi] = dot;
      part += p.r ? p.r[pi] * dot : dot;
    }
    if (p.has_w)
      for (int a = threadIdx.x; a < p.F; a += blockDim.x) part += E[a * p.es + Fk];
    if (p.logit || p.prob) {
      part = group_sum<32>(part);
      if (lane == 0) red[warp] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        float z = 0.f;
        for (int q = 0; q < nwarps; ++q) z += red[q];
        z += (p.bias ? p.bias[0] : 0.f) + (p.r0 ? p.r0[0] : 0.f);
        if (p.logit) p.logit[b] = z;
        if (p.prob) p.prob[b] = sigmoidf_exact(z);
      }
    }
    __syncthreads();
  }
}

// backward: transpose-within-sample of the pooled rows, scaled by the upstream grad.
__global__ void __launch_bounds__(256) field_pair_bwd_kernel(const PairParams p) {
  extern __shared__ __align__(16) float E[];          // [F][es] then inv_cnt[F]
  const int Fk = p.F * p.k;
  const int P = p.F * (p.F - 1) / 2;
  float* inv_cnt = E + p.F * p.es;
  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    for (int t = threadIdx.x; t < p.F * Fk; t += blockDim.x) {
      const int a = t / Fk, col = t % Fk;
      E[a * p.es + col] = p.pooled[(b * p.F + a) * (long long)Fk + col];
    }
    for (int a = threadIdx.x; a < p.F; a += blockDim.x) {
      float inv = 1.0f;
      if (p.mean) {
        const int cnt = bag_count(p, b, a);
        if (cnt > 1) inv = 1.0f / (float)cnt;
      }
      inv_cnt[a] = inv;
    }
    __syncthreads();
    const float dl = p.dlogit ? p.dlogit[b] : 0.f;
    for (int t = threadIdx.x; t < p.F * p.grad_ld; t += blockDim.x) {
      const int a = t / p.grad_ld, col = t % p.grad_ld;
      float g = 0.f;
      if (col < Fk) {
        const int c = col / p.k, d = col % p.k;
        if (c != a) {
          const int lo = a < c ? a : c, hi = a < c ? c : a;
          const int pi = lo * (2 * p.F - lo - 1) / 2 + (hi - lo - 1);
          const float other = E[c * p.es + a * p.k + d];
          if (p.dpairvec) g = p.dpairvec[(b * P + pi) * (long long)p.k + d] * other;
          else g = dl * (p.r ? p.r[pi] : 1.0f) * other;
        }
      } else if (col == Fk && p.has_w) {
        g = dl;
      }
      p.bag_grad[(b * p.F + a) * (long long)p.grad_ld + col] = g * inv_cnt[a];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------- PNN
struct PnnParams {
  const float* x; long long ldx; long long B; int F, k, type;
  const float* kernel; float* out; long long ldo;
  const float* g; long long ldg; float* dx; long long lddx; float* dkernel;
};

// one warp per sample; x[b] staged in shared memory
__global__ void __launch_bounds__(256) pnn_fwd_kernel(const PnnParams p) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Fk = p.F * p.k, P = p.F * (p.F - 1) / 2;
  float* xs = sm + warp * Fk;
  for (long long b0 = (long long)blockIdx.x * nwarps; b0 < p.B; b0 += (long long)gridDim.x * nwarps) {
    const long long b = b0 + warp;
    if (b < p.B)
      for (int t = lane; t < Fk; t += 32) xs[t] = p.x[b * p.ldx + t];
    __syncwarp();
    if (b < p.B) {
      for (int pi = lane; pi < P; pi += 32) {
        int i, j;
        pair_from_index(pi, p.F, i, j);
        const float* xi = xs + i * p.k;
        const float* xj = xs + j * p.k;
        float o = 0.f;
        if (p.type == 0) {
          for (int d = 0; d < p.k; ++d) o += xi[d] * xj[d];
        } else if (p.type == 1) {            // mat: K[a,p,c]; out = sum_a x_j[a] * (sum_c x_i[c] K[a,p,c])
          for (int a = 0; a < p.k; ++a) {
            const float* Kr = p.kernel + ((long long)a * P + pi) * p.k;
            float s = 0.f;
            for (int c = 0; c < p.k; ++c) s += xi[c] * __ldg(Kr + c);
            o += s * xj[a];
          }
        } else if (p.type == 2) {            // vec: K[p,c]
          for (int c = 0; c < p.k; ++c) o += xi[c] * xj[c] * __ldg(p.kernel + (long long)pi * p.k + c);
        } else {                             // num: K[p]
          float s = 0.f;
          for (int d = 0; d < p.k; ++d) s += xi[d] * xj[d];
          o = s * __ldg(p.kernel + pi);
        }
        p.out[b * p.ldo + pi] = o;
      }
    }
    __syncwarp();
  }
}

// dx (accumulated into p.dx); the kernel gradient dK is reduced over the batch without atomics by pnn_dk_kernel
// (pair_dense.cu: one CTA per (pair, batch slice), slices added in order -- deterministic)
__global__ void __launch_bounds__(256) pnn_bwd_kernel(const PnnParams p) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Fk = p.F * p.k, P = p.F * (p.F - 1) / 2;
  float* xs = sm + warp * (Fk + P);
  float* gs = xs + Fk;
  for (long long b0 = (long long)blockIdx.x * nwarps; b0 < p.B; b0 += (long long)gridDim.x * nwarps) {
    const long long b = b0 + warp;
    if (b < p.B) {
      for (int t = lane; t < Fk; t += 32) xs[t] = p.x[b * p.ldx + t];
      for (int t = lane; t < P; t += 32) gs[t] = p.g[b * p.ldg + t];
    }
    __syncwarp();
    if (b < p.B) {
      // dx[i,d] = sum_{j != i} G[p(i,j)] * dOut_p/dx_i[d]
      for (int t = lane; t < Fk; t += 32) {
        const int i = t / p.k, d = t % p.k;
        float acc = 0.f;
        for (int j = 0; j < p.F; ++j) {
          if (j == i) continue;
          const int lo = i < j ? i : j, hi = i < j ? j : i;
          const int pi = lo * (2 * p.F - lo - 1) / 2 + (hi - lo - 1);
          const float G = gs[pi];
          const float* xo = xs + j * p.k;
          if (p.type == 0) acc += G * xo[d];
          else if (p.type == 3) acc += G * __ldg(p.kernel + pi) * xo[d];
          else if (p.type == 2) acc += G * __ldg(p.kernel + (long long)pi * p.k + d) * xo[d];
          else {
            // mat: out_p = sum_{a,c} x_lo[c] K[a,p,c] x_hi[a]
            float s = 0.f;
            if (i == lo) { for (int a = 0; a < p.k; ++a) s += __ldg(p.kernel + ((long long)a * P + pi) * p.k + d) * xo[a]; }
            else         { for (int c = 0; c < p.k; ++c) s += xo[c] * __ldg(p.kernel + ((long long)d * P + pi) * p.k + c); }
            acc += G * s;
          }
        }
        p.dx[b * p.lddx + t] += acc;
      }
    }
    __syncwarp();
  }
}

static int fill_pair(const char* fn, etr_ctx* ctx, const etr_table* table, int k, int has_w, const etr_ids* ids,
                     PairParams* p) {
  if (!ctx || !ids || !ids->d_ids) { etr_set_error("%s: NULL argument", fn); return ETR_EINVAL; }
  memset(p, 0, sizeof(*p));
  p->F = ids->fields; p->k = k; p->has_w = has_w ? 1 : 0;
  const int Fk = p->F * k;
  if (table) {
    if (!table->d_data) { etr_set_error("%s: NULL table", fn); return ETR_EINVAL; }
    if (table->dtype != ETR_F32) { etr_set_error("%s: field-pair tables are fp32", fn); return ETR_EUNSUPPORTED; }
    if (Fk + p->has_w > table->width || table->stride % 4 != 0 || ((uintptr_t)table->d_data & 15)) {
      etr_set_error("%s: table width %d < F*k+has_w = %d, or rows not 16-byte aligned", fn, table->width, Fk + p->has_w);
      return ETR_EINVAL;
    }
    p->table = (const char*)table->d_data; p->rows = table->rows; p->row_bytes = table->stride * 4;
    p->nchunks = (Fk + p->has_w + 3) / 4;
  }
  p->ids = (const long long*)ids->d_ids; p->csr = ids->d_csr_offsets; p->B = ids->batch;
  p->L = ids->d_csr_offsets ? 0 : ids->bag; p->sb = ids->stride_b; p->sf = ids->stride_f; p->sl = ids->stride_l;
  p->pad = ids->pad_id; p->has_pad = ids->has_pad; p->mean = ids->pooling == ETR_POOL_MEAN;
  // smem row stride: >= 4*nchunks (whole chunks are stored), == 4 (mod 32) to spread banks
  int es = ((Fk + p->has_w + 3) / 4) * 4;
  while (es % 32 != 4) es += 4;
  p->es = es;
  p->err = ctx->d_err;
  return ETR_OK;
}

}  // namespace etr

using namespace etr;

extern "C" {

int etr_field_pair_forward(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t has_w, const etr_ids* ids,
                           const float* d_bias, const float* d_r, const float* d_r0, float* d_pairvec,
                           float* d_pairdot, float* d_logit, float* d_prob, float* d_pooled, void* stream) {
  PairParams p;
  ETR_CHECK_ARG(table != nullptr, "table is NULL");
  int st = fill_pair(__func__, ctx, table, k, has_w, ids, &p);
  if (st != ETR_OK) return st;
  if (p.B == 0) return ETR_OK;
  ETR_CHECK_ARG(p.F >= 2, "need at least two fields");
  p.bias = d_bias; p.r = d_r; p.r0 = d_r0; p.pairvec = d_pairvec; p.pairdot = d_pairdot; p.logit = d_logit;
  p.prob = d_prob; p.pooled = d_pooled;
  const size_t smem = (size_t)p.F * p.es * sizeof(float);
  if (smem > 220 * 1024 || p.nchunks > 128) {
    etr_set_error("etr_field_pair_forward: F=%d k=%d needs %zu B of shared memory per sample (max 220 KiB)", p.F, k, smem);
    return ETR_EUNSUPPORTED;
  }
  const int cpl = (p.nchunks + 31) / 32;
  const int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int grid = grid_for(p.B, 1, ctx->sm_count, per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  cudaStream_t s = (cudaStream_t)stream;
#define ETR_FP(CPL)                                                                                          \
  do {                                                                                                       \
    ETR_CUDA(cudaFuncSetAttribute(field_pair_fwd_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                  (int)smem));                                                               \
    field_pair_fwd_kernel<CPL><<<grid, 256, smem, s>>>(p);                                                   \
  } while (0)
  if (cpl <= 1) ETR_FP(1); else if (cpl == 2) ETR_FP(2); else if (cpl == 3) ETR_FP(3); else ETR_FP(4);
#undef ETR_FP
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_field_pair_backward(etr_ctx* ctx, int32_t k, int32_t has_w, const etr_ids* ids, const float* d_pooled,
                            const float* d_dlogit, const float* d_r, const float* d_dpairvec, float* d_bag_grad,
                            int32_t grad_ld, void* stream) {
  PairParams p;
  int st = fill_pair(__func__, ctx, nullptr, k, has_w, ids, &p);
  if (st != ETR_OK) return st;
  if (p.B == 0) return ETR_OK;
  ETR_CHECK_ARG(d_pooled && d_bag_grad && (d_dlogit || d_dpairvec), "NULL argument");
  ETR_CHECK_ARG(grad_ld >= p.F * k + p.has_w, "grad_ld too small");
  p.pooled = const_cast<float*>(d_pooled); p.dlogit = d_dlogit; p.r = d_r; p.dpairvec = d_dpairvec;
  p.bag_grad = d_bag_grad; p.grad_ld = grad_ld;
  const size_t smem = ((size_t)p.F * p.es + p.F) * sizeof(float);
  if (smem > 220 * 1024) { etr_set_error("etr_field_pair_backward: shared memory tile too large"); return ETR_EUNSUPPORTED; }
  const int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int grid = grid_for(p.B, 1, ctx->sm_count, per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  ETR_CUDA(cudaFuncSetAttribute(field_pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  field_pair_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

static int fill_pnn(const char* fn, etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields,
                    int32_t k, int32_t kernel_type, const float* d_kernel, PnnParams* p) {
  if (!ctx || !d_x) { etr_set_error("%s: NULL argument", fn); return ETR_EINVAL; }
  if (fields < 2 || k < 1 || kernel_type < 0 || kernel_type > 3 || ldx < (int64_t)fields * k) {
    etr_set_error("%s: bad shape", fn); return ETR_EINVAL;
  }
  if (kernel_type != 0 && !d_kernel) { etr_set_error("%s: kernel weights missing", fn); return ETR_EINVAL; }
  memset(p, 0, sizeof(*p));
  p->x = d_x; p->ldx = ldx; p->B = batch; p->F = fields; p->k = k; p->type = kernel_type; p->kernel = d_kernel;
  return ETR_OK;
}

int etr_pnn_forward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k,
                    int32_t kernel_type, const float* d_kernel, float* d_out, int64_t ldo, void* stream) {
  PnnParams p;
  int st = fill_pnn(__func__, ctx, d_x, ldx, batch, fields, k, kernel_type, d_kernel, &p);
  if (st != ETR_OK) return st;
  ETR_CHECK_ARG(d_out != nullptr, "d_out is NULL");
  if (batch == 0) return ETR_OK;
  p.out = d_out; p.ldo = ldo;
  const size_t smem = 8 * (size_t)fields * k * sizeof(float);
  if (smem > 200 * 1024) { etr_set_error("etr_pnn_forward: F*k too large for the shared-memory tile"); return ETR_EUNSUPPORTED; }
  ETR_CUDA(cudaFuncSetAttribute(pnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pnn_fwd_kernel<<<grid_for(batch, 8, ctx->sm_count, 4), 256, smem, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_pnn_backward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k,
                     int32_t kernel_type, const float* d_kernel, const float* d_g, int64_t ldg, float* d_dx,
                     int64_t lddx, float* d_dkernel, void* stream) {
  PnnParams p;
  int st = fill_pnn(__func__, ctx, d_x, ldx, batch, fields, k, kernel_type, d_kernel, &p);
  if (st != ETR_OK) return st;
  ETR_CHECK_ARG(d_g && d_dx, "NULL argument");
  if (batch == 0) return ETR_OK;
  p.g = d_g; p.ldg = ldg; p.dx = d_dx; p.lddx = lddx; p.dkernel = d_dkernel;
  const int P = fields * (fields - 1) / 2;
  const size_t smem = 8 * ((size_t)fields * k + P) * sizeof(float);
  if (smem > 200 * 1024) { etr_set_error("etr_pnn_backward: F*k too large for the shared-memory tile"); return ETR_EUNSUPPORTED; }
  ETR_CUDA(cudaFuncSetAttribute(pnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pnn_bwd_kernel<<<grid_for(batch, 8, ctx->sm_count, 4), 256, smem, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  if (d_dkernel && kernel_type != 0)
    return etr_pnn_kernel_grad(ctx, d_x, ldx, batch, fields, k, kernel_type, d_g, ldg, d_dkernel, stream);
  return ETR_OK;
}

}  // extern "C"
