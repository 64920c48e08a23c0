// K5: DCN cross-vector layer, all layers fused (3.DCN/CustomLayers.py:195-203):
//   x_{l+1} = x0 * <x_l, w_l> + b_l + x_l
// HBM-bound: one read of x0 and one write of the output per sample; w_l, b_l
// live in shared memory.  One warp per sample, NPL elements per lane held in
// registers (D <= 32*NPL).
//
// Backward uses the closed form x_l = c_l * x0 + Bc_l with c_l = 1 + sum_{j<l} s_j
// and Bc_l = sum_{j<l} b_j, so no activation is saved: the kernel recomputes the
// scalars s_l, emits dx0 and the per-sample scalars (ds_l, ds_l*c_l); the
// batch reductions for dw / db are then one deterministic split-K GEMM and one
// column sum on the host side (cross_vec_finish).
#include "etr_common.cuh"

namespace etr {

struct CrossVecParams {
  const float* x0; long long ldx; long long B; int D; int layers;
  const float* w; const float* b;            // [layers, D]
  float* out; long long ldo;
  const float* gout; long long ldg; float* dx0; long long lddx;
  float* scal;                               // [B, 2*layers]: ds_l, ds_l*c_l
};

constexpr int kMaxCrossLayers = 8;

template <int NPL>
__global__ void __launch_bounds__(256) cross_vec_fwd_kernel(const CrossVecParams p) {
  extern __shared__ float sm[];               // w[layers*D], b[layers*D]
  float* ws = sm;
  float* bs = sm + (size_t)p.layers * p.D;
  for (int t = threadIdx.x; t < p.layers * p.D; t += blockDim.x) { ws[t] = p.w[t]; bs[t] = p.b[t]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long b = warp_global; b < p.B; b += nwarps) {
    float x0[NPL], xl[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int d = lane + i * 32;
      x0[i] = d < p.D ? p.x0[b * p.ldx + d] : 0.f;
      xl[i] = x0[i];
    }
    for (int l = 0; l < p.layers; ++l) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int d = lane + i * 32;
        if (d < p.D) s += xl[i] * ws[l * p.D + d];
      }
      s = group_sum<32>(s);
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int d = lane + i * 32;
        if (d < p.D) xl[i] = x0[i] * s + bs[l * p.D + d] + xl[i];
      }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int d = lane + i * 32;
      if (d < p.D) p.out[b * p.ldo + d] = xl[i];
    }
  }
}

template <int NPL>
__global__ void __launch_bounds__(256) cross_vec_bwd_kernel(const CrossVecParams p) {
  extern __shared__ float sm[];               // w[layers*D], Bc[(layers)*D] (cumulative bias before layer l)
  float* ws = sm;
  float* bc = sm + (size_t)p.layers * p.D;
  for (int t = threadIdx.x; t < p.layers * p.D; t += blockDim.x) ws[t] = p.w[t];
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < p.layers; ++l) { bc[l * p.D + d] = acc; acc += p.b[l * p.D + d]; }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long b = warp_global; b < p.B; b += nwarps) {
    float x0[NPL], G[NPL], dx[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int d = lane + i * 32;
      x0[i] = d < p.D ? p.x0[b * p.ldx + d] : 0.f;
      G[i] = d < p.D ? p.gout[b * p.ldg + d] : 0.f;
      dx[i] = 0.f;
    }
    // forward scalars: s_l = <x_l, w_l>, x_l = c_l x0 + Bc_l
    float s[kMaxCrossLayers], c[kMaxCrossLayers];
    float cl = 1.0f;
    for (int l = 0; l < p.layers; ++l) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int d = lane + i * 32;
        if (d < p.D) t += (cl * x0[i] + bc[l * p.D + d]) * ws[l * p.D + d];
      }
      t = group_sum<32>(t);
      s[l] = t; c[l] = cl;
      cl += t;
    }
    for (int l = p.layers - 1; l >= 0; --l) {
      float ds = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) ds += G[i] * x0[i];
      ds = group_sum<32>(ds);
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int d = lane + i * 32;
        dx[i] += G[i] * s[l];
        if (d < p.D) G[i] += ds * ws[l * p.D + d];
      }
      if (lane == 0) {
        p.scal[b * 2 * p.layers + l] = ds;
        p.scal[b * 2 * p.layers + p.layers + l] = ds * c[l];
      }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int d = lane + i * 32;
      if (d < p.D) p.dx0[b * p.lddx + d] = dx[i] + G[i];
    }
  }
}

// dw_l = XtC[:, L + l] + Bc_l * sum_b ds_l ;  db_l = colsum(gout) + sum_{j>l} w_j * sum_b ds_j
// XtC [D, 2L] = X0^T scal (GEMM), sums [2L] = colsum(scal), gsum [D] = colsum(gout)
__global__ void __launch_bounds__(256) cross_vec_finish_kernel(const float* XtC, const float* sums, const float* gsum,
                                                               const float* w, const float* b, int D, int layers,
                                                               float* dw, float* db) {
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
    float bcum = 0.f;
    for (int l = 0; l < layers; ++l) {
      dw[l * D + d] = XtC[(long long)d * 2 * layers + layers + l] + bcum * sums[l];
      bcum += b[l * D + d];
    }
    float tail = 0.f;
    for (int l = layers - 1; l >= 0; --l) {
      db[l * D + d] = gsum[d] + tail;
      tail += w[l * D + d] * sums[l];
    }
  }
}

static int npl_for(int D) {
  const int need = (D + 31) / 32;
  int npl = 1;
  while (npl < need) npl <<= 1;
  return npl;
}

}  // namespace etr

using namespace etr;

extern "C" {

int etr_cross_vec_forward(etr_ctx* ctx, const float* d_x0, int64_t ldx, int64_t batch, int32_t D, int32_t layers,
                          const float* d_w, const float* d_b, float* d_out, int64_t ldo, void* stream) {
  ETR_CHECK_ARG(ctx && d_x0 && d_w && d_b && d_out, "NULL argument");
  ETR_CHECK_ARG(D > 0 && layers > 0 && layers <= kMaxCrossLayers, "bad D / layers (max 8 layers)");
  if (batch == 0) return ETR_OK;
  CrossVecParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = d_x0; p.ldx = ldx; p.B = batch; p.D = D; p.layers = layers; p.w = d_w; p.b = d_b; p.out = d_out; p.ldo = ldo;
  const int npl = npl_for(D);
  const size_t smem = 2 * (size_t)layers * D * sizeof(float);
  if (npl > 64 || smem > 200 * 1024) { etr_set_error("etr_cross_vec_forward: D=%d too wide (max 2048)", D); return ETR_EUNSUPPORTED; }
  const int grid = grid_for(batch, 8, ctx->sm_count, 4);
  cudaStream_t s = (cudaStream_t)stream;
#define ETR_CV(N)                                                                                             \
  do {                                                                                                        \
    ETR_CUDA(cudaFuncSetAttribute(cross_vec_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    cross_vec_fwd_kernel<N><<<grid, 256, smem, s>>>(p);                                                       \
  } while (0)
  switch (npl) {
    case 1: ETR_CV(1); break; case 2: ETR_CV(2); break; case 4: ETR_CV(4); break; case 8: ETR_CV(8); break;
    case 16: ETR_CV(16); break; case 32: ETR_CV(32); break; default: ETR_CV(64); break;
  }
#undef ETR_CV
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cross_vec_backward(etr_ctx* ctx, const float* d_x0, int64_t ldx, int64_t batch, int32_t D, int32_t layers,
                           const float* d_w, const float* d_b, const float* d_gout, int64_t ldg, float* d_dx0,
                           int64_t lddx, float* d_scal, void* stream) {
  ETR_CHECK_ARG(ctx && d_x0 && d_w && d_b && d_gout && d_dx0 && d_scal, "NULL argument");
  ETR_CHECK_ARG(D > 0 && layers > 0 && layers <= kMaxCrossLayers, "bad D / layers (max 8 layers)");
  if (batch == 0) return ETR_OK;
  CrossVecParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = d_x0; p.ldx = ldx; p.B = batch; p.D = D; p.layers = layers; p.w = d_w; p.b = d_b;
  p.gout = d_gout; p.ldg = ldg; p.dx0 = d_dx0; p.lddx = lddx; p.scal = d_scal;
  const int npl = npl_for(D);
  const size_t smem = 2 * (size_t)layers * D * sizeof(float);
  if (npl > 64 || smem > 200 * 1024) { etr_set_error("etr_cross_vec_backward: D=%d too wide (max 2048)", D); return ETR_EUNSUPPORTED; }
  const int grid = grid_for(batch, 8, ctx->sm_count, 2);
  cudaStream_t s = (cudaStream_t)stream;
#define ETR_CV(N)                                                                                             \
  do {                                                                                                        \
    ETR_CUDA(cudaFuncSetAttribute(cross_vec_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    cross_vec_bwd_kernel<N><<<grid, 256, smem, s>>>(p);                                                       \
  } while (0)
  switch (npl) {
    case 1: ETR_CV(1); break; case 2: ETR_CV(2); break; case 4: ETR_CV(4); break; case 8: ETR_CV(8); break;
    case 16: ETR_CV(16); break; case 32: ETR_CV(32); break; default: ETR_CV(64); break;
  }
#undef ETR_CV
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cross_vec_finish(etr_ctx* ctx, const float* d_XtC, const float* d_sums, const float* d_gsum,
                         const float* d_w, const float* d_b, int32_t D, int32_t layers, float* d_dw, float* d_db,
                         void* stream) {
  ETR_CHECK_ARG(ctx && d_XtC && d_sums && d_gsum && d_w && d_b && d_dw && d_db, "NULL argument");
  cross_vec_finish_kernel<<<grid_for(D, 256, ctx->sm_count, 1), 256, 0, (cudaStream_t)stream>>>(
      d_XtC, d_sums, d_gsum, d_w, d_b, D, layers, d_dw, d_db);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
