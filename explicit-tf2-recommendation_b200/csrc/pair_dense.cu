// Pairwise-interaction consumers of the same embedding rows, on the dense [B, F, k] tensor K1 leaves
// behind (the flattened embedding it writes for the tower is that tensor):
//   mode 0  AFM / xDeepFM-style pair vectors   out[b,p,:] = x_i (.) x_j                 3.DCN/CustomLayers.py:825-838
//   mode 1  FiBiNet bilinear interaction       out[b,p,:] = (x_i W_w) (.) x_j           3.DCN/CustomLayers.py:977-1009
//           w_kind 0 'all' (one W [k,k]), 1 'each' (W_i, i < F-1), 2 'interaction' (W_p, p < P)
//   mode 2  NFM bi-interaction pooling         out[b,:]   = 0.5 ((sum_f x_f)^2 - sum_f x_f^2)   3.DCN/CustomLayers.py:499-501
// pairs (i < j) in itertools.combinations order, like K4.  One warp per sample, x[b] staged in shared memory; the
// pair-vector output (P*k floats per sample: 20.8 KB at F = 26, k = 16) is written once, coalesced -- these kernels
// are bound by that write.  Weight gradients are reduced over the batch WITHOUT atomics: one CTA per (weight matrix,
// batch slice) accumulates its [k,k] block in registers in sample order, a finish kernel adds the slices in order.
// The same scheme replaces the fp32 atomics the PNN outer-product kernels used for dK (etr_pnn_backward).
#include <algorithm>

#include "etr_common.cuh"

namespace etr {

__device__ __forceinline__ void pd_pair_from_index(int p, int F, int& i, int& j) {
  int ii = 0, rem = p;
  while (rem >= F - 1 - ii) { rem -= F - 1 - ii; ++ii; }
  i = ii; j = ii + 1 + rem;
}

struct PairDenseParams {
  const float* x; long long ldx; long long B; int F, k, mode, w_kind;
  const float* W; float* out; long long ldo;
  const float* g; long long ldg; float* dx; long long lddx;
  float* partial; int slices;            // weight-gradient partials [slices][n_w][k][k]
};

__device__ __forceinline__ int pd_weight_of(const PairDenseParams& p, int pi, int i) {
  return p.w_kind == 0 ? 0 : (p.w_kind == 1 ? i : pi);
}

__global__ void __launch_bounds__(256) pair_dense_fwd_kernel(const PairDenseParams p) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Fk = p.F * p.k, P = p.F * (p.F - 1) / 2, k = p.k;
  float* xs = sm + warp * 2 * Fk;        // x[b]
  float* ys = xs + Fk;                   // bilinear 'all' / 'each': y_i = x_i W_(i)
  for (long long b0 = (long long)blockIdx.x * nwarps; b0 < p.B; b0 += (long long)gridDim.x * nwarps) {
    const long long b = b0 + warp;
    if (b < p.B)
      for (int t = lane; t < Fk; t += 32) xs[t] = p.x[b * p.ldx + t];
    __syncwarp();
    if (b < p.B) {
      if (p.mode == 2) {
        for (int c = lane; c < k; c += 32) {
          float s = 0.f, q = 0.f;
          for (int f = 0; f < p.F; ++f) { const float v = xs[f * k + c]; s += v; q += v * v; }
          p.out[b * p.ldo + c] = 0.5f * (s * s - q);
        }
      } else {
        if (p.mode == 1 && p.w_kind != 2) {
          for (int t = lane; t < Fk; t += 32) {
            const int i = t / k, c = t % k;
            const float* Wm = p.W + (long long)(p.w_kind == 0 ? 0 : (i < p.F - 1 ? i : 0)) * k * k;
            float s = 0.f;
            for (int a = 0; a < k; ++a) s += xs[i * k + a] * __ldg(Wm + a * k + c);
            ys[t] = s;
          }
          __syncwarp();
        }
        const int total = P * k;
        int pi_cached = -1, ci = 0, cj = 0;
        for (int t = lane; t < total; t += 32) {
          const int pi = t / k, c = t % k;
          if (pi != pi_cached) { pd_pair_from_index(pi, p.F, ci, cj); pi_cached = pi; }
          float left;
          if (p.mode == 0) left = xs[ci * k + c];
          else if (p.w_kind != 2) left = ys[ci * k + c];
          else {
            const float* Wm = p.W + (long long)pi * k * k;
            left = 0.f;
            for (int a = 0; a < k; ++a) left += xs[ci * k + a] * __ldg(Wm + a * k + c);
          }
          p.out[b * p.ldo + t] = left * xs[cj * k + c];
        }
      }
    }
    __syncwarp();
  }
}

// dx (ACCUMULATED into p.dx): one warp per sample, one thread per (field, column)
__global__ void __launch_bounds__(256) pair_dense_bwd_kernel(const PairDenseParams p) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Fk = p.F * p.k, P = p.F * (p.F - 1) / 2, k = p.k;
  const int gsz = p.mode == 2 ? k : P * k;
  float* xs = sm + warp * (2 * Fk + gsz);
  float* ys = xs + Fk;
  float* gs = ys + Fk;
  for (long long b0 = (long long)blockIdx.x * nwarps; b0 < p.B; b0 += (long long)gridDim.x * nwarps) {
    const long long b = b0 + warp;
    if (b < p.B) {
      for (int t = lane; t < Fk; t += 32) xs[t] = p.x[b * p.ldx + t];
      for (int t = lane; t < gsz; t += 32) gs[t] = p.g[b * p.ldg + t];
    }
    __syncwarp();
    if (b < p.B) {
      if (p.mode == 2) {
        for (int t = lane; t < Fk; t += 32) {
          const int c = t % k;
          float s = 0.f;
          for (int f = 0; f < p.F; ++f) s += xs[f * k + c];
          p.dx[b * p.lddx + t] += gs[c] * (s - xs[t]);
        }
      } else if (p.mode == 0) {
        for (int t = lane; t < Fk; t += 32) {
          const int i = t / k, c = t % k;
          float acc = 0.f;
          for (int j = 0; j < p.F; ++j) {
            if (j == i) continue;
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            const int pi = lo * (2 * p.F - lo - 1) / 2 + (hi - lo - 1);
            acc += gs[pi * k + c] * xs[j * k + c];
          }
          p.dx[b * p.lddx + t] += acc;
        }
      } else {
        // bilinear: out_p = (x_i W) (.) x_j.  As the LEFT factor of pair (i, j): dx_i[a] += sum_c G_p[c] x_j[c] W[a][c];
        // as the RIGHT factor of pair (l, i): dx_i[c] += G_p[c] (x_l W)[c]
        if (p.w_kind != 2) {
          for (int t = lane; t < Fk; t += 32) {
            const int i = t / k, c = t % k;
            const float* Wm = p.W + (long long)(p.w_kind == 0 ? 0 : (i < p.F - 1 ? i : 0)) * k * k;
            float s = 0.f;
            for (int a = 0; a < k; ++a) s += xs[i * k + a] * __ldg(Wm + a * k + c);
            ys[t] = s;
          }
          __syncwarp();
        }
        for (int t = lane; t < Fk; t += 32) {
          const int i = t / k, a = t % k;
          float acc = 0.f;
          for (int j = i + 1; j < p.F; ++j) {                      // i is the left factor
            const int pi = i * (2 * p.F - i - 1) / 2 + (j - i - 1);
            const float* Wm = p.W + (long long)pd_weight_of(p, pi, i) * k * k + a * k;
            float s = 0.f;
            for (int c = 0; c < k; ++c) s += gs[pi * k + c] * xs[j * k + c] * __ldg(Wm + c);
            acc += s;
          }
          for (int l = 0; l < i; ++l) {                            // i is the right factor
            const int pi = l * (2 * p.F - l - 1) / 2 + (i - l - 1);
            float y;
            if (p.w_kind != 2) y = ys[l * k + a];
            else {
              const float* Wm = p.W + (long long)pi * k * k;
              y = 0.f;
              for (int q = 0; q < k; ++q) y += xs[l * k + q] * __ldg(Wm + q * k + a);
            }
            acc += gs[pi * k + a] * y;
          }
          p.dx[b * p.lddx + t] += acc;
        }
      }
    }
    __syncwarp();
  }
}

// dW[w][a][c] = sum_b sum_{pairs p using w} x_i[a] * G_p[c] x_j[c]; block = (weight w, batch slice), thread = (a, c)
__global__ void __launch_bounds__(1024) pair_dense_dw_kernel(const PairDenseParams p) {
  const int k = p.k, P = p.F * (p.F - 1) / 2;
  const int w = blockIdx.x, sl = blockIdx.y;
  const int a = threadIdx.x / k, c = threadIdx.x % k;
  const long long per = (p.B + p.slices - 1) / p.slices;
  const long long bs = sl * per, be = bs + per < p.B ? bs + per : p.B;
  float acc = 0.f;
  if (a < k) {
    for (long long b = bs; b < be; ++b) {
      const float* xb = p.x + b * p.ldx;
      const float* gb = p.g + b * p.ldg;
      if (p.w_kind == 2) {
        int i, j;
        pd_pair_from_index(w, p.F, i, j);
        acc += xb[i * k + a] * gb[(long long)w * k + c] * xb[j * k + c];
      } else {
        const int i0 = p.w_kind == 1 ? w : 0, i1 = p.w_kind == 1 ? w + 1 : p.F - 1;
        for (int i = i0; i < i1; ++i) {
          float t = 0.f;
          for (int j = i + 1; j < p.F; ++j) {
            const int pi = i * (2 * p.F - i - 1) / 2 + (j - i - 1);
            t += gb[(long long)pi * k + c] * xb[j * k + c];
          }
          acc += xb[i * k + a] * t;
        }
      }
    }
    p.partial[((long long)sl * gridDim.x + w) * k * k + a * k + c] = acc;
  }
  (void)P;
}

__global__ void __launch_bounds__(256) slice_sum_kernel(const float* partial, long long n, int slices, float* out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < slices; ++q) s += partial[q * n + t];      // fixed order: deterministic
    out[t] = s;
  }
}

// ---- PNN outer-product kernel gradients without atomics (2.FM/CustomLayers.py:658-682 under tape.gradient):
//   mat  K[a,p,c]: dK = sum_b G_p x_i[c] x_j[a]     vec  K[p,c]: sum_b G_p x_i[c] x_j[c]     num  K[p]: sum_b G_p <x_i,x_j>
// block = (pair p, batch slice); mat: thread (a,c); vec: thread c; num: thread 0 .. k-1 then a fixed-order sum
struct PnnDkParams {
  const float* x; long long ldx; long long B; int F, k, type;
  const float* g; long long ldg; float* partial; int slices;
};
__global__ void __launch_bounds__(1024) pnn_dk_kernel(const PnnDkParams p) {
  __shared__ float red[1024];
  const int k = p.k, P = p.F * (p.F - 1) / 2;
  const int pi = blockIdx.x, sl = blockIdx.y;
  int i, j;
  pd_pair_from_index(pi, p.F, i, j);
  const long long per = (p.B + p.slices - 1) / p.slices;
  const long long bs = sl * per, be = bs + per < p.B ? bs + per : p.B;
  const int t = threadIdx.x;
  float acc = 0.f;
  if (p.type == 1) {                      // mat
    const int a = t / k, c = t % k;
    if (a < k)
      for (long long b = bs; b < be; ++b) {
        const float* xb = p.x + b * p.ldx;
        acc += p.g[b * p.ldg + pi] * xb[i * k + c] * xb[j * k + a];
      }
    if (a < k) p.partial[(long long)sl * k * P * k + ((long long)a * P + pi) * k + c] = acc;
  } else {                                // vec / num: thread c
    if (t < k)
      for (long long b = bs; b < be; ++b) {
        const float* xb = p.x + b * p.ldx;
        acc += p.g[b * p.ldg + pi] * xb[i * k + t] * xb[j * k + t];
      }
    if (p.type == 2) {
      if (t < k) p.partial[(long long)sl * P * k + (long long)pi * k + t] = acc;
    } else {
      red[t] = t < k ? acc : 0.f;
      __syncthreads();
      if (t == 0) {
        float s = 0.f;
        for (int d = 0; d < k; ++d) s += red[d];
        p.partial[(long long)sl * P + pi] = s;
      }
    }
  }
}

}  // namespace etr

using namespace etr;

static int pd_check(const char* fn, etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t B, int F, int k, int mode, int w_kind,
                    const float* d_W) {
  if (!ctx || !d_x) { etr_set_error("%s: NULL argument", fn); return ETR_EINVAL; }
  if (F < 2 || k < 1 || mode < 0 || mode > 2 || ldx < (int64_t)F * k || B < 0) { etr_set_error("%s: bad shape", fn); return ETR_EINVAL; }
  if (mode == 1 && (!d_W || w_kind < 0 || w_kind > 2)) { etr_set_error("%s: bilinear weights missing / bad w_kind", fn); return ETR_EINVAL; }
  if (mode == 1 && k > 32) { etr_set_error("%s: bilinear interaction covers k <= 32", fn); return ETR_EUNSUPPORTED; }
  return ETR_OK;
}

extern "C" {

int etr_pair_dense_forward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k, int32_t mode,
                           int32_t w_kind, const float* d_W, float* d_out, int64_t ldo, void* stream) {
  int st = pd_check(__func__, ctx, d_x, ldx, batch, fields, k, mode, w_kind, d_W);
  if (st != ETR_OK) return st;
  ETR_CHECK_ARG(d_out != nullptr, "d_out is NULL");
  if (batch == 0) return ETR_OK;
  PairDenseParams p;
  memset(&p, 0, sizeof(p));
  p.x = d_x; p.ldx = ldx; p.B = batch; p.F = fields; p.k = k; p.mode = mode; p.w_kind = w_kind; p.W = d_W; p.out = d_out; p.ldo = ldo;
  const size_t smem = 8 * 2 * (size_t)fields * k * sizeof(float);
  if (smem > 200 * 1024) { etr_set_error("%s: F*k too large for the shared-memory tile", __func__); return ETR_EUNSUPPORTED; }
  ETR_CUDA(cudaFuncSetAttribute(pair_dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_dense_fwd_kernel<<<grid_for(batch, 8, ctx->sm_count, 4), 256, smem, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_pair_dense_backward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k, int32_t mode,
                            int32_t w_kind, const float* d_W, const float* d_g, int64_t ldg, float* d_dx, int64_t lddx,
                            float* d_dW, void* stream) {
  int st = pd_check(__func__, ctx, d_x, ldx, batch, fields, k, mode, w_kind, d_W);
  if (st != ETR_OK) return st;
  ETR_CHECK_ARG(d_g && d_dx, "NULL argument");
  if (batch == 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  PairDenseParams p;
  memset(&p, 0, sizeof(p));
  p.x = d_x; p.ldx = ldx; p.B = batch; p.F = fields; p.k = k; p.mode = mode; p.w_kind = w_kind; p.W = d_W;
  p.g = d_g; p.ldg = ldg; p.dx = d_dx; p.lddx = lddx;
  const int P = fields * (fields - 1) / 2;
  const size_t per_warp = 2 * (size_t)fields * k + (mode == 2 ? k : (size_t)P * k);
  int warps = 8;
  while (warps > 1 && warps * per_warp * sizeof(float) > 200 * 1024) warps >>= 1;
  const size_t smem = warps * per_warp * sizeof(float);
  if (smem > 200 * 1024) { etr_set_error("%s: P*k too large for the shared-memory tile", __func__); return ETR_EUNSUPPORTED; }
  ETR_CUDA(cudaFuncSetAttribute(pair_dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_dense_bwd_kernel<<<grid_for(batch, warps, ctx->sm_count, 2), warps * 32, smem, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  if (mode == 1 && d_dW) {
    const int n_w = w_kind == 0 ? 1 : (w_kind == 1 ? fields - 1 : P);
    p.slices = (int)std::min<int64_t>(std::max<int64_t>(1, (int64_t)ctx->sm_count * 4 / n_w), std::max<int64_t>(1, batch / 64));
    const size_t bytes = (size_t)p.slices * n_w * k * k * sizeof(float);
    st = etr_ws_reserve(ctx, bytes);
    if (st != ETR_OK) return st;
    p.partial = (float*)ctx->d_ws;
    pair_dense_dw_kernel<<<dim3(n_w, p.slices), k * k, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
    const long long n = (long long)n_w * k * k;
    slice_sum_kernel<<<grid_for(n, 256, ctx->sm_count, 4), 256, 0, s>>>(p.partial, n, p.slices, d_dW);
    ETR_LAUNCH_CHECK(ctx);
  }
  return ETR_OK;
}

// deterministic replacement of the fp32 atomics of the PNN outer-product kernel gradients: writes d_dkernel
int etr_pnn_kernel_grad(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k,
                        int32_t kernel_type, const float* d_g, int64_t ldg, float* d_dkernel, void* stream) {
  ETR_CHECK_ARG(ctx && d_x && d_g && d_dkernel, "NULL argument");
  ETR_CHECK_ARG(fields >= 2 && k >= 1 && kernel_type >= 1 && kernel_type <= 3 && ldx >= (int64_t)fields * k, "bad shape");
  if (kernel_type == 1 && k > 32) { etr_set_error("%s: 'mat' kernels cover k <= 32", __func__); return ETR_EUNSUPPORTED; }
  if (k > 1024) { etr_set_error("%s: k too large", __func__); return ETR_EUNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
  const int P = fields * (fields - 1) / 2;
  const long long n = kernel_type == 1 ? (long long)k * P * k : (kernel_type == 2 ? (long long)P * k : P);
  if (batch == 0) { ETR_CUDA(cudaMemsetAsync(d_dkernel, 0, n * sizeof(float), s)); return ETR_OK; }
  PnnDkParams p;
  p.x = d_x; p.ldx = ldx; p.B = batch; p.F = fields; p.k = k; p.type = kernel_type; p.g = d_g; p.ldg = ldg;
  p.slices = (int)std::min<int64_t>(std::max<int64_t>(1, (int64_t)ctx->sm_count * 4 / P), std::max<int64_t>(1, batch / 64));
  int st = etr_ws_reserve(ctx, (size_t)p.slices * n * sizeof(float));
  if (st != ETR_OK) return st;
  p.partial = (float*)ctx->d_ws;
  const int threads = kernel_type == 1 ? k * k : ((k + 31) / 32) * 32;
  pnn_dk_kernel<<<dim3(P, p.slices), threads, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  slice_sum_kernel<<<grid_for(n, 256, ctx->sm_count, 4), 256, 0, s>>>(p.partial, n, p.slices, d_dkernel);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
