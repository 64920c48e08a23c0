// K7 (fp32 exact-parity path) + loss + small elementwise helpers.
//
// etr_gemm_f32 replaces MatMul + BiasAdd + activation of the MLP towers
// (2.FM/CustomLayers.py:72-84; 3.DCN/CustomLayers.py:163-167) and, through
// etr_cross_mat_layer_f32, the fp32 form of the DCN-matrix cross layer.  It is
// a register-tiled SIMT SGEMM (TM x TN outputs per thread, k-major shared
// tiles read with 128-bit LDS) with deterministic split-K for the wgrad shapes
// ([in,B] x [B,out]: tiny M,N, K = batch).
#include "etr_common.cuh"

namespace etr {

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ETR_ACT_RELU: return x > 0.f ? x : 0.f;
    case ETR_ACT_SIGMOID: return sigmoidf_exact(x);
    case ETR_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

struct GemmParams {
  const float* A; long long sam, sak;   // A(m,k) = A[m*sam + k*sak]
  const float* B; long long sbk, sbn;   // B(k,n) = B[k*sbk + n*sbn]
  float* C; long long ldc;
  const float* bias;
  float alpha, beta;
  int act;
  long long M, N, K;
  int splits; long long k_per_split;
  float* partial;                        // [splits, M, N] when splits > 1
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_f32_kernel(const GemmParams p) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PADM = BM + 4, PADN = BN + 4;
  __shared__ __align__(16) float As[BK][PADM];
  __shared__ __align__(16) float Bs[BK][PADN];
  const int tid = threadIdx.x;
  const int tn = tid % (BN / TN), tm = tid / (BN / TN);
  const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
  const long long kbeg = (long long)blockIdx.z * p.k_per_split;
  long long kend = kbeg + p.k_per_split;
  if (kend > p.K) kend = p.K;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const bool a_kcontig = (p.sak == 1);
  const bool b_ncontig = (p.sbn == 1);
  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile BM x BK
    for (int e = tid; e < BM * BK; e += NT) {
      int mm, kk;
      if (a_kcontig) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      const long long gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < kend) ? p.A[gm * p.sam + gk * p.sak] : 0.f;
    }
    for (int e = tid; e < BN * BK; e += NT) {
      int nn, kk;
      if (b_ncontig) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      const long long gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < p.N && gk < kend) ? p.B[gk * p.sbk + gn * p.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[kk][tm * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
      if (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&Bs[kk][tn * TN + j]);
          b[j] = t.x; b[(j + 1) % TN] = t.y; b[(j + 2) % TN] = t.z; b[(j + 3) % TN] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tn * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long gm = m0 + tm * TM + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const long long gn = n0 + tn * TN + j;
      if (gn >= p.N) continue;
      if (p.splits > 1) {
        p.partial[((long long)blockIdx.z * p.M + gm) * p.N + gn] = acc[i][j];
      } else {
        float x = p.alpha * acc[i][j];
        if (p.beta != 0.f) x += p.beta * p.C[gm * p.ldc + gn];
        if (p.bias) x += p.bias[gn];
        p.C[gm * p.ldc + gn] = apply_act(x, p.act);
      }
    }
  }
}

__global__ void __launch_bounds__(256) gemm_splitk_finish_kernel(const GemmParams p) {
  const long long total = p.M * p.N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long gm = t / p.N, gn = t % p.N;
    float s = 0.f;
    for (int z = 0; z < p.splits; ++z) s += p.partial[(long long)z * total + t];   // fixed order
    float x = p.alpha * s;
    if (p.beta != 0.f) x += p.beta * p.C[gm * p.ldc + gn];
    if (p.bias) x += p.bias[gn];
    p.C[gm * p.ldc + gn] = apply_act(x, p.act);
  }
}


// ---- skinny shapes of the MLP tail (e.g. [B,32]x[32,8], [B,8]x[8,1] and their dgrads):
// one thread per output row, NT accumulators in registers, the whole weight
// matrix broadcast from shared memory.  C = act(alpha*A B + beta*C + bias).
template <int NT>
__global__ void __launch_bounds__(256) gemm_skinny_rows_kernel(const GemmParams p) {
  extern __shared__ __align__(16) float Ws[];          // [K][NT], zero padded
  const int K = (int)p.K, N = (int)p.N;
  for (int e = threadIdx.x; e < K * NT; e += blockDim.x) {
    const int kk = e / NT, nn = e % NT;
    Ws[e] = nn < N ? p.B[kk * p.sbk + nn * p.sbn] : 0.f;
  }
  __syncthreads();
  const bool vec_a = (p.sam % 4 == 0) && ((((uintptr_t)p.A) & 15) == 0);
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < p.M; m += (long long)gridDim.x * blockDim.x) {
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
    const float* a = p.A + m * p.sam;
    int k = 0;
    if (vec_a) {
      for (; k + 4 <= K; k += 4) {
        const float4 x = *reinterpret_cast<const float4*>(a + k);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int n = 0; n < NT; ++n) acc[n] = fmaf(xs[q], Ws[(k + q) * NT + n], acc[n]);
        }
      }
    }
    for (; k < K; ++k) {
      const float x = a[k];
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n] = fmaf(x, Ws[k * NT + n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (n < N) {
        float x = p.alpha * acc[n];
        if (p.beta != 0.f) x += p.beta * p.C[m * p.ldc + n];
        if (p.bias) x += p.bias[n];
        p.C[m * p.ldc + n] = apply_act(x, p.act);
      }
    }
  }
}

// wgrad of the tail: C[M,N] = A^T B with M*N <= 1024 (M + N <= 88: the slab fits 48 KB) and K = batch.  Each CTA stages a slab of
// kSkinnySlab batch rows of A ([rows, M]) and B ([rows, N]) in shared memory (coalesced, no index
// division) and thread t = (m, n) reduces it; the slab partials are then summed by one warp per
// output in a fixed order (lane partition + xor tree) -- deterministic.
constexpr int kSkinnySlab = 128;
__global__ void __launch_bounds__(256) gemm_skinny_wgrad_kernel(const GemmParams p) {
  extern __shared__ __align__(16) float sm_w[];          // As[slab][M] | Bs[slab][N]
  const int M = (int)p.M, N = (int)p.N;
  float* As = sm_w;
  float* Bs = sm_w + kSkinnySlab * M;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long r0 = (long long)blockIdx.x * kSkinnySlab;
  int rows = kSkinnySlab;
  if (r0 + rows > p.K) rows = (int)(p.K - r0);
  for (int r = warp; r < kSkinnySlab; r += 8) {
    const bool in = r < rows;
    for (int mm = lane; mm < M; mm += 32) As[r * M + mm] = in ? p.A[mm * p.sam + (r0 + r) * p.sak] : 0.f;
    for (int nn = lane; nn < N; nn += 32) Bs[r * N + nn] = in ? p.B[(r0 + r) * p.sbk + nn * p.sbn] : 0.f;
  }
  __syncthreads();
  for (int o = t; o < M * N; o += 256) {               // up to 4 outputs per thread (M*N <= 1024)
    const int om = o / N, on = o % N;
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < kSkinnySlab; ++r) acc = fmaf(As[r * M + om], Bs[r * N + on], acc);
    p.partial[(long long)blockIdx.x * M * N + o] = acc;
  }
}
// one warp per output element: lane l sums partials l, l+32, ... in order, then a fixed xor tree
__global__ void __launch_bounds__(256) skinny_finish_kernel(const GemmParams p) {
  const int lane = threadIdx.x & 31;
  const long long total = p.M * p.N;
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= total) return;
  float s = 0.f;
  for (int z = lane; z < p.splits; z += 32) s += p.partial[(long long)z * total + o];
  s = group_sum<32>(s);
  if (lane == 0) {
    const long long gm = o / p.N, gn = o % p.N;
    float x = p.alpha * s;
    if (p.beta != 0.f) x += p.beta * p.C[gm * p.ldc + gn];
    if (p.bias) x += p.bias[gn];
    p.C[gm * p.ldc + gn] = apply_act(x, p.act);
  }
}

__global__ void __launch_bounds__(256) act_backward_kernel(float* dy, const float* y, long long n, int act) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float yy = y[t];
    float d = dy[t];
    switch (act) {
      case ETR_ACT_RELU: d = yy > 0.f ? d : 0.f; break;
      case ETR_ACT_SIGMOID: d = d * yy * (1.0f - yy); break;
      case ETR_ACT_TANH: d = d * (1.0f - yy * yy); break;
      default: break;
    }
    dy[t] = d;
  }
}

// column sums: stage 1 -- each block sums a slab of rows for CB (power of two <= 32)
// columns; the 256 threads are arranged as (256/CB) row lanes x CB columns.
constexpr int kColRowsPerBlock = 512;
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const float* X, long long M, long long N, long long ldx,
                                                            int CB, float* partial /* [slabs, N] */) {
  __shared__ float sm[256];
  const int cx = threadIdx.x % CB, ry = threadIdx.x / CB, RY = 256 / CB;
  const long long n = (long long)blockIdx.x * CB + cx;
  const long long r0 = (long long)blockIdx.y * kColRowsPerBlock;
  long long r1 = r0 + kColRowsPerBlock;
  if (r1 > M) r1 = M;
  float s = 0.f;
  if (n < N)
    for (long long r = r0 + ry; r < r1; r += RY) s += X[r * ldx + n];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float t = 0.f;
    for (int q = 0; q < RY; ++q) t += sm[q * CB + cx];
    partial[(long long)blockIdx.y * N + n] = t;
  }
}
// stage 2: one warp per column -- lane l sums slabs l, l+32, ... in order, then a fixed xor tree
__global__ void __launch_bounds__(256) colsum_stage2_kernel(const float* partial, long long slabs, long long N, float* out) {
  const int lane = threadIdx.x & 31;
  const long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float s = 0.f;
  for (long long q = lane; q < slabs; q += 32) s += partial[q * N + n];
  s = group_sum<32>(s);
  if (lane == 0) out[n] = s;
}

// Keras BCE on probabilities + d/dz through the sigmoid.
constexpr float kKerasEps = 1e-7f;
__global__ void __launch_bounds__(256) bce_kernel(const float* prob, const float* label, long long B, float* partial,
                                                  float* dlogit) {
  __shared__ float sm[8];
  float s = 0.f;
  const float invB = 1.0f / (float)B;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (long long)gridDim.x * blockDim.x) {
    const float p = prob[t], y = label[t];
    const float pc = fminf(fmaxf(p, kKerasEps), 1.0f - kKerasEps);
    s += -(y * logf(pc + kKerasEps) + (1.0f - y) * logf(1.0f - pc + kKerasEps));
    if (dlogit) {
      // clip passes the gradient only inside [eps, 1-eps]
      const bool inside = (p >= kKerasEps) && (p <= 1.0f - kKerasEps);
      const float dp = inside ? -(y / (pc + kKerasEps) - (1.0f - y) / (1.0f - pc + kKerasEps)) * invB : 0.f;
      dlogit[t] = dp * p * (1.0f - p);
    }
  }
  s = group_sum<32>(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += sm[q];
    partial[blockIdx.x] = t;
  }
}
__global__ void bce_finish_kernel(const float* partial, int n, long long B, float* loss) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += partial[i];
  *loss = s / (float)B;
}

__global__ void __launch_bounds__(256) add_sigmoid_kernel(const float* a, const float* b, long long n, float* logit,
                                                          float* prob) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float z = a[t] + (b ? b[t] : 0.f);
    if (logit) logit[t] = z;
    if (prob) prob[t] = sigmoidf_exact(z);
  }
}

__global__ void __launch_bounds__(256) cross_mat_bwd_elem_kernel(const float* g, const float* x0, const float* u,
                                                                 long long n, float* du, float* dx0) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float gg = g[t];
    du[t] = gg * x0[t];
    dx0[t] += gg * u[t];
  }
}

// out = x0 (.) u + xl  (u already holds xl W^T + b)
__global__ void __launch_bounds__(256) cross_mat_fwd_elem_kernel(const float* x0, const float* xl, long long ldx,
                                                                 const float* u, long long ldu, long long B, int D,
                                                                 float* out, long long ldo) {
  const long long total = B * D;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / D;
    const int d = (int)(t % D);
    out[b * ldo + d] = x0[b * ldx + d] * u[b * ldu + d] + xl[b * ldx + d];
  }
}

static int gemm_dispatch(etr_ctx* ctx, GemmParams& p, cudaStream_t s) {
  // skinny shapes first: rows x tiny weight matrix (forward / dgrad of the MLP tail)
  if (p.sak == 1 && p.N <= 32 && p.K <= 128 && p.M >= 1024) {
    const int nt = p.N <= 1 ? 1 : (p.N <= 8 ? 8 : (p.N <= 16 ? 16 : 32));
    const size_t smem = (size_t)p.K * nt * sizeof(float);
    const int grid = grid_for(p.M, 256, ctx->sm_count, 8);
    if (nt == 1) gemm_skinny_rows_kernel<1><<<grid, 256, smem, s>>>(p);
    else if (nt == 8) gemm_skinny_rows_kernel<8><<<grid, 256, smem, s>>>(p);
    else if (nt == 16) gemm_skinny_rows_kernel<16><<<grid, 256, smem, s>>>(p);
    else gemm_skinny_rows_kernel<32><<<grid, 256, smem, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
    return ETR_OK;
  }
  // wgrad of the tail: tiny output, reduction over the batch
  if (p.M * p.N <= 1024 && p.M <= 64 && p.N <= 32 && p.M + p.N <= 88 && p.K >= 4096) {
    p.splits = (int)ceil_div(p.K, kSkinnySlab);
    int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)p.splits * p.M * p.N);
    if (st != ETR_OK) return st;
    p.partial = (float*)ctx->d_ws;
    const size_t smem = (size_t)kSkinnySlab * (p.M + p.N) * sizeof(float);
    gemm_skinny_wgrad_kernel<<<p.splits, 256, smem, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
    skinny_finish_kernel<<<(unsigned)ceil_div(p.M * p.N, 8), 256, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
    return ETR_OK;
  }
  // tile selection by N; split-K when the MN grid cannot fill the machine
  int bm, bn;
  if (p.N <= 8) { bm = 256; bn = 8; } else if (p.N <= 32) { bm = 256; bn = 32; } else { bm = 128; bn = 64; }
  const long long gx = ceil_div(p.N, bn), gy = ceil_div(p.M, bm);
  int splits = 1;
  if (gx * gy < ctx->sm_count && p.K >= 4096) {
    splits = (int)((2LL * ctx->sm_count) / (gx * gy));
    const long long max_splits = p.K / 512;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  p.k_per_split = ceil_div(ceil_div(p.K, splits), 16) * 16;
  p.splits = (int)ceil_div(p.K, p.k_per_split);
  if (p.splits > 1) {
    int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)p.splits * p.M * p.N);
    if (st != ETR_OK) return st;
    p.partial = (float*)ctx->d_ws;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)p.splits);
  if (gy > 65535) { etr_set_error("gemm_f32: M too large for one launch"); return ETR_EUNSUPPORTED; }
  if (bn == 8) gemm_f32_kernel<256, 8, 8, 1><<<grid, 256, 0, s>>>(p);
  else if (bn == 32) gemm_f32_kernel<256, 32, 8, 4><<<grid, 256, 0, s>>>(p);
  else gemm_f32_kernel<128, 64, 8, 4><<<grid, 256, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  if (p.splits > 1) {
    gemm_splitk_finish_kernel<<<grid_for(p.M * p.N, 256, ctx->sm_count, 8), 256, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
  }
  return ETR_OK;
}

}  // namespace etr

using namespace etr;

extern "C" {

int etr_gemm_f32(etr_ctx* ctx, int32_t trans_a, int32_t trans_b, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* d_A, int64_t lda, const float* d_B, int64_t ldb, float beta, float* d_C,
                 int64_t ldc, const float* d_bias, int32_t act, void* stream) {
  ETR_CHECK_ARG(ctx && d_A && d_B && d_C, "NULL argument");
  ETR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative dimension");
  if (M == 0 || N == 0) return ETR_OK;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.A = d_A; p.B = d_B; p.C = d_C; p.ldc = ldc; p.bias = d_bias; p.alpha = alpha; p.beta = beta; p.act = act;
  p.M = M; p.N = N; p.K = K;
  if (!trans_a) { p.sam = lda; p.sak = 1; } else { p.sam = 1; p.sak = lda; }   // trans: A stored [K,M]
  if (!trans_b) { p.sbk = ldb; p.sbn = 1; } else { p.sbk = 1; p.sbn = ldb; }   // trans: B stored [N,K]
  return gemm_dispatch(ctx, p, (cudaStream_t)stream);
}

int etr_act_backward(etr_ctx* ctx, float* d_dy, const float* d_y, int64_t n, int32_t act, void* stream) {
  ETR_CHECK_ARG(ctx && d_dy && d_y, "NULL argument");
  if (n <= 0 || act == ETR_ACT_NONE) return ETR_OK;
  act_backward_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(d_dy, d_y, n, act);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_colsum_f32(etr_ctx* ctx, const float* d_X, int64_t M, int64_t N, int64_t ldx, float* d_out, void* stream) {
  ETR_CHECK_ARG(ctx && d_X && d_out, "NULL argument");
  if (N <= 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (M <= 0) { ETR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float) * N, s)); return ETR_OK; }
  const long long slabs = ceil_div(M, kColRowsPerBlock);
  ETR_CHECK_ARG(slabs <= 65535, "M too large");
  int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)slabs * N);
  if (st != ETR_OK) return st;
  int cb = 1;
  while (cb < 32 && cb < N) cb <<= 1;
  dim3 grid((unsigned)ceil_div(N, cb), (unsigned)slabs);
  colsum_stage1_kernel<<<grid, 256, 0, s>>>(d_X, M, N, ldx, cb, (float*)ctx->d_ws);
  ETR_LAUNCH_CHECK(ctx);
  colsum_stage2_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, s>>>((const float*)ctx->d_ws, slabs, N, d_out);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_bce_forward_backward(etr_ctx* ctx, const float* d_prob, const float* d_label, int64_t batch, float* d_loss,
                             float* d_dlogit, void* stream) {
  ETR_CHECK_ARG(ctx && d_prob && d_label && d_loss, "NULL argument");
  ETR_CHECK_ARG(batch > 0, "empty batch");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(batch, 256 * 4, ctx->sm_count, 2);
  int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)grid);
  if (st != ETR_OK) return st;
  bce_kernel<<<grid, 256, 0, s>>>(d_prob, d_label, batch, (float*)ctx->d_ws, d_dlogit);
  ETR_LAUNCH_CHECK(ctx);
  bce_finish_kernel<<<1, 1, 0, s>>>((const float*)ctx->d_ws, grid, batch, d_loss);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_add_sigmoid(etr_ctx* ctx, const float* d_a, const float* d_b, int64_t n, float* d_logit, float* d_prob,
                    void* stream) {
  ETR_CHECK_ARG(ctx && d_a, "NULL argument");
  if (n <= 0) return ETR_OK;
  add_sigmoid_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, d_logit, d_prob);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cross_mat_layer_f32(etr_ctx* ctx, const float* d_x0, const float* d_xl, int64_t ldx, int64_t batch,
                            int32_t D, const float* d_W, const float* d_b, float* d_out, int64_t ldo, float* d_u,
                            int64_t ldu, void* stream) {
  ETR_CHECK_ARG(ctx && d_x0 && d_xl && d_W && d_b && d_out && d_u, "NULL argument (d_u scratch [B,ldu] is required)");
  if (batch == 0) return ETR_OK;
  // U = xl W^T + b : B(k,n) = W[n,k]  -> trans_b
  int st = etr_gemm_f32(ctx, 0, 1, batch, D, D, 1.0f, d_xl, ldx, d_W, D, 0.0f, d_u, ldu, d_b, ETR_ACT_NONE, stream);
  if (st != ETR_OK) return st;
  cross_mat_fwd_elem_kernel<<<grid_for(batch * D, 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(
      d_x0, d_xl, ldx, d_u, ldu, batch, D, d_out, ldo);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cross_mat_bwd_elementwise(etr_ctx* ctx, const float* d_g, const float* d_x0, const float* d_u, int64_t n,
                                  float* d_du, float* d_dx0_accum, void* stream) {
  ETR_CHECK_ARG(ctx && d_g && d_x0 && d_u && d_du && d_dx0_accum, "NULL argument");
  if (n <= 0) return ETR_OK;
  cross_mat_bwd_elem_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(d_g, d_x0, d_u, n,
                                                                                                 d_du, d_dx0_accum);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
