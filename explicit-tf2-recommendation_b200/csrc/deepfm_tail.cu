// K7c: the narrow tail of the DeepFM tower with the loss, forward AND backward, in one kernel:
//   h2 = relu(h1 K2 + b2) [8] ; dnn = h2 K3 + b3 ; z = fm_logit + dnn ; p = sigmoid(z)
//   loss = Keras BinaryCrossentropy(p, y) (clip to [1e-7, 1-1e-7], mean over the batch)
//   dz = dL/dz ; dh2 = dz K3 * [h2>0] ; d1 = (dh2 K2^T) * [h1>0]         (h1 = relu output of layer 1)
//   dK3 = h2^T dz, db3 = sum dz, dK2 = h1^T dh2, db2 = sum dh2, d bias_fm = sum dz
// Replaces, for MLPLayer(mlp_dims=[32,8]) + MLPLayer([1]) (2.FM/CustomLayers.py:255-256, 301-305) and the
// train step's loss (2.FM/ModelManager.py:175-176), 14 small launches (two skinny GEMMs, add+sigmoid, BCE,
// two activation backwards, two weight-gradient GEMMs with their finish kernels, three column sums) that
// together moved ~8 MB tensors a dozen times: h1 [B,32] is read once and d1 [B,32] written once.
// A warp works on 32 rows: coalesced tile load -> shared memory (thread r owns row r; column reads give
// the h1^T operand of dK2), per-CTA partial sums, summed in CTA order by the finish kernel.
#include "etr_common.cuh"

namespace etr {
namespace tail {

constexpr int H1 = 32, H2 = 8;
constexpr int NPART = 2 + H2 + 1 + H2 + H1 * H2;     // loss, sum_dz, dK3[8], db3, db2[8], dK2[256]
constexpr float kKerasEps = 1e-7f;

struct Params {
  const float* h1; const float* fm_logit; const float* label;
  const float* K2; const float* b2; const float* K3; const float* b3;
  float* prob; float* dlogit; float* d1; float* part;
  long long B; float grad_scale;
};

__global__ void __launch_bounds__(256) deepfm_tail_kernel(const Params p) {
  __shared__ float tile[8][32][33];
  __shared__ float dh2s[8][32][H2];
  __shared__ float K2s[H1 * H2], b2s[H2], K3s[H2];
  float (*red)[NPART] = reinterpret_cast<float (*)[NPART]>(&tile[0][0][0]);     // reused after the row loop
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < H1 * H2; e += 256) K2s[e] = p.K2[e];
  if (threadIdx.x < H2) { b2s[threadIdx.x] = p.b2[threadIdx.x]; K3s[threadIdx.x] = p.K3[threadIdx.x]; }
  __syncthreads();
  const float b3 = p.b3[0];
  const float invB = 1.0f / (float)p.B;
  float loss = 0.f, sdz = 0.f, gK3[H2], gb2[H2], gK2[H2];      // gK2: row `lane` of dK2
#pragma unroll
  for (int j = 0; j < H2; ++j) gK3[j] = gb2[j] = gK2[j] = 0.f;
  const long long ntiles = (p.B + 31) / 32;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long tl = (long long)blockIdx.x * 8 + warp; tl < ntiles; tl += nwarps) {
    const long long r0 = tl * 32;
    // coalesced load of the 32 x 32 tile: 8 lanes per row, 4 rows per instruction
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = q * 4 + (lane >> 3), c = (lane & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < p.B) v = *reinterpret_cast<const float4*>(p.h1 + (r0 + r) * H1 + c);
      tile[warp][r][c] = v.x; tile[warp][r][c + 1] = v.y; tile[warp][r][c + 2] = v.z; tile[warp][r][c + 3] = v.w;
    }
    __syncwarp();
    const long long b = r0 + lane;
    const bool active = b < p.B;
    float h1[H1];
#pragma unroll
    for (int i = 0; i < H1; ++i) h1[i] = tile[warp][lane][i];
    float h2[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) h2[j] = b2s[j];
#pragma unroll
    for (int i = 0; i < H1; ++i)
#pragma unroll
      for (int j = 0; j < H2; ++j) h2[j] += h1[i] * K2s[i * H2 + j];
    float dnn = b3;
#pragma unroll
    for (int j = 0; j < H2; ++j) { h2[j] = fmaxf(h2[j], 0.f); dnn += h2[j] * K3s[j]; }
    float dz = 0.f;
    if (active) {
      const float z = p.fm_logit[b] + dnn;
      const float pr = sigmoidf_exact(z), y = p.label[b];
      const float pc = fminf(fmaxf(pr, kKerasEps), 1.0f - kKerasEps);
      loss += -(y * logf(pc + kKerasEps) + (1.0f - y) * logf(1.0f - pc + kKerasEps));
      const bool inside = (pr >= kKerasEps) && (pr <= 1.0f - kKerasEps);      // clip passes the gradient only inside
      const float dp = inside ? -(y / (pc + kKerasEps) - (1.0f - y) / (1.0f - pc + kKerasEps)) * invB : 0.f;
      dz = dp * pr * (1.0f - pr) * p.grad_scale;
      p.prob[b] = pr;
      p.dlogit[b] = dz;
    }
    sdz += dz;
    float dh2[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) {
      gK3[j] += h2[j] * dz;
      dh2[j] = h2[j] > 0.f ? dz * K3s[j] : 0.f;
      gb2[j] += dh2[j];
      dh2s[warp][lane][j] = dh2[j];
    }
    __syncwarp();
    // dK2[lane][j] += sum_r h1[r][lane] dh2[r][j]   (column read of the tile, broadcast read of dh2)
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float x = tile[warp][r][lane];
      const float4 a = *reinterpret_cast<const float4*>(&dh2s[warp][r][0]);
      const float4 c = *reinterpret_cast<const float4*>(&dh2s[warp][r][4]);
      gK2[0] += x * a.x; gK2[1] += x * a.y; gK2[2] += x * a.z; gK2[3] += x * a.w;
      gK2[4] += x * c.x; gK2[5] += x * c.y; gK2[6] += x * c.z; gK2[7] += x * c.w;
    }
    __syncwarp();
    // d1 = (dh2 K2^T) * [h1 > 0] -> tile -> coalesced store
#pragma unroll
    for (int i = 0; i < H1; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < H2; ++j) s += dh2[j] * K2s[i * H2 + j];
      tile[warp][lane][i] = h1[i] > 0.f ? s : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = q * 4 + (lane >> 3), c = (lane & 7) * 4;
      if (r0 + r < p.B)
        *reinterpret_cast<float4*>(p.d1 + (r0 + r) * H1 + c) =
            make_float4(tile[warp][r][c], tile[warp][r][c + 1], tile[warp][r][c + 2], tile[warp][r][c + 3]);
    }
    __syncwarp();
  }
  // per-CTA partials (fixed order: lanes by butterfly, warps 0..7)
  loss = group_sum<32>(loss); sdz = group_sum<32>(sdz);
#pragma unroll
  for (int j = 0; j < H2; ++j) { gK3[j] = group_sum<32>(gK3[j]); gb2[j] = group_sum<32>(gb2[j]); }
  __syncthreads();                                            // every warp is done with its tile
  if (lane == 0) {
    red[warp][0] = loss; red[warp][1] = sdz;
#pragma unroll
    for (int j = 0; j < H2; ++j) { red[warp][2 + j] = gK3[j]; red[warp][2 + H2 + 1 + j] = gb2[j]; }
    red[warp][2 + H2] = sdz;                                  // db3 = sum dz
  }
#pragma unroll
  for (int j = 0; j < H2; ++j) red[warp][2 + H2 + 1 + H2 + lane * H2 + j] = gK2[j];
  __syncthreads();
  for (int e = threadIdx.x; e < NPART; e += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][e];
    p.part[(size_t)blockIdx.x * NPART + e] = s;
  }
}

// a CTA owns 32 consecutive partial-sum slots; its 8 warps each add a contiguous range of the CTAs'
// partials, combined in warp order (deterministic)
__global__ void __launch_bounds__(256) deepfm_tail_finish_kernel(const float* part, int nparts, long long B, float* loss,
                                                                 float* g_bias_fm, float* gK3, float* gb3, float* gb2,
                                                                 float* gK2) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const int per = (nparts + 7) / 8;
  int c0 = warp * per, c1 = c0 + per;
  if (c1 > nparts) c1 = nparts;
  float t = 0.f;
  if (e < NPART) {
#pragma unroll 8
    for (int c = c0; c < c1; ++c) t += __ldg(part + (size_t)c * NPART + e);      // independent loads, ordered adds
  }
  sm[warp][lane] = t;
  __syncthreads();
  if (warp != 0 || e >= NPART) return;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += sm[w][lane];
  if (e == 0) *loss = s / (float)B;
  else if (e == 1) { if (g_bias_fm) *g_bias_fm = s; }
  else if (e < 2 + H2) gK3[e - 2] = s;
  else if (e == 2 + H2) gb3[0] = s;
  else if (e < 2 + H2 + 1 + H2) gb2[e - (2 + H2 + 1)] = s;
  else gK2[e - (2 + H2 + 1 + H2)] = s;
}

}  // namespace tail
}  // namespace etr

using namespace etr;

extern "C" {

int etr_deepfm_tail_train(etr_ctx* ctx, const float* d_h1, const float* d_fm_logit, const float* d_label, int64_t batch,
                          const float* d_K2, const float* d_b2, const float* d_K3, const float* d_b3, float grad_scale,
                          float* d_prob, float* d_dlogit, float* d_d1, float* d_loss, float* d_g_bias_fm, float* d_gK2,
                          float* d_gb2, float* d_gK3, float* d_gb3, void* stream) {
  ETR_CHECK_ARG(ctx && d_h1 && d_fm_logit && d_label && d_K2 && d_b2 && d_K3 && d_b3 && d_prob && d_dlogit && d_d1 &&
                    d_loss && d_gK2 && d_gb2 && d_gK3 && d_gb3, "NULL argument");
  ETR_CHECK_ARG(batch > 0, "empty batch");
  ETR_CHECK_ARG((((uintptr_t)d_h1 | (uintptr_t)d_d1) & 15) == 0, "h1 / d1 must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(batch, 256, ctx->sm_count, 2);
  int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)grid * tail::NPART);
  if (st != ETR_OK) return st;
  tail::Params p;
  p.h1 = d_h1; p.fm_logit = d_fm_logit; p.label = d_label; p.K2 = d_K2; p.b2 = d_b2; p.K3 = d_K3; p.b3 = d_b3;
  p.prob = d_prob; p.dlogit = d_dlogit; p.d1 = d_d1; p.part = (float*)ctx->d_ws; p.B = batch; p.grad_scale = grad_scale;
  tail::deepfm_tail_kernel<<<grid, 256, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  tail::deepfm_tail_finish_kernel<<<(tail::NPART + 31) / 32, 256, 0, s>>>((const float*)ctx->d_ws, grid, batch, d_loss, d_g_bias_fm, d_gK3,
                                                    d_gb3, d_gb2, d_gK2);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
