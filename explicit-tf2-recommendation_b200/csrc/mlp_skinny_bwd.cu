// K7b: backward of a WIDE-in / NARROW-out dense layer (the first MLP layer of DeepFM:
// [B, 432] x [432, 32]) in ONE pass over the batch:
//     dX[b, j]  = sum_o d[b, o] K[j, o]          (bf16, feeds the fused FM backward)
//     dK[j, o]  = sum_b X[b, j] d[b, o]          (fp32)
//     db[o]     = sum_b d[b, o]
// Replaces MatMul/BiasAdd backward of 2.FM/CustomLayers.py:72-84 as taken by tape.gradient
// (2.FM/ModelManager.py:176).  The previous path ran transpose(X) + cast(d^T) + split-K tcgen05 GEMM +
// finish for dK and cast(d) + cast(K) + a K=32 tcgen05 GEMM for dX: 163 us of kernels at c2 for
// ~120 MB of compulsory traffic.  Both products are memory-bound (32-deep / 32-wide), so this kernel
// reads each X slab once into shared memory (cp.async), keeps dK in registers across all slabs of a
// CTA and uses warp-level mma.sync (m16n8k16, bf16 in / fp32 accumulate) with ldmatrix(.trans) so
// that neither X nor d is ever transposed in memory.  Per-CTA dK / db partials are summed in CTA
// order by a finish kernel (deterministic).
#include "etr_common.cuh"

namespace etr {

namespace sk {

constexpr int SB = 32;                 // samples per slab (two X slabs in flight: the next one lands while this one is used)
constexpr int NOUT = 32;               // narrow side
constexpr int DS = NOUT + 8;           // smem row stride (bf16) of d and K rows: 80 B -> conflict-free ldmatrix
constexpr int STG = 32 + 8;            // staging row stride (bf16) of a 16 x 32 dX fragment block

__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((unsigned)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_t(unsigned (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((unsigned)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

struct Params {
  const __nv_bfloat16* X; long long ldx;     // [B, n_in] bf16, n_in % 16 == 0
  const float* d;                            // [B, NOUT] fp32 (pre-activation gradient)
  const float* K;                            // [n_in, NOUT] fp32
  __nv_bfloat16* dX; long long ld_dx;        // [B, n_in] bf16 (may be NULL)
  float* part;                               // [grid][n_in * NOUT + NOUT]  (dK partial | db partial)
  long long B; int n_in;
};

// MT: feature m-tiles (16 rows of dK) owned per warp = ceil(n_in / 16 / 8)
// NIN: n_in as a compile-time constant for the hot shape (0 = run-time): the index divisions of the staging loops and
// the tile bounds fold away (15.1 M warp instructions per c2 launch with run-time n_in, issue-bound at 16 warps / SM)
// Pipeline (end of round 2): 32-sample slabs, the X slab of the NEXT iteration is in flight (cp.async into the other
// buffer) and its d rows sit in registers while this slab's two products run; the one-stage form stalled on every slab
// (issue active 35 %, 16 warps / SM: profiles/r02_prof_c2_tower.md).
template <int MT, int NIN = 0>
__global__ void __launch_bounds__(256, 2) mlp_skinny_bwd_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int n_in = NIN ? NIN : p.n_in;
  const int XS = n_in + 8;                                            // smem row stride of an X row (bf16)
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(smem);         // [2][SB][XS]
  __nv_bfloat16* Ds = Xs + 2 * SB * XS;                                // [SB][DS]
  __nv_bfloat16* Ks = Ds + SB * DS;                                    // [n_in][DS]
  __nv_bfloat16* St = Ks + (size_t)n_in * DS;                          // [8 warps][16][STG]
  float* dbs = reinterpret_cast<float*>(St + 8 * 16 * STG);            // [8 warps][NOUT] column sums of d (4 rows each)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mtiles = n_in / 16;
  const int ntiles_x = n_in / 8;                                       // n-tiles of the dX product
  const int chunks_per_row = n_in / 8;

  // K -> bf16 in shared memory, once per CTA
  for (int e = tid; e < n_in * NOUT; e += 256) {
    const int j = e / NOUT, o = e % NOUT;
    Ks[j * DS + o] = __float2bfloat16_rn(p.K[e]);
  }

  float acc[MT][NOUT / 8][4];                                          // dK tiles of this warp, across all slabs
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int n = 0; n < NOUT / 8; ++n)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][n][q] = 0.f;
  float db_acc = 0.f;                                                  // thread tid < NOUT: column tid of d

  const long long nslabs = (p.B + SB - 1) / SB;
  const int dr = tid >> 3, dc = tid & 7;                               // this thread's float4 of a d slab: row dr, columns 4 dc..
  auto stage_x = [&](long long slab, int buf) {                        // cp.async, 16 B per request
    const long long b0 = slab * SB;
    __nv_bfloat16* dst = Xs + buf * SB * XS;
    for (int e = tid; e < SB * chunks_per_row; e += 256) {
      const int r = e / chunks_per_row, c = e % chunks_per_row;
      if (b0 + r < p.B) cp16(dst + r * XS + c * 8, p.X + (b0 + r) * p.ldx + c * 8);
      else *reinterpret_cast<uint4*>(dst + r * XS + c * 8) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto load_d = [&](long long slab) {
    const long long b = slab * SB + dr;
    return b < p.B ? *reinterpret_cast<const float4*>(p.d + b * NOUT + dc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 dreg = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((long long)blockIdx.x < nslabs) { stage_x(blockIdx.x, 0); dreg = load_d(blockIdx.x); }
  int buf = 0;
  for (long long slab = blockIdx.x; slab < nslabs; slab += gridDim.x, buf ^= 1) {
    const long long b0 = slab * SB;
    __syncthreads();                                                   // the previous slab is fully consumed
    // ---- this slab's d: registers -> bf16 rows in shared memory, column sums (fixed shuffle tree, then warp order)
    {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(dreg.x, dreg.y), hi = __floats2bfloat162_rn(dreg.z, dreg.w);
      *reinterpret_cast<uint2*>(Ds + dr * DS + dc * 4) =
          make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
      float4 s4 = dreg;                                                // lanes l, l^8, l^16, l^24: rows 4 warp .. 4 warp + 3
      s4.x += __shfl_xor_sync(0xffffffffu, s4.x, 8); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, 8);
      s4.z += __shfl_xor_sync(0xffffffffu, s4.z, 8); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, 8);
      s4.x += __shfl_xor_sync(0xffffffffu, s4.x, 16); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, 16);
      s4.z += __shfl_xor_sync(0xffffffffu, s4.z, 16); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, 16);
      if (lane < 8) *reinterpret_cast<float4*>(dbs + warp * NOUT + lane * 4) = s4;
    }
    // ---- the next slab: X into the other buffer, d into registers
    const long long nxt = slab + gridDim.x;
    if (nxt < nslabs) { stage_x(nxt, buf ^ 1); dreg = load_d(nxt); }
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");               // this slab's X has landed
    __syncthreads();
    if (tid < NOUT) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += dbs[w * NOUT + tid];
      db_acc += s;
    }
    const __nv_bfloat16* Xb = Xs + buf * SB * XS;

    // ---- (I) dX slab = d (SB x 32) * K^T: warp -> sample tile (warp & 1), a quarter of the n-tiles (warp >> 1)
    if (p.dX) {
      const int st = warp & 1, quarter = warp >> 1;
      unsigned a[NOUT / 16][4];
#pragma unroll
      for (int ks = 0; ks < NOUT / 16; ++ks)
        ldsm_x4(a[ks], Ds + (st * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * DS + ks * 16 + 8 * (lane >> 4));
      const int per = ((ntiles_x + 3) / 4 + 1) & ~1;                   // n-tiles per quarter, whole pairs
      const int nt0 = quarter * per;
      int nt1 = nt0 + per;
      if (nt1 > ntiles_x) nt1 = ntiles_x;
      __nv_bfloat16* stg = St + warp * 16 * STG;
      for (int nb = nt0; nb < nt1; nb += 4) {                         // blocks of 4 n-tiles = 32 columns
        float c[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {                               // pairs of n-tiles per ldmatrix.x4
          const int nt = nb + 2 * pr;
          if (nt < nt1) {
#pragma unroll
            for (int ks = 0; ks < NOUT / 16; ++ks) {
              unsigned b[4];
              // matrices: q=0 (n 0-7, k 0-7) q=1 (n 0-7, k 8-15) q=2 (n 8-15, k 0-7) q=3 (n 8-15, k 8-15)
              int nrow = nt * 8 + (lane & 7) + 8 * (lane >> 4);
              if (nrow >= n_in) nrow = n_in - 1;                       // odd tail: duplicate row, result unused
              ldsm_x4(b, Ks + nrow * DS + ks * 16 + 8 * ((lane >> 3) & 1));
              mma16816(c[2 * pr], a[ks], b[0], b[1]);
              mma16816(c[2 * pr + 1], a[ks], b[2], b[3]);
            }
          }
        }
        // fragments -> staging [16][32] -> 16-byte row segments
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(c[q][0], c[q][1]), hi = __floats2bfloat162_rn(c[q][2], c[q][3]);
          *reinterpret_cast<__nv_bfloat162*>(stg + g * STG + q * 8 + 2 * t) = lo;
          *reinterpret_cast<__nv_bfloat162*>(stg + (g + 8) * STG + q * 8 + 2 * t) = hi;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = h * 8 + (lane >> 2), cseg = lane & 3;         // 4 x 16 B per row of 32 columns
          const int col = nb * 8 + cseg * 8;
          const long long b = b0 + st * 16 + r;
          if (b < p.B && col < nt1 * 8)
            *reinterpret_cast<uint4*>(p.dX + b * p.ld_dx + col) = *reinterpret_cast<const uint4*>(stg + r * STG + cseg * 8);
        }
      }
    }

    // ---- (II) dK += X^T d over this slab: warp -> feature m-tiles warp, warp + 8, ...
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const int mt = warp + 8 * i;
      if (mt < mtiles) {
#pragma unroll
        for (int ks = 0; ks < SB / 16; ++ks) {
          unsigned a[4], b0r[4], b1r[4];
          // A = X^T tile: stored [k = sample][m = feature]; matrices q: k-off 8*(q/2), m-off 8*(q%2)
          ldsm_x4_t(a, Xb + (ks * 16 + (lane & 7) + 8 * (lane >> 4)) * XS + mt * 16 + 8 * ((lane >> 3) & 1));
          // B = d: stored [k = sample][n = o]; matrices q: k-off 8*(q%2), n-off 8*(q/2)
          ldsm_x4_t(b0r, Ds + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * DS + 0 + 8 * (lane >> 4));
          ldsm_x4_t(b1r, Ds + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * DS + 16 + 8 * (lane >> 4));
          mma16816(acc[i][0], a, b0r[0], b0r[1]);
          mma16816(acc[i][1], a, b0r[2], b0r[3]);
          mma16816(acc[i][2], a, b1r[0], b1r[1]);
          mma16816(acc[i][3], a, b1r[2], b1r[3]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // ---- per-CTA partials: dK [n_in][NOUT] then db [NOUT]
  float* out = p.part + (size_t)blockIdx.x * ((size_t)n_in * NOUT + NOUT);
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int mt = warp + 8 * i;
    if (mt < mtiles) {
#pragma unroll
      for (int n = 0; n < NOUT / 8; ++n) {
        const int col = n * 8 + 2 * t;
        *reinterpret_cast<float2*>(out + (size_t)(mt * 16 + g) * NOUT + col) = make_float2(acc[i][n][0], acc[i][n][1]);
        *reinterpret_cast<float2*>(out + (size_t)(mt * 16 + g + 8) * NOUT + col) = make_float2(acc[i][n][2], acc[i][n][3]);
      }
    }
  }
  if (tid < NOUT) out[(size_t)n_in * NOUT + tid] = db_acc;
}

// sum the per-CTA partials: a CTA owns 32 consecutive elements (one float per lane: 433 CTAs for the c2 layer instead of
// 109 with a float4 per lane, which left a quarter of the SMs idle and every warp with 37 dependent-latency rounds); its
// 8 warps each add a contiguous range of the partials (8 independent loads in flight), the 8 sub-sums are combined in
// warp order (deterministic; the per-element order is unchanged)
__global__ void __launch_bounds__(256) mlp_skinny_bwd_finish_kernel(const float* part, int nparts, int n_elems, int n_k,
                                                                    float* dK, float* db) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const int per = (nparts + 7) / 8;
  int c0 = warp * per, c1 = c0 + per;
  if (c1 > nparts) c1 = nparts;
  float s = 0.f;
  if (e < n_elems) {
#pragma unroll 8
    for (int c = c0; c < c1; ++c) s += __ldg(part + (size_t)c * n_elems + e);
  }
  sm[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && e < n_elems) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    if (e < n_k) dK[e] = t;
    else if (db) db[e - n_k] = t;
  }
}

}  // namespace sk
}  // namespace etr

using namespace etr;

extern "C" {

int etr_mlp_skinny_backward(etr_ctx* ctx, const void* d_X, int64_t ldx, const float* d_dy, const float* d_K, int64_t B,
                            int32_t n_in, int32_t n_out, void* d_dX, int64_t ld_dx, float* d_dK, float* d_db,
                            void* stream) {
  ETR_CHECK_ARG(ctx && d_X && d_dy && d_K && d_dK, "NULL argument");
  if (n_out != sk::NOUT || n_in % 16 != 0 || n_in < 16 || n_in > 16 * 8 * 4) {
    etr_set_error("etr_mlp_skinny_backward: needs n_out == 32 and n_in a multiple of 16 up to 512 (n_in=%d n_out=%d)", n_in, n_out);
    return ETR_EUNSUPPORTED;
  }
  ETR_CHECK_ARG(ldx % 8 == 0 && ((uintptr_t)d_X & 15) == 0, "X rows must be 16-byte aligned");
  ETR_CHECK_ARG(!d_dX || (ld_dx % 8 == 0 && ((uintptr_t)d_dX & 15) == 0), "dX rows must be 16-byte aligned");
  ETR_CHECK_ARG(((uintptr_t)d_dy & 15) == 0, "dy must be 16-byte aligned");
  if (B <= 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int XS = n_in + 8;
  const size_t smem = ((size_t)2 * sk::SB * XS + (size_t)sk::SB * sk::DS + (size_t)n_in * sk::DS + 8 * 16 * sk::STG) * 2 +
                      8 * sk::NOUT * sizeof(float);
  const long long nslabs = (B + sk::SB - 1) / sk::SB;
  long long grid = 2LL * ctx->sm_count;
  if (grid > nslabs) grid = nslabs;
  const size_t n_elems = (size_t)n_in * sk::NOUT + sk::NOUT;
  int st = etr_ws_reserve(ctx, (size_t)grid * n_elems * sizeof(float));
  if (st != ETR_OK) return st;
  sk::Params p;
  p.X = (const __nv_bfloat16*)d_X; p.ldx = ldx; p.d = d_dy; p.K = d_K; p.dX = (__nv_bfloat16*)d_dX; p.ld_dx = ld_dx;
  p.part = (float*)ctx->d_ws; p.B = B; p.n_in = n_in;
  const int mt = (n_in / 16 + 7) / 8;
  // the attribute is per device: set it on every call (as the other files do), not once per process
#define ETR_SK(MT)                                                                                                    \
  do {                                                                                                                \
    ETR_CUDA(cudaFuncSetAttribute(sk::mlp_skinny_bwd_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sk::mlp_skinny_bwd_kernel<MT><<<(int)grid, 256, smem, s>>>(p);                                                   \
  } while (0)
  if (n_in == 432) {                                  // DeepFM c2: 13 dense + 26 x 16 (+3 pad) inputs
    ETR_CUDA(cudaFuncSetAttribute(sk::mlp_skinny_bwd_kernel<4, 432>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sk::mlp_skinny_bwd_kernel<4, 432><<<(int)grid, 256, smem, s>>>(p);
  } else
  switch (mt) {
    case 1: ETR_SK(1); break;
    case 2: ETR_SK(2); break;
    case 3: ETR_SK(3); break;
    default: ETR_SK(4); break;
  }
#undef ETR_SK
  ETR_LAUNCH_CHECK(ctx);
  sk::mlp_skinny_bwd_finish_kernel<<<(int)((n_elems + 31) / 32), 256, 0, s>>>(
      (const float*)ctx->d_ws, (int)grid, (int)n_elems, n_in * sk::NOUT, d_dK, d_db);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
