// K6 / K7 tensor-core path: bf16 GEMM on the 5th-gen tensor cores.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )      A, B bf16, K contiguous ("TN")
//
// Operand tiles are brought in by TMA (cp.async.bulk.tensor, 128-byte swizzle)
// into a multi-stage shared-memory ring; one elected thread issues
// tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BLOCK_N x 16, fp32 accumulate)
// with the accumulator living in TMEM; four epilogue warps read it back with
// tcgen05.ld and apply the fused epilogue:
//   EPI_LINEAR : C = act(acc + bias)                      (MLP layers, dgrad)
//   EPI_CROSS  : u = acc + b ; out = x0 (.) u + xl        (DCN-matrix cross layer,
//                3.DCN/CustomLayers.py:300-303 -- y = W x  ==  X_l W^T, so W
//                itself, row-major [D,D], IS the [N,K] operand)
//   EPI_PARTIAL: fp32 partial tile for split-K (wgrad shapes, K = batch)
// Two CTAs are resident per SM (shared memory and TMEM are budgeted for it), so
// one CTA's epilogue overlaps the other's main loop.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer, warps 2..9 = epilogue: TMEM lane quarter = warp_id % 4, and the two
// warps of a quarter split the tile's 32-column chunks (even / odd) so that eight
// latency-bound epilogue chains run per CTA instead of four.
#include <cuda.h>

#include "etr_common.cuh"

namespace etr {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kThreads = 320;          // warp 0 TMA, warp 1 MMA/TMEM, warps 2..9 epilogue

enum { EPI_LINEAR = 0, EPI_CROSS = 1, EPI_PARTIAL = 2 };

struct EpiParams {
  int mode;
  long long M, N;
  void* C; long long ldc; int c_bf16;            // LINEAR: output
  const float* bias; int act;
  int accumulate;                                // LINEAR, fp32 C: C += act(acc + bias)
  const __nv_bfloat16* x0; const __nv_bfloat16* xl; long long ldx;   // CROSS inputs
  __nv_bfloat16* out; long long ldo;             // CROSS: x_{l+1}
  __nv_bfloat16* u; long long ldu;               // CROSS: optional U = xl W^T + b
  float* partial;                                // PARTIAL: [splits, M, N]
  int k_blocks_total; int k_blocks_per_split;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld32_issue(taddr, v);
  tmem_ld_wait();
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (sm_100 UMMA):
// start address >> 4 in bits [0,14); LBO unused for swizzled K-major (0);
// SBO = 8 rows * 128 B = 1024 B (>> 4) in bits [32,46); descriptor version 1 in
// bits [46,48); layout type SWIZZLE_128B = 2 in bits [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major, 128-byte-swizzled operand (the operand's M / N extent is the contiguous one in memory): the tile is a row of
// [64 mn x BLOCK_K k] boxes (one TMA box each, 8 KiB, rows of 128 bytes = 64 mn elements, one row per k), i.e. the
// canonical layout ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp):
// LBO = byte distance between 64-element mn blocks = 8192, SBO = byte distance between groups of 8 k rows = 1024.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9,
// 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t make_idesc(int umma_m, int umma_n, bool mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(umma_n >> 3) << 17) |
         ((uint32_t)(umma_m >> 4) << 24);
}

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case ETR_ACT_RELU: return x > 0.f ? x : 0.f;
    case ETR_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case ETR_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& w, float (&f)[8]) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[t]));
    f[2 * t] = x.x; f[2 * t + 1] = x.y;
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
  return make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                    *reinterpret_cast<uint32_t*>(&h3));
}
__device__ __noinline__ float act_slow(float x, int act) { return act_apply(x, act); }
template <int N>
__device__ __forceinline__ void act_vec(float (&a)[N], int act) {      // act is warp-uniform: one branch per vector
  if (act == ETR_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = a[i] > 0.f ? a[i] : 0.f;
  } else if (act == ETR_ACT_SIGMOID || act == ETR_ACT_TANH) {
#pragma unroll 1
    for (int i = 0; i < N; ++i) a[i] = act_slow(a[i], act);
  }
}
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// 8 consecutive bf16 / N consecutive fp32 of one row: one 16-byte access when the piece is whole and aligned,
// element by element (first nv only) on ragged column tails and odd pitches
__device__ __forceinline__ uint4 ldg_u4(const void* p) {          // the epilogue operands are global: spare the generic path
  uint4 w;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p));
  return w;
}
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 w;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "l"(p));
  return w;
}
__device__ __forceinline__ void stg_u4(void* p, const uint4& w) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, bool vec, int nv, float (&f)[8]) {
  if (vec) {
    unpack_bf16x8(ldg_u4(p), f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i < nv ? __bfloat162float(p[i]) : 0.f;
  }
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, bool vec, int nv, const float (&f)[8]) {
  if (vec) {
    stg_u4(p, pack_bf16x8(f));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) p[i] = __float2bfloat16_rn(f[i]);
  }
}

// The hot form of the cross epilogue: a whole 32 x 32 block, every pointer 16-byte aligned, all 32 rows inside M, x0 / xl
// preloaded -- straight-line code, no per-row or per-element predicates.  (__fmul_rn / __fadd_rn here and in the general
// form: no FMA contraction, so both give the same bits.)
template <bool HAS_U, bool HAS_X0>
__device__ __forceinline__ void epi_cross_lean(const EpiParams& ep, const float* Tl, const float (&b8)[8], long long r0,
                                               long long off, const uint4 (&px0)[4], const uint4 (&pxl)[4]) {
  constexpr int TS = 36;
  __nv_bfloat16* po = ep.out + r0 * ep.ldo + off;
  __nv_bfloat16* pu = HAS_U ? ep.u + r0 * ep.ldu + off : nullptr;
  const long long so = 8 * ep.ldo, su = 8 * ep.ldu;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t0 = *reinterpret_cast<const float4*>(Tl + 8 * j * TS);
    const float4 t1 = *reinterpret_cast<const float4*>(Tl + 8 * j * TS + 4);
    float a[8] = {__fadd_rn(t0.x, b8[0]), __fadd_rn(t0.y, b8[1]), __fadd_rn(t0.z, b8[2]), __fadd_rn(t0.w, b8[3]),
                  __fadd_rn(t1.x, b8[4]), __fadd_rn(t1.y, b8[5]), __fadd_rn(t1.z, b8[6]), __fadd_rn(t1.w, b8[7])};
    if (HAS_U) stg_u4(pu + j * su, pack_bf16x8(a));
    float x[8];
    if (HAS_X0) {
      unpack_bf16x8(px0[j], x);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __fmul_rn(a[i], x[i]);
    }
    unpack_bf16x8(pxl[j], x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __fadd_rn(a[i], x[i]);
    stg_u4(po + j * so, pack_bf16x8(a));
  }
}

// One 32-row x (<= 32)-column block of the accumulator (fp32 bits in v[], lane = row row_base + lane, columns
// col0 .. col0 + ncol) through the fused epilogue; shared by the per-tile and the persistent kernel.  The block is
// transposed ONCE through the warp's private 32 x 36 fp32 tile T (lane = row  ->  4 or 8 lanes per row); everything
// after that is elementwise in the coalesced layout: bias, activation, the cross combine with x0 / xl (loaded
// straight into that layout, or preloaded by epi_preload: px0 / pxl, only for whole aligned blocks), 16-byte stores.
// Ragged column tails (BLOCK_N = 240: the last block of a tile is 16 wide; N = 1677: 13) and pitches that are not
// 16-byte multiples take the same path with per-element accesses.
__device__ __forceinline__ bool epi_vec_ok(const __nv_bfloat16* base, long long ld, long long col0, int ncol) {
  return ncol == 32 && ((ld & 7) == 0) && al16(base + col0);
}
__device__ __forceinline__ void epi_preload(const __nv_bfloat16* base, long long ld, long long row_base, long long col0,
                                            long long M, int lane, uint4 (&w)[4]) {
  const int g4 = lane & 3, r4 = lane >> 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long rr = row_base + r4 + 8 * j;
    w[j] = make_uint4(0u, 0u, 0u, 0u);
    if (rr < M) w[j] = ldg_u4(base + rr * ld + col0 + 8 * g4);
  }
}
__device__ __forceinline__ void epi_block(const EpiParams& ep, float* T, int lane, long long row_base, long long col0, int ncol,
                                          const uint32_t (&v)[32], int split, bool pre, const uint4 (&px0)[4],
                                          const uint4 (&pxl)[4]) {
  constexpr int TS = 36;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; i += 4)
    *reinterpret_cast<uint4*>(&T[lane * TS + i]) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  __syncwarp();
  if (ep.mode == EPI_PARTIAL || (ep.mode == EPI_LINEAR && !ep.c_bf16)) {
    // fp32 out: 8 lanes x 16 bytes per row, 4 rows per instruction
    const int g8 = lane & 7, r8 = lane >> 3;
    const int nv = ncol - 4 * g8;                    // valid columns of this lane's piece (<= 0: none)
    if (nv <= 0) return;
    const bool linear = ep.mode == EPI_LINEAR;
    float* obase = linear ? reinterpret_cast<float*>(ep.C) : ep.partial + (long long)split * ep.M * ep.N;
    const long long ld = linear ? ep.ldc : ep.N;
    const long long off = col0 + 4 * g8;
    const bool vec = nv >= 4 && (ld & 3) == 0 && al16(obase + off);
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (linear && ep.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nv) b4[i] = ep.bias[off + i];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r8 + 4 * j;
      const long long rr = row_base + r;
      const float4 t = *reinterpret_cast<const float4*>(&T[r * TS + 4 * g8]);
      float a[4] = {t.x + b4[0], t.y + b4[1], t.z + b4[2], t.w + b4[3]};
      if (linear) act_vec<4>(a, ep.act);
      if (rr >= ep.M) continue;
      float* o = obase + rr * ld + off;
      if (vec) {
        if (linear && ep.accumulate) {
          const float4 c = *reinterpret_cast<const float4*>(o);
          a[0] += c.x; a[1] += c.y; a[2] += c.z; a[3] += c.w;
        }
        *reinterpret_cast<float4*>(o) = make_float4(a[0], a[1], a[2], a[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nv) o[i] = (linear && ep.accumulate) ? o[i] + a[i] : a[i];
      }
    }
    return;
  }
  // bf16 out: 4 lanes x 16 bytes per row, 8 rows per instruction
  const int g4 = lane & 3, r4 = lane >> 2;
  const int nv = ncol - 8 * g4;
  if (nv <= 0) return;
  const long long off = col0 + 8 * g4;
  const bool whole = nv >= 8;
  float b8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ep.bias) {
    if (whole && al16(ep.bias + off)) {
      const float4 p = ldg_f4(ep.bias + off), q = ldg_f4(ep.bias + off + 4);
      b8[0] = p.x; b8[1] = p.y; b8[2] = p.z; b8[3] = p.w; b8[4] = q.x; b8[5] = q.y; b8[6] = q.z; b8[7] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nv) b8[i] = ep.bias[off + i];
    }
  }
  const long long r0 = row_base + r4;                // this lane's rows: r0 + 8 j
  const float* Tl = T + r4 * TS + 8 * g4;
  if (ep.mode == EPI_LINEAR) {
    __nv_bfloat16* pc = reinterpret_cast<__nv_bfloat16*>(ep.C) + r0 * ep.ldc + off;
    const bool vc = whole && (ep.ldc & 7) == 0 && al16(pc);
#pragma unroll
    for (int j = 0; j < 4; ++j, pc += 8 * ep.ldc) {
      const float4 t0 = *reinterpret_cast<const float4*>(Tl + 8 * j * TS);
      const float4 t1 = *reinterpret_cast<const float4*>(Tl + 8 * j * TS + 4);
      float a[8] = {t0.x + b8[0], t0.y + b8[1], t0.z + b8[2], t0.w + b8[3], t1.x + b8[4], t1.y + b8[5], t1.z + b8[6], t1.w + b8[7]};
      act_vec<8>(a, ep.act);
      if (r0 + 8 * j < ep.M) store_bf16x8(pc, vc, nv, a);
    }
    return;
  }
  // EPI_CROSS: u = acc + b ; out = x0 * u + xl   (x0 == NULL: out = u + xl)
  const __nv_bfloat16* pxl_g = ep.xl + r0 * ep.ldx + off;
  const __nv_bfloat16* px0_g = ep.x0 ? ep.x0 + r0 * ep.ldx + off : nullptr;
  __nv_bfloat16* po = ep.out + r0 * ep.ldo + off;
  __nv_bfloat16* pu = ep.u ? ep.u + r0 * ep.ldu + off : nullptr;
  const bool vx = whole && (ep.ldx & 7) == 0 && al16(pxl_g) && (!px0_g || al16(px0_g));
  const bool vo = whole && (ep.ldo & 7) == 0 && al16(po);
  const bool vu = whole && pu && (ep.ldu & 7) == 0 && al16(pu);
  if (pre && ncol == 32 && vx && vo && (!pu || vu) && row_base + 32 <= ep.M) {
    if (pu) {
      if (px0_g) epi_cross_lean<true, true>(ep, Tl, b8, r0, off, px0, pxl);
      else epi_cross_lean<true, false>(ep, Tl, b8, r0, off, px0, pxl);
    } else {
      if (px0_g) epi_cross_lean<false, true>(ep, Tl, b8, r0, off, px0, pxl);
      else epi_cross_lean<false, false>(ep, Tl, b8, r0, off, px0, pxl);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t0 = *reinterpret_cast<const float4*>(Tl + 8 * j * TS);
    const float4 t1 = *reinterpret_cast<const float4*>(Tl + 8 * j * TS + 4);
    float a[8] = {__fadd_rn(t0.x, b8[0]), __fadd_rn(t0.y, b8[1]), __fadd_rn(t0.z, b8[2]), __fadd_rn(t0.w, b8[3]),
                  __fadd_rn(t1.x, b8[4]), __fadd_rn(t1.y, b8[5]), __fadd_rn(t1.z, b8[6]), __fadd_rn(t1.w, b8[7])};
    if (r0 + 8 * j < ep.M) {
      if (pu) store_bf16x8(pu, vu, nv, a);
      float x[8];
      if (px0_g) {
        if (pre) unpack_bf16x8(px0[j], x);
        else load_bf16x8(px0_g, vx, nv, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __fmul_rn(a[i], x[i]);
      }
      if (pre) unpack_bf16x8(pxl[j], x);
      else load_bf16x8(pxl_g, vx, nv, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __fadd_rn(a[i], x[i]);
      store_bf16x8(po, vo, nv, a);
    }
    pxl_g += 8 * ep.ldx;
    if (px0_g) px0_g += 8 * ep.ldx;
    po += 8 * ep.ldo;
    if (pu) pu += 8 * ep.ldu;
  }
}

// MN = false: A [M,K], B [N,K] (K contiguous, "TN").  MN = true: A [K,M], B [K,N] (M / N contiguous): C = A^T B, the
// weight-gradient shape dW = dY^T X with K = batch -- both operands are read as they are stored, no transposes.
template <int BLOCK_N, int STAGES, bool MN = false>
__global__ void __launch_bounds__(kThreads, 2)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const EpiParams ep) {
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KiB
  constexpr int NB_BOX = (BLOCK_N + 63) / 64;                // MN: 64-column boxes of the B tile
  constexpr uint32_t B_BYTES = MN ? NB_BOX * 64 * BLOCK_K * 2 : BLOCK_N * BLOCK_K * 2;
  constexpr uint32_t TMEM_COLS = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * ep.k_blocks_per_split;
  int kb_end = kb_begin + ep.k_blocks_per_split;
  if (kb_end > ep.k_blocks_total) kb_end = ep.k_blocks_total;
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        mbar_expect_tx(smem_u32(&full_bar[s]), A_BYTES + B_BYTES);
        const int k0 = (kb_begin + i) * BLOCK_K;
        if (MN) {
#pragma unroll
          for (int b = 0; b < BLOCK_M / 64; ++b)
            tma_load_2d(smem_u32(smem_a + s * A_BYTES + b * 8192), &map_a, smem_u32(&full_bar[s]), m_tile * BLOCK_M + b * 64, k0);
#pragma unroll
          for (int b = 0; b < NB_BOX; ++b)
            tma_load_2d(smem_u32(smem_b + s * B_BYTES + b * 8192), &map_b, smem_u32(&full_bar[s]), n_tile * BLOCK_N + b * 64, k0);
        } else {
          tma_load_2d(smem_u32(smem_a + s * A_BYTES), &map_a, smem_u32(&full_bar[s]), k0, m_tile * BLOCK_M);
          tma_load_2d(smem_u32(smem_b + s * B_BYTES), &map_b, smem_u32(&full_bar[s]), k0, n_tile * BLOCK_N);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, MN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        fence_after();
        const uint64_t da = MN ? make_mnmajor_sw128_desc(smem_u32(smem_a + s * A_BYTES)) : make_kmajor_sw128_desc(smem_u32(smem_a + s * A_BYTES));
        const uint64_t db = MN ? make_mnmajor_sw128_desc(smem_u32(smem_b + s * B_BYTES)) : make_kmajor_sw128_desc(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major: advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (>>4) address field;
          // MN-major: advance 16 k rows of 128 bytes = 2048 bytes: +128
          const uint64_t adv = MN ? (uint64_t)(k * 128) : (uint64_t)(k * 2);
          umma_f16(tmem_base, da + adv, db + adv, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));        // frees the smem stage once these MMAs retire
      }
      umma_commit(smem_u32(tmem_full_bar));            // accumulator complete
    }
  } else {
    // ===== epilogue warps 2..9 =====
    // Each warp owns TMEM lanes [32q, 32q+32) = 32 output rows.  tcgen05.ld hands every lane
    // 32 consecutive columns of ITS row; a 32x33 fp32 tile per warp (in the pipeline buffers,
    // which are drained once tmem_full fires) transposes each 32x32 block so that every global
    // load/store instruction touches ONE row's 32 consecutive elements (coalesced).
    const int q = warp & 3;                            // TMEM lane quarter this warp may read
    if (ep.mode == EPI_CROSS) {
      // While the main loop runs, pull this tile's x0 / xl rows into L2: the epilogue's loads of them
      // do not depend on the accumulator, so they should not pay DRAM latency after it is ready.
      const int et = threadIdx.x - 64;                 // 0..255 over the 8 epilogue warps
      const long long tile_col0 = (long long)n_tile * BLOCK_N;
      const int lines = (BLOCK_N * 2 + 127) / 128;     // 128-byte lines per tile row (bf16)
      for (int i = et; i < BLOCK_M * lines; i += 256) {
        const long long rr = (long long)m_tile * BLOCK_M + i / lines;
        const long long cc = tile_col0 + (long long)(i % lines) * 64;
        if (rr < ep.M && cc < ep.N) {
          if (ep.x0) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.x0 + rr * ep.ldx + cc));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.xl + rr * ep.ldx + cc));
        }
      }
    }
    mbar_wait(smem_u32(tmem_full_bar), 0);
    fence_after();
    // per-warp tile, 32 rows x 36 floats: the 36-float stride keeps every 128-bit shared access
    // 16-byte aligned and bank-conflict free in both directions (per quarter-warp)
    constexpr int TS = 36;
    const int half = (warp - 2) >> 2;                  // 0: even 32-column chunks, 1: odd ones
    float* T = reinterpret_cast<float*>(smem_a) + (warp - 2) * (32 * TS);
    const long long row_base = (long long)m_tile * BLOCK_M + q * 32;
#pragma unroll 1
    for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
      const long long col0 = (long long)n_tile * BLOCK_N + c0;
      if (col0 >= ep.N) break;                         // warp-uniform
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      int ncol = (ep.N - col0) < 32 ? (int)(ep.N - col0) : 32;
      if (ncol > BLOCK_N - c0) ncol = BLOCK_N - c0;     // BLOCK_N = 240: the last chunk is 16 columns wide
      const uint4 nopre[4] = {};
      epi_block(ep, T, lane, row_base, col0, ncol, v, split, false, nopre, nopre);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent form for the big K-major ("TN") products of the cross network (3.DCN/CustomLayers.py:300-303 and
// its backward): M = batch, N = K = D.  One CTA per SM walks a static list of output tiles; the accumulator is
// DOUBLE-BUFFERED in TMEM (2 x 256 columns), so the eight epilogue warps drain tile i (x0 / xl blocks preloaded
// into registers before the accumulator is waited for) while the MMA warp is already accumulating tile i + 1, and
// the smem ring never drains between tiles.  CG = 2 pairs two SMs (cluster of 2, tcgen05 cta_group::2): the pair
// owns a 256 x BLOCK_N tile, each CTA stages its own 128 rows of A and HALF of the B tile (BLOCK_N / 2 rows), the
// leader CTA issues UMMA 256 x BLOCK_N x 16 for both, and every B byte is fetched from L2 once per pair instead
// of once per CTA (the 128 x 240 one-CTA tile is shared-memory-fill-bound: TMA writes + MMA operand reads).
//   barriers: full[s]   (leader CTA only; tx bytes of BOTH CTAs' loads, armed by the leader's producer)
//             empty[s]  (per CTA; tcgen05.commit, multicast to both CTAs when CG = 2)
//             tmem_full[a]  (per CTA; commit after a tile's last k-block, multicast when CG = 2)
//             tmem_empty[a] (leader; one arrival per epilogue warp of the whole pair)
// Tile order: n fastest, so the clusters that run side by side read the same A rows (L2 hits).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  if (CG == 2) {
    // the mbarrier may live in the peer (leader) CTA: shared::cluster address
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
        : "memory");
  } else {
    tma_load_2d(dst, map, bar, c_inner, c_outer);
  }
}
template <int CG>
__device__ __forceinline__ void umma_f16_cg(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (CG == 2) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_f16(tmem_d, desc_a, desc_b, idesc, accumulate);
  }
}
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint32_t bar) {
  if (CG == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
  } else {
    umma_commit(bar);
  }
}

constexpr int kEpiWarps = 8;
constexpr int kEpiTileBytes = kEpiWarps * 32 * 36 * 4;       // the epilogue warps' private transpose tiles

template <int BLOCK_N, int STAGES, int CG, bool PIPE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const EpiParams ep,
                         int num_tiles, int n_tiles) {
  static_assert(BLOCK_N % 16 == 0 && (BLOCK_N / CG) % 8 == 0 && BLOCK_N <= 256, "UMMA N");
  constexpr int B_ROWS = BLOCK_N / CG;                        // rows of the B tile this CTA stages
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;         // 16 KiB
  constexpr uint32_t B_TX = B_ROWS * BLOCK_K * 2;
  constexpr uint32_t B_BYTES = (B_TX + 1023) & ~1023u;        // keep every stage 1024-byte aligned
  constexpr uint32_t ACC_COLS = 256;                          // column pitch of the two accumulators
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  float* smem_t = reinterpret_cast<float*>(smem_b + STAGES * B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_t) + kEpiTileBytes);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  const int nkb = ep.k_blocks_total;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&tmem_full_bar[a]), 1); mbar_init(smem_u32(&tmem_empty_bar[a]), CG * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    }
  }
  fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();            // the peer's barriers and TMEM exist before anything signals them
  fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs of a pair: each stages its own A rows and its half of the B tile) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int m_blk = t / n_tiles, n_blk = t % n_tiles;
        const int a_row = (m_blk * CG + (int)rank) * BLOCK_M;
        const int b_row = n_blk * BLOCK_N + (int)rank * B_ROWS;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          uint32_t fb = smem_u32(&full_bar[s]);
          if (rank == 0) mbar_expect_tx(fb, CG * (A_BYTES + B_TX));
          if (CG == 2) fb = mapa_shared(fb, 0);
          tma_load_2d_cg<CG>(smem_u32(smem_a + s * A_BYTES), &map_a, fb, kb * BLOCK_K, a_row);
          tma_load_2d_cg<CG>(smem_u32(smem_b + s * B_BYTES), &map_b, fb, kb * BLOCK_K, b_row);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA =====
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M * CG, BLOCK_N);
      uint32_t it = 0;
      int tl = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++tl) {
        const int acc = tl & 1;
        const uint32_t accph = (tl >> 1) & 1;
        mbar_wait_cluster(smem_u32(&tmem_empty_bar[acc]), accph ^ 1);     // the epilogue has drained this accumulator
        fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          fence_after();
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem_a + s * A_BYTES));
          const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16_cg<CG>(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_cg<CG>(smem_u32(&empty_bar[s]));
        }
        umma_commit_cg<CG>(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===== epilogue warps 2..9 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* T = smem_t + (warp - 2) * (32 * 36);
    const uint32_t empty_addr0 = CG == 2 ? mapa_shared(smem_u32(&tmem_empty_bar[0]), 0) : smem_u32(&tmem_empty_bar[0]);
    const bool cross = ep.mode == EPI_CROSS;
    int tl = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++tl) {
      const int m_blk = t / n_tiles, n_blk = t % n_tiles;
      const int acc = tl & 1;
      const uint32_t accph = (tl >> 1) & 1;
      const long long row_base = ((long long)m_blk * CG + rank) * BLOCK_M + q * 32;
      const long long tile_col0 = (long long)n_blk * BLOCK_N;
      uint4 nx0[4] = {}, nxl[4] = {};
      bool npre = false;
      auto chunk_cols = [&](int c0) {
        int ncol = (ep.N - (tile_col0 + c0)) < 32 ? (int)(ep.N - (tile_col0 + c0)) : 32;
        if (ncol > BLOCK_N - c0) ncol = BLOCK_N - c0;
        return ncol;
      };
      auto preload = [&](int c0) {
        npre = false;
        if (!cross || c0 >= BLOCK_N || tile_col0 + c0 >= ep.N) return;
        const long long col0 = tile_col0 + c0;
        const int ncol = chunk_cols(c0);
        if (!epi_vec_ok(ep.xl, ep.ldx, col0, ncol) || (ep.x0 && !epi_vec_ok(ep.x0, ep.ldx, col0, ncol))) return;
        if (ep.x0) epi_preload(ep.x0, ep.ldx, row_base, col0, ep.M, lane, nx0);
        epi_preload(ep.xl, ep.ldx, row_base, col0, ep.M, lane, nxl);
        npre = true;
      };
      preload(half * 32);                                   // in flight while the accumulator is still being computed
      mbar_wait(smem_u32(&tmem_full_bar[acc]), accph);
      fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * ACC_COLS;
      uint32_t vn[32];
      if (PIPE && tile_col0 + half * 32 < ep.N) {
        tmem_ld32_issue(t_acc + (uint32_t)(half * 32), vn);
        tmem_ld_wait();
      }
      bool released = false;
      auto release = [&]() {                                // all of this warp's reads of the accumulator are done
        released = true;
        fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(empty_addr0 + (uint32_t)acc * 8u);
          else mbar_arrive_local(empty_addr0 + (uint32_t)acc * 8u);
        }
      };
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
        const long long col0 = tile_col0 + c0;
        if (col0 >= ep.N) break;                            // warp-uniform
        uint4 cx0[4], cxl[4];
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) { cx0[j] = nx0[j]; cxl[j] = nxl[j]; }
        const bool cpre = npre;
        const bool more = c0 + 64 < BLOCK_N && col0 + 64 < ep.N;
        if (PIPE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = vn[i];
          if (more) tmem_ld32_issue(t_acc + (uint32_t)(c0 + 64), vn);    // next block: TMEM -> registers under this block's work
        } else {
          tmem_ld32(t_acc + (uint32_t)c0, v);
        }
        if (!more) release();                               // (the last block is already in registers)
        preload(c0 + 64);                                   // the next block's x0 / xl fly during this block's work too
        epi_block(ep, T, lane, row_base, col0, chunk_cols(c0), v, 0, cpre, cx0, cxl);
        if (PIPE && more) tmem_ld_wait();
      }
      if (!released) release();
    }
  }
  fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();            // the peer may still be read by the leader's MMAs / signalled by its commits
  if (warp == 1) {
    fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else tmem_dealloc(tmem_base, 512);
  }
}

// split-K finish: C = act(sum_z partial[z] + bias) (fixed order), fp32 or bf16 out
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* partial, int splits, long long M, long long N,
                                                            void* C, long long ldc, int c_bf16, const float* bias, int act,
                                                            float beta) {
  const long long total = M * N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / N, n = t % N;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(long long)z * total + t];
    if (bias) s += bias[n];
    if (c_bf16) {
      reinterpret_cast<__nv_bfloat16*>(C)[m * ldc + n] = __float2bfloat16_rn(act_apply(s, act));
    } else {
      float* c = reinterpret_cast<float*>(C) + m * ldc + n;
      if (beta != 0.f) s += beta * *c;
      *c = act_apply(s, act);
    }
  }
}

// fp32 [rows, cols] (ld_src) -> bf16, optionally transposed, into a zero-padded
// [rows_dst, ld_dst] buffer (dst rows >= rows (or cols when transposing)).
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* src, long long rows, long long cols, long long ld_src,
                                                        __nv_bfloat16* dst, long long ld_dst, int transpose) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  for (int i = ty; i < 32; i += 8) {
    const long long r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
  if (!transpose) {
    for (int i = ty; i < 32; i += 8) {
      const long long r = r0 + i, c = c0 + tx;
      if (r < rows && c < ld_dst) dst[r * ld_dst + c] = __float2bfloat16_rn(c < cols ? tile[i][tx] : 0.f);
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const long long orow = c0 + i, ocol = r0 + tx;          // dst[c][r]
      if (orow < cols && ocol < ld_dst) dst[orow * ld_dst + ocol] = __float2bfloat16_rn(ocol < rows ? tile[tx][i] : 0.f);
    }
  }
}
// bf16 [rows, cols] -> bf16 transposed [cols, ld_dst]; 64x64 tiles, 4-byte (bf16x2)
// global accesses on both sides (128 B per warp instruction), zero fill up to ld_dst.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* src, long long rows, long long cols,
                                                             long long ld_src, __nv_bfloat16* dst, long long ld_dst) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const bool pair_ok = (ld_src % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
  for (int i = ty; i < 64; i += 8) {
    const long long r = r0 + i, c = c0 + 2 * tx;
    __nv_bfloat16 a = __float2bfloat16_rn(0.f), b = a;
    if (r < rows) {
      if (pair_ok && c + 1 < cols) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + r * ld_src + c);
        a = v.x; b = v.y;
      } else {
        if (c < cols) a = src[r * ld_src + c];
        if (c + 1 < cols) b = src[r * ld_src + c + 1];
      }
    }
    tile[i][2 * tx] = a;
    tile[i][2 * tx + 1] = b;
  }
  __syncthreads();
  const bool opair_ok = (ld_dst % 2 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
  for (int i = ty; i < 64; i += 8) {
    const long long orow = c0 + i, ocol = r0 + 2 * tx;                // dst[c][r]
    if (orow >= cols) continue;
    __nv_bfloat162 v;
    v.x = tile[2 * tx][i];
    v.y = tile[2 * tx + 1][i];
    if (ocol >= rows) v.x = __float2bfloat16_rn(0.f);
    if (ocol + 1 >= rows) v.y = __float2bfloat16_rn(0.f);
    if (opair_ok && ocol + 1 < ld_dst) {
      *reinterpret_cast<__nv_bfloat162*>(dst + orow * ld_dst + ocol) = v;
    } else {
      if (ocol < ld_dst) dst[orow * ld_dst + ocol] = v.x;
      if (ocol + 1 < ld_dst) dst[orow * ld_dst + ocol + 1] = v.y;
    }
  }
}

// bf16 elementwise pieces of the cross-matrix backward (SURVEY a'):
//   du = G (.) x0 (bf16) ; dx0 += G (.) u (fp32 accumulate)
__global__ void __launch_bounds__(256) cross_bwd_elem_bf16_kernel(const __nv_bfloat16* G, long long cols2, long long ldg2,
                                                                  const __nv_bfloat16* x0, const __nv_bfloat16* u, long long n2,
                                                                  __nv_bfloat16* du, float* dx0, int init) {
  // G may be a column slice of a wider matrix (row pitch ldg2 bf16x2 pairs); x0, u, du, dx0 are dense [rows, cols]
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (long long)gridDim.x * blockDim.x) {
    const long long tg = (ldg2 == cols2) ? t : (t / cols2) * ldg2 + t % cols2;
    const float2 g = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(G)[tg]);
    const float2 a = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(x0)[t]);
    const float2 b = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(u)[t]);
    reinterpret_cast<__nv_bfloat162*>(du)[t] = __floats2bfloat162_rn(g.x * a.x, g.y * a.y);
    float2 d = make_float2(0.f, 0.f);
    if (!init) d = reinterpret_cast<float2*>(dx0)[t];
    d.x += g.x * b.x; d.y += g.y * b.y;
    reinterpret_cast<float2*>(dx0)[t] = d;
  }
}
// ---- one-pass forms of the elementwise half of the cross-matrix backward (round 2) ----------------------------
// (a) du = G (.) x0 (bf16) AND the column sums of du (db_l) in the same pass: a thread owns 8 consecutive columns
//     (16-byte accesses), the 8 warps of a CTA walk a 512-row slab, per-slab partial sums are combined in slab order.
constexpr int kDuSlab = 512;
__global__ void __launch_bounds__(256) cross_bwd_du_colsum_kernel(const __nv_bfloat16* G, long long ldg, const __nv_bfloat16* x0,
                                                                  long long rows, int cols, __nv_bfloat16* du, float* partial) {
  __shared__ float sm[8][32][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;              // 8-column chunk
  const bool on = c * 8 < cols;
  const long long r0 = (long long)blockIdx.y * kDuSlab;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int i = 0; i < kDuSlab / 8; ++i) {
    const long long r = r0 + warp + 8 * i;
    if (on && r < rows) {
      float g[8], a[8];
      unpack_bf16x8(ldg_u4(G + r * ldg + c * 8), g);
      unpack_bf16x8(ldg_u4(x0 + r * cols + c * 8), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= a[j];
      const uint4 w = pack_bf16x8(g);
      stg_u4(du + r * cols + c * 8, w);
      unpack_bf16x8(w, a);                           // db sums du AS STORED (bf16-rounded), like colsum(du) did
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[warp][lane][j] = acc[j];
  __syncthreads();
  if (warp == 0 && on) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sm[w][lane][j];
      partial[(long long)blockIdx.y * cols + c * 8 + j] = t;
    }
  }
}
// (b) dx0 = sum_l G_{l+1} (.) u_l  (l = L-1 .. 0, that order)  + G_0, fp32, written ONCE: replaces the fp32 read-modify-write of
//     dx0 in every layer's elementwise kernel and the final bf16 -> fp32 add.
struct Dx0Params {
  const __nv_bfloat16* G[9]; long long ldg[9];       // G[l + 1] pairs with u[l]; G[0] is the gradient that reaches x0 through x_0
  const __nv_bfloat16* u[8];
  const __nv_bfloat16* extra; long long ld_extra;    // optional: one more bf16 term (the other branch's gradient for the same input)
  int layers; long long rows; int cols; float* dx0;
};
__global__ void __launch_bounds__(256) cross_bwd_dx0_kernel(const Dx0Params p) {
  const int nch = p.cols / 8;
  const long long total = p.rows * nch;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / nch;
    const int c = (int)(t % nch);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = p.layers - 1; l >= 0; --l) {
      float g[8], u[8];
      unpack_bf16x8(ldg_u4(p.G[l + 1] + r * p.ldg[l + 1] + c * 8), g);
      unpack_bf16x8(ldg_u4(p.u[l] + r * p.cols + c * 8), u);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (l == p.layers - 1) ? g[j] * u[j] : acc[j] + g[j] * u[j];
    }
    float g0[8];
    unpack_bf16x8(ldg_u4(p.G[0] + r * p.ldg[0] + c * 8), g0);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += g0[j];
    if (p.extra) {
      unpack_bf16x8(ldg_u4(p.extra + r * p.ld_extra + c * 8), g0);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g0[j] = 0.f;
    }
    float* o = p.dx0 + r * p.cols + c * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0] + g0[0], acc[1] + g0[1], acc[2] + g0[2], acc[3] + g0[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] + g0[4], acc[5] + g0[5], acc[6] + g0[6], acc[7] + g0[7]);
  }
}
// out[b, n] = bf16(d[b] * k[n]): the input gradient of a Dense(1) layer (dz K^T is a rank-1 product: no GEMM needed)
__global__ void __launch_bounds__(256) outer_bf16_kernel(const float* d, const float* k, long long rows, int cols, __nv_bfloat16* out,
                                                         long long ldo) {
  const int nch = (cols + 7) / 8;
  const long long total = rows * nch;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / nch;
    const int c = (int)(t % nch) * 8;
    const float dv = d[r];
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (c + j < cols) ? dv * __ldg(k + c + j) : 0.f;
    __nv_bfloat16* o = out + r * ldo + c;
    if (c + 8 <= ldo && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      stg_u4(o, pack_bf16x8(a));
    } else {
      for (int j = 0; j < 8 && c + j < cols; ++j) o[j] = __float2bfloat16_rn(a[j]);
    }
  }
}
// y (fp32) += x (bf16)
__global__ void __launch_bounds__(256) add_bf16_into_f32_kernel(const __nv_bfloat16* x, long long n2, float* y) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (long long)gridDim.x * blockDim.x) {
    const float2 a = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(x)[t]);
    float2 d = reinterpret_cast<float2*>(y)[t];
    d.x += a.x; d.y += a.y;
    reinterpret_cast<float2*>(y)[t] = d;
  }
}
// column sums of a bf16 matrix: stage 1 of the two-pass deterministic reduction
__global__ void __launch_bounds__(256) colsum_bf16_stage1_kernel(const __nv_bfloat16* X, long long M, long long N, long long ldx,
                                                                 float* partial) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long n = (long long)blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * 2048;
  long long r1 = r0 + 2048;
  if (r1 > M) r1 = M;
  float s = 0.f;
  if (n < N)
    for (long long r = r0 + ry; r < r1; r += 8) s += __bfloat162float(X[r * ldx + n]);
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sm[q][cx];
    partial[(long long)blockIdx.y * N + n] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_stage2b_kernel(const float* partial, long long slabs, long long N, float* out) {
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (long long q = 0; q < slabs; ++q) s += partial[q * N + n];
    out[n] = s;
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor map: [rows, cols] with leading dimension ld (elements), box = [box_rows, 64 cols], 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { etr_set_error("cuTensorMapEncodeTiled is not available from the driver"); return ETR_ECUDA; }
  if (((uintptr_t)base & 15) || (ld * 2) % 16 != 0) {
    etr_set_error("tcgen05 GEMM operands need 16-byte aligned base and row pitch (ld %% 8 == 0)");
    return ETR_EINVAL;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { etr_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ETR_ECUDA; }
  return ETR_OK;
}

// MN-major operand [k_rows, mn_cols] (mn contiguous, leading dimension ld): box = [BLOCK_K k rows, 64 mn columns]
static int make_map_mn(CUtensorMap* map, const void* base, long long k_rows, long long mn_cols, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { etr_set_error("cuTensorMapEncodeTiled is not available from the driver"); return ETR_ECUDA; }
  if (((uintptr_t)base & 15) || (ld * 2) % 16 != 0) {
    etr_set_error("tcgen05 GEMM operands need 16-byte aligned base and row pitch (ld %% 8 == 0)");
    return ETR_EINVAL;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)mn_cols, (cuuint64_t)k_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)BLOCK_K};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { etr_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ETR_ECUDA; }
  return ETR_OK;
}

template <int BLOCK_N, int STAGES, bool MN = false>
static int launch_tile(etr_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const EpiParams& ep, dim3 grid,
                       cudaStream_t s) {
  constexpr size_t b_bytes = MN ? (size_t)((BLOCK_N + 63) / 64) * 64 * BLOCK_K * 2 : (size_t)BLOCK_N * BLOCK_K * 2;
  constexpr size_t smem = (size_t)STAGES * (BLOCK_M * BLOCK_K * 2 + b_bytes) + (2 * STAGES + 2) * 8 + 1024;
  // the opt-in is per device: set it on every call (one process may drive several GPUs)
  ETR_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, STAGES, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_bf16_tn_kernel<BLOCK_N, STAGES, MN><<<grid, kThreads, smem, s>>>(ma, mb, ep);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

template <int BLOCK_N, int STAGES, int CG, bool PIPE>
static int launch_persist2(etr_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const EpiParams& ep, int num_tiles,
                          int n_tiles, cudaStream_t s) {
  constexpr size_t b_bytes = ((size_t)(BLOCK_N / CG) * BLOCK_K * 2 + 1023) & ~(size_t)1023;
  constexpr size_t smem = (size_t)STAGES * (BLOCK_M * BLOCK_K * 2 + b_bytes) + kEpiTileBytes + (2 * STAGES + 4) * 8 + 16 + 1024;
  static_assert(smem <= 232448, "shared memory budget of one SM");
  auto kern = gemm_bf16_persist_kernel<BLOCK_N, STAGES, CG, PIPE>;
  ETR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int clusters = ctx->sm_count / CG;
  if (clusters > num_tiles) clusters = num_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * CG));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ETR_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, ep, num_tiles, n_tiles));
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

template <int BLOCK_N, int STAGES, int CG>
static int launch_persist(etr_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const EpiParams& ep, int num_tiles,
                          int n_tiles, cudaStream_t s) {
  const char* e = getenv("ETR_GEMM_EPI_PIPE");        // 1: the next block's tcgen05.ld is issued under the current block's work
  if (e && atoi(e) == 1) return launch_persist2<BLOCK_N, STAGES, CG, true>(ctx, ma, mb, ep, num_tiles, n_tiles, s);
  return launch_persist2<BLOCK_N, STAGES, CG, false>(ctx, ma, mb, ep, num_tiles, n_tiles, s);
}

// ETR_GEMM_PERSIST: 0 = per-tile kernel only, 1 = persistent one-CTA form, 2 (default) = persistent CTA-pair form
static int persist_mode() {
  const char* e = getenv("ETR_GEMM_PERSIST");      // read per call: tests switch it within one process
  const int mode = e ? atoi(e) : 2;
  return (mode < 0 || mode > 2) ? 2 : mode;
}

// BLOCK_N choice: widest tile that does not waste more than ~7% of the N extent.
static int pick_block_n(long long N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const long long t240 = ceil_div(N, 240) * 240, t256 = ceil_div(N, 256) * 256, t128 = ceil_div(N, 128) * 128;
  if (t240 <= t256 && t240 <= t128 + t128 / 16) return 240;
  if (t256 <= t128 + t128 / 16) return 256;
  return 128;
}

static int run_gemm(etr_ctx* ctx, long long M, long long N, long long K, const void* A, long long lda, const void* B,
                    long long ldb, EpiParams ep, int allow_split, cudaStream_t s, void* C, long long ldc, int c_bf16,
                    const float* bias, int act, float beta, bool mn = false) {
  int bn = pick_block_n(N);
  if (mn && bn < 64) bn = 64;                        // MN-major tiles are built from 64-column boxes
  CUtensorMap ma, mb;
  // big K-major products without split-K (cross layers forward / dgrad): persistent kernel, TMEM double-buffered
  const int pm = persist_mode();
  if (pm > 0 && !mn && bn >= 128 && (ep.mode == EPI_LINEAR || ep.mode == EPI_CROSS) &&
      ceil_div(M, BLOCK_M) * ceil_div(N, bn) >= 2LL * ctx->sm_count) {
    const int cg = pm;
    int st = make_map(&ma, A, M, K, lda, BLOCK_M);
    if (st != ETR_OK) return st;
    st = make_map(&mb, B, N, K, ldb, bn / cg);
    if (st != ETR_OK) return st;
    const long long mt2 = ceil_div(M, (long long)BLOCK_M * cg), nt2 = ceil_div(N, bn);
    if (mt2 * nt2 < (1LL << 30)) {
      ep.M = M; ep.N = N;
      ep.k_blocks_total = (int)ceil_div(K, BLOCK_K);
      ep.k_blocks_per_split = ep.k_blocks_total;
      const int tiles = (int)(mt2 * nt2), ntl = (int)nt2;
      if (cg == 2) {
        switch (bn) {
          case 128: return launch_persist<128, 5, 2>(ctx, ma, mb, ep, tiles, ntl, s);
          case 240: return launch_persist<240, 5, 2>(ctx, ma, mb, ep, tiles, ntl, s);
          default: return launch_persist<256, 5, 2>(ctx, ma, mb, ep, tiles, ntl, s);
        }
      }
      switch (bn) {
        case 128: return launch_persist<128, 5, 1>(ctx, ma, mb, ep, tiles, ntl, s);
        case 240: return launch_persist<240, 4, 1>(ctx, ma, mb, ep, tiles, ntl, s);
        default: return launch_persist<256, 3, 1>(ctx, ma, mb, ep, tiles, ntl, s);
      }
    }
  }
  int st = mn ? make_map_mn(&ma, A, K, M, lda) : make_map(&ma, A, M, K, lda, BLOCK_M);
  if (st != ETR_OK) return st;
  st = mn ? make_map_mn(&mb, B, K, N, ldb) : make_map(&mb, B, N, K, ldb, bn);
  if (st != ETR_OK) return st;
  const long long mt = ceil_div(M, BLOCK_M), nt = ceil_div(N, bn);
  const int kbt = (int)ceil_div(K, BLOCK_K);
  int splits = 1;
  if (allow_split && mt * nt < ctx->sm_count && kbt >= 64) {
    splits = (int)((2LL * ctx->sm_count) / (mt * nt));
    if (splits > kbt / 16) splits = kbt / 16;
    if (splits < 1) splits = 1;
  }
  ep.M = M; ep.N = N;
  ep.k_blocks_total = kbt;
  ep.k_blocks_per_split = (int)ceil_div(kbt, splits);
  splits = (int)ceil_div(kbt, ep.k_blocks_per_split);
  if (splits > 1) {
    st = etr_ws_reserve(ctx, sizeof(float) * (size_t)splits * M * N);
    if (st != ETR_OK) return st;
    ep.mode = EPI_PARTIAL;
    ep.partial = (float*)ctx->d_ws;
  }
  if (mt > 65535 || nt > 65535) { etr_set_error("tcgen05 GEMM: too many tiles"); return ETR_EUNSUPPORTED; }
  dim3 grid((unsigned)nt, (unsigned)mt, (unsigned)splits);
  if (mn) {
    switch (bn) {
      case 64: st = launch_tile<64, 4, true>(ctx, ma, mb, ep, grid, s); break;
      case 128: st = launch_tile<128, 3, true>(ctx, ma, mb, ep, grid, s); break;
      case 240: st = launch_tile<240, 2, true>(ctx, ma, mb, ep, grid, s); break;
      default: st = launch_tile<256, 2, true>(ctx, ma, mb, ep, grid, s); break;
    }
  } else {
    switch (bn) {
      case 32: st = launch_tile<32, 4>(ctx, ma, mb, ep, grid, s); break;
      case 64: st = launch_tile<64, 4>(ctx, ma, mb, ep, grid, s); break;
      case 128: st = launch_tile<128, 3>(ctx, ma, mb, ep, grid, s); break;
      case 240: st = launch_tile<240, 2>(ctx, ma, mb, ep, grid, s); break;
      default: st = launch_tile<256, 2>(ctx, ma, mb, ep, grid, s); break;
    }
  }
  if (st != ETR_OK) return st;
  if (splits > 1) {
    splitk_finish_kernel<<<grid_for(M * N, 256, ctx->sm_count, 8), 256, 0, s>>>((const float*)ctx->d_ws, splits, M, N, C,
                                                                              ldc, c_bf16, bias, act, beta);
    ETR_LAUNCH_CHECK(ctx);
  }
  return ETR_OK;
}

}  // namespace tc
}  // namespace etr

using namespace etr;

extern "C" {

int etr_gemm_bf16_tn(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda, const void* d_B,
                     int64_t ldb, void* d_C, int64_t ldc, int32_t c_dtype, const float* d_bias, int32_t act,
                     void* stream) {
  ETR_CHECK_ARG(ctx && d_A && d_B && d_C, "NULL argument");
  ETR_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  tc::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = tc::EPI_LINEAR;
  ep.C = d_C; ep.ldc = ldc; ep.c_bf16 = c_dtype == ETR_BF16; ep.bias = d_bias; ep.act = act;
  return tc::run_gemm(ctx, M, N, K, d_A, lda, d_B, ldb, ep, 1, (cudaStream_t)stream, d_C, ldc, ep.c_bf16, d_bias, act, 0.f);
}

int etr_gemm_bf16_nn_wgrad(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda, const void* d_B,
                           int64_t ldb, float* d_C, int64_t ldc, void* stream) {
  ETR_CHECK_ARG(ctx && d_A && d_B && d_C, "NULL argument");
  ETR_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  tc::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = tc::EPI_LINEAR;
  ep.C = d_C; ep.ldc = ldc; ep.c_bf16 = 0;
  return tc::run_gemm(ctx, M, N, K, d_A, lda, d_B, ldb, ep, 1, (cudaStream_t)stream, d_C, ldc, 0, nullptr, 0, 0.f, true);
}

int etr_gemm_bf16_tn_accumulate(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda, const void* d_B,
                                int64_t ldb, float* d_C, int64_t ldc, void* stream) {
  ETR_CHECK_ARG(ctx && d_A && d_B && d_C, "NULL argument");
  ETR_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  tc::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = tc::EPI_LINEAR;
  ep.C = d_C; ep.ldc = ldc; ep.c_bf16 = 0; ep.accumulate = 1;
  return tc::run_gemm(ctx, M, N, K, d_A, lda, d_B, ldb, ep, 0, (cudaStream_t)stream, d_C, ldc, 0, nullptr, 0, 0.f);
}

int etr_cross_mat_layer_bf16(etr_ctx* ctx, const void* d_x0, const void* d_xl, int64_t ldx, int64_t batch, int32_t D,
                             const void* d_W, int64_t ldw, const float* d_b, void* d_out, int64_t ldo, void* d_u,
                             int64_t ldu, void* stream) {
  ETR_CHECK_ARG(ctx && d_x0 && d_xl && d_W && d_b && d_out, "NULL argument");
  ETR_CHECK_ARG(batch > 0 && D > 0, "empty problem");
  tc::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = tc::EPI_CROSS;
  ep.bias = d_b;
  ep.x0 = (const __nv_bfloat16*)d_x0; ep.xl = (const __nv_bfloat16*)d_xl; ep.ldx = ldx;
  ep.out = (__nv_bfloat16*)d_out; ep.ldo = ldo; ep.u = (__nv_bfloat16*)d_u; ep.ldu = ldu;
  // U = X_l W^T : A = X_l [B, D], B operand = W [N = D, K = D] row-major
  return tc::run_gemm(ctx, batch, D, D, d_xl, ldx, d_W, ldw, ep, 0, (cudaStream_t)stream, nullptr, 0, 1, nullptr, 0, 0.f);
}

int etr_gemm_bf16_tn_residual(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda,
                              const void* d_B, int64_t ldb, const void* d_res, int64_t ldr, void* d_out, int64_t ldo,
                              void* stream) {
  ETR_CHECK_ARG(ctx && d_A && d_B && d_res && d_out, "NULL argument");
  ETR_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  tc::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = tc::EPI_CROSS;                  // x0 == NULL, bias == NULL: out = acc + res
  ep.xl = (const __nv_bfloat16*)d_res; ep.ldx = ldr;
  ep.out = (__nv_bfloat16*)d_out; ep.ldo = ldo;
  return tc::run_gemm(ctx, M, N, K, d_A, lda, d_B, ldb, ep, 0, (cudaStream_t)stream, nullptr, 0, 1, nullptr, 0, 0.f);
}

int etr_cross_mat_bwd_elementwise_bf16(etr_ctx* ctx, const void* d_g, int64_t ldg, const void* d_x0, const void* d_u,
                                       int64_t rows, int64_t cols, void* d_du, float* d_dx0_accum, int32_t init, void* stream) {
  ETR_CHECK_ARG(ctx && d_g && d_x0 && d_u && d_du && d_dx0_accum, "NULL argument");
  ETR_CHECK_ARG(cols % 2 == 0 && ldg % 2 == 0 && ldg >= cols && ((uintptr_t)d_g & 3) == 0, "cols / ldg must be even (bf16x2 accesses)");
  const int64_t n = rows * cols;
  if (n <= 0) return ETR_OK;
  tc::cross_bwd_elem_bf16_kernel<<<grid_for(n / 2, 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)d_g, cols / 2, ldg / 2, (const __nv_bfloat16*)d_x0, (const __nv_bfloat16*)d_u, n / 2,
      (__nv_bfloat16*)d_du, d_dx0_accum, init);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cross_mat_bwd_du_colsum_bf16(etr_ctx* ctx, const void* d_g, int64_t ldg, const void* d_x0, int64_t rows, int64_t cols,
                                     void* d_du, float* d_db, void* stream) {
  ETR_CHECK_ARG(ctx && d_g && d_x0 && d_du && d_db, "NULL argument");
  ETR_CHECK_ARG(cols % 8 == 0 && ldg % 8 == 0 && ldg >= cols && ((uintptr_t)d_g & 15) == 0 && ((uintptr_t)d_x0 & 15) == 0 &&
                    ((uintptr_t)d_du & 15) == 0, "cols / ldg must be multiples of 8 and the matrices 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (cols <= 0) return ETR_OK;
  if (rows <= 0) { ETR_CUDA(cudaMemsetAsync(d_db, 0, sizeof(float) * cols, s)); return ETR_OK; }
  const long long slabs = ceil_div(rows, tc::kDuSlab);
  ETR_CHECK_ARG(slabs <= 65535, "too many rows");
  int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)slabs * cols);
  if (st != ETR_OK) return st;
  dim3 grid((unsigned)ceil_div(cols / 8, 32), (unsigned)slabs);
  tc::cross_bwd_du_colsum_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)d_g, ldg, (const __nv_bfloat16*)d_x0, rows, (int)cols,
                                                      (__nv_bfloat16*)d_du, (float*)ctx->d_ws);
  ETR_LAUNCH_CHECK(ctx);
  tc::colsum_stage2b_kernel<<<grid_for(cols, 256, ctx->sm_count, 1), 256, 0, s>>>((const float*)ctx->d_ws, slabs, cols, d_db);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cross_mat_bwd_dx0_bf16(etr_ctx* ctx, int32_t layers, const void* const* h_G, const int64_t* h_ldg, const void* const* h_u,
                               const void* d_extra, int64_t ld_extra, int64_t rows, int64_t cols, float* d_dx0, void* stream) {
  ETR_CHECK_ARG(ctx && h_G && h_ldg && h_u && d_dx0, "NULL argument");
  ETR_CHECK_ARG(layers >= 1 && layers <= 8, "1..8 cross layers");
  ETR_CHECK_ARG(cols % 8 == 0 && ((uintptr_t)d_dx0 & 15) == 0, "cols must be a multiple of 8");
  if (rows <= 0 || cols <= 0) return ETR_OK;
  tc::Dx0Params p;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l <= layers; ++l) {
    ETR_CHECK_ARG(h_G[l] && h_ldg[l] % 8 == 0 && h_ldg[l] >= cols && ((uintptr_t)h_G[l] & 15) == 0, "G[l]: 16-byte aligned rows");
    p.G[l] = (const __nv_bfloat16*)h_G[l]; p.ldg[l] = h_ldg[l];
  }
  for (int l = 0; l < layers; ++l) {
    ETR_CHECK_ARG(h_u[l] && ((uintptr_t)h_u[l] & 15) == 0, "u[l]: 16-byte aligned");
    p.u[l] = (const __nv_bfloat16*)h_u[l];
  }
  ETR_CHECK_ARG(!d_extra || (ld_extra % 8 == 0 && ld_extra >= cols && ((uintptr_t)d_extra & 15) == 0), "extra: 16-byte aligned rows");
  p.extra = (const __nv_bfloat16*)d_extra; p.ld_extra = ld_extra;
  p.layers = layers; p.rows = rows; p.cols = (int)cols; p.dx0 = d_dx0;
  tc::cross_bwd_dx0_kernel<<<grid_for(rows * (cols / 8), 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_outer_bf16(etr_ctx* ctx, const float* d_d, const float* d_k, int64_t rows, int64_t cols, void* d_out, int64_t ldo, void* stream) {
  ETR_CHECK_ARG(ctx && d_d && d_k && d_out, "NULL argument");
  ETR_CHECK_ARG(ldo >= cols, "ldo < cols");
  if (rows <= 0 || cols <= 0) return ETR_OK;
  tc::outer_bf16_kernel<<<grid_for(rows * ((cols + 7) / 8), 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(
      d_d, d_k, rows, (int)cols, (__nv_bfloat16*)d_out, ldo);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_add_bf16_into_f32(etr_ctx* ctx, const void* d_x, int64_t n, float* d_y, void* stream) {
  ETR_CHECK_ARG(ctx && d_x && d_y, "NULL argument");
  ETR_CHECK_ARG(n % 2 == 0, "n must be even");
  if (n <= 0) return ETR_OK;
  tc::add_bf16_into_f32_kernel<<<grid_for(n / 2, 256, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)d_x, n / 2, d_y);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_colsum_bf16(etr_ctx* ctx, const void* d_X, int64_t M, int64_t N, int64_t ldx, float* d_out, void* stream) {
  ETR_CHECK_ARG(ctx && d_X && d_out, "NULL argument");
  if (N <= 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (M <= 0) { ETR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float) * N, s)); return ETR_OK; }
  const long long slabs = ceil_div(M, 2048);
  ETR_CHECK_ARG(slabs <= 65535, "M too large");
  int st = etr_ws_reserve(ctx, sizeof(float) * (size_t)slabs * N);
  if (st != ETR_OK) return st;
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)slabs);
  tc::colsum_bf16_stage1_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)d_X, M, N, ldx, (float*)ctx->d_ws);
  ETR_LAUNCH_CHECK(ctx);
  tc::colsum_stage2b_kernel<<<grid_for(N, 256, ctx->sm_count, 1), 256, 0, s>>>((const float*)ctx->d_ws, slabs, N, d_out);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_cast_bf16(etr_ctx* ctx, const float* d_src, int64_t rows, int64_t cols, int64_t ld_src, void* d_dst,
                  int64_t ld_dst, int32_t transpose, void* stream) {
  ETR_CHECK_ARG(ctx && d_src && d_dst, "NULL argument");
  if (rows <= 0 || cols <= 0) return ETR_OK;
  // cover the padded destination width too (zero fill up to ld_dst)
  const long long out_cols = transpose ? rows : cols;
  const long long span_c = transpose ? cols : (ld_dst > cols ? ld_dst : cols);
  const long long span_r = transpose ? (ld_dst > rows ? ld_dst : rows) : rows;
  (void)out_cols;
  dim3 grid((unsigned)ceil_div(span_c, 32), (unsigned)ceil_div(span_r, 32));
  tc::cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, rows, cols, ld_src, (__nv_bfloat16*)d_dst, ld_dst,
                                                              transpose);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_transpose_bf16(etr_ctx* ctx, const void* d_src, int64_t rows, int64_t cols, int64_t ld_src, void* d_dst,
                       int64_t ld_dst, void* stream) {
  ETR_CHECK_ARG(ctx && d_src && d_dst, "NULL argument");
  if (rows <= 0 || cols <= 0) return ETR_OK;
  const long long span_r = ld_dst > rows ? ld_dst : rows;
  dim3 grid((unsigned)ceil_div(cols, 64), (unsigned)ceil_div(span_r, 64));
  tc::transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)d_src, rows, cols, ld_src,
                                                                   (__nv_bfloat16*)d_dst, ld_dst);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
