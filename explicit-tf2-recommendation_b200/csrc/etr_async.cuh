// mbarrier / bulk-copy (1-D TMA) helpers and the approximate-MUFU Adam update shared by the record-layout
// kernels (fm_fused_apply.cu, fm_fused_flat.cu).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace etr {

__device__ __forceinline__ uint32_t rec_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rec_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rec_mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t rec_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void rec_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void rec_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_div(float a, float b) {
  float r;
  asm("div.approx.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void adam_update1_fast(float& var, float& m, float& v, const float g, float lr_t, float b1,
                                                  float b2, float eps) {
  m = b1 * m + (1.0f - b1) * g;
  v = b2 * v + (1.0f - b2) * g * g;
  var = var - fast_div(lr_t * m, fast_sqrt(v) + eps);
}
__device__ __forceinline__ void adam_update4_fast(float4& var, float4& m, float4& v, const float4 g, float lr_t, float b1,
                                                  float b2, float eps) {
  adam_update1_fast(var.x, m.x, v.x, g.x, lr_t, b1, b2, eps);
  adam_update1_fast(var.y, m.y, v.y, g.y, lr_t, b1, b2, eps);
  adam_update1_fast(var.z, m.z, v.z, g.z, lr_t, b1, b2, eps);
  adam_update1_fast(var.w, m.w, v.w, g.w, lr_t, b1, b2, eps);
}


}  // namespace etr
