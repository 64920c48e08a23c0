// Shared device/host helpers for libetr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/etr.h"

// A row-sharded table as seen from one rank: shard g (rows g, g+G, g+2G, ...) lives in GPU g's HBM
// and is mapped here through CUDA IPC (NVLink peer memory); base[rank] is the local shard.
struct EtrShardSet {
  const char* base[16];
  int world;
  int rank;
  long long rows_global;
};
constexpr int kMaxShardSets = 8;

struct etr_ctx {
  int device;
  int sm_count;
  EtrShardSet shard_sets[kMaxShardSets];
  int n_shard_sets;
  // device error word: [0] flag, [1] first offending id (best effort)
  unsigned long long* d_err;
  // reusable workspace (CUB temp storage, long-segment lists, partial sums)
  void* d_ws;
  size_t ws_bytes;
  long long launches;
  // fork / join inside one entry point (independent kernels of a call overlap; capturable: the side
  // stream always joins back before the call returns)
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
};

void etr_set_error(const char* fmt, ...);
int etr_ws_reserve(etr_ctx* ctx, size_t bytes);   // grows ctx->d_ws (may sync); returns status

#define ETR_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      etr_set_error("%s: %s", __func__, msg);                      \
      return ETR_EINVAL;                                           \
    }                                                              \
  } while (0)

#define ETR_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      etr_set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(e__), \
                    __FILE__, __LINE__);                                            \
      return ETR_ECUDA;                                                             \
    }                                                                               \
  } while (0)

#define ETR_LAUNCH_CHECK(ctx)                 \
  do {                                        \
    (ctx)->launches++;                        \
    ETR_CUDA(cudaGetLastError());             \
  } while (0)

namespace etr {

constexpr int kWarp = 32;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- 128-bit global accesses -------------------------------------------
// Table rows are read through the read-only path; they are re-used only via
// L2 (hot ids), never via L1, so do not allocate them in L1.
__device__ __forceinline__ float4 ldg_row16(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// streaming 128-bit store (outputs are written once, consumed by a later kernel)
__device__ __forceinline__ void stg_stream16(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// 8 elements of a row chunk, always widened to fp32 in registers.
template <typename Elem>
struct Chunk;
template <>
struct Chunk<float> {
  static constexpr int kElems = 4;
  float v[4];
  __device__ __forceinline__ void load(const void* p) {
    float4 r = ldg_row16(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = v[3] = 0.f; }
};
template <>
struct Chunk<__nv_bfloat16> {
  static constexpr int kElems = 8;
  float v[8];
  __device__ __forceinline__ void load(const void* p) {
    float4 r = ldg_row16(p);
    const uint32_t u[4] = {__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z),
                           __float_as_uint(r.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);            // low half  = element 2i
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);  // high half = element 2i+1
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
};

__device__ __forceinline__ float sigmoidf_exact(float z) { return 1.0f / (1.0f + expf(-z)); }

// device error word: err[0] = code of the FIRST error since the last poll (0 none, 1 embedding id out of range,
// 2 mailbox region overflow, 3 peer-barrier timeout, 4 touched-list overflow), err[1] = its detail (the id)
constexpr unsigned long long kErrBadId = 1, kErrMailbox = 2, kErrBarrier = 3, kErrTouched = 4;
__device__ __forceinline__ void flag_error(unsigned long long* err, unsigned long long code, long long detail) {
  if (atomicCAS(&err[0], 0ull, code) == 0ull) err[1] = (unsigned long long)detail;
}
__device__ __forceinline__ void flag_bad_id(unsigned long long* err, long long id) { flag_error(err, kErrBadId, id); }

template <int W>
__device__ __forceinline__ float group_sum(float x, unsigned mask = 0xffffffffu) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
  return x;
}

inline int grid_for(int64_t work_items, int per_block, int sm_count, int blocks_per_sm) {
  int64_t need = ceil_div(work_items, per_block);
  int64_t cap = (int64_t)sm_count * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace etr
