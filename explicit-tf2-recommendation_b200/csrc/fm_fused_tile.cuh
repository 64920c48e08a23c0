// Arguments of the tiled fused FM backward + Adam kernels (csrc/fm_fused_tile.cu), filled by fused_impl in
// csrc/fm_fused_apply.cu for the hot configuration: Adam apply on a local RECORD table, k = 16, no dflat or a bf16 dflat.
#pragma once
#include "etr_common.cuh"

namespace etr {

struct TileArgs {
  float* table;                       // RECORD base: row r at table + r * 64, [var 0..19 | m 20..39 | v 40..59 | pad]
  const int* sorted_bag; const int* seg_start; const long long* unique_ids; const int* n_unique;
  const float* dlogit; const float* sumv;
  const __nv_bfloat16* dflat;         // NULL or [B, flat_ld] bf16, already advanced by flat_col0
  long long flat_ld;
  int F; unsigned long long magic; int shift;       // b = (bag * magic) >> shift == bag / F
  float lr_t; const float* d_lr_t; float b1, b2, eps;
  long long n_slots;
  // push mode (peer-sharded table, deferred export): instead of the Adam update, row u = [P_0..P_15, sum_g, 0, 0, 0] is
  // stored into the owner's gradient mailbox, slot_of_u[u] = owner * cap + slot (NVLink peer stores)
  const int* slot_of_u; int cap; int gld; float* grads_mb[16];
};

// row descriptors + long-run items of a plan (depend on the ids only): prepared once per plan into a caller-owned
// buffer of fused_tile_prep_bytes(n_slots) bytes, reused by every apply of that plan
size_t fused_tile_prep_bytes(long long n_slots);
int fused_tile_prepare(etr_ctx* ctx, const int* d_seg_start, const long long* d_unique_ids, const int* d_n_unique, long long n_slots,
                       void* d_prep, cudaStream_t s);
// one kernel on `s`; d_prep == NULL: the lists are first built in the ctx workspace (one more launch)
int fused_tile_launch(etr_ctx* ctx, const TileArgs& a, const void* d_prep, cudaStream_t s);

}  // namespace etr
