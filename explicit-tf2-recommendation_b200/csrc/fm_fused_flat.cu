// K2 + K8 fused, occurrence-parallel form (round 2) -- FM / DeepFM, single-hot ids, k = 16, RECORD table.
//
// Replaces tape.gradient + Keras apply_gradients on IndexedSlices (2.FM/ModelManager.py:176-178) for the FM
// family like fm_fused_apply.cu does, but walks the sorted OCCURRENCE list instead of the unique-row list:
//
//   dv_r = sum_{(b,f) in run r} [ g_b S_b + dflat(b,f) ]  -  v_r sum g_b ,      dw_r = sum g_b
//
// Why: the row-parallel kernel gives every 4-lane group one row and lets it chase seg_start -> sorted_bag ->
// (g, S, dflat) through three dependent memory round trips per pair of occurrences; a warp is as slow as the
// longest of its 8 runs (ncu, r02_prof_apply_old_record: 15 000 cycles per 8-row pass, issue slots 38 % used,
// long-scoreboard stalls 12.6 warps per issue) and runs > 64 need three more kernels.  Here
//   * every warp owns ONE contiguous range of the sorted list (R occurrences, ranges dealt round-robin to the
//     CTAs), so the work per step is 8 occurrences whatever the run lengths are;
//   * ids and bag indices stream in coalesced, one 32-occurrence batch ahead; the (g, S, dflat) gathers of a
//     whole batch (4 steps) are issued together -- no load depends on another load of the same batch;
//   * runs are reduced by a segmented scan over the 8 lane groups (3 shuffle rounds) with a carry between
//     steps: fixed tree, deterministic, any run length;
//   * a finished row (its last occurrence seen) is QUEUED in shared memory: gradient (16 + 1 floats) and id go
//     to a slot, the row's 256-byte record [var|m|v] is fetched into the same slot by cp.async.bulk (1-D TMA,
//     mbarrier complete_tx) -- no register is held while it flies;
//   * when 8 rows are queued, the warp runs Adam on all of them at once (4 lanes per row, every lane busy) and
//     writes the records back with 128-bit stores;
//   * a run that crosses a range boundary leaves its partial sums in a 2-slot-per-range entry list; a second,
//     tiny kernel sums the pieces of each such run in range order and applies Adam (deterministic).
#include <algorithm>

#include "etr_common.cuh"
#include "etr_async.cuh"

namespace etr {

constexpr int kQSlots = 24;                 // queued rows per warp: three batches of 8 (one filling, one in flight, one being applied)
constexpr int kQRecBytes = 272;             // record slot stride (256-byte record + 16: at most 2-way bank conflicts on the 3 reads per flush)
constexpr int kQWarpBytes = kQSlots * kQRecBytes + kQSlots * 64 + kQSlots * 4 + kQSlots * 4 + 32;   // rec | grad | gs | key | 3 mbarriers
constexpr int kFlatRange = 128;             // occurrences per range (4 batches of 32); ranges are dealt round-robin to the warps

struct FlatEntry {            // partial sums of a run that crosses a range boundary (80 bytes)
  unsigned key;
  unsigned state;             // 0 empty, 1 last piece of its run, 2 the run continues into the next range
  float gs;
  float pad;
  float v[16];
};

struct FlatParams {
  char* table;                              // RECORD table base (256-byte records)
  long long rows;
  int F; unsigned long long magic; int shift;
  const unsigned* keys; const int* bags; long long n;
  const float* dlogit; const float* sumv;
  const void* dflat; long long flat_ld; int flat_col0;
  float lr_t; const float* d_lr_t; float b1, b2, eps;
  FlatEntry* ent; long long n_ranges;
};

__device__ __forceinline__ void flat_bag_to_bf(const FlatParams& p, int bag, int& b, int& f) {
  b = (int)(((unsigned long long)(unsigned)bag * p.magic) >> p.shift);
  f = bag - b * p.F;
}

// Adam on one queued / combined row: var, m, v are this lane's float4 column chunk of the record at ``rec`` (shared
// or global), P the summed gradient terms, gs = sum g; result stored to the global record ``grow``.
__device__ __forceinline__ void flat_adam_row(const FlatParams& p, const float* rec, float* grow, int gl, float4 P, float gs,
                                              float lr_t) {
  float4 var = *reinterpret_cast<const float4*>(rec + gl * 4);
  float4 m = *reinterpret_cast<const float4*>(rec + 20 + gl * 4);
  float4 v = *reinterpret_cast<const float4*>(rec + 40 + gl * 4);
  const float4 gr = make_float4(P.x - var.x * gs, P.y - var.y * gs, P.z - var.z * gs, P.w - var.w * gs);
  adam_update4_fast(var, m, v, gr, lr_t, p.b1, p.b2, p.eps);
  *reinterpret_cast<float4*>(grow + gl * 4) = var;
  *reinterpret_cast<float4*>(grow + 20 + gl * 4) = m;
  *reinterpret_cast<float4*>(grow + 40 + gl * 4) = v;
  if (gl == 0) {
    float wv = rec[16], wm = rec[36], wvv = rec[56];
    adam_update1_fast(wv, wm, wvv, gs, lr_t, p.b1, p.b2, p.eps);
    grow[16] = wv; grow[36] = wm; grow[56] = wvv;
  }
}

template <int DFLAT>      // 0: no dflat, 1: bf16, 2: fp32
__global__ void __launch_bounds__(256, 3) fm_fused_flat_kernel(const FlatParams p) {
  extern __shared__ __align__(128) unsigned char flat_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gl = lane & 3, g = lane >> 2;
  unsigned char* wbase = flat_smem + (size_t)warp * kQWarpBytes;
  unsigned char* q_rec = wbase;
  float* q_grad = reinterpret_cast<float*>(wbase + kQSlots * kQRecBytes);             // [24][16]
  float* q_gs = q_grad + kQSlots * 16;                                                // [24]
  unsigned* q_key = reinterpret_cast<unsigned*>(q_gs + kQSlots);                      // [24]
  unsigned long long* q_bar = reinterpret_cast<unsigned long long*>(q_key + kQSlots); // [3]
  if (lane == 0) {
    rec_mbar_init(rec_smem_u32(q_bar), 1);
    rec_mbar_init(rec_smem_u32(q_bar + 1), 1);
    rec_mbar_init(rec_smem_u32(q_bar + 2), 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const float lr_t = p.d_lr_t ? *p.d_lr_t : p.lr_t;
  const long long n = p.n;
  const unsigned rows = (unsigned)p.rows;
  const long long W = (long long)gridDim.x * (blockDim.x >> 5);
  const long long w0 = (long long)warp * gridDim.x + blockIdx.x;     // consecutive ranges go to different CTAs
  unsigned q_total = 0;             // rows queued so far: slot = q_total % 24, batch = (q_total / 8) % 3
  unsigned q_flushed = 0;           // batches applied so far
  unsigned par = 0;                 // bit b: mbarrier phase parity of batch b

  // Adam on the ``count`` (<= 8) queued rows of batch ``bq``
  auto flush = [&](int bq, int count) {
    const uint32_t bar = rec_smem_u32(q_bar + bq);
    if (lane == 0) rec_mbar_arrive_tx(bar, (uint32_t)count * 256u);
    while (!rec_mbar_try_wait(bar, (par >> bq) & 1u)) {}
    par ^= 1u << bq;
    if (g < count) {
      const int slot = bq * 8 + g;
      const float* rec = reinterpret_cast<const float*>(q_rec + slot * kQRecBytes);
      const float4 P = *reinterpret_cast<const float4*>(q_grad + slot * 16 + gl * 4);
      const float gs = q_gs[slot];
      const unsigned key = q_key[slot];
      flat_adam_row(p, rec, reinterpret_cast<float*>(p.table + (size_t)key * 256), gl, P, gs, lr_t);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // slot reads are done before the copy engine refills them
    __syncwarp();
    ++q_flushed;
  };

  // ids / bags of the first batch of this warp's first range (later ranges: prefetched by the previous range's last batch)
  long long r = w0;
  unsigned kc = 0xffffffffu, prev_key = 0xffffffffu;
  int bc = 0;
  if (r < p.n_ranges) {
    const long long i0 = r * kFlatRange;
    kc = (i0 + lane < n) ? __ldg(p.keys + i0 + lane) : 0xffffffffu;
    bc = (i0 + lane < n) ? __ldg(p.bags + i0 + lane) : 0;
    prev_key = i0 > 0 ? __ldg(p.keys + i0 - 1) : 0xffffffffu;
  }
  for (; r < p.n_ranges; r += W) {
    const long long i0 = r * kFlatRange;
    const long long i1 = (i0 + kFlatRange < n) ? i0 + kFlatRange : n;
    const bool first_open = i0 > 0 && prev_key == __shfl_sync(0xffffffffu, kc, 0) && prev_key < rows;
    bool tails_seen = false;
    // the open run (the one the previous step ended in): reduced part (carry_*, replicated over the groups) + per-group
    // private sums of steps that lay entirely inside it (acc_*), folded in only when the run ends
    float4 carry_v = make_float4(0.f, 0.f, 0.f, 0.f), acc_v = make_float4(0.f, 0.f, 0.f, 0.f);
    float carry_g = 0.f, acc_g = 0.f;
    bool acc_dirty = false;
    unsigned carry_key = 0xffffffffu;
    bool first_step = true;
    unsigned key_after = 0xffffffffu;      // id of the occurrence that follows this range
    for (long long base = i0; base < i1; base += 32) {
      // ids / bags of the NEXT batch: of this range, or -- at its last batch -- of this warp's next range
      const bool last_batch = base + 32 >= i1;
      const long long nbase = last_batch ? (r + W) * kFlatRange : base + 32;
      const long long nx = nbase + lane;
      const bool nvalid = (!last_batch || r + W < p.n_ranges) && nx < n;
      const unsigned kn = nvalid ? __ldg(p.keys + nx) : 0xffffffffu;
      const int bn = nvalid ? __ldg(p.bags + nx) : 0;
      unsigned pk_next = 0xffffffffu;
      if (last_batch) {
        key_after = i1 < n ? __ldg(p.keys + i1) : 0xffffffffu;
        if (r + W < p.n_ranges && nbase > 0 && nbase <= n) pk_next = __ldg(p.keys + nbase - 1);
      }
      // ---- issue every gather of this batch (4 steps x 8 occurrences): nothing below depends on another load
      unsigned key4[4];
      bool ok4[4];
      float gv[4];
      float4 sv[4];
      float4 dv[4];      // fp32 dflat slices (DFLAT == 2)
      uint2 dr[4];       // bf16 dflat slices, decoded when used (DFLAT == 1)
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int src = 8 * s + g;
        key4[s] = __shfl_sync(0xffffffffu, kc, src);
        const int bag = __shfl_sync(0xffffffffu, bc, src);
        ok4[s] = (base + src < i1) && key4[s] < rows;
        gv[s] = 0.f;
        sv[s] = dv[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        dr[s] = make_uint2(0u, 0u);
        if (ok4[s]) {
          int b, f;
          flat_bag_to_bf(p, bag, b, f);
          gv[s] = __ldg(p.dlogit + b);
          sv[s] = *reinterpret_cast<const float4*>(p.sumv + (long long)b * 16 + gl * 4);
          if (DFLAT == 1) {
            const long long e0 = (long long)b * p.flat_ld + p.flat_col0 + f * 16 + gl * 4;
            dr[s] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.dflat) + e0);
          } else if (DFLAT == 2) {
            const long long e0 = (long long)b * p.flat_ld + p.flat_col0 + f * 16 + gl * 4;
            dv[s] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dflat) + e0);
          }
        }
      }
      // ---- reduce: 4 steps of 8 occurrences
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int src = 8 * s + g;
        const unsigned key = key4[s];
        const bool ok = ok4[s];
        // id of the occurrence after this one (next lane group, next step, next batch, or past the range)
        const unsigned k_in = __shfl_sync(0xffffffffu, kc, (src + 1) & 31);
        const unsigned k_nx = last_batch ? key_after : __shfl_sync(0xffffffffu, kn, 0);
        const unsigned key_next = (base + src + 1 >= i1) ? key_after : ((src < 31) ? k_in : k_nx);
        // id of the occurrence before (previous group; group 0: the carry of the previous step)
        const unsigned k_pv = __shfl_up_sync(0xffffffffu, key, 4);
        const unsigned key_prev = (g == 0) ? carry_key : k_pv;
        bool h = (first_step && g == 0) || key != key_prev;       // head of a run (a range's first occurrence starts fresh)
        const bool tail = ok && key != key_next;
        float4 dd = dv[s];
        if (DFLAT == 1)
          dd = make_float4(__uint_as_float(dr[s].x << 16), __uint_as_float(dr[s].x & 0xffff0000u),
                           __uint_as_float(dr[s].y << 16), __uint_as_float(dr[s].y & 0xffff0000u));
        float4 c = make_float4(gv[s] * sv[s].x + dd.x, gv[s] * sv[s].y + dd.y, gv[s] * sv[s].z + dd.z, gv[s] * sv[s].w + dd.w);
        float gs = gv[s];
        first_step = false;
        if (!__any_sync(0xffffffffu, h || tail)) {
          // all 8 occurrences continue the open run and none ends it: private add, no shuffles
          acc_v.x += c.x; acc_v.y += c.y; acc_v.z += c.z; acc_v.w += c.w;
          acc_g += gs;
          acc_dirty = true;
          carry_key = __shfl_sync(0xffffffffu, key, 28);
          continue;
        }
        if (acc_dirty) {            // fold the private sums into the carry (fixed tree over the 8 groups)
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            acc_v.x += __shfl_xor_sync(0xffffffffu, acc_v.x, o); acc_v.y += __shfl_xor_sync(0xffffffffu, acc_v.y, o);
            acc_v.z += __shfl_xor_sync(0xffffffffu, acc_v.z, o); acc_v.w += __shfl_xor_sync(0xffffffffu, acc_v.w, o);
            acc_g += __shfl_xor_sync(0xffffffffu, acc_g, o);
          }
          carry_v.x += acc_v.x; carry_v.y += acc_v.y; carry_v.z += acc_v.z; carry_v.w += acc_v.w;
          carry_g += acc_g;
          acc_v = make_float4(0.f, 0.f, 0.f, 0.f);
          acc_g = 0.f;
          acc_dirty = false;
        }
        // segmented inclusive scan over the 8 lane groups (Hillis-Steele, fixed tree)
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
          const float px = __shfl_up_sync(0xffffffffu, c.x, 4 * d), py = __shfl_up_sync(0xffffffffu, c.y, 4 * d);
          const float pz = __shfl_up_sync(0xffffffffu, c.z, 4 * d), pw = __shfl_up_sync(0xffffffffu, c.w, 4 * d);
          const float pg = __shfl_up_sync(0xffffffffu, gs, 4 * d);
          const int ph = __shfl_up_sync(0xffffffffu, (int)h, 4 * d);
          if (g >= d && !h) {
            c.x += px; c.y += py; c.z += pz; c.w += pw;
            gs += pg;
            h = ph != 0;
          }
        }
        if (!h) {                     // no run head in groups 0..g: the run continues from the previous step
          c.x += carry_v.x; c.y += carry_v.y; c.z += carry_v.z; c.w += carry_v.w;
          gs += carry_g;
        }
        carry_v.x = __shfl_sync(0xffffffffu, c.x, 28 + gl); carry_v.y = __shfl_sync(0xffffffffu, c.y, 28 + gl);
        carry_v.z = __shfl_sync(0xffffffffu, c.z, 28 + gl); carry_v.w = __shfl_sync(0xffffffffu, c.w, 28 + gl);
        carry_g = __shfl_sync(0xffffffffu, gs, 28);
        carry_key = __shfl_sync(0xffffffffu, key, 28);
        // ---- finished rows: last occurrence of a run
        unsigned tm = __ballot_sync(0xffffffffu, tail && gl == 0);          // bit 4g set: group g ends a run
        if (tm) {
          if (!tails_seen && first_open) {
            // the range's first run started in an earlier range: its sum so far is a PIECE (the last one) -> entry 2r
            const int fg = (__ffs(tm) - 1) >> 2;
            if (g == fg) {
              FlatEntry* e = p.ent + 2 * r;
              *reinterpret_cast<float4*>(e->v + gl * 4) = c;
              if (gl == 0) { e->key = key; e->gs = gs; e->state = 1u; }
            }
            tm &= tm - 1;
          }
          tails_seen = true;
          const bool mine = (tm >> (4 * g)) & 1u;
          const int rank = __popc(tm & ((1u << (4 * g)) - 1u));
          const int cnt = __popc(tm);
          if (mine) {
            const int slot = (int)((q_total + (unsigned)rank) % kQSlots);
            *reinterpret_cast<float4*>(q_grad + slot * 16 + gl * 4) = c;
            if (gl == 0) {
              q_gs[slot] = gs;
              q_key[slot] = key;
              rec_bulk_g2s(rec_smem_u32(q_rec + slot * kQRecBytes), p.table + (size_t)key * 256, 256u,
                           rec_smem_u32(q_bar + (slot >> 3)));
            }
          }
          __syncwarp();
          q_total += (unsigned)cnt;
          // apply a batch once the batch AFTER it is complete too: its records were requested >= 8 rows ago
          if ((q_total >> 3) >= q_flushed + 2) flush((int)(q_flushed % 3u), 8);
        }
      }
      kc = kn;
      bc = bn;
      if (last_batch) prev_key = pk_next;
    }
    // ---- the range's last run continues into the next range: leave its partial sums behind
    if (i1 > i0) {
      const bool open_end = i1 < n && carry_key < rows && key_after == carry_key;
      if (open_end) {
        if (acc_dirty) {
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            acc_v.x += __shfl_xor_sync(0xffffffffu, acc_v.x, o); acc_v.y += __shfl_xor_sync(0xffffffffu, acc_v.y, o);
            acc_v.z += __shfl_xor_sync(0xffffffffu, acc_v.z, o); acc_v.w += __shfl_xor_sync(0xffffffffu, acc_v.w, o);
            acc_g += __shfl_xor_sync(0xffffffffu, acc_g, o);
          }
          carry_v.x += acc_v.x; carry_v.y += acc_v.y; carry_v.z += acc_v.z; carry_v.w += acc_v.w;
          carry_g += acc_g;
        }
        if (g == 7) {
          // entry 2r: a middle piece (the run came in AND goes out); entry 2r+1: the FIRST piece of a run that starts here
          FlatEntry* e = p.ent + 2 * r + ((!tails_seen && first_open) ? 0 : 1);
          *reinterpret_cast<float4*>(e->v + gl * 4) = carry_v;
          if (gl == 0) { e->key = carry_key; e->gs = carry_g; e->state = 2u; }
        }
      }
    }
  }
  // ---- drain the queue
  while (q_flushed * 8 + 8 <= q_total) flush((int)(q_flushed % 3u), 8);
  const int rest = (int)(q_total - q_flushed * 8);
  if (rest > 0) flush((int)(q_flushed % 3u), rest);
}

// Runs that cross range boundaries.  A run that starts in range r and leaves it open puts its FIRST piece in entry
// 2r+1 (state 2); every following range contributes entry 2r' (state 2 = the whole range lies inside the run, state 1 =
// the run ends there).  One warp per odd entry: sum the pieces (8 groups take every 8th one, fixed shuffle tree), Adam.
__global__ void __launch_bounds__(256) fm_fused_flat_fixup_kernel(const FlatParams p) {
  const int lane = threadIdx.x & 31, gl = lane & 3, g = lane >> 2;
  const long long r0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r0 >= p.n_ranges) return;
  const FlatEntry* E = p.ent;
  if (E[2 * r0 + 1].state != 2u) return;
  const unsigned key = E[2 * r0 + 1].key;
  // number of following pieces: ranges r0+1 .. r0+m, the last one has state 1 (128 entries examined per round trip)
  long long m = 0;
  for (;;) {
    unsigned st4[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const long long q = r0 + 1 + m + q4 * 32 + lane;
      st4[q4] = q < p.n_ranges ? E[2 * q].state : 1u;
    }
    bool done = false;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      if (done) continue;
      const unsigned stop = __ballot_sync(0xffffffffu, st4[q4] != 2u);
      if (stop) { m += q4 * 32 + __ffs(stop); done = true; }      // includes the final (state 1) piece
    }
    if (done) break;
    m += 128;
  }
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  float ts = 0.f;
  // pieces j = 0 (the head, entry 2 r0 + 1) .. m; group g takes j = g, g + 8, ...; four loads in flight per lane
  for (long long j0 = g; j0 <= m; j0 += 32) {
    float4 x4[4];
    float s4[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const long long j = j0 + 8 * q4;
      x4[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
      s4[q4] = 0.f;
      if (j <= m && r0 + j < p.n_ranges) {
        const FlatEntry* e = j == 0 ? &E[2 * r0 + 1] : &E[2 * (r0 + j)];
        if (j == 0 || (e->state != 0u && e->key == key)) {
          x4[q4] = *reinterpret_cast<const float4*>(e->v + gl * 4);
          s4[q4] = e->gs;
        }
      }
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      t.x += x4[q4].x; t.y += x4[q4].y; t.z += x4[q4].z; t.w += x4[q4].w;
      ts += s4[q4];
    }
  }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
    ts += __shfl_xor_sync(0xffffffffu, ts, o);
  }
  if (g == 0) {
    const float lr_t = p.d_lr_t ? *p.d_lr_t : p.lr_t;
    float* grow = reinterpret_cast<float*>(p.table + (size_t)key * 256);
    flat_adam_row(p, grow, grow, gl, t, ts, lr_t);
  }
}

}  // namespace etr

using namespace etr;

extern "C" int etr_fm_fused_flat_apply(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t fields, int64_t batch,
                                       const uint32_t* d_sorted_key, const int32_t* d_sorted_bag, int64_t n_slots,
                                       const float* d_dlogit, const float* d_sumv, const void* d_dflat, int32_t flat_dtype,
                                       int64_t flat_ld, int32_t flat_col0, float lr_t, const float* d_lr_t, float beta1,
                                       float beta2, float eps, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_sorted_key && d_sorted_bag && d_dlogit && d_sumv, "NULL argument");
  if (table->reserved != ETR_TABLE_RECORD || table->dtype != ETR_F32 || table->stride != 64 || k != 16 || table->width != 17 ||
      ((uintptr_t)table->d_data & 255)) {
    etr_set_error("etr_fm_fused_flat_apply: needs a RECORD table (fp32, k = 16, [V, 17] rows, 256-byte records)");
    return ETR_EUNSUPPORTED;
  }
  ETR_CHECK_ARG(fields > 0 && batch >= 0 && (long long)fields * batch == n_slots, "n_slots must equal batch*fields (single-hot)");
  ETR_CHECK_ARG(table->rows < 0xffffffffLL, "table rows must fit 32 bits");
  if (d_dflat) {
    const int osz = flat_dtype == ETR_BF16 ? 2 : 4;
    ETR_CHECK_ARG(flat_col0 % 4 == 0 && flat_ld % 4 == 0 && ((uintptr_t)d_dflat % (4 * osz)) == 0,
                  "dflat must be aligned for 4-element vector loads");
  }
  ETR_CHECK_ARG((((uintptr_t)d_sumv) & 15) == 0, "sumv must be 16-byte aligned");
  if (n_slots == 0) return ETR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  FlatParams p;
  memset(&p, 0, sizeof(p));
  p.table = (char*)table->d_data; p.rows = table->rows; p.F = fields;
  int lg = 0;
  while ((1 << lg) < fields) ++lg;
  p.shift = 32 + lg;
  p.magic = ((1ull << p.shift) + (unsigned long long)fields - 1) / (unsigned long long)fields;
  p.keys = d_sorted_key; p.bags = d_sorted_bag; p.n = n_slots;
  p.dlogit = d_dlogit; p.sumv = d_sumv; p.dflat = d_dflat; p.flat_ld = flat_ld; p.flat_col0 = flat_col0;
  p.lr_t = lr_t; p.d_lr_t = d_lr_t; p.b1 = beta1; p.b2 = beta2; p.eps = eps;
  // ranges of kFlatRange occurrences dealt round-robin to the warps (8 per CTA, up to 3 CTAs per SM): every warp gets a
  // mix of hot-id ranges (cheap: private adds) and unique-heavy ranges (a queued row + Adam per occurrence)
  const size_t smem = (size_t)8 * kQWarpBytes;
  const long long n_ranges = (n_slots + kFlatRange - 1) / kFlatRange;
  long long ncta = std::min<long long>((long long)ctx->sm_count * 3, (n_ranges + 7) / 8);
  if (ncta < 1) ncta = 1;
  p.n_ranges = n_ranges;
  const size_t ent_bytes = sizeof(FlatEntry) * 2 * (size_t)n_ranges;
  int st = etr_ws_reserve(ctx, ent_bytes);
  if (st != ETR_OK) return st;
  p.ent = (FlatEntry*)ctx->d_ws;
  ETR_CUDA(cudaMemsetAsync(p.ent, 0, ent_bytes, s));
  const int mode = d_dflat ? (flat_dtype == ETR_BF16 ? 1 : 2) : 0;
#define ETR_FLAT(M)                                                                                                  \
  do {                                                                                                               \
    ETR_CUDA(cudaFuncSetAttribute(fm_fused_flat_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    fm_fused_flat_kernel<M><<<(int)ncta, 256, smem, s>>>(p);                                                        \
  } while (0)
  if (mode == 1) ETR_FLAT(1); else if (mode == 2) ETR_FLAT(2); else ETR_FLAT(0);
#undef ETR_FLAT
  ETR_LAUNCH_CHECK(ctx);
  fm_fused_flat_fixup_kernel<<<(int)((n_ranges + 7) / 8), 256, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}
