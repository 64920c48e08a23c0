// K9: routing for row-sharded embedding tables (no reference type).
// owner(id) = id mod G, local row = id div G.  A rank's ids are split by owner
// with a STABLE partition (1-pass radix sort on the owner key), so routing is
// deterministic and the un-permute is an exact inverse:
//   send_rows[j]  local row (id div G) of the j-th id in owner-grouped order
//   send_pos[j]   original slot of that id            (int64: usable as gather ids)
//   inv_pos[slot] position j of the slot in the grouped order
//   counts[g]     how many ids go to owner g
#include <cub/device/device_radix_sort.cuh>

#include "etr_common.cuh"
#include "etr_async.cuh"

namespace etr {

__global__ void __launch_bounds__(256) shard_keys_kernel(const long long* ids, long long n, int world, long long rows_global,
                                                         unsigned* keys, int* slots, unsigned long long* err) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long id = ids[t];
    unsigned key;
    if ((unsigned long long)id >= (unsigned long long)rows_global) {
      flag_bad_id(err, id);
      key = 0;                      // routed to rank 0 as row 0; the error word reports it
    } else {
      key = (unsigned)(id % world);
    }
    keys[t] = key;
    slots[t] = (int)t;
  }
}

__global__ void __launch_bounds__(256) shard_finish_kernel(const long long* ids, const unsigned* sorted_keys,
                                                           const int* sorted_slots, long long n, int world,
                                                           long long rows_global, long long* send_rows,
                                                           long long* send_pos, long long* inv_pos, int* counts) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const int slot = sorted_slots[j];
    const long long id = ids[slot];
    const bool bad = (unsigned long long)id >= (unsigned long long)rows_global;
    send_rows[j] = bad ? 0 : id / world;
    send_pos[j] = slot;
    inv_pos[slot] = j;
    // run boundaries of the sorted owner keys give the counts
    const unsigned k = sorted_keys[j];
    if (j == n - 1 || sorted_keys[j + 1] != k) {
      // last element of owner k's run: count = (j+1) - start(k); start found by the previous owner's end
      long long start = 0;
      if (k > 0) {
        // binary search the first position with key >= k
        long long lo = 0, hi = j;
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if (sorted_keys[mid] < k) lo = mid + 1; else hi = mid;
        }
        start = lo;
      }
      counts[k] = (int)(j + 1 - start);
    }
  }
}

// ---- peer-memory path: push deduplicated gradient rows into the owners' mailboxes
struct PushParams {
  const long long* unique_ids; const int* n_unique; const float* unique_grad; int ld;
  int world; int cap;                    // slots per (owner, source) region
  long long* ids_mb[16];                 // owner g: this source rank's id region (peer pointer)
  float* grads_mb[16];                   // owner g: this source rank's gradient region
  int* counts_mb[16];                    // owner g: &counts[source rank]
  int* local_cnt;                        // [world] slots used so far per owner (this step)
  unsigned long long* err;
};

// Slots are reserved per CTA: lane-group leaders count into shared-memory bins (one per owner), one
// global atomicAdd per (CTA, owner) reserves the range.  (The first version did one global atomic
// per row on G counters: 471 k same-address atomics = 232 us at N = 2.)  Slot order inside a source
// region is arbitrary; nothing downstream depends on it (rows are unique within a source).
template <int LPR>
__global__ void __launch_bounds__(256) shard_push_kernel(const PushParams p) {
  constexpr int GPW = 32 / LPR;
  constexpr int GPB = 8 * GPW;                          // lane groups (rows) per CTA iteration
  __shared__ int bin[16], base[16];
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const int grp = (threadIdx.x >> 5) * GPW + g;
  const int nchunks = p.ld / 4;
  const int n_unique = *p.n_unique;
  for (long long u0 = (long long)blockIdx.x * GPB; u0 < n_unique; u0 += (long long)gridDim.x * GPB) {   // CTA-uniform
    if (threadIdx.x < 16) bin[threadIdx.x] = 0;
    __syncthreads();
    const long long u = u0 + grp;
    const bool active = u < n_unique;
    int owner = 0, off = 0;
    long long lrow = 0;
    if (active) {
      const long long id = p.unique_ids[u];
      owner = (int)(id % p.world);
      lrow = id / p.world;
      if (gl == 0) off = atomicAdd(&bin[owner], 1);
    }
    __syncthreads();
    if (threadIdx.x < p.world) base[threadIdx.x] = bin[threadIdx.x] ? atomicAdd(&p.local_cnt[threadIdx.x], bin[threadIdx.x]) : 0;
    __syncthreads();
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
    off = __shfl_sync(gmask, off, g * LPR);
    if (!active) continue;
    const int slot = base[owner] + off;
    if (slot >= p.cap) {
      if (gl == 0) flag_error(p.err, kErrMailbox, 0);            // mailbox region overflow
      continue;
    }
    if (gl == 0) p.ids_mb[owner][slot] = lrow;
    for (int c = gl; c < nchunks; c += LPR)
      *reinterpret_cast<float4*>(p.grads_mb[owner] + (long long)slot * p.ld + c * 4) =
          *reinterpret_cast<const float4*>(p.unique_grad + u * p.ld + c * 4);
  }
}
__global__ void shard_push_counts_kernel(const PushParams p) {
  const int g = threadIdx.x;
  if (g < p.world) {
    int c = p.local_cnt[g];
    if (c > p.cap) c = p.cap;
    *p.counts_mb[g] = c;
    p.local_cnt[g] = 0;                                // ready for the next step
  }
}
// ---- de-duplicated row exchange (peer form, forward): request -> serve -> virtual ids
// request: every unique id of this rank's batch (sorted plan) is routed to its owner's request mailbox
// (local row numbers; slots reserved per CTA) and remembers its (owner, slot); serve: the owner copies the
// requested rows out of its shard into the requester's response buffer at the same (owner, slot) -- slots
// of a region are consecutive, so the NVLink traffic is sequential full-line STORES instead of the
// request-bound 80-byte remote loads of the fused pull (profiles/r01_mgpu.md); vid map: occurrence ->
// response-buffer row, so the ordinary gather kernel runs on the response buffer as its table.
struct RequestParams {
  const long long* unique_ids; const int* n_unique;
  int world, cap;
  long long* req_mb[16];                 // owner g: this source's id region (peer pointer)
  int* counts_mb[16];                    // owner g: &counts[source rank]
  int* local_cnt;
  int* slot_of_u;                        // [max_unique]: owner * cap + slot
  unsigned long long* err;
};
__global__ void __launch_bounds__(256) shard_request_kernel(const RequestParams p) {
  __shared__ int bin[16], base[16];
  const int n_unique = *p.n_unique;
  for (long long u0 = (long long)blockIdx.x * 256; u0 < n_unique; u0 += (long long)gridDim.x * 256) {      // CTA-uniform
    if (threadIdx.x < 16) bin[threadIdx.x] = 0;
    __syncthreads();
    const long long u = u0 + threadIdx.x;
    const bool active = u < n_unique;
    int owner = 0, off = 0;
    long long lrow = 0;
    if (active) {
      const long long id = p.unique_ids[u];
      owner = (int)(id % p.world);
      lrow = id / p.world;
      off = atomicAdd(&bin[owner], 1);
    }
    __syncthreads();
    if (threadIdx.x < p.world) base[threadIdx.x] = bin[threadIdx.x] ? atomicAdd(&p.local_cnt[threadIdx.x], bin[threadIdx.x]) : 0;
    __syncthreads();
    if (active) {
      const int slot = base[owner] + off;
      if (slot >= p.cap) {
        flag_error(p.err, kErrMailbox, 0);                       // mailbox region overflow
        p.slot_of_u[u] = owner * p.cap;               // keep later kernels in range
      } else {
        p.req_mb[owner][slot] = lrow;
        p.slot_of_u[u] = owner * p.cap + slot;
      }
    }
    __syncthreads();
  }
}
__global__ void shard_request_counts_kernel(const RequestParams p) {
  const int g = threadIdx.x;
  if (g < p.world) {
    int c = p.local_cnt[g];
    if (c > p.cap) c = p.cap;
    *p.counts_mb[g] = c;
    p.local_cnt[g] = 0;
  }
}

struct ServeParams {
  const float* table; int stride; long long rows;
  const long long* req; const int* counts;   // own request mailbox [world][cap], counts[world]
  int world, cap, ld, rank;
  float* resp[16];                           // source g's response region for THIS owner: [cap][ld] (peer pointer)
  unsigned long long* err;
};
template <int LPR>
__global__ void __launch_bounds__(256) shard_serve_kernel(const ServeParams p) {
  constexpr int GPW = 32 / LPR;
  const int src = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const int nchunks = p.ld / 4;
  int n = p.counts[src];
  if (n > p.cap) n = p.cap;
  const long long* req = p.req + (long long)src * p.cap;
  float* out = p.resp[src];
  for (long long s = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * GPW + g; s < n; s += (long long)gridDim.x * 8 * GPW) {
    long long row = req[s];
    if ((unsigned long long)row >= (unsigned long long)p.rows) { flag_bad_id(p.err, row); row = 0; }
    for (int c = gl; c < nchunks; c += LPR)
      *reinterpret_cast<float4*>(out + s * p.ld + c * 4) = ldg_row16(p.table + row * p.stride + c * 4);
  }
}
// The same copy with the response region addressed as a flat array of 16-byte chunks: thread t of a CTA iteration moves
// chunk t of a 256-slot span (slot = chunk / chunks-per-row), so every warp store is 512 contiguous, line-aligned bytes
// (cap is a multiple of 64 slots and the span starts on a multiple of 256 slots: 256 * ld * 4 bytes is a whole number of
// lines).  The row-per-lane-group form above ends most warp stores inside a 32-byte sector (80-byte rows): partial
// writes that NVLink carries as separate, masked packets.
__global__ void __launch_bounds__(256) shard_serve_flat_kernel(const ServeParams p) {
  const int src = blockIdx.y;
  const int nchunks = p.ld / 4;
  int n = p.counts[src];
  if (n > p.cap) n = p.cap;
  const long long* req = p.req + (long long)src * p.cap;
  float4* out = reinterpret_cast<float4*>(p.resp[src]);
  const long long total = (long long)n * nchunks;
  for (long long q0 = (long long)blockIdx.x * (256 * nchunks); q0 < total; q0 += (long long)gridDim.x * (256 * nchunks)) {
#pragma unroll 5
    for (int it = 0; it < nchunks; ++it) {
      const long long q = q0 + it * 256 + threadIdx.x;
      if (q < total) {
        const long long s = q / nchunks;
        const int c = (int)(q - s * nchunks);
        long long row = req[s];
        if ((unsigned long long)row >= (unsigned long long)p.rows) { flag_bad_id(p.err, row); row = 0; }
        out[q] = ldg_row16(p.table + row * p.stride + c * 4);
      }
    }
  }
}

// vid[occurrence] = response-buffer row of the occurrence's id.  A thread per sorted position finds its
// run by binary search over seg_start (n_unique + 1 entries, L2-resident).
__global__ void __launch_bounds__(256) shard_vid_map_kernel(const int* sorted_bag, const int* seg_start, const int* n_unique,
                                                            const int* slot_of_u, long long* vid) {
  constexpr int PER = 8;                   // consecutive sorted positions per thread: one search, then a walk
  const int nu = *n_unique;
  const int n_valid = seg_start[nu];
  for (long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PER; p0 < n_valid;
       p0 += (long long)gridDim.x * blockDim.x * PER) {
    int lo = 0, hi = nu;                   // largest u with seg_start[u] <= p0
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (seg_start[mid] <= p0) lo = mid; else hi = mid;
    }
    int next = seg_start[lo + 1];
    int so = slot_of_u[lo];
    for (int q = 0; q < PER; ++q) {
      const int pos = (int)p0 + q;
      if (pos >= n_valid) break;
      while (pos >= next) { ++lo; next = seg_start[lo + 1]; so = slot_of_u[lo]; }
      vid[sorted_bag[pos]] = so;
    }
  }
}

// ---- owner side without a sort: dense gradient accumulator + touched-row list
// Source g's region holds rows that are unique within the region, so adding it into the accumulator
// gacc[local_rows, ld] needs no atomics; the regions are added by G consecutive launches in rank
// order (deterministic sum).  The LAST column of an accumulator row is its stamp: the first launch
// of a step that meets the row (stamp != epoch) OVERWRITES it, stamps it and puts it on the touched
// list (one global atomic per CTA iteration); later launches add.  So the accumulator is never
// cleared and costs one read + one write request per (row, source).
struct AccParams {
  const long long* ids; const float* grads; const int* counts;   // one source region
  int cap, ld;
  float* gacc; const unsigned* epoch;
  int* touched; int* n_touched; int max_touched;
  unsigned long long* err;
};
template <int LPR>
__global__ void __launch_bounds__(256) mailbox_accumulate_kernel(const AccParams p) {
  constexpr int GPW = 32 / LPR;
  constexpr int GPB = 8 * GPW;
  __shared__ int wcount[8], wbase[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPR, g = lane / LPR;
  const int nchunks = p.ld / 4;                       // LPR >= nchunks: one chunk per lane
  const int stamp_lane = nchunks - 1;
  int n = *p.counts;
  if (n > p.cap) n = p.cap;
  const unsigned epoch = *p.epoch;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
  for (int s0 = blockIdx.x * GPB; s0 < n; s0 += gridDim.x * GPB) {            // CTA-uniform trip count
    const int sl = s0 + warp * GPW + g;
    const bool active = sl < n;
    long long row = 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), x = a;
    if (active) row = p.ids[sl];
    const bool mine = active && gl < nchunks;
    if (mine) {
      a = *reinterpret_cast<const float4*>(p.gacc + row * p.ld + gl * 4);
      x = *reinterpret_cast<const float4*>(p.grads + (long long)sl * p.ld + gl * 4);
    }
    const unsigned st = __shfl_sync(gmask, __float_as_uint(a.w), g * LPR + stamp_lane);
    const bool fresh_row = active && st != epoch;
    if (mine) {
      if (fresh_row) a = x;
      else { a.x += x.x; a.y += x.y; a.z += x.z; if (gl != stamp_lane) a.w += x.w; }
      if (gl == stamp_lane) a.w = __uint_as_float(epoch);
      *reinterpret_cast<float4*>(p.gacc + row * p.ld + gl * 4) = a;
    }
    const bool fresh = fresh_row && gl == 0;
    const unsigned ball = __ballot_sync(0xffffffffu, fresh);
    if (lane == 0) wcount[warp] = __popc(ball);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) { wbase[w] = tot; tot += wcount[w]; }
      const int b0 = tot ? atomicAdd(p.n_touched, tot) : 0;
      for (int w = 0; w < 8; ++w) wbase[w] += b0;
    }
    __syncthreads();
    if (fresh) {
      const int t = wbase[warp] + __popc(ball & ((1u << lane) - 1u));
      if (t < p.max_touched) p.touched[t] = (int)row;
      else flag_error(p.err, kErrTouched, 0);
    }
    __syncthreads();
  }
}

// Adam on the touched rows; fm_k > 0: the pushed rows carry the DEFERRED FM gradient
// [P_0..P_{k-1}, sum_g] with P = sum g S + sum dflat, and the owner finishes
// dv = P - v * sum_g with its own copy of the row (so the exporting rank never reads it remotely).
struct TouchedAdamParams {
  float* table; float* m; float* v; int stride;
  float* gacc; int ld; const int* touched; const int* n_touched; int max_touched;
  int fm_k; const float* d_lr_t; float b1, b2, eps;
};
__global__ void __launch_bounds__(256) touched_adam_kernel(const TouchedAdamParams p) {
  const int nch = p.ld / 4;
  const float lr_t = *p.d_lr_t;
  int n = *p.n_touched;
  if (n > p.max_touched) n = p.max_touched;
  const long long total = (long long)n * nch;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / nch;
    const int c = (int)(t % nch);
    const long long row = p.touched[i];
    float4 g = *reinterpret_cast<const float4*>(p.gacc + row * p.ld + c * 4);
    if (c == nch - 1) g.w = 0.f;                             // the stamp column is not a gradient
    float4* pvar = reinterpret_cast<float4*>(p.table + row * p.stride + c * 4);
    float4* pm = reinterpret_cast<float4*>(p.m + row * p.stride + c * 4);
    float4* pv = reinterpret_cast<float4*>(p.v + row * p.stride + c * 4);
    float4 var = *pvar, m = *pm, v = *pv;
    if (p.fm_k > 0 && c * 4 < p.fm_k) {
      const float sg = p.gacc[row * p.ld + p.fm_k];
      g.x -= var.x * sg; g.y -= var.y * sg; g.z -= var.z * sg; g.w -= var.w * sg;
    }
    float* xv = &var.x; float* xm = &m.x; float* xvv = &v.x; const float* xg = &g.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      adam_update1_fast(xv[q], xm[q], xvv[q], xg[q], lr_t, p.b1, p.b2, p.eps);   // the fused single-GPU apply's expression
    }
    *pvar = var; *pm = m; *pv = v;
  }
}
// ---- owner side, round 2: who-contributes-to-which-row is known from the REQUESTS (the gradient rows return
// through the request slots), so it is worked out off the critical path while the forward runs:
//   owner_prep  : every request entry e = src * cap + s claims its row in a dense map (64-bit CAS of
//                 (step << 32) | e); the winner is the row's LEADER (bit 31 of mask[e]); the others record their
//                 slot at the leader (others[leader][src]) and OR their source bit into mask[leader];
//   owner_apply : after the gradient barrier, ONE pass over the entries: a leader adds its own mailbox row and
//                 the recorded ones in ascending source order (the same order as the region-by-region
//                 accumulation: bit-identical), finishes the deferred FM gradient and runs Adam on the row --
//                 no dense accumulator, no touched list, one read + one write of the row's state.  The pass
//                 clears mask[e] behind itself, so prep needs no memset.
struct OwnerPrepParams {
  const long long* req; const int* counts;       // [world][cap], [world]
  int world, cap; long long rows;
  unsigned long long* map;                       // [local rows]
  unsigned* step;                                // bumped by owner_step_kernel before the prep launch
  unsigned* mask; int* others;                   // [world * cap], [world * cap][world]
  unsigned long long* err;
};
__global__ void owner_step_kernel(unsigned* step) { *step += 1u; }
__global__ void __launch_bounds__(256) owner_prep_kernel(const OwnerPrepParams p) {
  const int src = blockIdx.y;
  int n = p.counts[src];
  if (n > p.cap) n = p.cap;
  const unsigned long long ep = (unsigned long long)(*p.step) << 32;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < n; s += gridDim.x * 256) {
    const unsigned e = (unsigned)src * (unsigned)p.cap + (unsigned)s;
    const long long row = p.req[e];
    if ((unsigned long long)row >= (unsigned long long)p.rows) continue;       // the serve kernel has flagged it
    unsigned long long* slot = p.map + row;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(slot);
    unsigned leader = e;
    for (;;) {
      if ((cur >> 32 << 32) == ep) { leader = (unsigned)cur; break; }
      const unsigned long long old = atomicCAS(slot, cur, ep | e);
      if (old == cur) break;
      cur = old;
    }
    if (leader == e) {
      atomicOr(p.mask + e, 0x80000000u | (1u << src));
    } else {
      p.others[(long long)leader * p.world + src] = s;
      atomicOr(p.mask + leader, 1u << src);
    }
  }
}

struct OwnerApplyParams {
  const long long* req; const int* counts; const float* grads;
  int world, cap, ld;
  unsigned* mask; const int* others;
  float* table; float* m; float* v; int stride; long long rows;
  int fm_k; const float* d_lr_t; float b1, b2, eps;
};
template <int LPR, int U>                                   // U: entries in flight per lane group
__global__ void __launch_bounds__(256) owner_apply_kernel(const OwnerApplyParams p) {
  constexpr int GPW = 32 / LPR;
  const int src = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR, g = lane / LPR;
  const int nch = p.ld / 4;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
  int n = p.counts[src];
  if (n > p.cap) n = p.cap;
  const float lr_t = *p.d_lr_t;
  const bool mine = gl < nch;
  const int sg_lane = g * LPR + p.fm_k / 4;
  const int step = gridDim.x * 8 * GPW;
  // every condition below is uniform per lane group (the shuffles are group-wide)
  // (LPR = 5, the 20-float FM row: six entries per warp, lanes 30 and 31 idle -- the pass is issue-bound, and the
  // power-of-two grouping spent 3 of 8 lanes on nothing)
  for (int s0 = g < GPW ? (blockIdx.x * 8 + (threadIdx.x >> 5)) * GPW + g : n; s0 < n; s0 += U * step) {
    unsigned mk[U];
    long long e[U], row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u * step;
      e[u] = (long long)src * p.cap + s;
      mk[u] = 0u; row[u] = 0;
      if (s < n) { mk[u] = p.mask[e[u]]; row[u] = p.req[e[u]]; }
    }
    __syncwarp(gmask);
    float4 var[U], mm[U], vv[U], acc[U];
    bool lead[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      lead[u] = (mk[u] & 0x80000000u) != 0u;                // else: a leader elsewhere sums this row (or never claimed)
      if (mk[u] != 0u && gl == 0) p.mask[e[u]] = 0u;        // self-cleaning
      var[u] = mm[u] = vv[u] = acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lead[u] && mine) {
        const long long off = row[u] * p.stride + gl * 4;
        var[u] = *reinterpret_cast<const float4*>(p.table + off);
        mm[u] = *reinterpret_cast<const float4*>(p.m + off);
        vv[u] = *reinterpret_cast<const float4*>(p.v + off);
        if ((mk[u] & 0xffffu) == (1u << src))               // the common case: one contributor
          acc[u] = __ldcs(reinterpret_cast<const float4*>(p.grads + e[u] * p.ld + gl * 4));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned srcs = mk[u] & 0xffffu;
      if (!lead[u] || srcs == (1u << src)) continue;
      bool first = true;
      while (srcs) {                                        // ascending source order = the region-by-region order
        const int q = __ffs(srcs) - 1;
        srcs &= srcs - 1;
        const long long eq = (q == src) ? e[u] : (long long)q * p.cap + p.others[e[u] * p.world + q];
        if (mine) {
          const float4 x = __ldcs(reinterpret_cast<const float4*>(p.grads + eq * p.ld + gl * 4));
          if (first) acc[u] = x;
          else { acc[u].x += x.x; acc[u].y += x.y; acc[u].z += x.z; acc[u].w += x.w; }
        }
        first = false;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!lead[u]) continue;
      if (p.fm_k > 0) {
        const float sg = __shfl_sync(gmask, acc[u].x, sg_lane);
        if (gl * 4 < p.fm_k) { acc[u].x -= var[u].x * sg; acc[u].y -= var[u].y * sg; acc[u].z -= var[u].z * sg; acc[u].w -= var[u].w * sg; }
      }
      if (mine) {
        float* xv = &var[u].x; float* xm = &mm[u].x; float* xvv = &vv[u].x; const float* xg = &acc[u].x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          adam_update1_fast(xv[q], xm[q], xvv[q], xg[q], lr_t, p.b1, p.b2, p.eps);
        }
        const long long off = row[u] * p.stride + gl * 4;
        *reinterpret_cast<float4*>(p.table + off) = var[u];
        *reinterpret_cast<float4*>(p.m + off) = mm[u];
        *reinterpret_cast<float4*>(p.v + off) = vv[u];
      }
    }
  }
}

// ---- cross-rank barrier over peer memory: rank r stores the new epoch into flags[r] of every
// peer (release, system scope) and waits until every peer's epoch has arrived in its own flags.
// One warp; bounded spin (a missing peer flags error -3 instead of hanging the GPU).
struct BarrierParams {
  unsigned* peer_flags[16];
  unsigned* my_flags;
  unsigned* epoch;
  int world, rank;
  unsigned long long* err;
};
__global__ void peer_barrier_kernel(const BarrierParams p) {
  const int g = threadIdx.x;
  const unsigned e = *reinterpret_cast<volatile unsigned*>(p.epoch) + 1u;
  __syncwarp();
  if (g == 0) *p.epoch = e;
  if (g < p.world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_flags[g] + p.rank), "r"(e) : "memory");
    const long long t0 = clock64();
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.my_flags + g) : "memory");
      if ((int)(v - e) >= 0) break;
      if (clock64() - t0 > 40000000000ll) {          // ~20 s: a peer never arrived
        flag_error(p.err, kErrBarrier, 0);
        break;
      }
      __nanosleep(64);
    }
  }
}

// ---- small all-reduce over peer memory (replicated dense gradients): every rank stores its vector
// into slot [rank] of every peer, barrier, then sums the G slots in rank order (deterministic, so
// the replicas stay bit-identical).
struct ArPushParams {
  float* slots[16];          // peer g's slot buffer [world][n]
  const float* src;
  long long n;
  int world, rank;
};
__global__ void __launch_bounds__(256) peer_allreduce_push_kernel(const ArPushParams p) {
  const long long total = p.n * p.world;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t / p.n);
    const long long i = t - (long long)g * p.n;
    p.slots[g][(long long)p.rank * p.n + i] = p.src[i];
  }
}
__global__ void __launch_bounds__(256) peer_allreduce_sum_kernel(const float* slots, long long n, int world, float* dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = slots[i];
    for (int g = 1; g < world; ++g) acc += slots[(long long)g * n + i];
    dst[i] = acc;
  }
}

}  // namespace etr

using namespace etr;

extern "C" {

int etr_peer_alloc(etr_ctx* ctx, int64_t bytes, void** d_ptr, void* handle64) {
  ETR_CHECK_ARG(ctx && d_ptr && handle64 && bytes > 0, "bad argument");
  ETR_CUDA(cudaMalloc(d_ptr, (size_t)bytes));
  ETR_CUDA(cudaMemset(*d_ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  ETR_CUDA(cudaIpcGetMemHandle(&h, *d_ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return ETR_OK;
}

int etr_peer_open(etr_ctx* ctx, const void* handle64, void** d_ptr) {
  ETR_CHECK_ARG(ctx && handle64 && d_ptr, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  ETR_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ETR_OK;
}

int etr_peer_close(etr_ctx* ctx, void* d_ptr) {
  ETR_CHECK_ARG(ctx, "bad argument");
  if (d_ptr) ETR_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return ETR_OK;
}

int etr_peer_free(etr_ctx* ctx, void* d_ptr) {
  ETR_CHECK_ARG(ctx, "bad argument");
  if (d_ptr) ETR_CUDA(cudaFree(d_ptr));
  return ETR_OK;
}

int etr_shard_set_create(etr_ctx* ctx, const void* const* h_shard_ptrs, int32_t world, int32_t rank,
                         int64_t rows_global, int32_t* out_id) {
  ETR_CHECK_ARG(ctx && h_shard_ptrs && out_id, "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && rank >= 0 && rank < world, "world must be in [1,16]");
  ETR_CHECK_ARG(ctx->n_shard_sets < kMaxShardSets, "too many shard sets on this ctx");
  EtrShardSet& ss = ctx->shard_sets[ctx->n_shard_sets];
  for (int g = 0; g < world; ++g) {
    ETR_CHECK_ARG(h_shard_ptrs[g] != nullptr, "NULL shard pointer");
    ss.base[g] = (const char*)h_shard_ptrs[g];
  }
  ss.world = world; ss.rank = rank; ss.rows_global = rows_global;
  *out_id = ++ctx->n_shard_sets;                        // ids start at 1; 0 in etr_table.reserved = not sharded
  return ETR_OK;
}

int etr_shard_push(etr_ctx* ctx, const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t max_unique,
                   const float* d_unique_grad, int32_t ld, int32_t world, int32_t cap,
                   int64_t* const* h_ids_mb, float* const* h_grads_mb, int32_t* const* h_counts_mb,
                   int32_t* d_local_cnt, void* stream) {
  ETR_CHECK_ARG(ctx && d_unique_ids && d_n_unique && d_unique_grad && h_ids_mb && h_grads_mb && h_counts_mb && d_local_cnt,
                "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && ld % 4 == 0 && ld / 4 <= 32 && cap > 0, "bad world / ld / cap");
  PushParams p;
  memset(&p, 0, sizeof(p));
  p.unique_ids = (const long long*)d_unique_ids; p.n_unique = d_n_unique; p.unique_grad = d_unique_grad; p.ld = ld;
  p.world = world; p.cap = cap; p.local_cnt = d_local_cnt; p.err = ctx->d_err;
  for (int g = 0; g < world; ++g) {
    p.ids_mb[g] = (long long*)h_ids_mb[g]; p.grads_mb[g] = h_grads_mb[g]; p.counts_mb[g] = h_counts_mb[g];
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (max_unique > 0) {
    int lpr = 1;
    while (lpr < ld / 4) lpr <<= 1;
    const int grid = grid_for(max_unique, 8 * (32 / lpr), ctx->sm_count, 6);
    switch (lpr) {
      case 1: shard_push_kernel<1><<<grid, 256, 0, s>>>(p); break;
      case 2: shard_push_kernel<2><<<grid, 256, 0, s>>>(p); break;
      case 4: shard_push_kernel<4><<<grid, 256, 0, s>>>(p); break;
      case 8: shard_push_kernel<8><<<grid, 256, 0, s>>>(p); break;
      case 16: shard_push_kernel<16><<<grid, 256, 0, s>>>(p); break;
      default: shard_push_kernel<32><<<grid, 256, 0, s>>>(p); break;
    }
    ETR_LAUNCH_CHECK(ctx);
  }
  shard_push_counts_kernel<<<1, 32, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_request(etr_ctx* ctx, const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t max_unique,
                      int32_t world, int32_t cap, int64_t* const* h_req_mb, int32_t* const* h_counts_mb,
                      int32_t* d_local_cnt, int32_t* d_slot_of_u, void* stream) {
  ETR_CHECK_ARG(ctx && d_unique_ids && d_n_unique && h_req_mb && h_counts_mb && d_local_cnt && d_slot_of_u, "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && cap > 0 && (long long)world * cap < 0x7fffffffLL, "bad world / cap");
  RequestParams p;
  memset(&p, 0, sizeof(p));
  p.unique_ids = (const long long*)d_unique_ids; p.n_unique = d_n_unique; p.world = world; p.cap = cap;
  p.local_cnt = d_local_cnt; p.slot_of_u = d_slot_of_u; p.err = ctx->d_err;
  for (int g = 0; g < world; ++g) { p.req_mb[g] = (long long*)h_req_mb[g]; p.counts_mb[g] = h_counts_mb[g]; }
  cudaStream_t s = (cudaStream_t)stream;
  if (max_unique > 0) {
    shard_request_kernel<<<grid_for(max_unique, 256, ctx->sm_count, 4), 256, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
  }
  shard_request_counts_kernel<<<1, 32, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_serve(etr_ctx* ctx, const etr_table* table, const int64_t* d_req, const int32_t* d_counts, int32_t world,
                    int32_t cap, float* const* h_resp, int32_t ld, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_req && d_counts && h_resp, "NULL argument");
  ETR_CHECK_ARG(table->dtype == ETR_F32 && world >= 1 && world <= 16 && cap > 0 && ld % 4 == 0 && ld <= table->stride &&
                    ld / 4 <= 32, "fp32 table, ld <= stride");
  ServeParams p;
  memset(&p, 0, sizeof(p));
  p.table = (const float*)table->d_data; p.stride = table->stride; p.rows = table->rows; p.req = (const long long*)d_req;
  p.counts = d_counts; p.world = world; p.cap = cap; p.ld = ld; p.err = ctx->d_err;
  for (int g = 0; g < world; ++g) { ETR_CHECK_ARG(h_resp[g] != nullptr, "NULL response pointer"); p.resp[g] = h_resp[g]; }
  int lpr = 1;
  while (lpr < ld / 4) lpr <<= 1;
  dim3 grid((unsigned)grid_for(cap, 8 * (32 / lpr), ctx->sm_count, 8 / (world < 8 ? world : 8) + 1), (unsigned)world);
  cudaStream_t s = (cudaStream_t)stream;
  static int flat = -1;
  if (flat < 0) { const char* e = getenv("ETR_SERVE_FLAT"); flat = !(e && atoi(e) == 0); }
  if (flat) {
    dim3 fgrid((unsigned)grid_for(cap, 256, ctx->sm_count, 8 / (world < 8 ? world : 8) + 1), (unsigned)world);
    shard_serve_flat_kernel<<<fgrid, 256, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
    return ETR_OK;
  }
  switch (lpr) {
    case 1: shard_serve_kernel<1><<<grid, 256, 0, s>>>(p); break;
    case 2: shard_serve_kernel<2><<<grid, 256, 0, s>>>(p); break;
    case 4: shard_serve_kernel<4><<<grid, 256, 0, s>>>(p); break;
    case 8: shard_serve_kernel<8><<<grid, 256, 0, s>>>(p); break;
    case 16: shard_serve_kernel<16><<<grid, 256, 0, s>>>(p); break;
    default: shard_serve_kernel<32><<<grid, 256, 0, s>>>(p); break;
  }
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_vid_map(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start, const int32_t* d_n_unique,
                      int64_t n_slots, const int32_t* d_slot_of_u, int64_t* d_vid, void* stream) {
  ETR_CHECK_ARG(ctx && d_sorted_bag && d_seg_start && d_n_unique && d_slot_of_u && d_vid, "NULL argument");
  if (n_slots <= 0) return ETR_OK;
  shard_vid_map_kernel<<<grid_for(n_slots, 256 * 8, ctx->sm_count, 8), 256, 0, (cudaStream_t)stream>>>(
      d_sorted_bag, d_seg_start, d_n_unique, d_slot_of_u, (long long*)d_vid);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_mailbox_accumulate(etr_ctx* ctx, const int64_t* d_ids, const float* d_grads, const int32_t* d_counts,
                                 int32_t world, int32_t cap, int32_t ld, float* d_gacc,
                                 const uint32_t* d_epoch, int32_t* d_touched, int32_t* d_n_touched, int32_t max_touched,
                                 void* stream) {
  ETR_CHECK_ARG(ctx && d_ids && d_grads && d_counts && d_gacc && d_epoch && d_touched && d_n_touched,
                "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && cap > 0 && ld % 4 == 0 && ld / 4 <= 32 && max_touched > 0, "bad world / cap / ld");
  cudaStream_t s = (cudaStream_t)stream;
  ETR_CUDA(cudaMemsetAsync(d_n_touched, 0, sizeof(int), s));
  int lpr = 1;
  while (lpr < ld / 4) lpr <<= 1;
  const int grid = grid_for(cap, 8 * (32 / lpr), ctx->sm_count, 6);
  for (int g = 0; g < world; ++g) {                    // rank order: the sum is deterministic
    AccParams p;
    p.ids = (const long long*)d_ids + (long long)g * cap; p.grads = d_grads + (long long)g * cap * ld; p.counts = d_counts + g;
    p.cap = cap; p.ld = ld; p.gacc = d_gacc; p.epoch = d_epoch;
    p.touched = d_touched; p.n_touched = d_n_touched; p.max_touched = max_touched; p.err = ctx->d_err;
    switch (lpr) {
      case 1: mailbox_accumulate_kernel<1><<<grid, 256, 0, s>>>(p); break;
      case 2: mailbox_accumulate_kernel<2><<<grid, 256, 0, s>>>(p); break;
      case 4: mailbox_accumulate_kernel<4><<<grid, 256, 0, s>>>(p); break;
      case 8: mailbox_accumulate_kernel<8><<<grid, 256, 0, s>>>(p); break;
      case 16: mailbox_accumulate_kernel<16><<<grid, 256, 0, s>>>(p); break;
      default: mailbox_accumulate_kernel<32><<<grid, 256, 0, s>>>(p); break;
    }
    ETR_LAUNCH_CHECK(ctx);
  }
  return ETR_OK;
}

int etr_shard_touched_adam(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, float* d_gacc, int32_t ld,
                           const int32_t* d_touched, const int32_t* d_n_touched, int32_t max_touched, int32_t fm_k,
                           const float* d_lr_t, float beta1, float beta2, float eps, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_m && d_v && d_gacc && d_touched && d_n_touched && d_lr_t, "NULL argument");
  ETR_CHECK_ARG(table->dtype == ETR_F32 && ld % 4 == 0 && ld <= table->stride && max_touched > 0, "fp32 table, ld <= stride");
  ETR_CHECK_ARG(fm_k == 0 || (fm_k % 4 == 0 && fm_k + 4 <= ld), "fm_k must be a multiple of 4 with a chunk behind it");
  TouchedAdamParams p;
  p.table = (float*)table->d_data; p.m = d_m; p.v = d_v; p.stride = table->stride; p.gacc = d_gacc; p.ld = ld;
  p.touched = d_touched; p.n_touched = d_n_touched; p.max_touched = max_touched; p.fm_k = fm_k; p.d_lr_t = d_lr_t;
  p.b1 = beta1; p.b2 = beta2; p.eps = eps;
  const int grid = grid_for((long long)max_touched * (ld / 4), 256, ctx->sm_count, 8);
  touched_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_owner_prep(etr_ctx* ctx, const int64_t* d_req, const int32_t* d_counts, int32_t world, int32_t cap,
                         int64_t local_rows, uint64_t* d_map, uint32_t* d_step, uint32_t* d_mask, int32_t* d_others,
                         void* stream) {
  ETR_CHECK_ARG(ctx && d_req && d_counts && d_map && d_step && d_mask && d_others, "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && cap > 0 && local_rows > 0 && (long long)world * cap < 0x7fffffffLL,
                "bad world / cap / rows");
  OwnerPrepParams p;
  p.req = (const long long*)d_req; p.counts = d_counts; p.world = world; p.cap = cap; p.rows = local_rows;
  p.map = (unsigned long long*)d_map; p.step = d_step; p.mask = d_mask; p.others = d_others; p.err = ctx->d_err;
  cudaStream_t s = (cudaStream_t)stream;
  owner_step_kernel<<<1, 1, 0, s>>>(d_step);
  ETR_LAUNCH_CHECK(ctx);
  dim3 grid((unsigned)grid_for(cap, 256, ctx->sm_count, 8 / (world < 8 ? world : 8) + 1), (unsigned)world);
  owner_prep_kernel<<<grid, 256, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_owner_apply(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, const int64_t* d_req,
                          const int32_t* d_counts, const float* d_grads, int32_t world, int32_t cap, int32_t ld,
                          uint32_t* d_mask, const int32_t* d_others, int32_t fm_k, const float* d_lr_t, float beta1,
                          float beta2, float eps, void* stream) {
  ETR_CHECK_ARG(ctx && table && table->d_data && d_m && d_v && d_req && d_counts && d_grads && d_mask && d_others && d_lr_t,
                "NULL argument");
  ETR_CHECK_ARG(table->dtype == ETR_F32 && world >= 1 && world <= 16 && cap > 0 && ld % 4 == 0 && ld <= table->stride &&
                    ld / 4 <= 32, "fp32 table, ld <= stride");
  ETR_CHECK_ARG(fm_k == 0 || (fm_k % 4 == 0 && fm_k + 4 <= ld), "fm_k must be a multiple of 4 with a chunk behind it");
  OwnerApplyParams p;
  p.req = (const long long*)d_req; p.counts = d_counts; p.grads = d_grads; p.world = world; p.cap = cap; p.ld = ld;
  p.mask = d_mask; p.others = d_others; p.table = (float*)table->d_data; p.m = d_m; p.v = d_v; p.stride = table->stride;
  p.rows = table->rows; p.fm_k = fm_k; p.d_lr_t = d_lr_t; p.b1 = beta1; p.b2 = beta2; p.eps = eps;
  int lpr = 1;
  while (lpr < ld / 4) lpr <<= 1;
  static int bps = -1, unroll = 1;                          // ETR_OWNER_BPS / ETR_OWNER_U: measurement knobs
  if (bps < 0) {
    const char* e = getenv("ETR_OWNER_BPS"); bps = e ? atoi(e) : 0;
    const char* u = getenv("ETR_OWNER_U"); unroll = (u && atoi(u) == 2) ? 2 : 1;
  }
  const int per_sm = bps > 0 ? bps : 16 / (world < 8 ? world : 8) + 1;
  cudaStream_t s = (cudaStream_t)stream;
  static int five = -1;
  if (five < 0) { const char* e = getenv("ETR_OWNER_LPR5"); five = !(e && atoi(e) == 0); }
  if (five && ld == 20 && unroll == 1) {
    dim3 g5((unsigned)grid_for(cap, 8 * 6, ctx->sm_count, per_sm), (unsigned)world);
    owner_apply_kernel<5, 1><<<g5, 256, 0, s>>>(p);
    ETR_LAUNCH_CHECK(ctx);
    return ETR_OK;
  }
  dim3 grid((unsigned)grid_for(cap, 8 * (32 / lpr) * unroll, ctx->sm_count, per_sm), (unsigned)world);
  if (lpr == 8 && unroll == 2) owner_apply_kernel<8, 2><<<grid, 256, 0, s>>>(p);
  else switch (lpr) {
    case 1: owner_apply_kernel<1, 1><<<grid, 256, 0, s>>>(p); break;
    case 2: owner_apply_kernel<2, 1><<<grid, 256, 0, s>>>(p); break;
    case 4: owner_apply_kernel<4, 1><<<grid, 256, 0, s>>>(p); break;
    case 8: owner_apply_kernel<8, 1><<<grid, 256, 0, s>>>(p); break;
    case 16: owner_apply_kernel<16, 1><<<grid, 256, 0, s>>>(p); break;
    default: owner_apply_kernel<32, 1><<<grid, 256, 0, s>>>(p); break;
  }
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_peer_barrier(etr_ctx* ctx, uint32_t* const* h_peer_flags, uint32_t* d_my_flags, uint32_t* d_epoch,
                     int32_t world, int32_t rank, void* stream) {
  ETR_CHECK_ARG(ctx && h_peer_flags && d_my_flags && d_epoch, "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && rank >= 0 && rank < world, "world must be in [1,16]");
  BarrierParams p;
  memset(&p, 0, sizeof(p));
  for (int g = 0; g < world; ++g) {
    ETR_CHECK_ARG(h_peer_flags[g] != nullptr, "NULL peer flag pointer");
    p.peer_flags[g] = h_peer_flags[g];
  }
  p.my_flags = d_my_flags; p.epoch = d_epoch; p.world = world; p.rank = rank; p.err = ctx->d_err;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_peer_allreduce_push(etr_ctx* ctx, const float* d_src, int64_t n, float* const* h_peer_slots, int32_t world,
                            int32_t rank, void* stream) {
  ETR_CHECK_ARG(ctx && d_src && h_peer_slots && n > 0, "bad argument");
  ETR_CHECK_ARG(world >= 1 && world <= 16 && rank >= 0 && rank < world, "world must be in [1,16]");
  ArPushParams p;
  memset(&p, 0, sizeof(p));
  for (int g = 0; g < world; ++g) {
    ETR_CHECK_ARG(h_peer_slots[g] != nullptr, "NULL peer slot pointer");
    p.slots[g] = h_peer_slots[g];
  }
  p.src = d_src; p.n = n; p.world = world; p.rank = rank;
  peer_allreduce_push_kernel<<<grid_for(n * world, 256, ctx->sm_count, 4), 256, 0, (cudaStream_t)stream>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_peer_allreduce_sum(etr_ctx* ctx, const float* d_slots, int64_t n, int32_t world, float* d_dst, void* stream) {
  ETR_CHECK_ARG(ctx && d_slots && d_dst && n > 0 && world >= 1, "bad argument");
  peer_allreduce_sum_kernel<<<grid_for(n, 256, ctx->sm_count, 4), 256, 0, (cudaStream_t)stream>>>(d_slots, n, world, d_dst);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_shard_partition(etr_ctx* ctx, const int64_t* d_ids, int64_t n, int32_t world, int64_t rows_global,
                        int64_t* d_send_rows, int64_t* d_send_pos, int64_t* d_inv_pos, int32_t* d_counts,
                        void* stream) {
  ETR_CHECK_ARG(ctx && d_ids && d_send_rows && d_send_pos && d_inv_pos && d_counts, "NULL argument");
  ETR_CHECK_ARG(world >= 1 && world <= 256, "world must be in [1,256]");
  ETR_CHECK_ARG(n >= 0 && n < 0x7fffffffLL, "n must fit int32");
  cudaStream_t s = (cudaStream_t)stream;
  ETR_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * world, s));
  if (n == 0) return ETR_OK;
  int end_bit = 1;
  while ((1 << end_bit) < world) ++end_bit;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, (int)n, 0, end_bit, s);
  const size_t arr = ((size_t)n * 4 + 255) & ~(size_t)255;
  int st = etr_ws_reserve(ctx, 4 * arr + sort_bytes + 256);
  if (st != ETR_OK) return st;
  char* ws = (char*)ctx->d_ws;
  unsigned* keys_in = (unsigned*)ws;
  unsigned* keys_out = (unsigned*)(ws + arr);
  int* slots_in = (int*)(ws + 2 * arr);
  int* slots_out = (int*)(ws + 3 * arr);
  void* tmp = ws + 4 * arr;
  const int grid = grid_for(n, 256, ctx->sm_count, 8);
  shard_keys_kernel<<<grid, 256, 0, s>>>((const long long*)d_ids, n, world, rows_global, keys_in, slots_in, ctx->d_err);
  ETR_LAUNCH_CHECK(ctx);
  ETR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, keys_in, keys_out, slots_in, slots_out, (int)n, 0, end_bit, s));
  shard_finish_kernel<<<grid, 256, 0, s>>>((const long long*)d_ids, keys_out, slots_out, n, world, rows_global,
                                           (long long*)d_send_rows, (long long*)d_send_pos, (long long*)d_inv_pos,
                                           d_counts);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
