// K2+K8 fused, tiled form for the hot configuration (Adam apply on a local RECORD table, k = 16): the same computation as
// fm_fused_apply.cu,
//   dv_r = sum_(b,f) [ g_b S_b + dflat[b, f] ] - v_r * sum g_b ,   dw_r = sum g_b ,   then row-wise Adam on the record,
// organised so that NO lane group ever walks a dependent  seg_start -> sorted_bag -> (g, S, dflat)  chain per occurrence
// (what bound the row-parallel kernel: profiles/r02_mb_apply.md) and the work of a warp is proportional to the
// occurrences it covers, not to the longest run among its rows.
//
//   * fm_tile_prepare_kernel (once per plan): one 16-byte descriptor per unique row (first sorted position, run length,
//     table row), padded to whole tiles of 8 rows; runs longer than `short_max` are cut into ITEMS of `item_len`
//     occurrences.
//   * fm_tile_kernel: warp w takes entries w, w + W, ... of the work list [items..., tiles...].
//     TILE = 8 consecutive unique rows, one per 4-lane group.  The tile's occurrences are one contiguous span of the
//     sorted list: the warp gathers (g, S, dflat) for 16 occurrences per round with every lane group busy, parks
//     g*S + dflat in shared memory in sorted order, each group sums its own run from there (a 30-cycle chain instead
//     of a DRAM one), then applies Adam to its record.  Per warp the loop is software-pipelined by cp.async:
//       tile n+2: descriptors            -> shared memory
//       tile n+1: 8 records, bag indices -> shared memory  (no registers held while they are in flight)
//       tile n  : gather -> park -> sum -> Adam -> write back
//     so the only exposed round trip of an iteration is the operand gather.
//     ITEM = up to item_len occurrences of one long run: bag indices coalesced, 4 gathers in flight per lane, fixed
//     shuffle tree.  A single-item run is updated by its warp directly; for a multi-item run the warp that arrives LAST
//     (per-run counter) sums the partial rows in item order and updates the record -- the result does not depend on which
//     warp that is, so the kernel is deterministic without a combine launch.
//
// Two properties of ptxas output shaped the loop (profiles/r02_prof_apply_tile*.md): (1) every __shfl_sync it cannot
// prove convergent is guarded by a BRA.DIV that first waits for ALL outstanding loads, (2) register loads that stay in
// flight across an iteration share the 6 scoreboards with the loads of the current iteration.  Either way a register
// prefetch is waited for right after it is issued -- hence no shuffles in the tile loop and cp.async (commit groups,
// not scoreboards) for everything that is prefetched.
//
// Replaces tape.gradient + Keras apply_gradients on IndexedSlices (2.FM/ModelManager.py:176-178) for the FM family.
#include <stdlib.h>

#include "etr_async.cuh"
#include "fm_fused_tile.cuh"

namespace etr {

struct TileRun { int base; int nchunks; };

struct TileParams {
  TileArgs a;
  int short_max, item_len;
  int* counters;                      // [0] multi-item runs, [1] items
  int4* rowinfo;                      // [8 * ceil(n_unique / 8)]: (first sorted position, len | -len if long | 0 pad, table row lo, hi)
  TileRun* runs; int* arrived; int4* items; float* partials;
  int max_runs, max_items;
};

struct Occ { float g; float4 s; uint2 d; };

__device__ __forceinline__ uint2 ldg_nc_u2(const void* ptr) {
  uint2 r;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(ptr));
  return r;
}

template <int DF>
__device__ __forceinline__ void occ_load(const TileArgs& a, int bag, int gl, Occ& o) {
  const int b = (int)(((unsigned long long)(unsigned)bag * a.magic) >> a.shift);
  const int f = bag - b * a.F;
  o.g = __ldg(a.dlogit + b);
  o.s = __ldg(reinterpret_cast<const float4*>(a.sumv + b * 16) + gl);
  o.d = make_uint2(0u, 0u);
  if (DF) o.d = ldg_nc_u2(a.dflat + (long long)b * a.flat_ld + f * 16 + gl * 4);
}
__device__ __forceinline__ float4 occ_value(const Occ& o) {
  return make_float4(fmaf(o.g, o.s.x, __uint_as_float(o.d.x << 16)), fmaf(o.g, o.s.y, __uint_as_float(o.d.x & 0xffff0000u)),
                     fmaf(o.g, o.s.z, __uint_as_float(o.d.y << 16)), fmaf(o.g, o.s.w, __uint_as_float(o.d.y & 0xffff0000u)));
}

// record state of one row in a 4-lane group: lane gl holds columns 4gl..4gl+3 of var / m / v
struct RecState { float4 var, m, v; float wx; };
__device__ __forceinline__ void rec_load(const float* rec, int gl, RecState& st) {
  st.var = *reinterpret_cast<const float4*>(rec + gl * 4);
  st.m = *reinterpret_cast<const float4*>(rec + 20 + gl * 4);
  st.v = *reinterpret_cast<const float4*>(rec + 40 + gl * 4);
  st.wx = 0.f;
  if (gl) st.wx = rec[20 * gl - 4];                        // lanes 1, 2, 3: w (float 16), its m (36), its v (56)
}
__device__ __forceinline__ void rec_finish(float* rec, int gl, RecState& st, float4 P, float sum_g, float w, float wm, float wv,
                                           const TileArgs& a, float lr_t) {
  const float4 gr = make_float4(P.x - st.var.x * sum_g, P.y - st.var.y * sum_g, P.z - st.var.z * sum_g, P.w - st.var.w * sum_g);
  adam_update4_fast(st.var, st.m, st.v, gr, lr_t, a.b1, a.b2, a.eps);
  *reinterpret_cast<float4*>(rec + gl * 4) = st.var;
  *reinterpret_cast<float4*>(rec + 20 + gl * 4) = st.m;
  *reinterpret_cast<float4*>(rec + 40 + gl * 4) = st.v;
  if (gl == 0) {
    adam_update1_fast(w, wm, wv, sum_g, lr_t, a.b1, a.b2, a.eps);
    rec[16] = w; rec[36] = wm; rec[56] = wv;
  }
}

// push mode: the finished row goes to its owner's mailbox slot
__device__ __forceinline__ void row_push(const TileArgs& a, int so, int gl, float4 P, float sum_g) {
  const int owner = so / a.cap;
  float* dst = a.grads_mb[owner] + (long long)(so - owner * a.cap) * a.gld;
  *reinterpret_cast<float4*>(dst + gl * 4) = P;
  if (gl == 0) *reinterpret_cast<float4*>(dst + 16) = make_float4(sum_g, 0.f, 0.f, 0.f);
}

// one thread per unique row (and per pad row of the last tile): descriptor + the items of a long run.
// items[i] = (run u, first sorted position, length, slot), slot = -1 for a single-item run, else the index of its
// (base, nchunks) entry and arrival counter
__global__ void __launch_bounds__(256) fm_tile_prepare_kernel(const TileParams p) {
  const int n_unique = *p.a.n_unique;
  const int n_rows = (n_unique + 7) & ~7;
  const int n_valid = p.a.seg_start[n_unique];
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_rows; u += gridDim.x * blockDim.x) {
    if (u >= n_unique) { p.rowinfo[u] = make_int4(n_valid, 0, 0, 0); continue; }
    const int s0 = p.a.seg_start[u];
    const int len = p.a.seg_start[u + 1] - s0;
    const long long uid = p.a.unique_ids[u];
    p.rowinfo[u] = make_int4(s0, len > p.short_max ? -len : len, (int)(uid & 0xffffffffLL), (int)(uid >> 32));
    if (len > p.short_max) {
      const int nch = (len + p.item_len - 1) / p.item_len;
      const int base = atomicAdd(&p.counters[1], nch);
      int slot = -1;
      if (nch > 1) {
        slot = atomicAdd(&p.counters[0], 1);
        if (slot < p.max_runs) { p.runs[slot] = TileRun{base, nch}; p.arrived[slot] = 0; }
      }
      if (base + nch <= p.max_items && slot < p.max_runs)
        for (int c = 0; c < nch; ++c)
          p.items[base + c] = make_int4(u, s0 + c * p.item_len, min(p.item_len, len - c * p.item_len), slot);
    }
  }
}

__device__ __forceinline__ void tree8(float4& t, float& ts) {     // fixed combine of the 8 lane groups of a warp
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
    ts += __shfl_xor_sync(0xffffffffu, ts, o);
  }
}

template <int DF, int PUSH>
__device__ __forceinline__ void tile_item(const TileParams& p, int it, int lane, float lr_t) {
  const TileArgs& a = p.a;
  const unsigned full = 0xffffffffu;
  const int gl = lane & 3, g = lane >> 2;
  const int4 item = p.items[it];
  const int lo = item.y, n_it = item.z, slot = item.w;
  float* rec = PUSH ? a.table : a.table + __ldg(a.unique_ids + item.x) * 64;
  RecState st;
  st.var = st.m = st.v = make_float4(0.f, 0.f, 0.f, 0.f);
  st.wx = 0.f;
  if (!PUSH && slot < 0 && lane < 4) rec_load(rec, gl, st);       // single-item run: the record travels during the gather
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float sum_g = 0.f;
  for (int pos = 0; pos < n_it; pos += 128) {
    const int n = min(128, n_it - pos);
    int bag[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bag[q] = (q * 32 + lane < n) ? __ldg(a.sorted_bag + lo + pos + q * 32 + lane) : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q * 32 < n) {                                  // warp-uniform
        Occ o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int bg = __shfl_sync(full, bag[q], t * 8 + g);
          if (q * 32 + t * 8 + g < n) occ_load<DF>(a, bg, gl, o[t]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (q * 32 + t * 8 + g < n) {
            const float4 c = occ_value(o[t]);
            acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
            sum_g += o[t].g;
          }
        }
      }
    }
  }
  tree8(acc, sum_g);
  if (slot < 0) {
    if (PUSH) {
      if (lane < 4) row_push(a, __ldg(a.slot_of_u + item.x), gl, acc, sum_g);
      return;
    }
    const float w = __shfl_sync(full, st.wx, 1), wm = __shfl_sync(full, st.wx, 2), wv = __shfl_sync(full, st.wx, 3);
    if (lane < 4) rec_finish(rec, gl, st, acc, sum_g, w, wm, wv, a, lr_t);
    return;
  }
  float* dst = p.partials + (long long)it * 20;
  if (lane < 4) {
    *reinterpret_cast<float4*>(dst + gl * 4) = acc;
    if (gl == 0) dst[16] = sum_g;
  }
  __threadfence();
  __syncwarp();
  int old = 0;
  if (lane == 0) old = atomicAdd(&p.arrived[slot], 1);
  old = __shfl_sync(full, old, 0);
  const TileRun run = p.runs[slot];
  if (old == run.nchunks - 1) {                          // last item of the run to finish: combine in item order
    __threadfence();
    if (lane == 0) p.arrived[slot] = 0;                  // self-cleaning: the lists are reused by every apply of this plan
    if (!PUSH && lane < 4) rec_load(rec, gl, st);
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    float ts = 0.f;
    for (int c = g; c < run.nchunks; c += 8) {
      const float* src = p.partials + (long long)(run.base + c) * 20;
      const float4 x = __ldcg(reinterpret_cast<const float4*>(src) + gl);
      t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
      ts += __ldcg(src + 16);
    }
    tree8(t, ts);
    if (PUSH) {
      if (lane < 4) row_push(a, __ldg(a.slot_of_u + item.x), gl, t, ts);
      return;
    }
    const float w = __shfl_sync(full, st.wx, 1), wm = __shfl_sync(full, st.wx, 2), wv = __shfl_sync(full, st.wx, 3);
    if (lane < 4) rec_finish(rec, gl, st, t, ts, w, wm, wv, a, lr_t);
  }
}

// per-warp shared memory; every array is indexed by `lane` (= 4 * group + gl) or by a sorted position
struct WarpSmem {
  float4 P[128];                      // [j * 4 + gl]: g*S + dflat of the (up to 32) occurrences of a round, sorted order
  float G[32];
  float4 var[2][32], m[2][32], v[2][32];      // two tiles of records (cp.async destinations)
  float w[2][32];                     // [.][4 * g + (0 | 1 | 2 | 3)] = (w, w, its m, its v)
  int4 hdr[2][8];                     // row descriptors of a tile
  int bag[2][32];                     // bag indices of a tile's first 32 occurrences
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rec_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rec_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// RPG: occurrences per lane group and round (2: rounds of 16, 4: rounds of 32 -- a typical c2 tile, 29 occurrences, is then
// ONE exposed operand round trip instead of two, at 14 more registers)
template <int DF, int OCC, int PUSH, int RPG = 2>
__global__ void __launch_bounds__(128, OCC) fm_tile_kernel(const TileParams p) {
  constexpr int RND = 8 * RPG;
  __shared__ WarpSmem smem[4];
  const TileArgs& a = p.a;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gl = lane & 3, g = lane >> 2;
  WarpSmem& sm = smem[wid];
  const int n_unique = __ldg(a.n_unique);
  const float lr_t = a.d_lr_t ? __ldg(a.d_lr_t) : a.lr_t;
  const int n_tiles = (n_unique + 7) >> 3;
  const int last_slot = (int)a.n_slots - 1;
  int n_items = p.counters[1];
  if (n_items > p.max_items || p.counters[0] > p.max_runs) n_items = 0;      // cannot happen (the bounds are exact)
  const int W = gridDim.x * 4;
  int idx = blockIdx.x * 4 + wid;
  for (; idx < n_items; idx += W) tile_item<DF, PUSH>(p, idx, lane, lr_t);
  int tile = idx - n_items;
  if (tile >= n_tiles) return;

  // records + bag indices of the tile whose descriptors sit in sm.hdr[hb] -> buffers `nb`.  Unconditional: a pad or
  // long row fetches a record nobody reads, bag indices past the span are clamped to a valid slot.
  auto issue_rec_bags = [&](int t, int hb, int nb, int4& h, int& S0, int& S1) {
    h = sm.hdr[hb][g];
    const int4 first = sm.hdr[hb][0], last = sm.hdr[hb][7];
    S0 = first.x; S1 = last.x + abs(last.y);
    if (!PUSH) {
      const long long uid = ((long long)h.w << 32) | (unsigned)h.z;
      const float* rec = a.table + uid * 64;
      cp_async16(&sm.var[nb][lane], rec + gl * 4);
      cp_async16(&sm.m[nb][lane], rec + 20 + gl * 4);
      cp_async16(&sm.v[nb][lane], rec + 40 + gl * 4);
      cp_async4(&sm.w[nb][lane], rec + (gl ? 20 * gl - 4 : 16));
    } else if (lane >= 16 && lane < 24) {                  // push mode: the 8 mailbox slots instead of the 8 records
      cp_async4(&sm.w[nb][lane - 16], a.slot_of_u + min(t * 8 + lane - 16, last_slot));
    }
    if (lane < RND) cp_async4(&sm.bag[nb][lane], a.sorted_bag + min(S0 + lane, last_slot));
  };
  auto issue_hdr = [&](int t, int hb) {
    if (t < n_tiles && lane < 8) cp_async16(&sm.hdr[hb][lane], p.rowinfo + t * 8 + lane);
  };

  issue_hdr(tile, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
  int4 hc;
  int S0c, S1c;
  issue_rec_bags(tile, 0, 0, hc, S0c, S1c);
  issue_hdr(tile + W, 1);
  cp_async_commit();
  int buf = 0, hb = 1;
  for (; tile < n_tiles; tile += W) {
    cp_async_wait<0>();                                    // this tile's records + bag indices, the next tile's descriptors
    __syncwarp();
    int4 hn = make_int4(0, 0, 0, 0);
    int S0n = 0, S1n = 0;
    if (tile + W < n_tiles) {
      issue_rec_bags(tile + W, hb, buf ^ 1, hn, S0n, S1n);
      issue_hdr(tile + 2 * W, hb ^ 1);
    }
    cp_async_commit();
    // ---- this tile
    const int s0 = hc.x, len = hc.y;                       // len < 0: long run (an item), 0: pad row
    const int s1 = s0 + len;
    const bool any_long = __any_sync(full, len < 0);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float sum_g = 0.f;
    // one span (the whole tile) unless the tile holds a long run: then one span per short row
    const int nspan = any_long ? 8 : 1;
    for (int q = 0; q < nspan; ++q) {
      int lo = S0c, hi = S1c;
      int bag[RPG];
#pragma unroll
      for (int r = 0; r < RPG; ++r) bag[r] = 0;
      if (any_long) {
        lo = __shfl_sync(full, s0, q * 4);
        const int l = __shfl_sync(full, len, q * 4);
        if (l <= 0) continue;
        hi = lo + l;
#pragma unroll
        for (int r = 0; r < RPG; ++r)
          if (lo + r * 8 + g < hi) bag[r] = __ldg(a.sorted_bag + lo + r * 8 + g);
      } else {
#pragma unroll
        for (int r = 0; r < RPG; ++r) bag[r] = sm.bag[buf][r * 8 + g];
      }
      for (int pos = lo; pos < hi; pos += RND) {          // a round: RND occurrences, RPG per lane group
        const int n = hi - pos;                            // (>= RND: a full round)
        int nbag[RPG];                                     // the next round's bag indices travel during this one
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
          nbag[r] = 0;
          if (pos + RND + r * 8 + g < hi) nbag[r] = __ldg(a.sorted_bag + pos + RND + r * 8 + g);
        }
        Occ o[RPG];
#pragma unroll
        for (int r = 0; r < RPG; ++r)
          if (r * 8 + g < n) occ_load<DF>(a, bag[r], gl, o[r]);
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
          if (r * 8 + g < n) {
            sm.P[r * 32 + lane] = occ_value(o[r]);
            if (gl == 0) sm.G[r * 8 + g] = o[r].g;
          }
        }
        __syncwarp();
        const int ja = max(s0, pos) - pos, je = min(min(s1, hi), pos + RND) - pos;     // empty unless len > 0
        for (int j = ja; j < je; ++j) {
          const float4 t = sm.P[j * 4 + gl];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          sum_g += sm.G[j];
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < RPG; ++r) bag[r] = nbag[r];
      }
    }
    if (PUSH) {
      // the 8 finished rows go out as ONE 80-byte request each: staged in shared memory (the park buffer is free), then
      // five adjacent lanes store the five 16-byte chunks of a row (40 chunks: two warp stores) -- half the peer write
      // requests of the direct form (row_push: 64 bytes by the row's four lanes + a separate 16-byte store for sum_g).
      // Measured at N = 4: 132.9 -> 134.1 us, i.e. no difference; the growth of the push with N is not the request
      // count (profiles/r02_mgpu.md).  Kept: one request per row is the cleaner traffic pattern.
      float4* stg = sm.P;                                  // [8 rows][5 chunks]
      int* sslot = reinterpret_cast<int*>(sm.G);           // [8]
      stg[g * 5 + gl] = acc;
      if (gl == 0) {
        stg[g * 5 + 4] = make_float4(sum_g, 0.f, 0.f, 0.f);
        sslot[g] = len > 0 ? __float_as_int(sm.w[buf][g]) : -1;
      }
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int idx = rr * 32 + lane;
        if (idx < 40) {
          const int r = idx / 5, c = idx - r * 5;
          const int so = sslot[r];
          if (so >= 0) {
            const int owner = so / a.cap;
            float* dst = a.grads_mb[owner] + (long long)(so - owner * a.cap) * a.gld;
            *reinterpret_cast<float4*>(dst + c * 4) = stg[idx];
          }
        }
      }
      __syncwarp();
    } else if (len > 0) {
      const long long uid = ((long long)hc.w << 32) | (unsigned)hc.z;
      float* rec = a.table + uid * 64;
      RecState st;
      st.var = sm.var[buf][lane]; st.m = sm.m[buf][lane]; st.v = sm.v[buf][lane];
      st.wx = 0.f;
      float w = 0.f, wm = 0.f, wv = 0.f;
      if (gl == 0) { w = sm.w[buf][lane + 1]; wm = sm.w[buf][lane + 2]; wv = sm.w[buf][lane + 3]; }
      rec_finish(rec, gl, st, acc, sum_g, w, wm, wv, a, lr_t);
    }
    hc = hn; S0c = S0n; S1c = S1n; buf ^= 1; hb ^= 1;
  }
  cp_async_wait<0>();
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// tuning knobs (read once): rows up to T occurrences are handled inside tiles, longer runs as items of ITEM occurrences;
// OCC = CTAs (128 threads) per SM
static int g_T = -1, g_ITEM, g_OCC;
static void tile_knobs() {
  if (g_T >= 0) return;
  g_ITEM = env_int("ETR_TILE_ITEM", 256);
  g_OCC = env_int("ETR_TILE_OCC", 6);     // with rounds of 32 (ETR_TILE_RPG=4, the default): 80 registers, 6 CTAs / SM
  int t = env_int("ETR_TILE_T", 32);
  if (t < 1) t = 1;
  if (g_ITEM < t + 1) g_ITEM = t + 1;
  g_T = t;
}

// layout of the prepared lists inside one buffer (the ctx workspace, or a caller-owned buffer that lives with the plan)
static size_t tile_layout(long long n_slots, char* base, TileParams& p) {
  tile_knobs();
  p.short_max = g_T; p.item_len = g_ITEM;
  // exact bounds: a long run has > T occurrences, a multi-item run > ITEM
  p.max_runs = (int)(n_slots / (g_ITEM + 1) + 1);
  p.max_items = (int)(n_slots / g_ITEM + n_slots / (g_T + 1) + 2);
  auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_cnt = 256, b_runs = a256(sizeof(TileRun) * (size_t)p.max_runs), b_arr = a256(sizeof(int) * (size_t)p.max_runs),
               b_items = a256(sizeof(int4) * (size_t)p.max_items), b_part = a256(sizeof(float) * 20 * (size_t)p.max_items),
               b_rows = a256(sizeof(int4) * (size_t)(n_slots + 8));
  p.counters = (int*)base;
  p.runs = (TileRun*)(base + b_cnt);
  p.arrived = (int*)(base + b_cnt + b_runs);
  p.items = (int4*)(base + b_cnt + b_runs + b_arr);
  p.partials = (float*)(base + b_cnt + b_runs + b_arr + b_items);
  p.rowinfo = (int4*)(base + b_cnt + b_runs + b_arr + b_items + b_part);
  return b_cnt + b_runs + b_arr + b_items + b_part + b_rows;
}

size_t fused_tile_prep_bytes(long long n_slots) {
  TileParams p;
  return tile_layout(n_slots, nullptr, p);
}

int fused_tile_prepare(etr_ctx* ctx, const int* d_seg_start, const long long* d_unique_ids, const int* d_n_unique, long long n_slots,
                       void* d_prep, cudaStream_t s) {
  TileParams p;
  memset(&p, 0, sizeof(p));
  tile_layout(n_slots, (char*)d_prep, p);
  p.a.seg_start = d_seg_start; p.a.unique_ids = d_unique_ids; p.a.n_unique = d_n_unique; p.a.n_slots = n_slots;
  ETR_CUDA(cudaMemsetAsync(p.counters, 0, 2 * sizeof(int), s));
  fm_tile_prepare_kernel<<<grid_for(n_slots, 256, ctx->sm_count, 4), 256, 0, s>>>(p);
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int fused_tile_launch(etr_ctx* ctx, const TileArgs& a, const void* d_prep, cudaStream_t s) {
  TileParams p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  if (d_prep) {
    tile_layout(a.n_slots, (char*)d_prep, p);
  } else {                            // no prepared lists: build them in the ctx workspace first
    int st = etr_ws_reserve(ctx, fused_tile_prep_bytes(a.n_slots));
    if (st != ETR_OK) return st;
    st = fused_tile_prepare(ctx, a.seg_start, a.unique_ids, a.n_unique, a.n_slots, ctx->d_ws, s);
    if (st != ETR_OK) return st;
    tile_layout(a.n_slots, (char*)ctx->d_ws, p);
  }
  const int OCC = a.slot_of_u ? 7 : g_OCC;          // the push form is compiled for 7 CTAs / SM only
  const int grid = grid_for(a.n_slots, 32, ctx->sm_count, OCC);
  static int rpg = -1;
  if (rpg < 0) rpg = env_int("ETR_TILE_RPG", 4) == 2 ? 2 : 4;   // apply mode; the mailbox-push form keeps rounds of 16
#define ETR_TILE(DF) do { \
    if (rpg == 4 && !a.slot_of_u) { \
      if (OCC <= 5) fm_tile_kernel<DF, 5, 0, 4><<<grid, 128, 0, s>>>(p); else if (OCC == 6) fm_tile_kernel<DF, 6, 0, 4><<<grid, 128, 0, s>>>(p); \
      else fm_tile_kernel<DF, 7, 0, 4><<<grid, 128, 0, s>>>(p); \
    } else if (a.slot_of_u) fm_tile_kernel<DF, 7, 1><<<grid, 128, 0, s>>>(p); \
    else if (OCC <= 5) fm_tile_kernel<DF, 5, 0><<<grid, 128, 0, s>>>(p); else if (OCC == 6) fm_tile_kernel<DF, 6, 0><<<grid, 128, 0, s>>>(p); \
    else if (OCC == 7) fm_tile_kernel<DF, 7, 0><<<grid, 128, 0, s>>>(p); else fm_tile_kernel<DF, 8, 0><<<grid, 128, 0, s>>>(p); } while (0)
  if (a.dflat) ETR_TILE(1); else ETR_TILE(0);
#undef ETR_TILE
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // namespace etr
