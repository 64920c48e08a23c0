// Masked / weighted pooling of looked-up rows behind the Embedding drop-in.
//
// The sequence models of the reference gather a padded behaviour series and then weight, mask and sum it with
// separate eager ops, materialising [B, L, C*k] twice:
//   X_series = embed(ids[B, L*C]) -> [B, L, C*k];  pooled = sum_l mask[b,l] * score[b,l] * X_series[b,l,:]
//   (7.SIM/CustomLayers.py:88-95,107-118; 5.DIN/CustomLayers.py:258-283; plain mean over L: 5.DIN/...:662)
// and FiBiNet++ scales every looked-up row by a feature value: embed(keys[B,F]) * values[B,F,None]
// (11.FiBiNet++/CustomLayers.py:124-126).  Both are   out[b, (l,) c, :] = w[b,l] * table[ids[b,l,c], :]   with or
// without the sum over l, so one kernel pair covers them: the rows are fetched once with 128-bit loads, scaled in
// registers and either accumulated over l (reduce) or written in place; [B, L, C*k] is never materialised in reduce mode.
//   weight w[b,l] = (d_weights ? d_weights[b,l] : 1) * (d_mask ? d_mask[b,l] != 0 : 1) * (has_pad ? ids[b,l,0] != pad : 1)
// (the reference masks on the FIRST series feature's id, 7.SIM/CustomLayers.py:116).
// Backward: per-occurrence gradient rows  g[b,l,c,:] = w[b,l] * dOut[b,(l,)c,:]  (table row layout, ld = grad_ld) for the
// sorted-ID segment reduction, and  dw[b,l] = mask * sum_c <dOut[b,(l,)c,:], table[ids[b,l,c],:]>  for the attention scores.
#include "etr_common.cuh"

namespace etr {

struct SeqPoolParams {
  const char* table; long long rows; int row_bytes; int k; int bf16;
  const long long* ids; long long B; int L, C;
  const float* weights; const unsigned char* mask; long long pad; int has_pad; int reduce;
  float* out;                          // reduce: [B, C*k]; else [B, L, C*k]
  const float* dout; float* occ_grad; int grad_ld; float* dw;
  unsigned long long* err;
};

__device__ __forceinline__ float sp_weight(const SeqPoolParams& p, long long b, int l) {
  float w = p.weights ? p.weights[b * p.L + l] : 1.0f;
  if (p.mask && p.mask[b * p.L + l] == 0) w = 0.f;
  if (p.has_pad && p.ids[(b * p.L + l) * p.C] == p.pad) w = 0.f;
  return w;
}

// one lane group (k/4 lanes) per (b, c); loops over l with 4 independent row loads in flight
template <typename Elem, int LPR>
__global__ void __launch_bounds__(256) seq_pool_fwd_kernel(const SeqPoolParams p) {
  constexpr int EPC = Chunk<Elem>::kElems;            // elements per 16-byte chunk
  const int lane = threadIdx.x & 31, gl = lane % LPR, g = lane / LPR;
  const long long group = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / LPR) + g;
  const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5) * (32 / LPR);
  const long long total = p.B * p.C;
  for (long long t = group; t < total; t += ngroups) {
    const long long b = t / p.C;
    const int c = (int)(t % p.C);
    float acc[EPC];
#pragma unroll
    for (int e = 0; e < EPC; ++e) acc[e] = 0.f;
    for (int l0 = 0; l0 < p.L; l0 += 4) {
      Chunk<Elem> ch[4];
      float w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int l = l0 + q;
        w[q] = 0.f;
        ch[q].zero();
        if (l < p.L) {
          w[q] = sp_weight(p, b, l);
          const long long id = p.ids[(b * p.L + l) * p.C + c];
          const bool pad_slot = p.has_pad && id == p.pad;
          if ((unsigned long long)id >= (unsigned long long)p.rows) {
            if (!pad_slot) { if (gl == 0) flag_bad_id(p.err, id); }
            w[q] = 0.f;
          } else if (w[q] != 0.f || !p.reduce) {
            ch[q].load(p.table + id * p.row_bytes + gl * 16);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int l = l0 + q;
        if (l >= p.L) break;
        if (p.reduce) {
#pragma unroll
          for (int e = 0; e < EPC; ++e) acc[e] += w[q] * ch[q].v[e];
        } else {
          float* dst = p.out + ((b * p.L + l) * p.C + c) * (long long)p.k + gl * EPC;
#pragma unroll
          for (int e = 0; e < EPC; ++e) dst[e] = w[q] * ch[q].v[e];
        }
      }
    }
    if (p.reduce) {
      float* dst = p.out + (b * p.C + c) * (long long)p.k + gl * EPC;
#pragma unroll
      for (int e = 0; e < EPC; ++e) dst[e] = acc[e];
    }
  }
}

// one lane group per (b, l): gradient rows of its C occurrences and the weight gradient dw[b,l]
template <typename Elem, int LPR>
__global__ void __launch_bounds__(256) seq_pool_bwd_kernel(const SeqPoolParams p) {
  constexpr int EPC = Chunk<Elem>::kElems;
  const int lane = threadIdx.x & 31, gl = lane % LPR, g = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
  const long long group = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / LPR) + g;
  const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5) * (32 / LPR);
  const long long total = p.B * p.L;
  // warp-uniform trip count (group_sum shuffles need the whole group; groups past the end idle with w = 0)
  const long long rounds = (total + ngroups - 1) / ngroups;
  for (long long it = 0; it < rounds; ++it) {
    const long long t = group + it * ngroups;
    const bool live = t < total;
    const long long b = live ? t / p.L : 0;
    const int l = live ? (int)(t % p.L) : 0;
    const float w = live ? sp_weight(p, b, l) : 0.f;
    float dot = 0.f;
    for (int c = 0; c < p.C; ++c) {
      float d[EPC];
      const float* src = p.dout + (p.reduce ? (b * p.C + c) : ((b * p.L + l) * p.C + c)) * (long long)p.k + gl * EPC;
#pragma unroll
      for (int e = 0; e < EPC; ++e) d[e] = live ? src[e] : 0.f;
      if (live && p.occ_grad) {
        float* dst = p.occ_grad + ((b * p.L + l) * p.C + c) * (long long)p.grad_ld + gl * EPC;
#pragma unroll
        for (int e = 0; e < EPC; ++e) dst[e] = w * d[e];
        // padding columns behind k (table row layout) are zero
        for (int e = p.k + gl; e < p.grad_ld; e += LPR) p.occ_grad[((b * p.L + l) * p.C + c) * (long long)p.grad_ld + e] = 0.f;
      }
      if (p.dw) {
        float s = 0.f;
        if (live) {
          const long long id = p.ids[(b * p.L + l) * p.C + c];
          if ((unsigned long long)id < (unsigned long long)p.rows) {
            Chunk<Elem> ch;
            ch.load(p.table + id * p.row_bytes + gl * 16);
#pragma unroll
            for (int e = 0; e < EPC; ++e) s += d[e] * ch.v[e];
          }
        }
        dot += group_sum<LPR>(s, gmask);
      }
    }
    if (live && p.dw && gl == 0) {
      // d/dw of w_user * mask: the mask (and the pad mask) multiply the user weight
      float mk = 1.f;
      if (p.mask && p.mask[b * p.L + l] == 0) mk = 0.f;
      if (p.has_pad && p.ids[(b * p.L + l) * p.C] == p.pad) mk = 0.f;
      p.dw[b * p.L + l] = mk * dot;
    }
  }
}

template <typename Elem>
static int sp_launch(etr_ctx* ctx, const SeqPoolParams& p, bool backward, cudaStream_t s) {
  constexpr int EPC = Chunk<Elem>::kElems;
  const int lpr_need = p.k / EPC;
  int lpr = 1;
  while (lpr < lpr_need) lpr <<= 1;
  if (lpr != lpr_need || lpr > 32) {
    etr_set_error("sequence pooling: k must be %d x a power of two, at most %d (k = %d)", EPC, 32 * EPC, p.k);
    return ETR_EUNSUPPORTED;
  }
  const long long items = backward ? p.B * p.L : p.B * p.C;
  const int grid = grid_for(items, 8 * (32 / lpr), ctx->sm_count, 8);
#define ETR_SP(LPR)                                                                    \
  do {                                                                                 \
    if (backward) seq_pool_bwd_kernel<Elem, LPR><<<grid, 256, 0, s>>>(p);              \
    else seq_pool_fwd_kernel<Elem, LPR><<<grid, 256, 0, s>>>(p);                       \
  } while (0)
  switch (lpr) {
    case 1: ETR_SP(1); break;
    case 2: ETR_SP(2); break;
    case 4: ETR_SP(4); break;
    case 8: ETR_SP(8); break;
    case 16: ETR_SP(16); break;
    default: ETR_SP(32); break;
  }
#undef ETR_SP
  return ETR_OK;
}

static int sp_fill(const char* fn, etr_ctx* ctx, const etr_table* table, int k, const int64_t* d_ids, int64_t B, int L, int C,
                   const float* d_weights, const uint8_t* d_mask, int64_t pad_id, int has_pad, int reduce, SeqPoolParams* p) {
  if (!ctx || !table || !table->d_data || !d_ids) { etr_set_error("%s: NULL argument", fn); return ETR_EINVAL; }
  const int esize = table->dtype == ETR_BF16 ? 2 : 4;
  if (k <= 0 || k > table->width || (table->stride * esize) % 16 != 0 || ((uintptr_t)table->d_data & 15) || B < 0 || L < 1 || C < 1) {
    etr_set_error("%s: bad shape (k=%d width=%d L=%d C=%d)", fn, k, table->width, L, C);
    return ETR_EINVAL;
  }
  memset(p, 0, sizeof(*p));
  p->table = (const char*)table->d_data; p->rows = table->rows; p->row_bytes = table->stride * esize; p->k = k;
  p->bf16 = table->dtype == ETR_BF16;
  p->ids = (const long long*)d_ids; p->B = B; p->L = L; p->C = C; p->weights = d_weights; p->mask = d_mask;
  p->pad = pad_id; p->has_pad = has_pad; p->reduce = reduce; p->err = ctx->d_err;
  return ETR_OK;
}

}  // namespace etr

using namespace etr;

extern "C" {

int etr_sequence_pool_forward(etr_ctx* ctx, const etr_table* table, int32_t k, const int64_t* d_ids, int64_t batch, int32_t L,
                              int32_t C, const float* d_weights, const uint8_t* d_mask, int64_t pad_id, int32_t has_pad,
                              int32_t reduce, float* d_out, void* stream) {
  SeqPoolParams p;
  int st = sp_fill(__func__, ctx, table, k, d_ids, batch, L, C, d_weights, d_mask, pad_id, has_pad, reduce, &p);
  if (st != ETR_OK) return st;
  ETR_CHECK_ARG(d_out != nullptr && ((uintptr_t)d_out & 15) == 0, "d_out must be 16-byte aligned");
  if (batch == 0) return ETR_OK;
  p.out = d_out;
  st = p.bf16 ? sp_launch<__nv_bfloat16>(ctx, p, false, (cudaStream_t)stream) : sp_launch<float>(ctx, p, false, (cudaStream_t)stream);
  if (st != ETR_OK) return st;
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

int etr_sequence_pool_backward(etr_ctx* ctx, const etr_table* table, int32_t k, const int64_t* d_ids, int64_t batch, int32_t L,
                               int32_t C, const float* d_weights, const uint8_t* d_mask, int64_t pad_id, int32_t has_pad,
                               int32_t reduce, const float* d_dout, float* d_occ_grad, int32_t grad_ld, float* d_dweights,
                               void* stream) {
  SeqPoolParams p;
  int st = sp_fill(__func__, ctx, table, k, d_ids, batch, L, C, d_weights, d_mask, pad_id, has_pad, reduce, &p);
  if (st != ETR_OK) return st;
  ETR_CHECK_ARG(d_dout && (d_occ_grad || d_dweights), "nothing to do");
  ETR_CHECK_ARG(!d_occ_grad || (grad_ld >= k && grad_ld % 4 == 0), "grad_ld must be a multiple of 4 and >= k");
  if (batch == 0) return ETR_OK;
  p.dout = d_dout; p.occ_grad = d_occ_grad; p.grad_ld = grad_ld; p.dw = d_dweights;
  st = p.bf16 ? sp_launch<__nv_bfloat16>(ctx, p, true, (cudaStream_t)stream) : sp_launch<float>(ctx, p, true, (cudaStream_t)stream);
  if (st != ETR_OK) return st;
  ETR_LAUNCH_CHECK(ctx);
  return ETR_OK;
}

}  // extern "C"
