// Context, error reporting and workspace of libetr.so.
#include <stdarg.h>

#include <new>

#include "etr_common.cuh"

static thread_local char g_err[512] = "";

void etr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int etr_ws_reserve(etr_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return ETR_OK;
  // grow geometrically; freeing synchronises the device, which is fine: growth
  // only happens on the first call at a new (larger) shape.
  size_t want = bytes + bytes / 4 + (1u << 20);
  if (ctx->d_ws) {
    ETR_CUDA(cudaDeviceSynchronize());
    ETR_CUDA(cudaFree(ctx->d_ws));
    ctx->d_ws = nullptr;
    ctx->ws_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&ctx->d_ws, want);
  if (e != cudaSuccess) {
    etr_set_error("etr_ws_reserve: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return ETR_ENOMEM;
  }
  ctx->ws_bytes = want;
  return ETR_OK;
}

extern "C" {

int etr_version(void) { return ETR_VERSION; }

const char* etr_last_error(void) { return g_err; }

int etr_ctx_create(int device, etr_ctx** out) {
  ETR_CHECK_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  int n = 0;
  ETR_CUDA(cudaGetDeviceCount(&n));
  ETR_CHECK_ARG(device >= 0 && device < n, "no such CUDA device");
  ETR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ETR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    etr_set_error("etr_ctx_create: device %d is sm_%d%d; libetr is built for sm_100a (B200) only",
                  device, prop.major, prop.minor);
    return ETR_EUNSUPPORTED;
  }
  etr_ctx* c = new (std::nothrow) etr_ctx();
  if (!c) return ETR_ENOMEM;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->d_ws = nullptr;
  c->ws_bytes = 0;
  c->launches = 0;
  c->n_shard_sets = 0;
  cudaError_t e = cudaMalloc(&c->d_err, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    delete c;
    etr_set_error("etr_ctx_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return ETR_ENOMEM;
  }
  cudaMemset(c->d_err, 0, 2 * sizeof(unsigned long long));
  c->side = nullptr; c->ev_fork = nullptr; c->ev_join = nullptr;
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    etr_set_error("etr_ctx_create: could not create the side stream / events");
    etr_ctx_destroy(c);
    return ETR_ECUDA;
  }
  *out = c;
  return ETR_OK;
}

int etr_ctx_destroy(etr_ctx* ctx) {
  if (!ctx) return ETR_OK;
  cudaSetDevice(ctx->device);
  if (ctx->d_ws) cudaFree(ctx->d_ws);
  if (ctx->d_err) cudaFree(ctx->d_err);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  delete ctx;
  return ETR_OK;
}

static int decode_error_word(const unsigned long long* h, int64_t* bad_id) {
  if (h[0] == 0) return ETR_OK;
  if (bad_id) *bad_id = (int64_t)h[1];
  switch (h[0]) {
    case 2:
      etr_set_error("sharded step: a request / gradient mailbox region overflowed (ids skewed in id mod G beyond the "
                    "mailbox capacity); rows of this step were dropped -- enlarge the capacity and redo the step");
      return ETR_EOVERFLOW;
    case 3:
      etr_set_error("sharded step: a peer barrier timed out (a rank did not arrive within ~20 s)");
      return ETR_ETIMEOUT;
    case 4:
      etr_set_error("sharded step: the owner's touched-row list overflowed; rows of this step were dropped");
      return ETR_EOVERFLOW;
    default:
      etr_set_error("embedding id %lld out of range (TF-CPU raises InvalidArgumentError here)", (long long)h[1]);
      return ETR_ERANGE;
  }
}

int etr_ctx_poll_error(etr_ctx* ctx, void* stream, int64_t* bad_id) {
  ETR_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long h[2] = {0, 0};
  ETR_CUDA(cudaMemcpyAsync(h, ctx->d_err, sizeof(h), cudaMemcpyDeviceToHost, s));
  ETR_CUDA(cudaStreamSynchronize(s));
  if (h[0] != 0) ETR_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(h), s));
  return decode_error_word(h, bad_id);
}

int etr_ctx_peek_error_async(etr_ctx* ctx, void* stream, uint64_t* h_pinned2) {
  ETR_CHECK_ARG(ctx != nullptr && h_pinned2 != nullptr, "NULL argument");
  ETR_CUDA(cudaMemcpyAsync(h_pinned2, ctx->d_err, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return ETR_OK;
}

int etr_ctx_decode_error(etr_ctx* ctx, const uint64_t* h_word2, void* stream, int64_t* bad_id) {
  ETR_CHECK_ARG(ctx != nullptr && h_word2 != nullptr, "NULL argument");
  const unsigned long long h[2] = {h_word2[0], h_word2[1]};
  if (h[0] != 0) ETR_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(h), (cudaStream_t)stream));
  return decode_error_word(h, bad_id);
}

int64_t etr_ctx_launch_count(etr_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
