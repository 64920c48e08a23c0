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

}  // namespace etr

using namespace etr;

extern "C" {

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
