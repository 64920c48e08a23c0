// What does HBM3e give for SCATTERED 256-byte records?  The fused FM apply (fm_tile_kernel) and the owner apply of the
// sharded step both read and re-write one 256-byte [var|m|v] record per unique row of the batch (~417 k - 471 k ascending
// rows out of 16.9 M - 33.8 M).  This program measures that access pattern alone: read-only, write-only and
// read-modify-write of n records picked from a table of T records, ascending or shuffled, L2 flushed before every launch.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/mb/mb_rmw scripts/mb/mb_rmw.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// half a warp per record: 16 lanes x 16 bytes
template <int MODE>   // 0 read, 1 write, 2 read-modify-write
__global__ void __launch_bounds__(256) rec_kernel(float4* tab, const int* idx, int n, float4* sink) {
  const int lane16 = threadIdx.x & 15;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4; r < n; r += ((long long)gridDim.x * blockDim.x) >> 4) {
    float4* p = tab + (long long)idx[r] * 16 + lane16;
    if (MODE == 0) { float4 x = *p; acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
    if (MODE == 1) { *p = make_float4((float)r, 1.f, 2.f, 3.f); }
    if (MODE == 2) { float4 x = *p; x.x = x.x * 0.999f + 1e-3f; x.y += 1.f; x.z *= 0.5f; x.w -= 1.f; *p = x; }
  }
  if (MODE == 0 && acc.x == 123.456f) *sink = acc;
}
__global__ void fill(float4* p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int main(int argc, char** argv) {
  const long long T = argc > 1 ? atoll(argv[1]) : 16880000;       // records in the table (4.3 GB)
  float4 *tab, *sink; int* d_idx; char* flush;
  CK(cudaMalloc(&tab, T * 256)); CK(cudaMalloc(&sink, 256)); CK(cudaMalloc(&flush, 512 << 20));
  fill<<<2048, 256>>>(tab, T * 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  std::mt19937_64 rng(7);
  for (int n : {417000, 1668000, 6672000}) {
    std::vector<int> idx(n);
    { std::vector<char> used(T, 0); int c = 0; while (c < n) { long long r = rng() % T; if (!used[r]) { used[r] = 1; idx[c++] = (int)r; } } }
    CK(cudaMalloc(&d_idx, n * 4));
    for (int order = 0; order < 2; ++order) {
      if (order == 0) std::sort(idx.begin(), idx.end()); else std::shuffle(idx.begin(), idx.end(), rng);
      CK(cudaMemcpy(d_idx, idx.data(), n * 4, cudaMemcpyHostToDevice));
      for (int mode = 0; mode < 3; ++mode) {
        for (int grid : {148 * 4, 148 * 8}) {
          float best = 1e9f, sum = 0.f; const int reps = 7;
          for (int it = 0; it < reps; ++it) {
            CK(cudaMemsetAsync(flush, it, 512 << 20));
            cudaEventRecord(a);
            if (mode == 0) rec_kernel<0><<<grid, 256>>>(tab, d_idx, n, sink);
            if (mode == 1) rec_kernel<1><<<grid, 256>>>(tab, d_idx, n, sink);
            if (mode == 2) rec_kernel<2><<<grid, 256>>>(tab, d_idx, n, sink);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it >= 2) { sum += ms; best = std::min(best, ms); }
          }
          const float us = sum / (reps - 2) * 1e3f;
          const double bytes = (double)n * 256 * (mode == 2 ? 2 : 1);
          printf("n=%8d %-8s %-5s grid=%5d: %7.1f us  %6.0f GB/s (%s)\n", n, order == 0 ? "sorted" : "shuffled",
                 mode == 0 ? "read" : mode == 1 ? "write" : "rmw", grid, us, bytes / us / 1e3, mode == 2 ? "read + write bytes" : "one way");
        }
      }
    }
    cudaFree(d_idx);
  }
  return 0;
}
