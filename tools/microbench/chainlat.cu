// chainlat.cu — how many cycles does ONE link of the block-hash chain cost a lone warp?
// (chain_finalize_kernel measured ~200 cycles per link although ptxas' stall counts add up to ~115.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fusioninfer_b200/csrc -o tools/microbench/chainlat tools/microbench/chainlat.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "xxh64.cuh"
using namespace fi;

// A: registers only — the bare dependency chain
__global__ void pure_chain(uint64_t* out, int links, long long* cyc) {
  uint64_t h = threadIdx.x * 0x9E3779B97F4A7C15ull + blockIdx.x;
  uint64_t pre = h ^ 0x1234567ull;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < links; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) h = chain_step(pre + k, h);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = h;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// B: two independent chains per thread (does ILP help, i.e. is the chain latency- or issue-bound?)
__global__ void two_chains(uint64_t* out, int links, long long* cyc) {
  uint64_t h = threadIdx.x * 0x9E3779B97F4A7C15ull + blockIdx.x, g = ~h;
  uint64_t pre = h ^ 0x1234567ull;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < links; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      h = chain_step(pre + k, h);
      g = chain_step(pre - k, g);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = h ^ g;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// C: the same link written on 32-bit halves by hand, with the multiplies by constants spelled as mul.lo / mad
__device__ __forceinline__ uint64_t mul64c(uint64_t x, uint64_t c) {
  const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), cl = (uint32_t)c, ch = (uint32_t)(c >> 32);
  const uint64_t lo = (uint64_t)xl * cl;
  const uint32_t hi = (uint32_t)(lo >> 32) + xl * ch + xh * cl;
  return ((uint64_t)hi << 32) | (uint32_t)lo;
}
__device__ __forceinline__ uint64_t chain_step_manual(uint64_t pre, uint64_t prev) {
  uint64_t r = mul64c(prev, XP2);
  r = (r << 31) | (r >> 33);
  r = mul64c(r, XP1);
  uint64_t h = pre ^ r;
  h = (h << 27) | (h >> 37);
  h = mul64c(h, XP1) + XP4;
  h ^= h >> 33;
  h = mul64c(h, XP2);
  h ^= h >> 29;
  h = mul64c(h, XP3);
  h ^= h >> 32;
  return h;
}
__global__ void manual_chain(uint64_t* out, int links, long long* cyc) {
  uint64_t h = threadIdx.x * 0x9E3779B97F4A7C15ull + blockIdx.x;
  uint64_t pre = h ^ 0x1234567ull;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < links; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) h = chain_step_manual(pre + k, h);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = h;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
void run(const char* name, K kern, int grid, int block, int links) {
  uint64_t* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(uint64_t) * grid * block);
  cudaMallocManaged(&cyc, sizeof(long long) * grid);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  kern<<<grid, block>>>(out, links, cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  kern<<<grid, block>>>(out, links, cyc);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
  printf("%-14s grid %4d x %4d threads, %d links: %.1f cycles/link (clock64), %.2f us wall -> %.1f ns/link\n", name, grid, block,
         links, (double)mx / links, ms * 1e3, ms * 1e6 / links);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  const int links = 4096;
  for (int block : {32, 128, 256, 512}) {
    run("pure_chain", pure_chain, 1, block, links);
    run("pure_chain", pure_chain, 128, block, links);
  }
  run("two_chains", two_chains, 1, 32, links);
  run("two_chains", two_chains, 128, 128, links);
  run("manual_chain", manual_chain, 1, 32, links);
  run("manual_chain", manual_chain, 128, 128, links);
  return 0;
}
