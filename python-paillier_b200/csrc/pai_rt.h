// pai_rt.h -- the few runtime services the engine needs (memory, copies, kernel launch).
//
// Product build (nvcc): thin wrappers over the CUDA runtime; kernels are launched as
//   k_body<Body><<<grid, nthr, smem, stream>>>(body)
// Test-only build (-DPAI_HOSTSIM, g++): the same calls run on the CPU with plain malloc/memcpy and a
// loop nest over (cta, tid).  That build exists so that the orchestration code in pai_engine.cu
// (context creation, constant assembly, workspace sizing, multi-kernel ops) can be exercised in the
// GPU-less build container.  It is compiled into tests/hostsim/ only and is never loaded by the
// product package.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>
#if !defined(PAI_HOSTSIM)
#include <cuda_runtime.h>
#endif
#include "../../include/paillier_b200.h"
#include "pai_cta.cuh"

namespace pai {

static std::atomic<long> g_launches{0};
static thread_local std::string g_err;

#if !defined(PAI_HOSTSIM)
// ------------------------------------------------------------------------------------ CUDA
typedef cudaStream_t rt_stream;

#define RT_CHECK(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e_);                          \
      return PAI_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

static inline int rt_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
static inline int rt_set_device(int dev) { RT_CHECK(cudaSetDevice(dev)); return 0; }
// restores the caller's current device when an entry point returns (contexts pin their own device; a host
// program that also uses torch or other CUDA libraries must not see its current device change under it)
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
static inline int rt_malloc(void** p, size_t bytes) { RT_CHECK(cudaMalloc(p, bytes ? bytes : 16)); return 0; }
static inline void rt_free(void* p) { if (p) cudaFree(p); }
static inline int rt_h2d(void* d, const void* h, size_t n, rt_stream s) { RT_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); return 0; }
static inline int rt_d2h(void* h, const void* d, size_t n, rt_stream s) { RT_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); return 0; }
static inline int rt_d2d(void* d, const void* s_, size_t n, rt_stream s) { RT_CHECK(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); return 0; }
static inline int rt_memset(void* d, int v, size_t n, rt_stream s) { RT_CHECK(cudaMemsetAsync(d, v, n, s)); return 0; }
static inline int rt_sync(rt_stream s) { RT_CHECK(cudaStreamSynchronize(s)); return 0; }
static inline int rt_sm_count(int dev) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n > 0 ? n : 1; }
static inline size_t rt_max_smem(int dev) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); return (size_t)n; }

// threads per CTA a body may be launched with (sets the register budget: 256 -> 255 registers, 512 -> 128); bodies that
// want wider CTAs specialise this (pai_engine.cu: the tensor-core kernels of small digit moduli run 4 groups of 128)
template <class Body>
struct BodyMaxThreads { static const int v = 256; };

template <class Body>
__global__ void __launch_bounds__(BodyMaxThreads<Body>::v) k_body(Body b) {
  extern __shared__ u4 smem[];
  CtaId id{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  cta_load_consts(smem, id, b.consts, b.const_quads);
  __syncthreads();
  b.run(smem, id);
}

// resident CTAs per SM for this body at (nthr, smem)
template <class Body>
static inline int rt_occupancy(int nthr, size_t smem) {
  cudaFuncSetAttribute(k_body<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_body<Body>, nthr, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  return occ;
}
template <class Body>
static inline int rt_launch(const Body& b, int grid, int nthr, size_t smem, rt_stream s) {
  RT_CHECK(cudaFuncSetAttribute(k_body<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_body<Body><<<grid, nthr, smem, s>>>(b);
  g_launches++;
  RT_CHECK(cudaGetLastError());
  return 0;
}

// warp-cooperative bodies (pai_coop.cuh): every thread runs the body, lanes cooperate through shuffles
template <class Body>
static inline int rt_launch_coop(const Body& b, int grid, int nthr, size_t smem, rt_stream s) { return rt_launch(b, grid, nthr, smem, s); }
// group-cooperative bodies (pai_tc.cuh): every thread runs the body, 128-thread groups cooperate through the tensor cores
template <class Body>
static inline int rt_launch_group(const Body& b, int grid, int nthr, size_t smem, rt_stream s) { return rt_launch(b, grid, nthr, smem, s); }

#else
// ------------------------------------------------------------------------------------ host simulation (tests only)
typedef void* rt_stream;
#define RT_CHECK(expr) do { if ((expr) != 0) return PAI_E_CUDA; } while (0)
static inline int rt_device_count() { return 1; }
static inline int rt_set_device(int) { return 0; }
struct DeviceGuard {};
static inline int rt_malloc(void** p, size_t bytes) { *p = aligned_alloc(64, ((bytes ? bytes : 16) + 63) / 64 * 64); return *p ? 0 : -2; }
static inline void rt_free(void* p) { free(p); }
static inline int rt_h2d(void* d, const void* h, size_t n, rt_stream) { memcpy(d, h, n); return 0; }
static inline int rt_d2h(void* h, const void* d, size_t n, rt_stream) { memcpy(h, d, n); return 0; }
static inline int rt_d2d(void* d, const void* s_, size_t n, rt_stream) { memcpy(d, s_, n); return 0; }
static inline int rt_memset(void* d, int v, size_t n, rt_stream) { memset(d, v, n); return 0; }
static inline int rt_sync(rt_stream) { return 0; }
static inline int rt_sm_count(int) { return 2; }
static inline size_t rt_max_smem(int) { return 232448; }
template <class Body>
static inline int rt_occupancy(int, size_t smem) { return smem <= 232448 ? 1 : 0; }
template <class Body>
static inline int rt_launch(const Body& b, int grid, int nthr, size_t smem, rt_stream) {
  std::vector<u4> sm(smem / 16 + 1);
  for (int cta = 0; cta < grid; cta++) {
    for (int tid = 0; tid < nthr; tid++) { CtaId id{tid, nthr, cta, grid}; cta_load_consts(sm.data(), id, b.consts, b.const_quads); }
    for (int tid = 0; tid < nthr; tid++) { CtaId id{tid, nthr, cta, grid}; b.run(sm.data(), id); }
  }
  g_launches++;
  return 0;
}
// warp-cooperative bodies: one call per warp walks its 32 lanes in lockstep (lane arrays, pai_coop.cuh)
template <class Body>
static inline int rt_launch_coop(const Body& b, int grid, int nthr, size_t smem, rt_stream) {
  std::vector<u4> sm(smem / 16 + 1);
  for (int cta = 0; cta < grid; cta++)
    for (int warp = 0; warp < nthr / 32; warp++) { CtaId id{warp * 32, nthr, cta, grid}; b.run(sm.data(), id); }
  g_launches++;
  return 0;
}
// group-cooperative bodies: one call per CTA walks the rows of its group phase by phase (row arrays, pai_tc.cuh)
template <class Body>
static inline int rt_launch_group(const Body& b, int grid, int nthr, size_t smem, rt_stream) {
  std::vector<u4> sm(smem / 16 + 16);
  for (int cta = 0; cta < grid; cta++) {
    for (int tid = 0; tid < nthr; tid++) { CtaId id{tid, nthr, cta, grid}; cta_load_consts(sm.data(), id, b.consts, b.const_quads); }
    CtaId id{0, nthr, cta, grid};
    b.run(sm.data(), id);
  }
  g_launches++;
  return 0;
}
#endif

}  // namespace pai
