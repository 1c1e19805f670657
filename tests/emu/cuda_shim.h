// cuda_shim.h -- TEST INFRASTRUCTURE: compiles the product's CUDA sources as host C++ so the
// kernels' index arithmetic can be checked against the oracle in a container without a GPU.
//
// Not a CPU fallback: it is only ever compiled into tests/emu/libb200he_emu.so by
// tests/emu/build_emu.sh with -DB200HE_EMU; the shipped library (libb200he.so) is built by nvcc
// without that macro and the Python binding refuses to load anything else.
//
// A thread block is emulated by one fiber (ucontext) per CUDA thread on a single OS thread;
// __syncthreads() yields to a round-robin scheduler, so barrier semantics are exact for
// well-formed kernels.  Blocks of a grid run one after another.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{ x, y }; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
unsigned char *block_smem();
unsigned char *cluster_smem(unsigned rank);   // shared memory of block `rank` of the running cluster (DSMEM)
unsigned cluster_rank();
void sync();
void cluster_sync();
void run_grid(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body, unsigned cluster = 1);
}
#define __syncthreads() emu::sync()

// ---- the slice of the CUDA runtime API the library uses ----
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }

// ---- FP64 intrinsics used by the FP64-domain transforms (modarith.cuh); host doubles are IEEE, fma() is exact ----
#include <fenv.h>
#include <math.h>
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
__attribute__((noinline)) static double __fma_rd(double a, double b, double c)
{
    const int old = fegetround();
    fesetround(FE_DOWNWARD);
    volatile double va = a, vb = b, vc = c;
    volatile double r = fma(va, vb, vc);
    fesetround(old);
    return r;
}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }

#define B200HE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::run_grid((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define B200HE_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) \
    emu::run_grid((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); }, (unsigned)(cluster))
