// xchg_ab.cu -- A/B of the innermost register exchange of the CTA-local NTT (N = 8192): between the third and the fourth
// pass, groups of 8 threads transpose an 8 x 8 matrix of coefficient PAIRS (16 bytes each); thread t holds column pair t
// of all 8 rows and needs row t of all 8 column pairs (ntt_core.cuh: Pass<13,2> -> Pass<13,3>).
//   A  shared memory, as in the tree: 8 STS.128 + __syncwarp + 8 LDS.128 per thread (XOR-swizzled, conflict free)
//   B  warp shuffles (the north star's wording): 3 butterfly stages over lane ^ 1, ^ 2, ^ 4; a stage moves half of the
//      thread's pairs = 4 pairs x 4 words = 16 SHFL.BFLY, plus 32 selects
// Same occupancy as the transforms (one 512-thread CTA per SM, 64 KiB of shared memory reserved, 128 registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_ab tools/xchg_ab.cu && ./xchg_ab
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64;

__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 7) << 1); }

// light, data-dependent work between exchanges so that nothing can be hoisted or eliminated
__device__ __forceinline__ void mix(u64 (&x)[16], u64 k)
{
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = x[i] * 0x9E3779B97F4A7C15ull + k + i;
}

template <int MODE> __global__ void   // 0: shared memory, 1: warp shuffles, 2: no exchange (the mixing alone)
 __launch_bounds__(512, 1) k_xchg(u64 *out, int rounds)
{
    extern __shared__ __align__(16) unsigned char raw[];
    u64 *sm = reinterpret_cast<u64 *>(raw);
    const int tid = threadIdx.x, t8 = tid & 7, blk = (tid >> 3) * 128;
    u64 x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (u64)tid * 16 + i + blockIdx.x;
    for (int r = 0; r < rounds; r++) {
        mix(x, r);
        if (MODE == 0) {
            // register pair (2j, 2j+1) = row j, column pair t8  ->  row-major 8 x 16 block in shared memory
#pragma unroll
            for (int j = 0; j < 8; j++) *reinterpret_cast<ulonglong2 *>(sm + swz(blk + j * 16 + t8 * 2)) = make_ulonglong2(x[2 * j], x[2 * j + 1]);
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 8; p++) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + swz(blk + t8 * 16 + p * 2));
                x[2 * p] = v.x;
                x[2 * p + 1] = v.y;
            }
            __syncwarp();
        } else if (MODE == 1) {
            // 8 x 8 transpose of pairs over the lanes of an aligned group of 8: stage b swaps the off-diagonal 2^b blocks
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const int bit = 1 << b;
                const bool up = (t8 & bit) != 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (j & bit) continue;   // handle (j, j | bit) together
                    // lane without the bit sends row j|bit and keeps row j; lane with the bit sends row j and keeps row j|bit
                    u64 s0 = up ? x[2 * j] : x[2 * (j | bit)], s1 = up ? x[2 * j + 1] : x[2 * (j | bit) + 1];
                    s0 = __shfl_xor_sync(0xffffffffu, s0, bit);
                    s1 = __shfl_xor_sync(0xffffffffu, s1, bit);
                    if (up) { x[2 * j] = s0; x[2 * j + 1] = s1; }
                    else { x[2 * (j | bit)] = s0; x[2 * (j | bit) + 1] = s1; }
                }
            }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    out[(size_t)blockIdx.x * 512 + tid] = acc;
}

int main()
{
    const int rounds = 4096, grid = 148 * 4;
    u64 *out;
    cudaMalloc(&out, (size_t)grid * 512 * 8);
    cudaFuncSetAttribute(k_xchg<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_xchg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_xchg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    {   // both variants implement the same permutation: identical outputs after a few rounds
        u64 *o2, *h1 = new u64[(size_t)grid * 512], *h2 = new u64[(size_t)grid * 512];
        cudaMalloc(&o2, (size_t)grid * 512 * 8);
        k_xchg<0><<<grid, 512, 65536>>>(out, 3);
        k_xchg<1><<<grid, 512, 65536>>>(o2, 3);
        cudaMemcpy(h1, out, (size_t)grid * 512 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(h2, o2, (size_t)grid * 512 * 8, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < (size_t)grid * 512; i++)
            if (h1[i] != h2[i]) { printf("{\"error\": \"variants disagree at %zu\"}\n", i); return 1; }
        cudaFree(o2);
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms[3] = { 0, 0, 0 };
    for (int rep = 0; rep < 3; rep++)
        for (int v = 0; v < 3; v++) {
            cudaEventRecord(a);
            if (v == 0) k_xchg<0><<<grid, 512, 65536>>>(out, rounds);
            else if (v == 1) k_xchg<1><<<grid, 512, 65536>>>(out, rounds);
            else k_xchg<2><<<grid, 512, 65536>>>(out, rounds);   // the register mixing alone
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float t;
            cudaEventElapsedTime(&t, a, b);
            if (rep == 2) ms[v] = t;
        }
    if (cudaGetLastError() != cudaSuccess) { printf("{\"error\": \"launch failed\"}\n"); return 1; }
    // one exchange of one CTA, in SM cycles at 1.965 GHz (4 CTAs per SM run back to back)
    const double per = 1.965e9 / 1e3 / (4.0 * rounds);
    printf("{\"exchange\": \"8-lane transpose of 16 x u64 per thread, 512-thread CTA, one CTA per SM\", \"rounds\": %d, "
           "\"shared_memory_ms\": %.3f, \"warp_shuffle_ms\": %.3f, \"cycles_per_exchange_shared_memory\": %.0f, \"cycles_per_exchange_warp_shuffle\": %.0f, "
           "\"mixing_alone_ms\": %.3f, \"note\": \"cycles = (variant - mixing alone) per exchange of one CTA\"}\n",
           rounds, ms[0], ms[1], (ms[0] - ms[2]) * per, (ms[1] - ms[2]) * per, ms[2]);
    return 0;
}
