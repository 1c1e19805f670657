// imad_peak.cu -- integer-pipe microbenchmark for sm_100a ("IMAD peak denominator").
//
// Measures, register-resident and at the occupancy the NTT kernels run at (512 threads per SM and
// 1024 threads per SM), the issue rate of the instructions the 64-bit modular butterfly is built from:
//   mad.lo.u32 (IMAD), mad.hi.u32 (IMAD.HI.U32), mad.wide.u32 (IMAD.WIDE.U32), add.cc/addc (IADD3 / IADD3.X),
// and the throughput of complete lazy Cooley-Tukey butterflies in the formulations considered in DESIGN.md.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak tools/imad_peak.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;

#define ITERS 2048
#define ILP 8

template <int OP> __global__ void __launch_bounds__(512) k_op(u32 *out, u32 a0, u32 b0)
{
    u32 acc[ILP], a = a0 + threadIdx.x, b = b0 | 1;
    u64 wacc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x + i; wacc[i] = acc[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            // a multiplicand is always loop-carried data: ptxas hoists products of loop-invariant operands
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[i]) : "r"((u32)wacc[i]), "r"(b));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a));
            if (OP == 4) asm volatile("add.u64 %0, %0, %1;" : "+l"(wacc[i]) : "l"((u64)a << 20 | b));
            if (OP == 5) {   // 1 wide + 2 lo (the low 64 bits of a 64x64 product)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[i]) : "r"((u32)wacc[i]), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
            }
            if (OP == 6) {   // 1 lo + 1 add32 (do the pipes overlap?)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(*(u32 *)&wacc[i]) : "r"(a));
            }
            if (OP == 7) {   // 1 wide + 2 add32
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[i]) : "r"((u32)wacc[i]), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b));
            }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i] + (u32)wacc[i] + (u32)(wacc[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- complete butterflies: X = x + w*y, Y = x - w*y + 2q (lazy), w*y by Shoup ----
__device__ __forceinline__ u64 mulhi_exact(u64 a, u64 b) { return __umul64hi(a, b); }
// high word of a*b without the lo*lo partial product and without the low halves of the cross products: >= exact - 2
__device__ __forceinline__ u64 mulhi_approx_wide(u64 a, u64 b)
{
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 m1 = (u64)a1 * b0, m2 = (u64)a0 * b1;
    return (u64)a1 * b1 + (m1 >> 32) + (m2 >> 32);
}
__device__ __forceinline__ u64 mulhi_approx_hi(u64 a, u64 b)
{
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    return (u64)a1 * b1 + (u64)__umulhi(a1, b0) + (u64)__umulhi(a0, b1);
}
// two cross products summed first (one carry), then folded
__device__ __forceinline__ u64 mulhi_approx_sum(u64 a, u64 b)
{
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 m = (u64)a1 * b0;
    u64 m2 = (u64)a0 * b1;
    u64 s = m + m2;
    u64 c = s < m;
    return (u64)a1 * b1 + (s >> 32) + (c << 32);
}
template <int V> __device__ __forceinline__ u64 shoup_mad(u64 y, u64 w, u64 ws, u64 nq, u64 x)
{
    u64 h;
    if (V == 0) h = mulhi_exact(y, ws);
    if (V == 1) h = mulhi_approx_wide(y, ws);
    if (V == 2) h = mulhi_approx_hi(y, ws);
    if (V == 3) h = mulhi_approx_sum(y, ws);
    return y * w + h * nq + x;
}
template <int V> __global__ void __launch_bounds__(512) k_bfly(u64 *out, u64 w, u64 ws, u64 q)
{
    u64 x[16];
    const u64 nq = 0 - q, two_q = 2 * q;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + i;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 1 << s;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j & half) continue;
                u64 &a = x[j], &b = x[j + half];
                const u64 v = shoup_mad<V>(b, w + s, ws + s, nq, 0);
                b = a + two_q - v;
                a = a + v;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] &= 0x0fffffffffffffffull;   // keep the values bounded (1 LOP3 per 4 butterflies)
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same 4-stage pass with the twiddle pattern of the real kernels: 1, 2, 4, 8 distinct (w, w') pairs per stage,
// 15 per pass, fetched from (L1-resident) global memory every pass -- register pressure like the last NTT pass
__global__ void __launch_bounds__(512) k_bfly_tw(u64 *out, const ulonglong2 *__restrict__ tw, u64 q)
{
    u64 x[16];
    const u64 nq = 0 - q, two_q = 2 * q;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + i;
    for (int it = 0; it < ITERS / 4; it++) {
        ulonglong2 t[15];
#pragma unroll
        for (int i = 0; i < 15; i++) t[i] = __ldg(tw + ((threadIdx.x + it) & 31) * 16 + i);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 8 >> s;
#pragma unroll
            for (int blk = 0; blk < (1 << s); blk++) {
                const ulonglong2 w = t[(1 << s) - 1 + blk];
#pragma unroll
                for (int jj = 0; jj < half; jj++) {
                    u64 &a = x[blk * 2 * half + jj], &b = x[blk * 2 * half + jj + half];
                    const u64 v = shoup_mad<0>(b, w.x, w.y, nq, 0);
                    b = a + two_q - v;
                    a = a + v;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] &= 0x0fffffffffffffffull;
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 0: per-stage twiddle and the modulus constants in VECTOR registers (loaded from global memory, as the kernels
//         did before the constants moved to kernel parameters); 1: twiddle in vector registers, modulus constants
//         uniform; 2: twiddle uniform, modulus constants in vector registers
template <int MODE> __global__ void __launch_bounds__(512) k_bfly_vec(u64 *out, const ulonglong2 *__restrict__ tw, const u64 *__restrict__ mod, u64 qq)
{
    u64 x[16];
    const u64 q = MODE == 1 ? qq : mod[0];
    const u64 nq = 0 - q, two_q = 2 * q;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + i;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            ulonglong2 w;
            if (MODE == 2) w = make_ulonglong2(qq + s + it, qq * 3 + s);
            else w = __ldg(tw + ((threadIdx.x >> 5) + it + s) % 32);
            const int half = 1 << s;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j & half) continue;
                u64 &a = x[j], &b = x[j + half];
                const u64 v = shoup_mad<0>(b, w.x, w.y, nq, 0);
                b = a + two_q - v;
                a = a + v;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] &= 0x0fffffffffffffffull;
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static float time_ms(F &&f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, clk);
    u32 *out;
    cudaMalloc(&out, (size_t)sms * 8 * 512 * 8);
    const char *names[] = { "mad.lo.u32", "mad.hi.u32", "mad.wide.u32", "add.u32", "add.u64", "wide+2lo", "lo+add32", "wide+add32+xor" };
    const int per[] = { 1, 1, 1, 1, 1, 3, 2, 3 };
    for (int cps = 1; cps <= 2; cps++) {
        const int grid = sms * cps * 4;   // 4 waves
#define RUN(OP)                                                                                               \
        {                                                                                                     \
            float ms = time_ms([&] { k_op<OP><<<grid, 512>>>(out, 3, 5); });                                  \
            double inst = (double)grid * 512 * ITERS * ILP * per[OP];                                        \
            printf(" \"%s@%dthr/SM\": {\"ms\": %.4f, \"Ginst_per_s\": %.1f, \"lanes_per_clk_per_sm\": %.2f},\n", names[OP], 512 * cps, ms, \
                   inst / ms / 1e6, inst / (ms * 1e-3) / sms / (clk * 1e3));                                  \
        }
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    }
    const u64 q = 0xffffffffffe8001ull, w = 0x123456789abcdefull % q;
    const u64 ws = (u64)(((unsigned __int128)w << 64) / q);
    const char *vn[] = { "exact_mulhi", "approx_3wide", "approx_wide+2hi", "approx_sum" };
#define RUNB(V)                                                                                               \
    {                                                                                                         \
        const int grid = sms * 4;                                                                             \
        float ms = time_ms([&] { k_bfly<V><<<grid, 512>>>((u64 *)out, w, ws, q); });                          \
        double bf = (double)grid * 512 * (ITERS / 4) * 32;                                                    \
        printf(" \"bfly_%s\": {\"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", vn[V], ms, bf / ms / 1e6, \
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));                                              \
    }
    RUNB(0) RUNB(1) RUNB(2) RUNB(3)
    {
        ulonglong2 *tw;
        cudaMalloc(&tw, 32 * 16 * sizeof(ulonglong2));
        cudaMemset(tw, 0x5a, 32 * 16 * sizeof(ulonglong2));
        const int grid = sms * 4;
        float ms = time_ms([&] { k_bfly_tw<<<grid, 512>>>((u64 *)out, tw, q); });
        double bf = (double)grid * 512 * (ITERS / 4) * 32;
        u64 *mod;
        cudaMalloc(&mod, 64);
        cudaMemcpy(mod, &q, 8, cudaMemcpyHostToDevice);
#define RUNV(MODE, NAME)                                                                                      \
        {                                                                                                     \
            float ms2 = time_ms([&] { k_bfly_vec<MODE><<<grid, 512>>>((u64 *)out, tw, mod, q); });            \
            double bf2 = (double)grid * 512 * (ITERS / 4) * 32;                                               \
            printf(" \"" NAME "\": {\"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", ms2, bf2 / ms2 / 1e6, \
                   (ms2 * 1e-3) * (clk * 1e3) * sms * 4 / (bf2 / 32));                                        \
        }
        RUNV(0, "bfly_vector_twiddle_vector_modulus")
        RUNV(1, "bfly_vector_twiddle_uniform_modulus")
        RUNV(2, "bfly_uniform_twiddle_vector_modulus")
        printf(" \"bfly_exact_mulhi_15_twiddles_per_pass\": {\"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", ms, bf / ms / 1e6,
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));
    }
    printf(" \"note\": \"clock_khz is the attribute (max boost); lanes/clk uses it\"}\n");
    return 0;
}
