// fp64_peak.cu -- does the FP64 pipe of B200 (sm_100a keeps the full-rate DFMA unit that sm_103a dropped) pay for
// modular butterflies?  Measures, register-resident at 512 threads per SM like tools/imad_peak.cu:
//   * issue rates of fma.rn.f64 / add.f64 / mul.f64, alone and interleaved with mad.wide.u32 / mad.lo.u32
//     (separate pipes: do they overlap?);
//   * "dp" butterfly: a complete lazy Cooley-Tukey butterfly for primes below 2^47 computed entirely in FP64
//     (values are exact integers in doubles; the product is split into (h, l) by an FMA, the quotient comes from a
//     precomputed w/q and the 1.5*2^52 rounding constant): 8 FP64 instructions, no integer multiply;
//   * "hyb" butterfly for 60-bit primes: the Shoup quotient floor(y*w/q) is estimated on the FP64 pipe from the two
//     32-bit halves of y (twiddle kept as w and w*2^32 mod q), the two low products and qhat*q stay on the integer pipe:
//     3 IMAD.WIDE + 3 IMAD + 4 FP64 instead of 5-6 IMAD.WIDE + 4 IMAD.
// Every butterfly variant is validated against exact 128-bit host arithmetic (residues compared mod q).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

#define ITERS 2048
#define ILP 8
#define MAGIC 6755399441055744.0      /* 1.5 * 2^52 */
#define TWO52 4503599627370496.0

template <int OP> __global__ void __launch_bounds__(512) k_dop(double *out, double a0, double b0, u32 ib)
{
    double acc[ILP];
    u64 wacc[ILP];
    u32 iacc[ILP];
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x + i; wacc[i] = threadIdx.x + i; iacc[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0 || OP >= 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(b), "d"(a));
            if (OP == 1) asm volatile("add.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(a));
            if (OP == 2) asm volatile("mul.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(b));
            if (OP == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[i]) : "r"((u32)wacc[i]), "r"(ib));
            if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(iacc[i]) : "r"(ib), "r"(ib + 7));
            if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(iacc[i]) : "r"(ib));
            if (OP == 6) {   // 1 DFMA : 2 wide (the hybrid butterfly's ratio is 4 : 3 + 3 lo)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[i]) : "r"((u32)wacc[i]), "r"(ib));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(iacc[i]) : "r"(ib), "r"(ib + 7));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i] + (double)wacc[i] + (double)iacc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- pure-FP64 butterfly (q < 2^47) ----------------
// w*y mod q, symmetric lazy result: |v| <= 0.75 q for |y| < 2^51
__device__ __forceinline__ double dp_mulmod(double y, double w, double wq, double nq)
{
    const double h = __dmul_rn(y, w);
    const double l = __fma_rn(y, w, -h);          // y*w = h + l exactly
    const double t = __fma_rn(y, wq, MAGIC);      // MAGIC + rint(y * w/q)
    const double qh = __dadd_rn(t, -MAGIC);
    const double d = __fma_rn(qh, nq, h);         // exact: |h - qh q| < 2^53
    return __dadd_rn(d, l);
}
// x -> x - rint(x/q) q, |result| <= q/2 (+1)
__device__ __forceinline__ double dp_reduce(double x, double qinv, double nq)
{
    const double t = __fma_rn(x, qinv, MAGIC);
    return __fma_rn(__dadd_rn(t, -MAGIC), nq, x);
}
// lazy symmetric value -> canonical integer in [0, q)
__device__ __forceinline__ u64 dp_canon(double x, double q, double qinv_up, double nq)
{
    const double a = __dadd_rn(x, 64.0 * q);                                  // positive; exact
    const double t = __fma_rd(a, qinv_up, MAGIC);                             // MAGIC + floor(a/q) exactly (qinv rounded up)
    const double r = __fma_rn(__dadd_rn(t, -MAGIC), nq, __dadd_rn(a, TWO52)); // 2^52 + (a mod q)
    return (u64)__double_as_longlong(r) & 0x000fffffffffffffull;
}
struct DpTw { double w[4], wq[4]; };
__global__ void __launch_bounds__(512) k_bfly_dp(u64 *out, DpTw tw, double q, double qinv, double qinv_up, int iters, int dump)
{
    double x[16];
    const double nq = -q;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const u64 v = ((u64)(threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + (u64)i * 0xD1B54A32D192ED03ull) >> 20;   // 44 bits
        x[i] = __longlong_as_double(v | 0x4330000000000000ull) - TWO52;
        x[i] = dp_reduce(x[i], qinv, nq);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 1 << s;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j & half) continue;
                double &a = x[j], &b = x[j + half];
                const double v = dp_mulmod(b, tw.w[s], tw.wq[s], nq);
                b = __dadd_rn(a, -v);
                a = __dadd_rn(a, v);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = dp_reduce(x[i], qinv, nq);   // stands for the once-per-pass housekeeping
    }
    if (dump) {
#pragma unroll
        for (int i = 0; i < 16; i++) out[(size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i] = dp_canon(x[i], q, qinv_up, nq);
    } else {
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += x[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (u64)__double_as_longlong(s);
    }
}

// ---------------- hybrid butterfly (60-bit q): FP64 quotient, integer products ----------------
struct HybTw { u64 w[4], w2[4]; double wq[4], w2q[4]; };
// returns V = (w*y mod q) + 2q + e*q, |e| <= 1.01: V in (0.99q, 3.01q), for ANY 64-bit y
__device__ __forceinline__ u64 hyb_mul(u64 y, u64 w, u64 w2, double wq, double w2q, u64 nq, u64 two_q)
{
    const u32 y0 = (u32)y, y1 = (u32)(y >> 32);
    const double d0 = __hiloint2double(0x43300000, y0) - TWO52;
    const double d1 = __hiloint2double(0x43300000, y1) - TWO52;
    double t = __fma_rn(d0, wq, MAGIC);
    t = __fma_rn(d1, w2q, t);                                        // MAGIC + qhat, qhat < 2^33
    const u64 qhat = (u64)(u32)__double2loint(t) | ((u64)(__double2hiint(t) & 1) << 32);
    return (u64)y0 * w + (u64)y1 * w2 + qhat * nq + two_q;           // low 64 bits
}
template <int MASK> __global__ void __launch_bounds__(512) k_bfly_hyb(u64 *out, HybTw tw, u64 q, int iters, int dump)
{
    u64 x[16];
    const u64 nq = 0 - q, two_q = 2 * q, four_q = 4 * q;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (((u64)(threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + (u64)i * 0xD1B54A32D192ED03ull) >> 5) % q;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 1 << s;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j & half) continue;
                u64 &a = x[j], &b = x[j + half];
                const u64 v = hyb_mul(b, tw.w[s], tw.w2[s], tw.wq[s], tw.w2q[s], nq, two_q);
                b = a + four_q - v;
                a = a + v;
            }
        }
        if (MASK) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] &= 0x0fffffffffffffffull;
        }
    }
    if (dump) {
#pragma unroll
        for (int i = 0; i < 16; i++) out[(size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i] = x[i] % q;
    } else {
        u64 s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s ^= x[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}
// integer Shoup reference butterfly (tools/imad_peak.cu's exact_mulhi) on the same inputs, for an in-run comparison
__global__ void __launch_bounds__(512) k_bfly_int(u64 *out, HybTw tw, u64 q, int iters)
{
    u64 x[16];
    const u64 nq = 0 - q, two_q = 2 * q;
    u64 ws[4];
#pragma unroll
    for (int s = 0; s < 4; s++) ws[s] = tw.w2[s];   // any 64-bit constant: timing only
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int half = 1 << s;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j & half) continue;
                u64 &a = x[j], &b = x[j + half];
                const u64 v = b * tw.w[s] + __umul64hi(b, ws[s]) * nq;
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

static u64 mulmod_h(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, clk);
    u64 *out;
    cudaMalloc(&out, (size_t)sms * 4 * 512 * 16 * 8);
    const char *names[] = { "fma.f64", "add.f64", "mul.f64", "fma.f64+mad.wide", "fma.f64+mad.lo", "fma.f64+add.u32", "fma.f64+mad.wide+mad.lo" };
    const int grid = sms * 4;
#define RUN(OP)                                                                                                   \
    {                                                                                                             \
        float ms = time_ms([&] { k_dop<OP><<<grid, 512>>>((double *)out, 1.0000001, 0.9999999, 12345u); });       \
        double inst = (double)grid * 512 * ITERS * ILP;                                                           \
        printf(" \"%s\": {\"ms\": %.4f, \"fp64_lanes_per_clk_per_sm\": %.2f, \"cycles_per_warp_group_per_smsp\": %.2f},\n", names[OP], ms, \
               inst / (ms * 1e-3) / sms / (clk * 1e3), (ms * 1e-3) * (clk * 1e3) * sms * 4 / (inst / 32));        \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)

    // ---- pure FP64 butterfly, q = 45-bit SEAL prime ----
    {
        const u64 q = 0x1ffffff8c001ull;
        DpTw tw;
        u64 wi[4];
        for (int s = 0; s < 4; s++) {
            wi[s] = (0x123456789abcdefull * (s + 3)) % q;
            tw.w[s] = (double)wi[s];
            tw.wq[s] = (double)((long double)wi[s] / (long double)q);
        }
        const double qd = (double)q, qinv = 1.0 / qd;
        double qinv_up = qinv;
        if ((long double)qinv_up * (long double)q < 1.0L) qinv_up = __builtin_nextafter(qinv_up, 2.0);
        // validation: block 0, 8 iterations
        const int vit = 8;
        k_bfly_dp<<<1, 512>>>(out, tw, qd, qinv, qinv_up, vit, 1);
        std::vector<u64> got(512 * 16);
        cudaMemcpy(got.data(), out, got.size() * 8, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int t = 0; t < 512; t++) {
            u64 x[16];
            for (int i = 0; i < 16; i++) x[i] = ((((u64)(t + 1) * 0x9E3779B97F4A7C15ull + (u64)i * 0xD1B54A32D192ED03ull) >> 20)) % q;
            for (int it = 0; it < vit; it++)
                for (int s = 0; s < 4; s++) {
                    const int half = 1 << s;
                    for (int j = 0; j < 16; j++) {
                        if (j & half) continue;
                        const u64 v = mulmod_h(x[j + half], wi[s], q), a = x[j];
                        x[j] = (a + v) % q;
                        x[j + half] = (a + q - v) % q;
                    }
                }
            for (int i = 0; i < 16; i++) bad += got[t * 16 + i] != x[i];
        }
        float ms = time_ms([&] { k_bfly_dp<<<grid, 512>>>(out, tw, qd, qinv, qinv_up, ITERS / 4, 0); });
        double bf = (double)grid * 512 * (ITERS / 4) * 32;
        printf(" \"bfly_dp_45bit\": {\"mismatches\": %ld, \"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f, \"note\": \"includes one 3-op reduction per value per 4 stages\"},\n",
               bad, ms, bf / ms / 1e6, (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));
    }
    // ---- hybrid butterfly, q = 60-bit SEAL prime ----
    {
        const u64 q = 0xffffffffffe8001ull;
        HybTw tw;
        for (int s = 0; s < 4; s++) {
            tw.w[s] = (0x123456789abcdefull * (s + 3)) % q;
            tw.w2[s] = (u64)(((u128)tw.w[s] << 32) % q);
            tw.wq[s] = (double)((long double)tw.w[s] / (long double)q);
            tw.w2q[s] = (double)((long double)tw.w2[s] / (long double)q);
        }
        k_bfly_hyb<0><<<1, 512>>>(out, tw, q, 1, 1);   // one 4-stage pass from canonical inputs, no masking
        std::vector<u64> got(512 * 16);
        cudaMemcpy(got.data(), out, got.size() * 8, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int t = 0; t < 512; t++) {
            u64 x[16];
            for (int i = 0; i < 16; i++) x[i] = ((((u64)(t + 1) * 0x9E3779B97F4A7C15ull + (u64)i * 0xD1B54A32D192ED03ull) >> 5)) % q;
            for (int s = 0; s < 4; s++) {
                const int half = 1 << s;
                for (int j = 0; j < 16; j++) {
                    if (j & half) continue;
                    const u64 v = mulmod_h(x[j + half], tw.w[s], q), a = x[j];
                    x[j] = (a + v) % q;
                    x[j + half] = (a + q - v) % q;
                }
            }
            for (int i = 0; i < 16; i++) bad += got[t * 16 + i] != x[i];
        }
        float ms = time_ms([&] { k_bfly_hyb<1><<<grid, 512>>>(out, tw, q, ITERS / 4, 0); });
        double bf = (double)grid * 512 * (ITERS / 4) * 32;
        printf(" \"bfly_hybrid_60bit\": {\"mismatches\": %ld, \"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", bad, ms, bf / ms / 1e6,
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));
        ms = time_ms([&] { k_bfly_int<<<grid, 512>>>(out, tw, q, ITERS / 4); });
        printf(" \"bfly_int_shoup_60bit\": {\"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", ms, bf / ms / 1e6,
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));
    }
    // ---- the same butterflies at the occupancy of the NTT kernels: ONE 512-thread CTA per SM (4 warps per scheduler),
    // forced by a 200 KB dynamic shared-memory request; the rows above run 4 CTAs per SM (16 warps per scheduler)
    {
        const int smem = 200 * 1024;
        cudaFuncSetAttribute(k_bfly_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_bfly_hyb<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_bfly_int, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const u64 q45 = 0x1ffffff8c001ull, q60 = 0xffffffffffe8001ull;
        DpTw dt;
        HybTw ht;
        for (int s = 0; s < 4; s++) {
            const u64 w = (0x123456789abcdefull * (s + 3)) % q45;
            dt.w[s] = (double)w;
            dt.wq[s] = (double)((long double)w / (long double)q45);
            ht.w[s] = (0x123456789abcdefull * (s + 3)) % q60;
            ht.w2[s] = (u64)(((u128)ht.w[s] << 32) % q60);
            ht.wq[s] = (double)((long double)ht.w[s] / (long double)q60);
            ht.w2q[s] = (double)((long double)ht.w2[s] / (long double)q60);
        }
        const double qd = (double)q45, qinv = 1.0 / qd;
        const double bf = (double)grid * 512 * (ITERS / 4) * 32;
        float ms = time_ms([&] { k_bfly_dp<<<grid, 512, smem>>>(out, dt, qd, qinv, qinv, ITERS / 4, 0); });
        printf(" \"bfly_dp_45bit_one_cta_per_sm\": {\"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", ms, bf / ms / 1e6,
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));
        ms = time_ms([&] { k_bfly_int<<<grid, 512, smem>>>(out, ht, q60, ITERS / 4); });
        printf(" \"bfly_int_shoup_60bit_one_cta_per_sm\": {\"ms\": %.4f, \"Gbfly_per_s\": %.1f, \"cycles_per_warp_bfly_per_smsp\": %.2f},\n", ms, bf / ms / 1e6,
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / (bf / 32));
    }
    printf(" \"note\": \"clock_khz is the attribute (max boost)\"}\n");
    return 0;
}
