// cluster_occupancy.cu -- how many thread-block clusters of 1 / 2 / 4 / 8 CTAs the B200 keeps resident at once for a kernel with
// the NTT kernels' footprint (512 threads, 128 registers, 64 KiB .. 196 KiB of dynamic shared memory: one CTA per SM).
// A cluster lives inside one GPC; with one CTA per SM a GPC of n SMs holds floor(n / size) clusters, so sizes that do not
// divide the GPCs' SM counts leave SMs idle.   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy tools/cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(unsigned long long *out)
{
    extern __shared__ unsigned long long sm[];
    unsigned long long x[48];   // keeps the register count near the real kernels'
#pragma unroll
    for (int i = 0; i < 48; i++) x[i] = threadIdx.x * 48 + i;
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 48; i++) x[i] = x[i] * 6364136223846793005ull + x[(i + 1) % 48];
    sm[threadIdx.x] = x[0];
    __syncthreads();
    unsigned long long a = sm[(threadIdx.x + 1) & 511];
#pragma unroll
    for (int i = 0; i < 48; i++) a ^= x[i];
    out[blockIdx.x * 512 + threadIdx.x] = a;
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, p.multiProcessorCount);
    const int smem[2] = { 65536, 196608 + 16 };
    for (int s = 0; s < 2; s++) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem[s]);
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int c = 1; c <= 8; c *= 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(148 * 8);
            cfg.blockDim = dim3(512);
            cfg.dynamicSmemBytes = smem[s];
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c;
            at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf(", \"smem_%d_cluster_%d\": {\"max_active_clusters\": %d, \"resident_ctas\": %d%s}", smem[s], c, n, n * c, e == cudaSuccess ? "" : ", \"error\": true");
        }
    }
    printf("}\n");
    return 0;
}
