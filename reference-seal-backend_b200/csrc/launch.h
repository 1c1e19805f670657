// launch.h -- launch helpers and the entry points of the kernel families.
//
// The NTT-bearing kernels are templates over the CTA-local transform size and the cluster exponent; every family
// (transforms, key-switch inner product, mod-down) is instantiated in its own translation unit -- tu_ntt.cu, tu_ks.cu,
// tu_moddown.cu -- so the ~50 fully unrolled instances compile in parallel, and b200he.cu (host logic, the C ABI, the
// small elementwise kernels) reaches them through the launch_* functions below.  A launcher only enqueues; errors
// surface through cudaGetLastError() at the call site.
#pragma once
#include "kernels.cuh"

#ifndef B200HE_EMU
#define B200HE_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// launch with `cluster` consecutive CTAs per thread-block cluster (a limb split over 2^c CTAs, kernels.cuh)
template <class... KArgs, class... Args>
static inline void launch_cluster(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, unsigned cluster, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define B200HE_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) launch_cluster(kernel, (grid), (block), (smem), (stream), (cluster), __VA_ARGS__)
#endif

namespace b200he {

// geometry of a context's transforms: CTA-local transform 2^lognl, limb split over 2^c CTAs (c > 0 only with lognl = 13)
struct Geo {
    int lognl, c;
    cudaStream_t stream;
    // side stream + fork / join events of the context (nullptr: none): k_ks_inner's two launches for four-chunk limbs run next
    // to each other, so that the clusters of two of one fill the SMs that the clusters of four of the other cannot use
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

// dispatch on the CTA-local transform size LG (unsplit limbs)
#define NTT_DISPATCH(g, STMT)                                     \
    switch ((g).lognl) {                                          \
    case 10: { constexpr int LG = 10; STMT; } break;              \
    case 11: { constexpr int LG = 11; STMT; } break;              \
    case 12: { constexpr int LG = 12; STMT; } break;              \
    default: { constexpr int LG = 13; STMT; } break;              \
    }
// dispatch on LG and the cluster exponent CC (limbs larger than 8192 coefficients are split over 2 or 4 CTAs of 8192)
#define KERNEL_DISPATCH(g, STMT)                                                        \
    switch ((g).lognl * 4 + (g).c) {                                                    \
    case 40: { constexpr int LG = 10, CC = 0; STMT; } break;                            \
    case 44: { constexpr int LG = 11, CC = 0; STMT; } break;                            \
    case 48: { constexpr int LG = 12, CC = 0; STMT; } break;                            \
    case 52: { constexpr int LG = 13, CC = 0; STMT; } break;                            \
    case 53: { constexpr int LG = 13, CC = 1; STMT; } break;                            \
    default: { constexpr int LG = 13, CC = 2; STMT; } break;                            \
    }

// ---- tu_ntt.cu: K1 / K2
// forward transforms of nlimbs limbs; persistent_ctas > 0 selects the persistent variant (unsplit limbs only) with that grid
void launch_ntt_fwd(const Geo &g, const Tables &T, const u64 *src, u64 *dst, size_t src_outer, size_t dst_outer, int L, int mod_base, size_t nlimbs,
                    unsigned persistent_ctas);
// kind: KIND_INT / KIND_DP when the launch holds limbs of one kind only, KIND_BOTH otherwise
void launch_ntt_inv(const Geo &g, int kind, const Tables &T, const u64 *src, u64 *dst, size_t src_outer, size_t dst_outer, int L, int mod_base, int mode,
                    const InvFuse &F, size_t nlimbs);
int smem_attrs_ntt(const Geo &g);   // cudaError_t
// ---- tu_ks.cu: K6 step 2
void launch_ks_inner(const Geo &g, const Tables &T, const KsInnerArgs &A, size_t units);
int smem_attrs_ks(const Geo &g);
// ---- tu_moddown.cu: K6 step 3 / K9 (NTT form)
void launch_moddown_kind(const Geo &g, int kind, const Tables &T, const ModDownArgs &D, size_t units);
int smem_attrs_moddown(const Geo &g);

}   // namespace b200he
