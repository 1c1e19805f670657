// tu_moddown.cu -- instantiations and launcher of the NTT-form mod-down kernels k_moddown / k_moddown_dp / k_moddown_mix
// (K6 step 3, K9).
#include "launch.h"

namespace b200he {

void launch_moddown_kind(const Geo &g, int kind, const Tables &T, const ModDownArgs &D, size_t units)
{
    const unsigned grid = (unsigned)(units << g.c);
    // clusters of two exchange nothing in the forward direction (load_fwd_split computes their cross stage from global
    // memory): the two CTAs of a limb are launched as ordinary CTAs, free to land on any SM
    // limbs of four chunks: clusters of TWO (the stage that pairs chunk r with r ^ 2 also comes from global memory, the
    // other one is exchanged inside the pair; clusters of four would keep only 132 of the 148 SMs busy)
    const unsigned cl = g.c == 1 ? 1u : g.c == 2 ? 2u : 1u;
    if (kind == KIND_INT) {
        KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_moddown<LG, CC>), grid, NttCfg<LG>::THREADS, (D.gal ? ModDownCfg<LG>::SMEM_BYTES_GAL : ModDownCfg<LG>::SMEM_BYTES), g.stream, cl, T, D));
    } else if (kind == KIND_DP) {
        KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_moddown_dp<LG, CC>), grid, NttCfg<LG>::THREADS, (D.gal ? ModDownCfg<LG>::SMEM_BYTES_GAL : ModDownCfg<LG>::SMEM_BYTES), g.stream, cl, T, D));
    } else {
        KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_moddown_mix<LG, CC>), grid, NttCfg<LG>::THREADS, (D.gal ? ModDownCfg<LG>::SMEM_BYTES_GAL : ModDownCfg<LG>::SMEM_BYTES), g.stream, cl, T, D));
    }
}

template <int LG, int CC> static int attrs()
{
#ifndef B200HE_EMU
    cudaError_t e = cudaFuncSetAttribute(k_moddown<LG, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModDownCfg<LG>::SMEM_BYTES_GAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_moddown_dp<LG, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModDownCfg<LG>::SMEM_BYTES_GAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_moddown_mix<LG, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModDownCfg<LG>::SMEM_BYTES_GAL);
    return (int)e;
#else
    return 0;
#endif
}
int smem_attrs_moddown(const Geo &g)
{
    int rc = 0;
    KERNEL_DISPATCH(g, (rc = attrs<LG, CC>()));
    return rc;
}

}   // namespace b200he
