// tu_ntt.cu -- instantiations and launchers of the transform kernels k_ntt_fwd / k_ntt_fwd_p / k_ntt_inv (K1 / K2).
#include "launch.h"

namespace b200he {

void launch_ntt_fwd(const Geo &g, const Tables &T, const u64 *src, u64 *dst, size_t src_outer, size_t dst_outer, int L, int mod_base, size_t nlimbs,
                    unsigned persistent_ctas)
{
    if (persistent_ctas && g.c == 0) {
        NTT_DISPATCH(g, B200HE_LAUNCH(k_ntt_fwd_p<LG>, persistent_ctas, NttCfg<LG>::THREADS, NttFwdPCfg<LG>::SMEM_BYTES, g.stream, T, src, dst, src_outer,
                                      dst_outer, L, mod_base, (int)nlimbs));
        return;
    }
    KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_ntt_fwd<LG, CC>), (unsigned)(nlimbs << g.c), NttCfg<LG>::THREADS, NttCfg<LG>::SMEM_BYTES, g.stream,
                                             (g.c == 2 ? 2u : 1u) /* forward: limbs of two chunks exchange nothing, limbs of four chunks exchange inside pairs (load_fwd_split) */, T,
                                             src, dst, src_outer, dst_outer, L, mod_base));
}

void launch_ntt_inv(const Geo &g, int kind, const Tables &T, const u64 *src, u64 *dst, size_t src_outer, size_t dst_outer, int L, int mod_base, int mode,
                    const InvFuse &F, size_t nlimbs)
{
    const unsigned grid = (unsigned)(nlimbs << g.c);
    if (kind == KIND_INT) {
        KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_ntt_inv<LG, CC, KIND_INT>), grid, NttCfg<LG>::THREADS, NttCfg<LG>::SMEM_BYTES, g.stream, 1u << g.c,
                                                 T, src, dst, src_outer, dst_outer, L, mod_base, mode, F));
    } else if (kind == KIND_DP) {
        KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_ntt_inv<LG, CC, KIND_DP>), grid, NttCfg<LG>::THREADS, NttCfg<LG>::SMEM_BYTES, g.stream, 1u << g.c,
                                                 T, src, dst, src_outer, dst_outer, L, mod_base, mode, F));
    } else {
        KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_ntt_inv<LG, CC, KIND_BOTH>), grid, NttCfg<LG>::THREADS, NttCfg<LG>::SMEM_BYTES, g.stream, 1u << g.c,
                                                 T, src, dst, src_outer, dst_outer, L, mod_base, mode, F));
    }
}

template <int LG, int CC> static int attrs()
{
#ifndef B200HE_EMU
    const int bytes = NttCfg<LG>::SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(k_ntt_fwd<LG, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && CC == 0) e = cudaFuncSetAttribute(k_ntt_fwd_p<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttFwdPCfg<LG>::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute((k_ntt_inv<LG, CC, KIND_BOTH>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute((k_ntt_inv<LG, CC, KIND_INT>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute((k_ntt_inv<LG, CC, KIND_DP>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return (int)e;
#else
    return 0;
#endif
}
int smem_attrs_ntt(const Geo &g)
{
    int rc = 0;
    KERNEL_DISPATCH(g, (rc = attrs<LG, CC>()));
    return rc;
}

}   // namespace b200he
