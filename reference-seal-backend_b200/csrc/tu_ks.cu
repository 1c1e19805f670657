// tu_ks.cu -- instantiations and launcher of the key-switch inner product k_ks_inner (K6 step 2).
#include "launch.h"

namespace b200he {

void launch_ks_inner(const Geo &g, const Tables &T, const KsInnerArgs &A, size_t units)
{
    KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_ks_inner<LG, CC>), (unsigned)(units << g.c), NttCfg<LG>::THREADS, KsCfg<LG>::SMEM_BYTES, g.stream, 1u << g.c, T, A));
}

template <int LG, int CC> static int attrs()
{
#ifndef B200HE_EMU
    return (int)cudaFuncSetAttribute(k_ks_inner<LG, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, KsCfg<LG>::SMEM_BYTES);
#else
    return 0;
#endif
}
int smem_attrs_ks(const Geo &g)
{
    int rc = 0;
    KERNEL_DISPATCH(g, (rc = attrs<LG, CC>()));
    return rc;
}

}   // namespace b200he
