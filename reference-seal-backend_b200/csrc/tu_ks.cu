// tu_ks.cu -- instantiations and launcher of the key-switch inner product k_ks_inner (K6 step 2).
#include <stdlib.h>

#include "launch.h"

namespace b200he {

void launch_ks_inner(const Geo &g, const Tables &T, const KsInnerArgs &A, size_t units)
{
    // Limbs of four chunks: the special-prime units (0 .. B-1: forward transforms, inner product and the fused inverse
    // transforms) in clusters of four, everything else in clusters of two (k_ks_inner, PAIRS).  On the context's side stream
    // the first launch runs next to the second -- clusters of four leave 16 SMs idle, clusters of two take them.  Measured
    // (k_ks_inner per operate()): C5, B = 1024: one launch 160.6 ms, two launches back to back 150.1, overlapped 143.7;
    // MatMult Row, B = 100: 354.7 / 383 / 359.6 -- small batches stay with one launch (B200HE_KS_SPLIT_MIN).
    static const int split_min = []() { const char *e = getenv("B200HE_KS_SPLIT_MIN"); return e ? atoi(e) : 200; }();   // (tests set 1)
    if (g.c == 2 && g.lognl == 13 && units > (size_t)A.B && A.B >= split_min) {
        KsInnerArgs S = A, D = A;
        S.unit0 = 0;
        D.unit0 = (int)A.B;
        cudaStream_t ss = g.stream;
#ifndef B200HE_EMU
        if (g.aux) {
            cudaEventRecord(g.fork, g.stream);
            cudaStreamWaitEvent(g.aux, g.fork, 0);
            ss = g.aux;
        }
#endif
        B200HE_LAUNCH_CLUSTER((k_ks_inner<13, 2, false>), (unsigned)((size_t)A.B << 2), NttCfg<13>::THREADS, KsCfg<13>::SMEM_BYTES, ss, 4u, T, S);
        B200HE_LAUNCH_CLUSTER((k_ks_inner<13, 2, true>), (unsigned)((units - (size_t)A.B) << 2), NttCfg<13>::THREADS, KsCfg<13>::SMEM_BYTES, g.stream, 2u, T, D);
#ifndef B200HE_EMU
        if (g.aux) {
            cudaEventRecord(g.join, g.aux);
            cudaStreamWaitEvent(g.stream, g.join, 0);
        }
#endif
        return;
    }
    KERNEL_DISPATCH(g, B200HE_LAUNCH_CLUSTER((k_ks_inner<LG, CC>), (unsigned)(units << g.c), NttCfg<LG>::THREADS, KsCfg<LG>::SMEM_BYTES, g.stream, 1u << g.c, T, A));
}

template <int LG, int CC> static int attrs()
{
#ifndef B200HE_EMU
    int rc = (int)cudaFuncSetAttribute(k_ks_inner<LG, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, KsCfg<LG>::SMEM_BYTES);
    if (CC == 2 && !rc) rc = (int)cudaFuncSetAttribute(k_ks_inner<LG, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KsCfg<LG>::SMEM_BYTES);
    return rc;
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
