// kernels.cuh -- all template kernels of the NTT-bearing families (see kernels_common.cuh for the overview).
#pragma once
#include "kernels_common.cuh"
#include "kernels_ntt.cuh"
#include "kernels_ks.cuh"
#include "kernels_moddown.cuh"
