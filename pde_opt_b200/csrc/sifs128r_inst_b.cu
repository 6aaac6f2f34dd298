// Instantiations of the one-field-per-CTA 128x128 SIFS kernel (double well, runtime-switch variants); see capi.cu.
#include "sifs128r_launch.h"

cudaError_t pdeopt_sifs128r_launch_b(int variant, const SifsParams& p, cudaStream_t st) {
  switch (variant) {
    case 0: return launch_r<EQ_AC, MU_RUNTIME, MOB_RUNTIME>(p, st);
    case 3: return launch_r<EQ_CH, MU_DOUBLE_WELL, MOB_CONST>(p, st);
    case 4: return launch_r<EQ_CH, MU_RUNTIME, MOB_RUNTIME>(p, st);
    default: return cudaErrorInvalidValue;
  }
}
