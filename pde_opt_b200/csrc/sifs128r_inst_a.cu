// Instantiations of the one-field-per-CTA 128x128 SIFS kernel (logarithmic potential); see capi.cu.
#include "sifs128r_launch.h"

cudaError_t pdeopt_sifs128r_launch_a(int variant, const SifsParams& p, cudaStream_t st) {
  switch (variant) {
    case 1: return launch_r<EQ_CH, MU_LOG, MOB_DEGENERATE>(p, st);
    case 2: return launch_r<EQ_CH, MU_LOG, MOB_CONST>(p, st);
    default: return cudaErrorInvalidValue;
  }
}
