/* kernel instantiations for the wide models: Landmarks (d = 16, d' = 8) and the d' = 8 Wiener process that drives
 * it; see bb_wide.cuh */
#include <stdlib.h>

#include "bb_wide_mma.cuh"
bb_chain_launch_fn bb_lookup_landmarks(int gk, int gm, int auxc, int rng) {
  (void)gm;
  const char* e = getenv("BB_WIDE_LANES");
  if (e && atoi(e) == 1) return bb_lookup_wide<MLandmarks>(gk, auxc, rng);
  /* guided launches: the tensor-core kernel (bb_wide_mma.cuh; compared with the oracle at the contract tolerance);
   * BB_WIDE_MMA=0 selects the four-lane kernel without DMMA (bit-identical to the oracle's GPU-order build) */
  const char* mm = getenv("BB_WIDE_MMA");
  bb_chain_launch_fn f = (mm && atoi(mm) == 0) ? nullptr : bb_lookup_landmarks4m(gk, auxc, rng);
  if (!f) f = bb_lookup_landmarks4(gk, auxc, rng);
  return f ? f : bb_lookup_wide<MLandmarks>(gk, auxc, rng); /* sample! + solve! fused (rng 2) stays one thread per chain */
}
bb_chain_launch_fn bb_lookup_wiener_wide(int d, int rng) {
  if (d == 8) return bb_lookup_wide<MWiener<8>>(0, 1, rng);
  return nullptr;
}
