/* kernel instantiations for the Wiener process as a target (b = 0, sigma = I): plain Euler, sampling, innovations */
#include "bb_second.cuh"
bb_chain_launch_fn bb_lookup_wiener(int d, int rng) {
  if (d == 1) return bb_lookup_unguided<MWiener<1>>(rng);
  if (d == 2) return bb_lookup_unguided<MWiener<2>>(rng);
  if (d == 3) return bb_lookup_unguided<MWiener<3>>(rng);
  return nullptr;
}
bb_chain_launch_fn bb_lookup2_wiener(int d, int mode) {
  if (d == 1) return bb_lookup_second<MWiener<1>>(0, 0, 1, mode);
  if (d == 2) return bb_lookup_second<MWiener<2>>(0, 0, 1, mode);
  if (d == 3) return bb_lookup_second<MWiener<3>>(0, 0, 1, mode);
  return nullptr;
}
