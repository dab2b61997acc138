/* kernel instantiations for one target model; see bb_chain.cuh, bb_second.cuh */
#include "bb_second.cuh"
bb_chain_launch_fn bb_lookup_linpro3(int gk, int gm, int auxc, int rng) { return bb_lookup_model<MLinPro<3>>(gk, gm, auxc, rng); }
bb_chain_launch_fn bb_lookup2_linpro3(int gk, int gm, int auxc, int mode) { return bb_lookup_second<MLinPro<3>>(gk, gm, auxc, mode); }
