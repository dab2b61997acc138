/* kernel instantiations for one target model; see bb_chain.cuh, bb_second.cuh */
#include "bb_second.cuh"
bb_chain_launch_fn bb_lookup_fhn_diag(int gk, int gm, int auxc, int rng) { return bb_lookup_model<MFhnDiag>(gk, gm, auxc, rng); }
bb_chain_launch_fn bb_lookup2_fhn_diag(int gk, int gm, int auxc, int mode) { return bb_lookup_second<MFhnDiag>(gk, gm, auxc, mode); }
