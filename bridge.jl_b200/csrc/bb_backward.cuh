/*
 * bb_backward.cuh -- device building blocks of the proposal constructors (backward ODEs), shared by the
 * one-system kernels of bb_backward.cu and the per-chain kernels of bb_theta.cu:
 *   small dense algebra in the oracle's operation order, kernelr3 (src/ode.jl:44-49), the right-hand sides of
 *   src/partialbridgenuH.jl, src/gode.jl, src/partialbridge.jl, one backward step of partialbridgeodeνH!
 *   (R3: src/partialbridgenuH.jl:21-55; Lyap: :86-103, src/lyap.jl:2-6) and the observation update
 *   (src/guip.jl:221-243, partialbridge_bolus3.jl:128-137).
 */
#pragma once
#include <math.h>

#include "../../include/bridge_b200.h"

namespace bbk_common {
/* auxiliary process values of interval i, Ralston stage k */
struct aux_dev {
  const double* B;
  const double* beta;
  const double* a;
  const double* a_left;
  int is_const;
};
}  // namespace bbk_common

/* default flavour: explicit fused multiply-adds (bit-identical to liboracle_fma.so) */
#define BBK_NS bbk
#define BBK_MA(a, b, c) fma((a), (b), (c))
#include "bb_backward_body.inc"
#undef BBK_NS
#undef BBK_MA
