"""bridge.jl_b200 -- B200-native (sm_100a) implementation of the Bridge.jl hot path.

Import as `import bridge_jl_b200` (the loader at the repository root maps that name onto this
directory, whose own name is not a Python identifier).  Contents:
    csrc/   CUDA kernels + the C ABI (include/bridge_b200.h)  -> lib/libbridge_b200.so
    _cabi   ctypes binding of the C ABI
    api     host-side mirror of the reference interface (SamplePath, solve, llikelihood, ...)
"""
from ._cabi import (BridgeError, CUR, PROP, W, X, SYMBOLS, LIB_PATH, ARITH_REFERENCE, ARITH_FUSED)  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (Context, default_context, PathEnsemble, SamplePath, VSamplePath, samplepath, sample, sample_,
                  seed_, solve, solve_, bridge_, llikelihood, lptilde, innovations_, pcn_, theta_mcmc_, gpupdate, gpupdate_νH,
                  EulerMaruyama, EulerMaruyama_, Euler, StratonovichEuler, StochasticHeun, StochasticRungeKutta, Mdb, LeftRule, R3, Lyap, ContinuousTimeProcess, Wiener, OrnsteinUhlenbeck,
                  LinPro, FitzHughNagumo, FitzhughDiffusion, IntegratedDiffusion, NclarDiffusion, Lorenz, Landmarks, LandmarksTilde, BolusDiffusion,
                  LinearAux, LinearAppr, linearappr, bderiv, UserProcess, check_user_source, PartialBridgeνH, PartialBridgenuH, partialbridgeνH, partialbridgenuH, GuidedBridge,
                  PartialBridge, GuideTables, PartialBridgeνHChain)  # noqa: F401
