"""hiop_b200 -- B200-native (sm_100a) engine for HiOp's KKT assemble + factor + solve hot path.

The compute path is the C-ABI shared library hiop_b200/libhiopb200.so (hand-written CUDA, declared in
include/hiopb200.h). This package is only the Python host-side mirror of the reference's plug-in interfaces
(hiopKKTLinSysLowRank / hiopHessianLowRank / hiopLinSolverSymDense / hiopVector) on top of that library.
There is NO CPU fallback: using any engine class without the built library or without a GPU raises.
"""
__version__ = "0.1.0"
