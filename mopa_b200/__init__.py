"""mopa_b200: B200-native (sm_100a) implementation of MoPA's 3D-branch hot path, the SparseConvNet stack behind
UNetSCN. `mopa_b200.scn` mirrors the sparseconvnet module surface; kernels live in libmopa_scn.so (C ABI in
include/mopa_scn.h). GPU only: there is no CPU fallback."""
__version__ = "0.1.0"
