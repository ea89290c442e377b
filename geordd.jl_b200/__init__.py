"""geordd.jl_b200 -- B200-native (sm_100a) implementation of the GeoRDD.jl data-parallel hot path.

Layout:
  csrc/      hand-written CUDA kernels (DMMA mainloop, dataflow Cholesky / batched triangular solves,
             covariance assembly, statistics) and the C ABI (include/geordd_b200.h)
  lib/       libgeordd_b200.so (built in-tree by build.py; git-ignored)
  capi.py    ctypes binding of the C ABI (twin of julia/GeoRDDB200.jl)
  geordd.py  host-side mirror of the GeoRDD.jl entry points (cliff_face, boot_chi2test, ...)

The directory name contains a dot, so the package is imported through the repo-root loader
`georddb200.load()` under the module name `geordd_jl_b200`.
"""
from . import capi  # noqa: F401
from . import geordd  # noqa: F401
from . import parallel  # noqa: F401
from .geordd import *  # noqa: F401,F403
