"""getdist_b200 -- B200-native (sm_100a CUDA) backend for GetDist's FFT-KDE and weighted-statistics hot path.

The CUDA library (getdist_b200/lib/libgdk.so, C-ABI in include/gdk.h) is built in-tree by
``python -m getdist_b200.build`` / ``__graft_entry__.build()``.  Importing the package does not need a GPU;
constructing ``MCSamples`` does, and fails loudly without one (there is no CPU fallback)."""
from .densities import DensitiesError, Density1D, Density2D  # noqa: F401
from .mcsamples import (BandwidthError, MCSamples, MCSamplesError, ParamBounds, ParamError, SettingError,  # noqa: F401
                        WeightedSampleError)

__version__ = "0.1.0"
