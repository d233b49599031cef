"""fullrmc_b200 -- B200-native (sm_100a) backend for fullrmc's pair-histogram hot path.

Layout (only what the path needs):

* ``csrc/``  hand-written CUDA kernels + the C ABI (``include/fullrmc_b200.h``),
  built in-tree into ``lib/libfullrmc_b200.so`` by ``csrc/build.sh``.
* ``Core/``  drop-in modules with the reference's extension-module names, function
  names, keyword names, dtypes and return shapes: ``pairs_distances``,
  ``pairs_histograms``, ``reciprocal_space`` (reference: ``fullrmc.Core.<name>``).
* ``store``  the stateful device-resident path (coordinate store + running histograms).
* ``constraints``  device-backed mirrors of PairDistributionConstraint,
  PairCorrelationConstraint, StructureFactorConstraint's five hot methods.

There is no CPU fallback: importing works anywhere, but every compute call raises
``RuntimeError`` if the CUDA library or a CUDA device is missing.
"""
__version__ = "0.1.0"

from ._lib import FullrmcB200Error, library_path, load_library, set_edge_spill  # noqa: F401
