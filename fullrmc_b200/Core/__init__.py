"""Drop-in replacements for the reference's compiled extension modules
(``fullrmc.Core.pairs_distances``, ``fullrmc.Core.pairs_histograms``,
``fullrmc.Core.reciprocal_space``): same function names, keyword names, dtypes and
return shapes, computed by the CUDA library through the C ABI."""
