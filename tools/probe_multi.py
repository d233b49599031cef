"""Debug probe: wall time of ONE host-buffer full-histogram call spread over 1..N GPUs (cfg5)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic, _lib as L
from fullrmc_b200.Core import pairs_histograms as ph
s = synthetic.cfg5()
g = synthetic.RGrid(0.0, 0.02, 1000)
kw = dict(s.hist_kwargs(), **g.kwargs())
ndev = int(L.load_library().frmc_device_count())
ref = None
for nd in [d for d in (1, 2, 4, 8) if d <= ndev]:
    devs = list(range(nd))
    hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, _devices=devs, **kw)
    t0 = time.perf_counter()
    for _ in range(5):
        hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, _devices=devs, **kw)
    dt = (time.perf_counter() - t0) / 5
    if ref is None:
        ref = (hi, he)
    print("%d devices: %8.2f ms per call  identical=%s  reduce=%s" % (nd, 1e3 * dt, np.array_equal(hi, ref[0]) and np.array_equal(he, ref[1]),
                                                                     L.load_library().frmc_multi_reduce_path().decode()), flush=True)
