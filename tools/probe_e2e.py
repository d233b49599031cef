"""Debug probe: wall time of the stateless host-buffer full histogram (cfg5, cfg4), device vs host layout."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic, _lib as L
from fullrmc_b200.Core import pairs_histograms as ph

for name, s in (("cfg5", synthetic.cfg5()), ("cfg4", synthetic.cfg4())):
    g = synthetic.RGrid(0.0, 0.02, 1000)
    kw = dict(s.hist_kwargs(), **g.kwargs())
    ref = None
    for dev_layout in (True, False):
        L.set_device_layout(dev_layout)
        ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
        t0 = time.perf_counter()
        for _ in range(5):
            hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
        dt = (time.perf_counter() - t0) / 5
        if ref is None:
            ref = (hi, he)
        print("%s device_layout=%-5s %8.2f ms per call   identical=%s" % (name, dev_layout, 1e3 * dt, np.array_equal(hi, ref[0]) and np.array_equal(he, ref[1])), flush=True)
    L.set_device_layout(True)
