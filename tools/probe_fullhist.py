"""Debug probe: compute_data time of the device store for a few regimes, culling on and off."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic, _lib as L
from fullrmc_b200.store import DeviceStore

def run(name, s, rmax, hs):
    for cull in (True, False):
        L.set_block_culling(cull)
        st = DeviceStore(s.boxCoords, s.basis, s.isPBC, s.moleculeIndex, s.elementIndex, s.numberOfElements)
        st.add_grid(0.0, rmax, rmax / hs, hs)
        st.compute_data(); st.compute_data()
        st.set_timing(True)
        for _ in range(5): st.compute_data()
        ms, n = st.get_timing("full")
        n_at = s.boxCoords.shape[0]
        print("%-34s culling=%-5s %8.3f ms  swept %.4f  R=%s" % (name, cull, ms / n, st.swept_pairs / (n_at * (n_at - 1) / 2), "?"))
        st.close()
    L.set_block_culling(True)

run("cfg4 100k triclinic rmax 20", synthetic.cfg4(), 20.0, 1000)
run("cfg5-like 200k cubic rmax 20", synthetic.cfg5(200000), 20.0, 1000)
run("cfg5-like 100k cubic rmax 45 (dense)", synthetic.cfg5(100000), 45.0, 1000)
run("cfg5-like 400k cubic rmax 10", synthetic.cfg5(400000), 10.0, 500)
