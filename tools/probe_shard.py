"""Debug probe: kernel time of one shard of the cfg5 full histogram (what one GPU of N does), N = 1, 2, 4, 8.
usage: python tools/probe_shard.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic
from fullrmc_b200.store import DeviceStore

s = synthetic.cfg5()
st = DeviceStore(s.boxCoords, s.basis, s.isPBC, s.moleculeIndex, s.elementIndex, s.numberOfElements)
st.add_grid(0.0, 20.0, 0.02, 1000)
base = None
for n in (1, 2, 4, 8):
    worst = 0.0
    for shard in range(n):
        st.set_timing(False)
        st.compute_data_shard(shard, n); st.compute_data_shard(shard, n)
        st.set_timing(True)
        for _ in range(4):
            st.compute_data_shard(shard, n)
        ms, k = st.get_timing("full")
        worst = max(worst, ms / k)
        st.set_timing(False)
    base = base or worst
    print("shards %d: slowest shard %.3f ms (pipeline incl. box and list kernels), ideal %.3f, efficiency %.3f" % (n, worst, base / n, base / n / worst), flush=True)
st.close()
