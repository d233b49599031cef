"""Debug probe: compute_data time of the device store for a few regimes, culling on and off."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic, _lib as L
from fullrmc_b200.store import DeviceStore

def run(name, s, rmax, hs, both=True):
    for cull in ((True, False) if both else (True,)):
        L.set_block_culling(cull)
        st = DeviceStore(s.boxCoords, s.basis, s.isPBC, s.moleculeIndex, s.elementIndex, s.numberOfElements)
        st.add_grid(0.0, rmax, rmax / hs, hs)
        st.compute_data_shard(0, 1); st.compute_data_shard(0, 1)
        st.set_timing(True)
        for _ in range(5): st.compute_data_shard(0, 1)
        ms, n = st.get_timing("full")
        n_at = s.boxCoords.shape[0]
        hi, he = st.export_data(0)
        hits = float(hi.sum(dtype=np.float64) + he.sum(dtype=np.float64))
        print("%-38s culling=%-5s %8.3f ms  evals %.4g (%.4f of all)  hits %.4g (%.3f of evals)  %.0f G evals/s" % (
            name, cull, ms / n, st.swept_pairs, st.swept_pairs / (n_at * (n_at - 1) / 2), hits, hits / max(st.swept_pairs, 1),
            st.swept_pairs / (ms / n * 1e-3) / 1e9), flush=True)
        st.close()
    L.set_block_culling(True)

if __name__ == "__main__":
    run("cfg5 1M cubic rmax 20", synthetic.cfg5(), 20.0, 1000, both="--all" in sys.argv)
    if "--cfg5only" in sys.argv:
        sys.exit(0)
    run("cfg4 100k triclinic rmax 20", synthetic.cfg4(), 20.0, 1000)
    run("cfg5-like 200k cubic rmax 20", synthetic.cfg5(200000), 20.0, 1000)
    run("cfg5-like 100k cubic rmax 45 (dense)", synthetic.cfg5(100000), 45.0, 1000)
    run("cfg5-like 400k cubic rmax 10", synthetic.cfg5(400000), 10.0, 500)
