import os, sys
sys.path.insert(0, "/root/repo")
from fullrmc_b200 import synthetic
from fullrmc_b200.store import DeviceStore
s = synthetic.cfg5()
st = DeviceStore(s.boxCoords, s.basis, s.isPBC, s.moleculeIndex, s.elementIndex, s.numberOfElements)
st.add_grid(0.0, 20.0, 0.02, 1000)
for _ in range(3):
    st.compute_data_shard(3, 8)
print("done")
