"""One whole-system and one per-move coordination-number call on cfg4 (100 k atoms), for profiler runs:
    ncu --set full --clock-control none -k regex:coordnum -c 2 -o gpurun_out/coordnum python tools/probe_coordination.py
and, without a profiler, the device time of the kernel (CUDA events are not visible from here: wall clock of the
C call minus nothing -- the stateless call includes its copies)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic  # noqa: E402
from fullrmc_b200.Core import atomic_coordination as ac  # noqa: E402

s = synthetic.cfg4()
n, el = s.numberOfAtoms, s.elementIndex
pairs = [(0, 1), (2, 2), (3, 4)]
cores = [np.nonzero(el == a)[0].astype(np.int32) for a, _ in pairs]
shells = [np.nonzero(el == b)[0].astype(np.int32) for _, b in pairs]
as_core, in_shell = [[] for _ in range(n)], [[] for _ in range(n)]
for d in range(3):
    for i in cores[d]:
        as_core[i].append(d)
    for i in shells[d]:
        in_shell[i].append(d)
kw = dict(boxCoords=s.boxCoords, basis=s.basis, isPBC=s.isPBC, coresIndexes=cores, shellsIndexes=shells, lowerShells=[np.float32(1.5)] * 3,
          upperShells=[np.float32(3.5)] * 3, asCoreDefIdxs=as_core, inShellDefIdxs=in_shell)
for rep in range(2):
    data = np.zeros(3, np.float32)
    t0 = time.perf_counter()
    ac.all_atoms_coord_number_coords(coordNumData=data, **kw)
    t1 = time.perf_counter()
    one = np.zeros(3, np.float32)
    ac.multi_atoms_coord_number_coords(indexes=np.array([n // 2], np.int32), coordNumData=one, **kw)
    t2 = time.perf_counter()
    print("all atoms %.2f ms  %s   one atom %.3f ms  %s" % (1e3 * (t1 - t0), data / 2, 1e3 * (t2 - t1), one))
