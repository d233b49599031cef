"""Debug probe: phase timeline of the batch kernel (globaltimer stamps of CTA 0, last launch of a run).
usage: FRMC_BATCH_STAMPS=1 python tools/probe_batch.py [cfg4|cfg5] [n_proposals] [tolerance]"""
import os, sys, time
if not os.environ.get("FRMC_NO_STAMPS"):
    os.environ.setdefault("FRMC_BATCH_STAMPS", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic, _lib as L
from fullrmc_b200.store import DeviceStore
from fullrmc_b200.model import ModelSpec

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
F32 = np.float32
s = synthetic.cfg4() if which == "cfg4" else synthetic.cfg5()
grid = synthetic.RGrid(0.0, 0.02, 1000)
q = synthetic.q_values(nq=400)
common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
              rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
rng = np.random.default_rng(101)
smooth = lambda m, c: (c + 0.02 * np.convolve(rng.standard_normal(m + 20), np.ones(21) / 21.0, "valid")).astype(F32)
store = DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, s.numberOfElements)
g = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
store.add_model(g, ModelSpec("PDF", experimental=smooth(1000, 0.0), **common))
store.add_model(g, ModelSpec("SQ", experimental=smooth(400, 1.0), q_values=q, **common))
chi = store.compute_data()
total = np.sum([F32(x) for x in chi], dtype=F32)
idx = rng.integers(0, s.numberOfAtoms, n).astype(np.int32)
inv = np.linalg.inv(s.basis.astype(np.float64))
moved = (s.boxCoords[idx] + (rng.normal(0.0, 0.1, (n, 3)) @ inv).astype(F32)).astype(F32)
rand = rng.random(n).astype(F32)
w = store.run_batch(idx[:200], moved[:200], total, rand[:200], tolerance=tol)
t0 = time.perf_counter()
out = store.run_batch(idx[200:], moved[200:], w["total"], rand[200:], tolerance=tol)
wall = time.perf_counter() - t0
m = n - 200
launches, rounds, props = store.batch_stats()
print("%s: %d proposals, %d accepted, device %.2f us/eval, wall %.2f us/eval, %d launches, %d rounds (all runs)" % (
    which, m, int((out["decisions"] > 0).sum()), 1e3 * out["device_ms"] / m, 1e6 * wall / m, launches, rounds))
if os.environ.get("FRMC_NO_STAMPS"):
    store.close(); sys.exit(0)
st = np.zeros(4 + 5 * 64 + 64 * 128, np.int64)
L.check(store._lib.frmc_store_batch_stamps(store._handle, st.ctypes.data_as(L.c_i64p), st.shape[0]), "stamps")
print("last launch: clear %.2f us, delta pass %.2f us, rounds+end %.2f us, total %.2f us" % (
    (st[1] - st[0]) / 1e3, (st[2] - st[1]) / 1e3, (st[3] - st[2]) / 1e3, (st[3] - st[0]) / 1e3))
prev = st[2]
acc = {"own_epilogue": [], "wait_all": [], "decide": [], "commit": [], "commit_barrier": [], "round": []}
for r in range(64):
    a = st[4 + 5 * r: 9 + 5 * r]
    if a[0] == 0:
        break
    acc["own_epilogue"].append(a[0] - prev); acc["wait_all"].append(a[1] - a[0]); acc["decide"].append(a[2] - a[1])
    end = a[2]
    if a[3]:
        acc["commit"].append(a[3] - a[2]); acc["commit_barrier"].append(a[4] - a[3]); end = a[4]
    acc["round"].append(end - prev)
    prev = end
print("rounds in last launch: %d (with commit: %d)" % (len(acc["round"]), len(acc["commit"])))
for k, v in acc.items():
    if v:
        print("  %-15s mean %.2f us  min %.2f  max %.2f" % (k, np.mean(v) / 1e3, np.min(v) / 1e3, np.max(v) / 1e3))
# CTA 1 (first S(Q) slab of group 0): where its evaluation spends the round
sq = {"plan": [], "tables": [], "G(r)": [], "S(Q) rows": [], "fence+barrier": []}
for r in range(len(acc["round"])):
    e = st[4 + 5 * 64 + 128 * r: 4 + 5 * 64 + 128 * (r + 1)]
    if e[120] == 0 or e[75] == 0:
        continue
    sq["plan"].append(e[120] - e[121]); sq["tables"].append(e[73] - e[120]); sq["G(r)"].append(e[74] - e[73])
    sq["S(Q) rows"].append(e[75] - e[74]); sq["fence+barrier"].append(st[4 + 5 * r + 1] - e[75])
for k, v in sq.items():
    if v:
        print("  S(Q) CTA %-14s mean %.2f us  min %.2f  max %.2f" % (k, np.mean(v) / 1e3, np.min(v) / 1e3, np.max(v) / 1e3))
store.close()
