"""Small multi-block workloads through every kernel of the library, for compute-sanitizer runs:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
    SAN_SKIP_STORE=1 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from fullrmc_b200 import synthetic
from fullrmc_b200.Core import pairs_histograms as ph, atomic_distances as ad, atomic_coordination as ac
from fullrmc_b200.store import DeviceStore
from fullrmc_b200.model import ModelSpec
# small but multi-block systems through every new kernel
for n, basis, pbc in ((9000, np.diag([70.0, 66.0, 72.0]).astype(np.float32), True), (40000, np.array([[90, 0, 0], [11, 86, 0], [-8, 14, 88]], np.float32), True)):
    s = synthetic.random_system(n, 3, basis, n_elements=3, molecule_size=4, isPBC=pbc)
    kw = dict(s.hist_kwargs(), minDistance=np.float32(0.0), maxDistance=np.float32(7.0), bin=np.float32(0.05), histSize=140)
    hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    print("full hist", n, hi.sum() + he.sum())
    lo = np.zeros((3, 3, 1), np.float32); up = np.full((3, 3, 1), 2.0, np.float32)
    r = ad.full_atomic_distances_coords(s.boxCoords, s.basis, pbc, s.moleculeIndex, s.elementIndex, 3, lo, up, intraMolecular=False)
    print("atomdist", int(r[2].sum()))
    cores = [np.nonzero(s.elementIndex == 0)[0].astype(np.int32), np.nonzero(s.elementIndex == 1)[0].astype(np.int32)]
    shells = [np.nonzero(s.elementIndex == 2)[0].astype(np.int32), cores[1]]
    as_core = [[int(e)] if e < 2 else [] for e in s.elementIndex]; in_shell = [[0] if e == 2 else ([1] if e == 1 else []) for e in s.elementIndex]
    cn = np.zeros(2, np.float32)
    ac.multi_atoms_coord_number_coords(np.arange(0, n, 7, dtype=np.int32), s.boxCoords, s.basis, pbc, cores, shells, [0.5, 1.0], [3.0, 4.0], as_core, in_shell, cn)
    print("coordination", cn)
# nodes of more than 20 000 points: the multi-CTA splits of the device-side ordering (dl_whist / dl_wcut / dl_wscatter), full
# leaves (bitonic sort), the sweep-record pass (order inside the sub-blocks, chunk boxes) and the chunk-level culling
s = synthetic.random_system(50000, 9, np.diag([120.0, 118.0, 121.0]).astype(np.float32), n_elements=2)
kw = dict(s.hist_kwargs(), minDistance=np.float32(0.0), maxDistance=np.float32(3.0), bin=np.float32(0.05), histSize=60)
hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
print("full hist (wide splits)", 50000, hi.sum() + he.sum())
if os.environ.get("SAN_SKIP_STORE"):
    print("done (stateless kernels only)"); sys.exit(0)
s = synthetic.random_system(20000, 5, np.diag([60.0, 60.0, 60.0]).astype(np.float32), n_elements=2)
grid = synthetic.RGrid(0.0, 0.05, 200)
q = synthetic.q_values(nq=64)
st = DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 2)
g = st.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume, rho0=s.numberDensity,
              shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
st.add_model(g, ModelSpec("PDF", experimental=np.zeros(200, np.float32), **common))
st.add_model(g, ModelSpec("SQ", experimental=np.ones(64, np.float32), q_values=q, **common))
print("chi2", st.compute_data())
rng = np.random.default_rng(1)
for mode in (False, True):
    st.set_persistent(mode)
    prev = None
    for it in range(12):
        i = rng.integers(0, 20000, 1).astype(np.int32)
        chi = st.step(prev, i, (s.boxCoords[i] + rng.normal(0, 0.01, (1, 3))).astype(np.float32)).copy()
        prev = it % 2 == 0
    (st.accept if prev else st.reject)()
    print("persistent", mode, chi)
# runs of proposals resolved on the device (batch_kernel): single atoms, a repeated atom, neighbours
st.set_persistent(False)
n = 80
idx = rng.integers(0, 20000, n).astype(np.int32)
idx[10] = idx[9]; idx[30] = idx[2]
moved = (s.boxCoords[idx] + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
total = np.float32(np.sum(st.committed_chi2()[:2], dtype=np.float32))
out = st.run_batch(idx, moved, total, rng.random(n).astype(np.float32), tolerance=0.3)
print("batch", int((out["decisions"] > 0).sum()), "accepted of", n, "in", st.batch_stats()[1], "rounds")
# steps generated on the device (generate_batch_kernel + batch_kernel<GEN>), on-the-fly pair corrections (dense launch)
b64 = s.basis.astype(np.float64)
st.set_groups(None)
st.set_real_coords((st.get_coords().astype(np.float64) @ b64).astype(np.float32), np.linalg.inv(b64).astype(np.float32))
total = np.float32(np.sum(st.committed_chi2()[:2], dtype=np.float32))
out = st.run_generated(96, 5, 0, 0.3, total)
print("generated", int((out["decisions"] > 0).sum()), "accepted of 96")
# atom removal (amputation evaluation, accept, constants), the distance pass on the store
chi = st.propose_amputation(17)
st.accept_amputation()
chi = st.propose_amputation(40); st.reject_amputation()
print("amputation", chi, st.numberOfAtoms)
st.close()
s2 = synthetic.random_system(9000, 3, np.diag([70.0, 66.0, 72.0]).astype(np.float32), n_elements=3, molecule_size=4)
with DeviceStore(s2.boxCoords, s2.basis, True, s2.moleculeIndex, s2.elementIndex, 3) as st2:
    cid = st2.distance_add(s2.elementIndex, 3, np.zeros((3, 3, 1), np.float32), np.full((3, 3, 1), 2.5, np.float32), reduceDistanceToUpper=True)
    idx = np.arange(8, 12, dtype=np.int32)
    counts, sums = st2.distance_move(cid, idx, (s2.boxCoords[idx] + np.float32(0.01)).astype(np.float32))
    print("store distance pass", int(counts.sum()))
    by_el = [np.flatnonzero(s2.elementIndex == e).astype(np.int32) for e in range(3)]
    kid = st2.coordination_add([by_el[0], by_el[1], np.arange(9000, dtype=np.int32)], [by_el[1], by_el[2], by_el[0]], [1.0, 0.0, 0.5], [3.0, 2.5, np.inf])
    for rep in range(3):
        cn = st2.coordination_move(kid, idx, (s2.boxCoords[idx] + np.float32(0.01 * (rep + 1))).astype(np.float32))
    print("store coordination pass", cn.tolist())
print("done")
