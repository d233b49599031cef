"""Parity fixtures AT THE SIZES bench.py measures (BASELINE.json configs[3] and [4]):

    tests/golden/full_cfg4.npz, full_cfg5.npz          full pair histograms of the 100 000-atom triclinic
                                                       and the 1 000 000-atom cubic synthetic boxes
    tests/golden/constraints_cfg4.npz, _cfg5.npz       200-move Metropolis trajectories of the UNMODIFIED
                                                       reference PairDistributionConstraint +
                                                       StructureFactorConstraint on those boxes

Run in the build container (needs /root/reference):   python tests/gen_golden_large.py [cfg4] [cfg5]

Where the numbers come from
* cfg4 (5e9 pairs): everything is the reference itself -- its compiled full_pairs_histograms_coords and the
  unmodified constraint classes (about five minutes on one core).
* cfg5 (5e11 pairs, 4.3 h on one core through the reference's Python row loop): the full histograms come from
  the C restatement oracle/pairhist_oracle.c, which tests/test_oracle.py pins bit-for-bit to the compiled
  reference, row-sharded over the host cores (35 min on 7 cores per r-grid; run separately and cached under
  /tmp/gold by the two scripts quoted in the docstring of cached_full()); it is cross-checked here against the
  compiled reference on 64 uniformly spaced rows of the same system, and those reference rows are stored too.
  The per-move trajectory is the reference's own class code (compute_before_move / compute_after_move /
  accept_move / reject_move, PairDistributionConstraints.py:1044-1166, StructureFactorConstraints.py:975-1096)
  with the reference's compiled kernels; only the O(N^2) call inside compute_data is answered from the cache.

The per-atom arrays are not stored (12 MB of random floats): the fixtures hold the recipe of
fullrmc_b200.synthetic (name, n, seed) and tests/test_golden_large.py regenerates them.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as H  # noqa: E402
import gen_golden_constraints as G  # noqa: E402

CACHE = os.environ.get("FRMC_GOLD_CACHE", "/tmp/gold")
N_STEPS = 200
N_ROWS = 64


class Singles(object):
    """groups = one atom each, without materialising a million lists"""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return [int(i)]


def cached_full(name, grid_tag):
    """cfg5 full histograms computed by oracle/pairhist_oracle.c over 7 threads:
         orc.full_pairs_histograms_coords(boxCoords=synthetic.cfg5().boxCoords, ncores=7, return_overflow=True,
                                          minDistance=0, maxDistance=20, bin=0.02, histSize=1000)             (pdf grid)
         same with maxDistance=19.98, histSize=999 and orc.set_emulate_spill(True)                            (sf grid)
    """
    path = os.path.join(CACHE, "%s_full_%sraw.npz" % (name, grid_tag))
    assert os.path.exists(path), "run the oracle pass first: %s" % path
    z = np.load(path)
    return z["intra"], z["inter"], int(z["overflow"])


def make(name, fullrmc):
    from fullrmc_b200 import synthetic
    from fullrmc.Constraints import PairDistributionConstraints as PDM, StructureFactorConstraints as SFM
    from fullrmc.Core import pairs_histograms as ref_ph
    out_dir = os.path.join(ROOT, "tests", "golden")
    system = {"cfg4": synthetic.cfg4, "cfg5": synthetic.cfg5}[name]()
    n = system.numberOfAtoms
    seed = {"cfg4": 4, "cfg5": 5}[name]
    grid = synthetic.RGrid(0.0, 0.02, 1000)
    kw = dict(system.hist_kwargs(), **grid.kwargs())
    elements = [e.lower() for e in system.elements]
    arrays = (system.boxCoords, system.basis, True, system.moleculeIndex, system.elementIndex, elements)

    # ---- the full histogram on the bench grid
    t0 = time.time()
    rows = np.linspace(0, n - 2, N_ROWS).astype(np.int32)
    ri, re_ = ref_ph.multiple_pairs_histograms_coords(indexes=rows, boxCoords=system.boxCoords, allAtoms=False, ncores=1, **kw)
    if name == "cfg4":
        hi, he = ref_ph.full_pairs_histograms_coords(boxCoords=system.boxCoords, ncores=1, **kw)
        source = "compiled reference (Extensions/pairs_histograms.pyx:289-335)"
        from oracle import pairhist as orc
        oi, oe, ov = orc.full_pairs_histograms_coords(boxCoords=system.boxCoords, ncores=orc.max_threads(), return_overflow=True, **kw)
        assert np.array_equal(oi, hi) and np.array_equal(oe, he), "C oracle differs from the compiled reference at 100k atoms"
        memo = {}
    else:
        hi, he, ov = cached_full(name, "")
        source = "oracle/pairhist_oracle.c (pinned to the compiled reference), 64 rows cross-checked against the compiled reference"
        from oracle import pairhist as orc
        oi, oe = orc.multiple_pairs_histograms_coords(indexes=rows, boxCoords=system.boxCoords, allAtoms=False, **kw)
        assert np.array_equal(oi, ri) and np.array_equal(oe, re_), "C oracle rows differ from the compiled reference at 1M atoms"
        si, se, _ = cached_full(name, "sf_")
        memo = {(0.0, 20.0, 1000): (hi, he), (0.0, float(np.float32(19.98)), 999): (si, se)}
    print("%s full histogram: %d in-range pairs, %d edge overflows, %.0f s  [%s]" % (name, int(hi.sum(dtype=np.float64) + he.sum(dtype=np.float64)),
                                                                                   ov, time.time() - t0, source))
    np.savez_compressed(os.path.join(out_dir, "full_%s.npz" % name), recipe_name=np.array(name), recipe_n=np.int64(n),
                        recipe_seed=np.int64(seed), minDistance=grid.minDistance, maxDistance=grid.maxDistance, bin=grid.bin,
                        histSize=np.int32(grid.hs), intra=hi, inter=he, edge_overflow=np.uint64(ov), source=np.array(source),
                        rows=rows, rows_intra=ri, rows_inter=re_)

    # ---- the Metropolis trajectory of the reference classes
    real_full = PDM.full_pairs_histograms_coords

    def memo_full(boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements, minDistance, maxDistance, bin,
                  histSize, ncores=1):
        key = (float(minDistance), float(maxDistance), int(histSize))
        if boxCoords.shape[0] == n and key in memo:
            assert np.float32(bin) == np.float32(0.02)
            return memo[key][0].copy(), memo[key][1].copy()
        return real_full(boxCoords=boxCoords, basis=basis, isPBC=isPBC, moleculeIndex=moleculeIndex, elementIndex=elementIndex,
                         numberOfElements=numberOfElements, minDistance=minDistance, maxDistance=maxDistance, bin=bin,
                         histSize=histSize, ncores=ncores)

    if memo:
        PDM.full_pairs_histograms_coords = memo_full
        SFM.full_pairs_histograms_coords = memo_full
    exp_g = synthetic.smooth_target(grid.hs, 101, 0.0)
    q = synthetic.q_values()
    exp_s = synthetic.smooth_target(q.shape[0], 102, 1.0)

    def constraints(E):
        r = grid.shellCenters
        pdf = PDM.PairDistributionConstraint(experimentalData=np.stack([r, exp_g], 1).astype(np.float32), weighting="atomicNumber")
        sf = SFM.StructureFactorConstraint(experimentalData=np.stack([q, exp_s], 1).astype(np.float32), weighting="atomicNumber",
                                           rmin=0.0, rmax=19.98, dr=0.02)
        return [(pdf, "PDF"), (sf, "SQ")]

    try:
        G.run_case(name, fullrmc, arrays, constraints, Singles(n), N_STEPS, 7, 0.1, out_dir, recipe=(name, n, seed))
    finally:
        PDM.full_pairs_histograms_coords = real_full
        SFM.full_pairs_histograms_coords = real_full


def main():
    fullrmc = H.load_reference()
    assert fullrmc is not None, "needs /root/reference"
    for name in (sys.argv[1:] or ["cfg4", "cfg5"]):
        make(name, fullrmc)


if __name__ == "__main__":
    G.ONLY = set()
    main()
