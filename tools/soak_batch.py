"""Randomised soak of the device-resolved runs of proposals (frmc_run_batch: one launch working through all batches of a run)
against the sequential device path (one launch per proposal, the engine's rule on the host), on the geometry cases of
tests/cases.py with random seeds, run lengths, tolerances, group repeats and model sets.
usage: python tools/soak_batch.py [n_runs] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_store as TS
import test_gpu_batch as TB

F32 = np.float32
n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng0 = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4242)
names = ["ortho_atomic", "tri_molecular", "tri_unwrapped", "ortho_unwrapped", "ibc_nanoparticle", "coincident_empty_class", "cfg4_small"]
kindsets = [["PDF", "SQ"], ["PCF"], ["RSQ", "PDF"], ["PDF"], ["PDF", "PCF"], ["SQ"]]
bad = 0
for run in range(n_runs):
    name = names[int(rng0.integers(len(names)))]
    kinds = kindsets[int(rng0.integers(len(kindsets)))]
    n = int(rng0.integers(150, 700))
    tol = float(rng0.choice([0.0, 0.1, 0.35, 1.0]))
    rep = int(rng0.choice([0, 0, 5, 11, 23]))
    sigma = float(rng0.choice([0.01, 0.03, 0.1]))
    mkw = dict(with_weights=bool(rng0.random() < 0.3), with_shape=bool(rng0.random() < 0.3), scale=float(rng0.choice([1.0, 0.93, 1.07])))
    seed = int(rng0.integers(1, 10**6))
    case = TS.CASES[name]
    seq, _ = TS._build(case, kinds, np.random.default_rng(seed), **mkw)
    bat, _ = TS._build(case, kinds, np.random.default_rng(seed), **mkw)
    nm = len(kinds)
    var2 = np.array([1.0, 0.37, 2.5][:nm], F32)
    rng = np.random.default_rng(seed + 1)
    total0 = TB.host_total(seq.compute_data(), var2)
    bat.compute_data()
    props = TB.make_proposals(case, rng, n, sigma, rep)
    rand = rng.random(n).astype(F32)
    chis, decs, total, used = TB.sequential_run(seq, props, total0, rand, tol, var2)
    idx, moved, sizes = TB.flatten(props)
    out = bat.run_batch(idx, moved, total0, rand, tolerance=tol, group_sizes=sizes, variance_squared=var2)
    ok = bool(np.array_equal(out["decisions"], decs) and np.array_equal(out["chi2"], chis) and F32(out["total"]) == F32(total)
              and out["rand_used"] == used)
    try:
        TB.compare_stores(seq, bat, nm)
    except AssertionError as err:
        ok = False; print("   ", err)
    launches = bat.batch_stats()[0]
    seq.close(); bat.close()
    bad += 0 if ok else 1
    print("run %2d %-22s %-8s n %3d tol %.2f repeat %2d accepted %3d launches %d  %s" % (
        run, name, "+".join(kinds), n, tol, rep, int((decs > 0).sum()), launches, "ok" if ok else "MISMATCH"), flush=True)
print("soak: %d runs, %d mismatches" % (n_runs, bad))
sys.exit(1 if bad else 0)
