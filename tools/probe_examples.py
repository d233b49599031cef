"""Debug probe: device time per step of run_generated on the shipped-example fixtures (tests/golden/generated_*.npz)."""
import os, sys
os.environ.setdefault("FRMC_BATCH_STAMPS", "1")          # phase timeline of the last launch (costs a few percent)
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fullrmc_b200
from test_generated_runs import _load, _device_store, _amplitude
F32 = np.float32
fullrmc_b200.set_edge_spill(True)
for name in sys.argv[1:] or ["niti", "niti_sf", "thf", "siox", "synth"]:
    g = _load(os.path.join(ROOT, "tests", "golden"), name)
    backend, cons = _device_store(g)
    st = backend.store
    total = F32(np.sum(backend._compute_data(), dtype=F32))
    n = 1500
    out = st.run_generated(200, 11, 0, _amplitude(g), total)
    s0 = st.batch_stats()
    out = st.run_generated(n, 11, 200, _amplitude(g), out["total"])
    s1 = st.batch_stats()
    print("%-8s %d atoms: %.2f us/step device, %d accepted of %d, %d launches, %d rounds (%.1f steps/round)" % (
        name, g["boxCoords"].shape[0], 1e3 * out["device_ms"] / n, int((out["decisions"] > 0).sum()), n, s1[0] - s0[0], s1[1] - s0[1],
        n / max(1, s1[1] - s0[1])))
    from fullrmc_b200 import _lib as L
    stp = np.zeros(4 + 5 * 64 + 64 * 128, np.int64)
    L.check(st._lib.frmc_store_batch_stamps(st._handle, stp.ctypes.data_as(L.c_i64p), stp.shape[0]), "stamps")
    print("   last launch: clear %.1f us, delta pass %.1f us, rounds+end %.1f us" % ((stp[1] - stp[0]) / 1e3, (stp[2] - stp[1]) / 1e3, (stp[3] - stp[2]) / 1e3))
    prev = stp[2]
    acc = {"own_epilogue": [], "wait_all": [], "decide": [], "commit": [], "commit_barrier": [], "round": []}
    for r in range(64):
        a = stp[4 + 5 * r: 9 + 5 * r]
        if a[0] == 0:
            break
        acc["own_epilogue"].append(a[0] - prev); acc["wait_all"].append(a[1] - a[0]); acc["decide"].append(a[2] - a[1])
        end = a[2]
        if a[3]:
            acc["commit"].append(a[3] - a[2]); acc["commit_barrier"].append(a[4] - a[3]); end = a[4]
        acc["round"].append(end - prev); prev = end
    print("   " + "  ".join("%s %.1f" % (k, np.mean(v) / 1e3) for k, v in acc.items() if v) + "  (us, mean over %d rounds)" % len(acc["round"]))
    backend.close()
