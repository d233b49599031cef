"""Five-method mirror of AtomicCoordinationNumberConstraint (fullrmc_b200/constraints_coordination.py; SURVEY.md
section 8f rank 3) replayed against trajectories of the UNMODIFIED reference class on the SiOx (non-periodic) and NiTi
(periodic) example inputs (tests/gen_golden_coordination_constraint.py).

CPU: the mirror with the oracle's counting functions injected -- pins the host arithmetic (halving, data - before +
after, the float32 standard error).  GPU: the same replay on the CUDA functions.  Data are integer counts (exact); the
standard error is compared exactly as well (same float32 expression, term by term)."""
import os

import numpy as np
import pytest

from gen_golden_atomic_coordination import unpack_lists
from gen_golden_coordination_constraint import unpack_csr

CASES = ("siox", "niti")


def _replay(case, golden_dir, kernels):
    from fullrmc_b200.constraints_coordination import DeviceAtomicCoordinationNumberConstraint
    g = np.load(os.path.join(golden_dir, "coordination_constraint_%s.npz" % case))
    box = g["boxCoords"].copy()
    c = DeviceAtomicCoordinationNumberConstraint(box, g["basis"], bool(g["isPBC"]), unpack_lists(g, "cores"), unpack_lists(g, "shells"),
                                                 g["lowerShells"], g["upperShells"], g["minAtoms"], g["maxAtoms"], g["weights"], kernels=kernels)
    # the per-atom definition lists the mirror derives are the reference's
    for mine, ref in ((c.asCoreDefIdxs, unpack_csr(g, "asCore")), (c.inShellDefIdxs, unpack_csr(g, "inShell"))):
        assert len(mine) == len(ref) and all(list(a) == list(b) for a, b in zip(mine, ref))
    data, err = c.compute_data()
    assert data.dtype == np.float32 and np.array_equal(data, g["start_data"])
    assert np.float32(err) == g["start_stdErr"]
    n_acc = 0
    for step in range(len(g["steps/idx"])):
        idx = g["steps/idx"][step:step + 1].astype(np.int32)
        moved = g["steps/moved"][step:step + 1]
        before = box.copy()
        c.compute_before_move(realIndexes=idx, relativeIndexes=idx)
        c.compute_after_move(realIndexes=idx, relativeIndexes=idx, movedBoxCoordinates=moved)
        assert np.array_equal(box, before)                         # the temporary write is undone
        assert np.float32(c.afterMoveStandardError) == g["steps/stdErr_after"][step], (case, step)
        if g["steps/accepted"][step]:
            c.accept_move(realIndexes=idx, relativeIndexes=idx)
            box[idx] = moved
            n_acc += 1
        else:
            c.reject_move(realIndexes=idx, relativeIndexes=idx)
        assert np.array_equal(c.data, g["steps/data"][step]), (case, step)
    assert 0 < n_acc < len(g["steps/idx"]) and c.accepted == n_acc and c.tried == len(g["steps/idx"])
    assert np.float32(c.standardError) == g["final_stdErr"]
    recount, _ = c.compute_data(update=False)
    assert np.array_equal(recount, g["final_recount"])


@pytest.mark.parametrize("case", CASES)
def test_mirror_host_arithmetic_on_oracle_counts(case, golden_dir):
    from oracle import coordination
    _replay(case, golden_dir, coordination)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_replays_reference_trajectory(case, golden_dir):
    from fullrmc_b200.Core import atomic_coordination
    _replay(case, golden_dir, atomic_coordination)
