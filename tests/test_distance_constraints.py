"""Golden trajectories of the UNMODIFIED reference InterMolecularDistanceConstraint / IntraMolecularDistanceConstraint
(tests/gen_golden_distance_constraints.py, THF and SiOx example inputs) replayed through the device-backed mirror
(fullrmc_b200.constraints_distance): on the CPU with the oracle kernels (pins the mirror's host arithmetic), on the
GPU with the CUDA kernels.  Bar: data arrays, standard errors and rejection flags identical at every step."""
import os

import numpy as np
import pytest

NAMES = ["thf_inter", "thf_intra", "siox_inter"]
F32 = np.float32


def _replay(g, kernels):
    from fullrmc_b200.constraints_distance import DeviceMolecularDistanceConstraint
    box = g["boxCoords"].copy()
    c = DeviceMolecularDistanceConstraint(box, g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["typesIndex"], int(g["numberOfTypes"]),
                                          g["lowerLimitArray"], g["upperLimitArray"], g["typePairsIndex"],
                                          interMolecular=bool(g["interMolecular"]), flexible=bool(g["flexible"]), kernels=kernels)
    data, err = c.compute_data()
    assert np.array_equal(data["number"], g["start_number"]) and np.array_equal(data["distanceSum"], g["start_distanceSum"])
    assert F32(err) == F32(g["start_stdErr"])
    for s in range(g["steps/idx"].shape[0]):
        k = int(g["steps/k"][s])
        idx = g["steps/idx"][s, :k].astype(np.int32)
        moved = np.ascontiguousarray(g["steps/moved"][s, :k])
        c.compute_before_move(idx, idx)
        c.compute_after_move(idx, idx, moved)
        assert F32(c.afterMoveStandardError) == F32(g["steps/stdErr_after"][s]), "step %d" % s
        assert c.should_step_get_rejected(c.afterMoveStandardError) == bool(g["steps/rejected"][s]), "step %d" % s
        if bool(g["steps/accepted"][s]):
            c.accept_move(idx, idx)
            box[idx] = moved                                          # the engine moves the atoms (Engine.py:3337-3338)
        else:
            c.reject_move(idx, idx)
        assert np.array_equal(c.data["number"], g["steps/number"][s]) and np.array_equal(c.data["distanceSum"], g["steps/distanceSum"][s])
    assert F32(c.standardError) == F32(g["final_stdErr"])
    assert np.array_equal(c._get_constraint_value(), g["final_value"])


@pytest.mark.parametrize("name", NAMES)
def test_mirror_with_oracle_kernels_reproduces_reference_classes(name, golden_dir, orc):
    _replay(np.load(os.path.join(golden_dir, "distance_constraints_%s.npz" % name)), orc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_mirror_on_device_reproduces_reference_classes(name, golden_dir):
    from fullrmc_b200.Core import atomic_distances
    _replay(np.load(os.path.join(golden_dir, "distance_constraints_%s.npz" % name)), atomic_distances)
