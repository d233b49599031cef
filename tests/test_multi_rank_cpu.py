"""CPU tests of the N>1 path: tile-list sharding (host logic of the CUDA library, no device
needed) and the all-reduce plumbing with a world_size-2 gloo process group."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,nEl", [(1, 1), (300, 2), (5000, 3), (70000, 5)])
def test_shards_partition_the_upper_triangle(n, nEl):
    from fullrmc_b200 import parallel
    rng = np.random.default_rng(n)
    el = rng.integers(0, nEl, n).astype(np.int32)
    whole_items, whole_pairs = parallel.shard_pairs(n, el, nEl, 0, 1)
    assert whole_pairs == n * (n - 1) // 2
    for nshards in (2, 3, 8):
        items = pairs = 0
        per = []
        for shard in range(nshards):
            ni, npr = parallel.shard_pairs(n, el, nEl, shard, nshards)
            items += ni; pairs += npr; per.append(npr)
        assert items == whole_items and pairs == whole_pairs
        if n >= 70000:      # enough work items for the interleaving to balance: within 5 % of the mean
            assert max(per) <= 1.05 * (whole_pairs / nshards)


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fullrmc_b200 import parallel
    from oracle import pairhist as orc
    import cases as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert parallel.rank_world() == (rank, world, rank)
    case = [c for c in C.make_cases() if c["name"] == "tri_molecular"][0]
    kw = {k: case[k] for k in ("basis", "isPBC", "moleculeIndex", "elementIndex", "numberOfElements", "minDistance",
                               "maxDistance", "bin", "histSize")}
    n = case["boxCoords"].shape[0]
    rows = np.arange(rank, n, world, dtype=np.int32)          # this rank's rows of the upper triangle
    hi, he = orc.multiple_pairs_histograms_coords(indexes=rows, boxCoords=case["boxCoords"], allAtoms=False, **kw)
    counts = torch.from_numpy(np.stack([hi, he]).astype(np.int64))
    parallel.allreduce_sum_(counts)                           # the collective the GPU path issues over NCCL
    np.save(os.path.join(tmpdir, "rank%d.npy" % rank), counts.numpy())
    dist.destroy_process_group()


def test_integer_allreduce_of_partial_histograms_gloo(tmp_path, orc):
    import torch.multiprocessing as mp
    import cases as C
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    case = [c for c in C.make_cases() if c["name"] == "tri_molecular"][0]
    kw = {k: case[k] for k in ("basis", "isPBC", "moleculeIndex", "elementIndex", "numberOfElements", "minDistance",
                               "maxDistance", "bin", "histSize")}
    fi, fe = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"], **kw)
    want = np.stack([fi, fe]).astype(np.int64)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert np.array_equal(got, want)          # every rank holds the full histogram, bit-identical
