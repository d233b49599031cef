"""One-process-per-GPU plumbing for the sharded full histogram (torch.distributed, NCCL on the
GPU box, gloo in the CPU tests).

The full-system histogram shards over the tile work list (csrc/fullhist.cu: work item index
modulo world size); every rank owns a replica of the 16 B/atom store, computes its partial
INTEGER histogram and one all-reduce(sum) of 2*nEl^2*hs int64 counters (400 KB for the 1M-atom
config) makes every rank hold the full result, bit-identical for any world size.  The per-move
path does not shard: replicas only.
"""
import os

import numpy as np


def rank_world():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def allreduce_sum_(tensor):
    """In-place sum over ranks when a process group with more than one rank is up; no-op otherwise."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor)
    return tensor


class _DevArray(object):
    """zero-copy view of a raw device pointer for torch.as_tensor"""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def counts_tensor(store, grid):
    """The grid's committed int64 counts [2][nEl*nEl][hs] as a torch CUDA tensor aliasing the store."""
    import torch
    ptr, n = store.counts_pointer(grid)
    return torch.as_tensor(_DevArray(ptr, n), device="cuda:%d" % store.device)


def compute_data_sharded(store, rank=None, world=None, tensors=None, events=None):
    """compute_data over all ranks: this rank's tile shard, all-reduce of every grid's counts on the
    store's stream, then the epilogue.  Returns chi^2 per model (identical on every rank).
    events: optional list that receives one (start, end) pair of CUDA events around the all-reduce
    (recorded on the store's stream; the caller reads them after a synchronisation)."""
    import torch
    import torch.distributed as dist
    r, w, _ = rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    if world > 1 and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == world):
        # a partial histogram must never reach the epilogue: chi^2 would be plausible and wrong
        raise RuntimeError("compute_data_sharded(world=%d) needs an initialised process group of that size "
                           "(torch.distributed.init_process_group)" % world)
    store.compute_data_shard(rank, world)
    if world > 1:
        ext = torch.cuda.ExternalStream(store.stream, device=store.device)
        with torch.cuda.stream(ext):
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(ext)
            for g in range(len(store._grids)):
                allreduce_sum_(tensors[g] if tensors is not None else counts_tensor(store, g))
            if events is not None:
                e1.record(ext)
                events.append((e0, e1))
    return store.finalize_data()


def shard_pairs(n_atoms, elementIndex, numberOfElements, shard, nshards, sm_count=148):
    """(work items, atom pairs) the tile scheduler gives to one shard -- host-only, no GPU needed."""
    import ctypes
    from . import _lib as L
    lib = L.load_library()
    el = np.ascontiguousarray(elementIndex, dtype=np.int32)
    ni, npairs = ctypes.c_int64(0), ctypes.c_int64(0)
    L.check(lib.frmc_debug_work_items(int(n_atoms), L.ptr(el, L.c_i32p), int(numberOfElements), int(shard), int(nshards),
                                      int(sm_count), ctypes.byref(ni), ctypes.byref(npairs)), "debug_work_items")
    return int(ni.value), int(npairs.value)
