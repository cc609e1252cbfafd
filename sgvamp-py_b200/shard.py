"""Row-partitioned multi-GPU plumbing (one process or thread per GPU).

The data path needs no host collective: every reduction of the solver is completed across ranks
inside the CUDA kernels through peer memory (include/sgvamp_b200.h, "multi-GPU").  The host side
only has to (1) split the marker rows, (2) hand every rank the others' arena handles once, and
(3) gather the per-rank output slices at the end.  Those three things are what this module does;
it is pure host logic and is tested on CPU with gloo (tests/test_shard_cpu.py).

A ``Shard`` offers ``rank``, ``world``, ``allgather(obj) -> list`` and ``barrier()``.
"""
from __future__ import annotations

import threading

import numpy as np


def partition_rows(M, world, align=4):
    """Contiguous, balanced row ranges; interior boundaries are multiples of ``align``."""
    b = [0]
    for r in range(1, world):
        x = (M * r // world) // align * align
        b.append(max(x, b[-1]))
    b.append(M)
    return [(b[r], b[r + 1]) for r in range(world)]


def block_starts(indptr, indices):
    """Start rows of the diagonal blocks of a symmetric block-diagonal CSR matrix (plus M at the end):
    a new block starts at row i when no earlier row reaches column i or beyond and no later row
    reaches back before i."""
    M = len(indptr) - 1
    indptr = np.asarray(indptr, dtype=np.int64)
    lo = np.full(M, np.iinfo(np.int64).max, dtype=np.int64)
    hi = np.full(M, -1, dtype=np.int64)
    nz = np.diff(indptr) > 0
    first, last = indptr[:-1][nz], indptr[1:][nz] - 1
    idx = np.asarray(indices, dtype=np.int64)
    # rows are not assumed sorted: use reduceat over the row segments
    if idx.size:
        lo[nz] = np.minimum.reduceat(idx, first)
        hi[nz] = np.maximum.reduceat(idx, first)
    lo = np.minimum(lo, np.arange(M))
    hi = np.maximum(hi, np.arange(M))
    runmax = np.maximum.accumulate(hi)
    sufmin = np.minimum.accumulate(lo[::-1])[::-1]
    cut = np.ones(M, dtype=bool)
    cut[1:] = (runmax[:-1] < np.arange(1, M)) & (sufmin[1:] >= np.arange(1, M))
    return np.concatenate([np.flatnonzero(cut), [M]]).astype(np.int64)


def partition_blocks(starts, world):
    """Contiguous assignment of LD blocks to ranks balanced by stored values (sum of m_b^2): returns
    the row range of each rank; every boundary is a block boundary (scalar-only exchange)."""
    starts = np.asarray(starts, dtype=np.int64)
    m = np.diff(starts).astype(np.float64)
    cost = np.concatenate([[0.0], np.cumsum(m * m)])
    nb = len(m)
    if nb < world:
        raise Exception("%d LD blocks cannot be sharded over %d ranks" % (nb, world))
    cuts = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cost, cost[-1] * r / world))
        b = min(max(b, cuts[-1] + 1), nb - (world - r))
        # pick the nearer of the two neighbouring block boundaries
        if b - 1 > cuts[-1] and abs(cost[b - 1] - cost[-1] * r / world) < abs(cost[b] - cost[-1] * r / world):
            b -= 1
        cuts.append(b)
    cuts.append(nb)
    return [(int(starts[cuts[r]]), int(starts[cuts[r + 1]])) for r in range(world)]


def slice_rows_csr(R, lo, hi):
    """Rows [lo, hi) of a scipy CSR matrix with GLOBAL column indices (what a rank uploads)."""
    sub = R[lo:hi]
    return sub.indptr.astype(np.int64), sub.indices.astype(np.int32), sub.data


def local_bandwidth(indptr, indices, lo, sorted_indices=False):
    """max |col - row| over the local rows (global column indices).  With sorted column indices only
    the first and last entry of every row are looked at (O(rows), not O(nnz))."""
    if len(indices) == 0:
        return 0
    indptr = np.asarray(indptr, dtype=np.int64)
    n = len(indptr) - 1
    if sorted_indices:
        nz = np.flatnonzero(np.diff(indptr) > 0)
        rows = nz + lo
        first = np.asarray(indices)[indptr[nz]].astype(np.int64)
        last = np.asarray(indices)[indptr[nz + 1] - 1].astype(np.int64)
        return int(max((rows - first).max(), (last - rows).max(), 0))
    rows = np.repeat(np.arange(n, dtype=np.int64) + lo, np.diff(indptr))
    return int(np.abs(np.asarray(indices).astype(np.int64) - rows).max())


def gather_rows(shard, local, bounds):
    """Concatenate per-rank row slices (any leading dims, rows last) into the global array."""
    parts = shard.allgather(np.ascontiguousarray(local))
    return np.concatenate(parts, axis=-1)


class SoloShard:
    rank, world = 0, 1

    def allgather(self, obj):
        return [obj]

    def barrier(self):
        pass


class TorchShard:
    """torch.distributed as the plumbing (NCCL or gloo default group); objects travel pickled."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allgather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)


class TorchComm:
    """The `comm` duck type of the reference solver (`Get_rank`, `Get_size`, `bcast(obj, root=)`;
    src/sgvamp.py:202,230-233) over torch.distributed, for the reference's deployment shape
    "one rank per cohort" on a multi-GPU box: `torchrun --nproc-per-node K main.py ...` (NCCL on GPUs,
    gloo on CPU).  ndarrays travel as tensors (on the GPU with NCCL), anything else pickled."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.world

    def bcast(self, obj, root=0):
        # every rank knows whether a vector or a scalar travels (src/sgvamp.py:232 scalar gam1, :233 the r1 vector),
        # but only the root has the value: the header (kind, shape) goes first
        head = [None]
        if self.rank == root:
            head[0] = ("nd", obj.shape, str(obj.dtype)) if isinstance(obj, np.ndarray) else ("py",)
        self.dist.broadcast_object_list(head, src=root, group=self.group)
        if head[0][0] == "py":
            box = [obj if self.rank == root else None]
            self.dist.broadcast_object_list(box, src=root, group=self.group)
            return box[0]
        _, shape, dtype = head[0]
        if self.rank == root:
            t = self.torch.from_numpy(np.ascontiguousarray(obj)).to(self.device)
        else:
            t = self.torch.empty(shape, dtype=getattr(self.torch, dtype), device=self.device)
        self.dist.broadcast(t, src=root, group=self.group)
        return t.cpu().numpy()


class _DevArray:
    """A raw device pointer as a CUDA-array-interface object (zero-copy view for torch.as_tensor)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _allgather_r1(self, handle, gam1_mine):
    """Device-side form of the reference's 2K broadcasts (src/sgvamp.py:228-233) for one rank per cohort on GPUs: ONE
    NCCL all-gather writes every cohort's r1 straight into the library's K x M block of r1 vectors (in place: this rank's
    slice is already there), a second one collects the K gam1 scalars.  Returns gam1s (host), which also orders the r1
    gather before the caller's next enqueue.  None when the communicator is not NCCL on GPUs."""
    if self.dist.get_backend(self.group) != "nccl":
        return None
    torch = self.torch
    ptr, stride = handle.r1_block()
    K = self.world
    dev = torch.device("cuda", torch.cuda.current_device())
    blk = torch.as_tensor(_DevArray(ptr, K * stride), device=dev)
    handle.sync()                                        # the handle's stream may not be torch's current stream
    self.dist.all_gather_into_tensor(blk, blk[self.rank * stride:(self.rank + 1) * stride], group=self.group)
    g = torch.tensor([float(gam1_mine)], dtype=torch.float64, device=dev)
    out = torch.empty(K, dtype=torch.float64, device=dev)
    self.dist.all_gather_into_tensor(out, g, group=self.group)
    return out.cpu().numpy()                             # synchronises: the r1 gather (same stream, earlier) is complete


TorchComm.allgather_r1 = _allgather_r1


class ThreadShard:
    """In-process ranks (one host thread per GPU) for tests: barrier-based allgather."""

    class _Group:
        def __init__(self, world):
            self.world = world
            self.bar = threading.Barrier(world, timeout=180)   # a rank that died must not hang the others
            self.slots = [None] * world

    def __init__(self, group, rank):
        self.g, self.rank, self.world = group, rank, group.world

    @staticmethod
    def make(world):
        g = ThreadShard._Group(world)
        return [ThreadShard(g, r) for r in range(world)]

    def allgather(self, obj):
        self.g.slots[self.rank] = obj
        self.g.bar.wait()
        out = list(self.g.slots)
        self.g.bar.wait()
        return out

    def barrier(self):
        self.g.bar.wait()


def attach_peers(handle, shard):
    """Give every rank a mapping of every other rank's symmetric arena (CUDA IPC between
    processes, direct peer access between threads of one process).

    Ranks normally own one GPU each and complete cross-rank reductions inside their kernels.  If
    two ranks share a GPU (in-process test ranks on a box with fewer GPUs than ranks) their kernels
    are not guaranteed to run concurrently, so all ranks switch to completing the reductions with
    a host barrier (sgv_set_host_barrier); separate processes sharing a GPU are refused."""
    if shard.world == 1:
        return
    local = isinstance(shard, ThreadShard)
    if local:
        peers = shard.allgather((handle, handle.M, handle.device_id()))
    else:
        peers = shard.allgather((handle.ipc_export(), handle.M, handle.device_id()))
    devs = [p[2] for p in peers]
    shared = len(set(devs)) < len(devs)
    if shared and not local:
        raise Exception("ranks %s share a GPU: the multi-process row partition needs one GPU per rank" % devs)
    for q, (hq, rows, _dev) in enumerate(peers):
        if q != shard.rank:
            if local:
                handle.peer_attach_local(q, hq)
            else:
                handle.ipc_import(q, hq, rows)
    if local:
        handle.set_host_barrier(shared)
    shard.barrier()
