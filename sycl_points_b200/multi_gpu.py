"""Multi-GPU registration: one process per GPU (torch.distributed), source points sharded, target
cloud + index replicated on every rank.

Two exchange paths for the per-iteration sums row (21 H terms, 6 b terms, error, inlier count):

  mode="p2p"   the product path — spx_registration_align_sharded_*: ONE cooperative kernel per
               align per GPU; the row is stored into every peer's mailbox over NVLink from inside
               the kernel and folded in rank order (no collective launch, no host round trip);
  mode="nccl"  the baseline — shard_linearize -> torch.distributed.all_reduce -> shard_update, one
               collective per iteration on the queue's stream.

Both leave the identical pose on every rank.  The reference has no multi-device path at all
(SURVEY.md §2.2); what is mirrored here is Registration::align's contract
(I/algorithms/registration/registration.hpp:201-276) on a sharded source.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import RegistrationResultC, check
from .api import (DeviceQueue, ExecutionOptions, KDTree, PointCloudShared, Registration, RegistrationParams,
                  RegistrationResult, OptimizationMethod, _T16, _hostf, _ptr)

SUMS_LEN = _lib.SPX_SUMS_LEN
S_B, S_ERR, S_INL = 21, 27, 28  # layout of the sums row (include/spx.h, SPX_SUMS_LEN)


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous ceil(n / world) ranges; trailing ranks may be short or empty."""
    if world < 1:
        raise ValueError("world must be >= 1")
    step = -(-n // world) if n > 0 else 0
    return [(min(r * step, n), min((r + 1) * step, n)) for r in range(world)]


def shard_of(n: int, rank: int, world: int) -> tuple[int, int]:
    return shard_bounds(n, world)[rank]


def shard_indices(n: int, rank: int, world: int, block: int = 1024) -> np.ndarray:
    """Block-cyclic shard: blocks of `block` consecutive points dealt round-robin to the ranks.  Voxel-
    ordered clouds put dense ground and sparse walls in different contiguous ranges, so contiguous
    shards (shard_bounds) can be badly unbalanced in work; block-cyclic shards are not.  The sum over
    ranks is the same set of points either way."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    i = np.arange(n, dtype=np.int64)
    return i[(i // block) % world == rank]


def sums_to_Hb(s) -> tuple[np.ndarray, np.ndarray, float, int]:
    """Unpack a sums row into the symmetric 6x6 H, b, error, inlier count (float32 like the reference)."""
    s = np.asarray(s, np.float64)
    H = np.zeros((6, 6), np.float32)
    t = 0
    for a in range(6):
        for c in range(a, 6):
            H[a, c] = H[c, a] = np.float32(s[t])
            t += 1
    return H, s[S_B:S_B + 6].astype(np.float32), float(np.float32(s[S_ERR])), int(s[S_INL] + 0.5)


class Communicator:
    """spx_comm: this rank's NVLink mailbox, wired to the peers' through CUDA IPC handles that are
    all-gathered over the given torch.distributed group (any backend)."""

    def __init__(self, queue: DeviceQueue, rank: int | None = None, world: int | None = None, group=None):
        import torch.distributed as dist
        self.queue = queue
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        h = C.c_void_p()
        check(_lib.lib().spx_comm_create(queue.handle, self.rank, self.world, C.byref(h)))
        self._h = h
        if self.world > 1:
            buf = C.create_string_buffer(64)
            check(_lib.lib().spx_comm_ipc_handle(self._h, buf))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(buf.raw), group=group)
            check(_lib.lib().spx_comm_connect_ipc(self._h, b"".join(handles)))
            dist.barrier(group=group)  # nobody launches before every mailbox is mapped everywhere

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().spx_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalCommunicators:
    """One process driving several GPUs: a communicator per queue, wired by peer access."""

    def __init__(self, queues: list[DeviceQueue]):
        self.queues = queues
        self.handles = []
        for r, q in enumerate(queues):
            h = C.c_void_p()
            check(_lib.lib().spx_comm_create(q.handle, r, len(queues), C.byref(h)))
            self.handles.append(h)
        arr = (C.c_void_p * len(queues))(*[h.value for h in self.handles])
        check(_lib.lib().spx_comm_connect_local(arr, len(queues)))

    def close(self):
        for h in self.handles:
            _lib.lib().spx_comm_destroy(h)
        self.handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedAlignLoop:
    """The reduce-then-identical-update protocol of the sharded Gauss-Newton align, independent of
    where the per-shard sums come from (hooks).  `all_reduce(row)` must sum the row over ranks in
    place.  Every rank runs the same loop on the same reduced sums, so no pose broadcast is needed.
    Mirrors registration.hpp:227-272 with optimize_gauss_newton (:803-828)."""

    def __init__(self, max_iterations: int, all_reduce):
        self.max_iterations = max_iterations
        self.all_reduce = all_reduce

    # hooks -----------------------------------------------------------------
    def begin(self):
        raise NotImplementedError

    def linearize_shard(self):
        """-> the rank's partial sums row (anything all_reduce accepts)"""
        raise NotImplementedError

    def update(self, row) -> bool | None:
        """apply the GN step from the reduced row; return True to stop early (None = unknown: the
        device-resident variants keep launching and later launches no-op)"""
        raise NotImplementedError

    def finish(self):
        raise NotImplementedError

    def run(self):
        self.begin()
        for _ in range(self.max_iterations):
            row = self.linearize_shard()
            self.all_reduce(row)
            if self.update(row):
                break
        return self.finish()


class ShardedRegistration:
    """Registration over a source sharded across the ranks of a process group."""

    def __init__(self, queue: DeviceQueue, params: RegistrationParams | None = None, comm: Communicator | None = None,
                 mode: str = "p2p", group=None):
        if mode not in ("p2p", "nccl"):
            raise ValueError("mode must be 'p2p' or 'nccl'")
        self.queue = queue
        self.params = params if params is not None else RegistrationParams()
        if self.params.optimization_method != OptimizationMethod.GAUSS_NEWTON:
            raise RuntimeError("[ShardedRegistration] the sharded path is Gauss-Newton only")
        self.mode = mode
        self.group = group
        self.comm = comm
        if mode == "p2p" and comm is None:
            self.comm = Communicator(queue, group=group)
        self.reg = Registration(queue, self.params)
        self._sums = None

    def _common_args(self, source: PointCloudShared, target: PointCloudShared, tree: KDTree, T0, options):
        scale = options.robust_scale if options is not None else -1.0
        t16 = _T16(T0)
        return (source.points.ptr if source.size() else None, _ptr(source.covs) if source.has_cov() else None,
                source.size(), target.points.ptr, _ptr(target.covs) if target.has_cov() else None,
                _ptr(target.normals) if target.has_normal() else None, target.size(), tree.handle, _hostf(t16),
                float(scale)), t16

    def align(self, source_shard: PointCloudShared, target: PointCloudShared, target_knn: KDTree, initial_guess=None,
              options: ExecutionOptions | None = None) -> RegistrationResult:
        T0 = np.eye(4, dtype=np.float32) if initial_guess is None else np.asarray(initial_guess, np.float32)
        self.reg._loss()
        Pc = self.params.to_c()
        L = _lib.lib()
        check(L.spx_registration_set_params(self.reg._h, C.byref(Pc)))
        args, keep = self._common_args(source_shard, target, target_knn, T0, options)
        R = RegistrationResultC()
        if self.mode == "p2p":
            check(L.spx_registration_align_sharded_launch(self.reg._h, self.comm.handle, *args))
            check(L.spx_registration_align_sharded_finish(self.reg._h, C.byref(R)))
            return RegistrationResult.from_c(R)
        # NCCL baseline: the sums row lives in a torch tensor so that all_reduce can take it; the
        # queue must have been created on torch's current stream (DeviceQueue(dev, cuda_stream=...)).
        import torch
        import torch.distributed as dist
        if self._sums is None:
            self._sums = torch.zeros(SUMS_LEN, dtype=torch.float64, device=torch.device("cuda", self.queue.device))
        sums = self._sums
        reg_h = self.reg._h
        outer = self

        class _Loop(ShardedAlignLoop):
            def begin(self):
                check(L.spx_registration_shard_begin(reg_h, *args))

            def linearize_shard(self):
                check(L.spx_registration_shard_linearize(reg_h, C.c_void_p(sums.data_ptr())))
                return sums

            def update(self, row):
                check(L.spx_registration_shard_update(reg_h, C.c_void_p(row.data_ptr())))
                return None

            def finish(self):
                check(L.spx_registration_shard_finish(reg_h, C.byref(R)))
                return RegistrationResult.from_c(R)

        return _Loop(self.params.max_iterations, lambda t: dist.all_reduce(t, group=outer.group)).run()

    def last_timing(self) -> dict:
        return self.reg.last_timing()


# ------------------------------------------------------------------ sharded KNN / covariance
def all_gather_rows(local: np.ndarray, n_total: int, world: int, group=None) -> np.ndarray:
    """Concatenate the ranks' contiguous ceil(n / world) row shards (shard_bounds order) into the full
    array on every rank.  Host arrays through the process group's object gather: works on any backend
    (gloo on CPU, nccl)."""
    import torch.distributed as dist
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    out = np.concatenate([p for p in parts if len(p)], axis=0) if n_total else local[:0]
    assert len(out) == n_total, (len(out), n_total)
    return out


class ShardedKNN:
    """KNN with the QUERIES split over the ranks and the targets (+ their index) replicated — independent
    units, no data-path collective; results stay sharded unless `gather` is asked for.
    Per-shard contract = the single-GPU searches: knn_search_bruteforce (bruteforce.hpp:24-96) and
    KNNBase::knn_search (knn.hpp:22-24, kdtree.hpp:203-224): exact, ordered by (dist, index).

    `search(queries_rows, k) -> (idx, dist)` is the per-shard computation; the default runs the CUDA path
    (method "index" or "bruteforce") on this rank's queue, tests inject the oracle."""

    def __init__(self, rank: int, world: int, group=None, search=None):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("need 0 <= rank < world")
        self.rank, self.world, self.group = rank, world, group
        self._search = search

    @classmethod
    def on_device(cls, queue: DeviceQueue, targets: PointCloudShared, rank: int, world: int, group=None,
                  method: str = "index") -> "ShardedKNN":
        from .api import KNNResult, knn_search_bruteforce
        if method not in ("index", "bruteforce"):
            raise ValueError("method must be 'index' or 'bruteforce'")
        tree = KDTree.build(queue, targets) if method == "index" else None

        def search(rows: np.ndarray, k: int):
            qc = PointCloudShared(queue, rows)
            res = tree.knn_search(qc, k) if tree is not None else knn_search_bruteforce(queue, qc, targets, k)
            return res.indices_host(), res.distances_host()

        self = cls(rank, world, group, search)
        self.queue, self.targets, self.tree, self.method = queue, targets, tree, method
        return self

    def shard(self, n: int) -> tuple[int, int]:
        return shard_of(n, self.rank, self.world)

    def knn_search(self, queries: np.ndarray, k: int, gather: bool = False):
        """queries: the FULL (n, 4) host array, identical on every rank; this rank searches its rows.
        Returns (idx, dist, (lo, hi)) for the shard, or the gathered full arrays when gather=True."""
        lo, hi = self.shard(len(queries))
        idx, dist_ = self._search(np.ascontiguousarray(queries[lo:hi]), k)
        idx = np.asarray(idx, np.int32).reshape(hi - lo, k)
        dist_ = np.asarray(dist_, np.float32).reshape(hi - lo, k)
        if not gather:
            return idx, dist_, (lo, hi)
        return (all_gather_rows(idx, len(queries), self.world, self.group),
                all_gather_rows(dist_, len(queries), self.world, self.group), (0, len(queries)))


class ShardedCovariance:
    """covariance::estimate (covariance.hpp:16-47,260-311) with the POINTS split over the ranks: every rank
    holds the full cloud (the gathers need it) and its index, computes the k-neighbour covariances of its own
    contiguous range, and — when the next stage needs them replicated (the target side of GICP) — all-gathers
    the 64-byte matrices.  `compute(points, lo, hi, k) -> (hi-lo, 4, 4)` is injectable like ShardedKNN.search."""

    def __init__(self, rank: int, world: int, group=None, compute=None):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("need 0 <= rank < world")
        self.rank, self.world, self.group = rank, world, group
        self._compute = compute

    @classmethod
    def on_device(cls, queue: DeviceQueue, rank: int, world: int, group=None) -> "ShardedCovariance":
        from .api import covariance

        def compute(points: np.ndarray, lo: int, hi: int, k: int):
            from .api import DeviceArray
            cloud = PointCloudShared(queue, points)            # the full cloud, replicated: the gathers read it
            tree = KDTree.build(queue, cloud)
            nn = tree.knn_search(PointCloudShared(queue, points[lo:hi]), k)  # this rank's rows against the full cloud
            covs = DeviceArray(queue, (hi - lo, 16), np.float32)
            if hi > lo:
                check(_lib.lib().spx_covariance(queue.handle, cloud.points.ptr, hi - lo, nn.indices.ptr, k, covs.ptr))
            tree.close()
            return covs.download().reshape(-1, 4, 4).transpose(0, 2, 1).copy()

        return cls(rank, world, group, compute)

    def estimate(self, points: np.ndarray, k: int, gather: bool = True):
        lo, hi = shard_of(len(points), self.rank, self.world)
        covs = np.asarray(self._compute(points, lo, hi, k), np.float32).reshape(hi - lo, 4, 4)
        if not gather:
            return covs, (lo, hi)
        return all_gather_rows(covs, len(points), self.world, self.group), (0, len(points))
