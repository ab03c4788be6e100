"""Particle shards across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Rank g owns particles
[g*n_local, (g+1)*n_local); the map, the polar table and the scan are replicated (the scan is
rasterised redundantly on every rank — cheaper than a broadcast).  Per step the shards are all-gathered in two
pieces: raw weight + last_dist (8 B/particle) first, then the seven state rows (28 B/particle) on the
communication stream WHILE every rank normalises the N weights in GLOBAL order (redundantly, so weights and
resampled indices are bit-identical for every world size); when the states have arrived the rank builds the
order-exact prefix, draws its own slice [i0, i1) of the M systematic samples and gathers those states out of
the gathered rows.  (ShardedFilter.step_single_collective is the same update with ONE all-gather of the whole
36 B/particle block in front of the normalisation — 0.45 ms at 8 x 1e6 particles that nothing hides.)

The reference has no counterpart (its only parallelism is for_each(par) over particles,
particle_filter.cpp:104); the semantics are those of ParticleFilter::update (:94-189) on the
concatenated particle set.
"""
from __future__ import annotations

SHARD_ROWS = 9   # TDR_SHARD_ROWS (include/tdr.h)


def shard_range(n_total: int, rank: int, world: int):
    """particles owned by `rank` when n_total particles are dealt in contiguous, near-equal shards"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_slice(M: int, rank: int, world: int):
    """outputs [i0, i1) of the M systematic samples drawn by `rank` (contiguous; systematic resampling is
    monotone, so a contiguous output slice reads a contiguous source range)"""
    return shard_range(M, rank, world)


class ShardedFilter:
    """ParticleFilter::update over particle shards.  `ctx` holds this rank's shard; `stream` is the
    torch.cuda.ExternalStream wrapping ctx.stream so the NCCL collective is ordered with the kernels."""

    def __init__(self, ctx, stream, rank: int, world: int, group=None):
        import torch
        self.torch, self.ctx, self.stream, self.rank, self.world, self.group = torch, ctx, stream, rank, world, group
        self.send = self.recv = None
        if hasattr(ctx, "pf_set_shard_count"):
            ctx.pf_set_shard_count(world)      # kernel choices by the global particle count: results independent of `world`

    def _buffers(self, n_local):
        torch = self.torch
        if self.send is None or self.send.numel() != SHARD_ROWS * n_local:
            dev = torch.device("cuda", self.ctx.device)
            self.send = torch.empty(SHARD_ROWS * n_local, dtype=torch.float32, device=dev)
            self.recv = torch.empty(self.world * SHARD_ROWS * n_local, dtype=torch.float32, device=dev)

    def _all_gather(self, n_local, with_weights):
        import torch.distributed as dist
        self._buffers(n_local)
        self.ctx.pf_export_shard(self.send.data_ptr(), self.send.numel(), with_weights)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)

    def step(self, res, ang_res, n_theta, n_r, u, M_total):
        """render + score the local shard; all-gather weights + last_dist; all-gather the states behind the global
        normalisation; resample this rank's slice"""
        import torch.distributed as dist
        torch, ctx = self.torch, self.ctx
        n_local = ctx.pf_count()
        assert M_total % self.world == 0, "equal shards: the particle count must be a multiple of the number of ranks"
        i0, i1 = sample_slice(M_total, self.rank, self.world)
        if getattr(self, "_split_n", None) != n_local:
            dev = torch.device("cuda", ctx.device)
            self.wl = torch.empty(2 * n_local, dtype=torch.float32, device=dev)
            self.st = torch.empty(7 * n_local, dtype=torch.float32, device=dev)
            self.wl_all = torch.empty(self.world * 2 * n_local, dtype=torch.float32, device=dev)
            self.st_all = torch.empty(self.world * 7 * n_local, dtype=torch.float32, device=dev)
            self._split_n = n_local
        with torch.cuda.stream(self.stream):
            ctx.scan_render_polar(res, ang_res, n_theta, n_r, want=False)
            ctx.pf_score(res, want=False)
            ctx.pf_export_split(self.wl.data_ptr(), self.st.data_ptr())
            dist.all_gather_into_tensor(self.wl_all, self.wl, group=self.group)
            work = dist.all_gather_into_tensor(self.st_all, self.st, group=self.group, async_op=True)   # NCCL's stream
            ctx.pf_normalize_gathered(self.wl_all.data_ptr(), self.world, n_local)                      # overlaps it
            work.wait()                                                                                 # context stream waits
            ctx.pf_resample_gathered(self.st_all.data_ptr(), self.world, n_local, u, M_total, i0, i1)

    def step_single_collective(self, res, ang_res, n_theta, n_r, u, M_total):
        """render + score the local shard, all-gather, normalise globally, resample this rank's slice"""
        ctx = self.ctx
        n_local = ctx.pf_count()
        i0, i1 = sample_slice(M_total, self.rank, self.world)
        with self.torch.cuda.stream(self.stream):
            ctx.scan_render_polar(res, ang_res, n_theta, n_r, want=False)
            ctx.pf_score(res, want=False)
            self._all_gather(n_local, True)
            ctx.pf_update_gathered(self.recv.data_ptr(), self.world, n_local, u, M_total, i0, i1)

    def pose(self, want_ml=True):
        """mean / covariance / ML pose over the whole (all-gathered) resampled set; needs equal shards"""
        ctx = self.ctx
        n_local = ctx.pf_count()
        with self.torch.cuda.stream(self.stream):
            self._all_gather(n_local, False)
            return ctx.pf_pose_gathered(self.recv.data_ptr(), self.world, n_local, want_ml)


class LibraryShardedFilter:
    """The same update with everything below the C ABI (csrc/shard.cu: tdr_shard_*): the library owns the NCCL
    communicator (created from an id that rank 0 makes and torch.distributed merely carries to the other ranks) and the
    peer-mapped state slots.  Per scan: ONE all-gather of 8 B per particle; resampled states are read from the owning
    rank over NVLink.  What a C++ node would call; this class is the harness over it."""

    def __init__(self, ctx, rank: int, world: int, particles_per_rank: int, group=None):
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, world
        box = [ctx.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        ctx.shard_init(rank, world, box[0], particles_per_rank)

    def step(self, res, ang_res, n_theta, n_r, u, M_total):
        assert M_total % self.world == 0, "equal shards: the particle count must be a multiple of the number of ranks"
        self.ctx.shard_step(res, ang_res, n_theta, n_r, u, M_total)

    def pose(self, want_ml=True):
        return self.ctx.shard_pose(want_ml)

    def close(self):
        self.ctx.shard_finalize()


class FusedGridGather:
    """cfg4: the exhaustive grid sharded over the GPUs of one box with the weight all-gather FUSED into the score
    kernel.  Every rank owns a full (n_total x n_shifts) cost array that its peers have mapped through CUDA IPC
    over NVLink; the kernel's epilogue stores each cost of the local shard into all of them, so when the kernels
    have finished every rank holds every cost — no NCCL all-gather, the transfer overlaps the gather / MMA pipeline
    tile by tile.  One tiny all-reduce on the context stream is the cross-rank barrier."""

    def __init__(self, ctx, rank: int, world: int, n_total: int, n_shifts: int, group=None):
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.n_total, self.n_shifts = n_total, n_shifts
        self.full_ptr, handle = ctx.grid_peer_alloc(n_total * n_shifts)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        ptrs = [self.full_ptr if r == rank else ctx.grid_peer_open(handles[r]) for r in range(world)]
        self.lo, self.hi = shard_range(n_total, rank, world)
        ctx.grid_peer_set(ptrs, self.lo)

    def numel(self):
        return self.n_total * self.n_shifts

    def key_tensor(self, torch, device):
        """the context's 8-byte (min cost, first global flat index) key as a one-element int64 torch tensor (an alias,
        no copy): `dist.all_reduce(t, op=MIN)` on the context stream is the cross-rank barrier of the fused all-gather
        AND the arg-min reduction; decode the result with Context.grid_key_decode."""
        ptr, _ = self.ctx.dev_ptr(4)                  # TDR_BUF_GRID_BEST_KEY

        class _Alias:
            __cuda_array_interface__ = {"shape": (1,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Alias(), device=device)

    def close(self):
        self.ctx.grid_peer_clear()


# ---- layout helpers (numpy): the shard block exactly as k_pack_shard / k_unpack_all lay it out.
# Used by the CPU (gloo) tests of the protocol and as executable documentation of the wire format.
_ROWS = ("weight", "init_x_px", "init_y_px", "dx_m", "dy_m", "theta", "scale", "have_init", "last_dist")


def pack_block_numpy(states, last_dist, weights):
    import numpy as np
    n = len(states)
    blk = np.empty((SHARD_ROWS, n), dtype=np.float32)
    blk[0] = weights if weights is not None else 0
    for k, name in enumerate(_ROWS[1:7], start=1):
        blk[k] = states[name]
    blk[7] = (states["have_init"] != 0).astype(np.float32)
    blk[8] = last_dist
    return blk.reshape(-1)


def unpack_blocks_numpy(gathered, world, n_local, state_dtype):
    """gathered: world blocks in rank order -> (states[N], last_dist[N], weights[N]) in global particle order"""
    import numpy as np
    g = np.asarray(gathered, dtype=np.float32).reshape(world, SHARD_ROWS, n_local)
    N = world * n_local
    st = np.zeros(N, dtype=state_dtype)
    for k, name in enumerate(_ROWS[1:7], start=1):
        st[name] = g[:, k, :].reshape(N)
    st["have_init"] = (g[:, 7, :].reshape(N) != 0).astype(np.uint8)
    return st, g[:, 8, :].reshape(N).copy(), g[:, 0, :].reshape(N).copy()


# the two-collective wire format of ShardedFilter.step (k_pack_split / k_unpack_wl / k_unpack_states):
#   wl[2][n]     = raw weight, last_dist                                  (all-gathered first: all the normalisation needs)
#   states[7][n] = init_x, init_y, dx, dy, theta, scale, have_init as 0/1 (all-gathered behind the normalisation)
def pack_split_numpy(states, last_dist, weights):
    import numpy as np
    n = len(states)
    wl = np.empty((2, n), dtype=np.float32)
    wl[0], wl[1] = weights, last_dist
    st = np.empty((7, n), dtype=np.float32)
    for k, name in enumerate(_ROWS[1:7]):
        st[k] = states[name]
    st[6] = (states["have_init"] != 0).astype(np.float32)
    return wl.reshape(-1), st.reshape(-1)


def unpack_split_numpy(wl_all, st_all, world, n_local, state_dtype):
    """gathered pieces in rank order -> (states[N], last_dist[N], weights[N]) in global particle order"""
    import numpy as np
    N = world * n_local
    wl = np.asarray(wl_all, dtype=np.float32).reshape(world, 2, n_local)
    g = np.asarray(st_all, dtype=np.float32).reshape(world, 7, n_local)
    st = np.zeros(N, dtype=state_dtype)
    for k, name in enumerate(_ROWS[1:7]):
        st[name] = g[:, k, :].reshape(N)
    st["have_init"] = (g[:, 6, :].reshape(N) != 0).astype(np.uint8)
    return st, wl[:, 1, :].reshape(N).copy(), wl[:, 0, :].reshape(N).copy()
