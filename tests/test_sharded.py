"""Multi-GPU host logic on the CPU: world_size-2 (and 3) gloo runs of the shard / all-gather / slice protocol of
top_down_renderer_b200/sharded.py, with the oracle standing in for the device kernels.  The property under test:
the sharded update yields the SAME weights, indices and resampled states as the single-process update."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from top_down_renderer_b200 import sharded, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_and_slice_ranges_partition():
    for n in (1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [sharded.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
            assert sharded.sample_slice(n, 0, world) == r[0]


def test_block_layout_roundtrip():
    st, ld = synth.particles_tracking(10, (50.0, 60.0), 0.3, seed=1)
    st["have_init"][::3] = 0
    w = np.arange(10, dtype=np.float32)
    blk = sharded.pack_block_numpy(st, ld, w)
    assert blk.shape == (sharded.SHARD_ROWS * 10,)
    st2, ld2, w2 = sharded.unpack_blocks_numpy(np.concatenate([blk, blk]), 2, 10, synth.STATE_DTYPE)
    assert np.array_equal(st2[:10], st) and np.array_equal(st2[10:], st) and np.array_equal(ld2[:10], ld)
    assert np.array_equal(w2[10:], w)


def _world():
    C, H, W = 4, 200, 240
    cm = synth.make_class_map(H, W, C, seed=31)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C)
    ang = np.float32(2 * np.pi / 100)
    pose, heading = synth.default_pose(cm, seed=31)
    pts = synth.make_scan(cm, pose, heading, seed=31, n_rings=16, n_az=128)
    layers, mask = orc.compute_dists(orc.class_image_to_layers(img, lut, C, 1.0), 1.0)
    scan = orc.render_polar(pts, 2.0, ang, 100, 25, lut, C)
    tab = orc.polar_table(100, 25, ang, 1.0)
    thetas, shifts = orc.search_list(100)
    fp = orc.make_params(C, regularization=0.7, map_width=W, map_height=H)
    return dict(layers=layers, mask=mask, scan=scan, tab=tab, thetas=thetas, shifts=shifts, fp=fp, pose=pose, heading=heading)


def _single(n_total, u, M):
    wd = _world()
    st, ld = synth.particles_tracking(n_total, wd["pose"], wd["heading"], seed=5)
    st["have_init"][::4] = 0
    w = orc.score_all(st, wd["fp"], wd["layers"], wd["mask"], 1.0, wd["tab"], 100, 25, wd["scan"], 2.0, wd["thetas"], wd["shifts"], n_threads=1)
    wn, arg, _ = orc.normalize(w, ld)
    idx = orc.resample_fast(wn, u, M)
    return wn, arg, idx, st[idx], ld[idx]


def test_split_layout_roundtrip():
    st, ld = synth.particles_tracking(10, (50.0, 60.0), 0.3, seed=1)
    st["have_init"][::3] = 0
    w = np.arange(10, dtype=np.float32)
    wl, sb = sharded.pack_split_numpy(st, ld, w)
    assert wl.shape == (20,) and sb.shape == (70,)
    st2, ld2, w2 = sharded.unpack_split_numpy(np.concatenate([wl, wl]), np.concatenate([sb, sb]), 2, 10, synth.STATE_DTYPE)
    assert np.array_equal(st2[:10], st) and np.array_equal(st2[10:], st) and np.array_equal(ld2[10:], ld) and np.array_equal(w2[:10], w)


def _rank_main_split(rank, world, port, n_local, u, M, q):
    """the two-collective step of ShardedFilter.step: weights + last_dist first, the states asynchronously while the
    global normalisation runs, joined before the resampling"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wd = _world()
        n_total = n_local * world
        st_all, ld_all = synth.particles_tracking(n_total, wd["pose"], wd["heading"], seed=5)
        st_all["have_init"][::4] = 0
        lo, hi = sharded.shard_range(n_total, rank, world)
        st, ld = st_all[lo:hi].copy(), ld_all[lo:hi].copy()
        w = orc.score_all(st, wd["fp"], wd["layers"], wd["mask"], 1.0, wd["tab"], 100, 25, wd["scan"], 2.0, wd["thetas"], wd["shifts"], n_threads=1)
        wl, sb = (torch.from_numpy(a) for a in sharded.pack_split_numpy(st, ld, w))
        wl_all = torch.empty(world * wl.numel(), dtype=torch.float32)
        st_all_t = torch.empty(world * sb.numel(), dtype=torch.float32)
        dist.all_gather_into_tensor(wl_all, wl)                                   # 8 B / particle
        work = dist.all_gather_into_tensor(st_all_t, sb, async_op=True)          # 28 B / particle, in flight ...
        g_w = wl_all.numpy().reshape(world, 2, n_local)[:, 0, :].reshape(-1).copy()
        g_ld = wl_all.numpy().reshape(world, 2, n_local)[:, 1, :].reshape(-1).copy()
        wn, arg, _ = orc.normalize(g_w, g_ld)                                     # ... while every rank normalises globally
        work.wait()
        g_st, g_ld2, g_w2 = sharded.unpack_split_numpy(wl_all.numpy(), st_all_t.numpy(), world, n_local, synth.STATE_DTYPE)
        assert np.array_equal(g_ld2, g_ld) and np.array_equal(g_w2.view(np.uint32), g_w.view(np.uint32))
        i0, i1 = sharded.sample_slice(M, rank, world)
        idx = orc.resample_fast(wn, u, M)[i0:i1]
        q.put((rank, wn, arg, idx, g_st[idx], g_ld[idx]))
    finally:
        dist.destroy_process_group()


def _rank_main(rank, world, port, n_local, u, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wd = _world()
        n_total = n_local * world
        st_all, ld_all = synth.particles_tracking(n_total, wd["pose"], wd["heading"], seed=5)
        st_all["have_init"][::4] = 0
        lo, hi = sharded.shard_range(n_total, rank, world)
        st, ld = st_all[lo:hi].copy(), ld_all[lo:hi].copy()
        # "score kernel" on the local shard (oracle stands in for the device)
        w = orc.score_all(st, wd["fp"], wd["layers"], wd["mask"], 1.0, wd["tab"], 100, 25, wd["scan"], 2.0, wd["thetas"], wd["shifts"], n_threads=1)
        send = torch.from_numpy(sharded.pack_block_numpy(st, ld, w))
        recv = torch.empty(world * send.numel(), dtype=torch.float32)
        dist.all_gather_into_tensor(recv, send)                       # the ONE collective of the step
        g_st, g_ld, g_w = sharded.unpack_blocks_numpy(recv.numpy(), world, n_local, synth.STATE_DTYPE)
        wn, arg, _ = orc.normalize(g_w, g_ld)                         # redundantly on every rank, global order
        i0, i1 = sharded.sample_slice(M, rank, world)
        idx = orc.resample_fast(wn, u, M)[i0:i1]                      # this rank's output slice
        q.put((rank, wn, arg, idx, g_st[idx], g_ld[idx]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,split", [(2, False), (3, False), (2, True), (3, True)])
def test_sharded_update_equals_single_process(world, split):
    n_local, M = 60, 150
    u = orc.uniform_draw(9)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main_split if split else _rank_main, args=(r, world, port, n_local, u, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wn, arg, idx, st_new, ld_new = _single(n_local * world, u, M)
    for r in res:
        assert np.array_equal(r[1].view(np.uint32), wn.view(np.uint32)) and r[2] == arg   # identical on every rank
    assert np.array_equal(np.concatenate([r[3] for r in res]), idx)
    assert np.array_equal(np.concatenate([r[4] for r in res]), st_new)
    assert np.array_equal(np.concatenate([r[5] for r in res]), ld_new)


def test_grid_mailbox_exchange_protocol_model():
    """The protocol of tdr_grid_peer_exchange (csrc/api.cu: k_grid_exchange) as a state machine under random interleavings:
    every rank MINs its key into slot e & 1 of EVERY rank's mailbox, then bumps that rank's arrival counter; it waits until
    its own counter shows ranks x (uses of the slot), reads the minimum and resets the key slot for exchange e + 2.
    Counters never reset.  Claim: whatever the interleaving, every rank reads the minimum over all ranks' keys of THAT
    exchange — in particular no key of exchange e + 2 can land before the owner's reset of exchange e."""
    import random
    INF = 1 << 62
    for trial in range(300):
        rnd = random.Random(trial)
        G, E = rnd.choice([2, 3, 4, 8]), 6
        keys = [[rnd.randrange(1, 1 << 30) for _ in range(G)] for _ in range(E)]          # keys[e][rank]
        box_key = [[INF, INF] for _ in range(G)]
        box_cnt = [[0, 0] for _ in range(G)]
        got = [[None] * G for _ in range(E)]

        def rank_prog(r):
            for e in range(E):
                slot = e & 1
                for d in range(G):                       # lanes of the one warp: any order between ranks, in order per lane
                    box_key[d][slot] = min(box_key[d][slot], keys[e][r]); yield
                    box_cnt[d][slot] += 1; yield
                target = G * (e // 2 + 1)
                while box_cnt[r][slot] < target:
                    yield
                got[e][r] = box_key[r][slot]; yield
                box_key[r][slot] = INF; yield            # ready for exchange e + 2

        progs = [rank_prog(r) for r in range(G)]
        live = list(range(G))
        steps = 0
        while live:
            r = rnd.choice(live)
            try:
                next(progs[r])
            except StopIteration:
                live.remove(r)
            steps += 1
            assert steps < 2_000_000, "deadlock"
        for e in range(E):
            assert got[e] == [min(keys[e])] * G, (trial, e, got[e], min(keys[e]))


def test_shard_export_slot_protocol_model():
    """tdr_shard_step's state exchange (csrc/shard.cu) as a state machine: per scan t every rank packs its states into its
    export slot t & 1, joins the all-gather of the weights (which completes on a rank only once EVERY rank has joined it),
    then reads the drawn particles' states out of its peers' slot t & 1.  Claim: with two slots and no other
    synchronisation a reader never sees a slot that its owner has already overwritten with scan t + 2."""
    import random
    for trial in range(300):
        rnd = random.Random(1000 + trial)
        G, T = rnd.choice([2, 3, 4, 8]), 7
        slots = [[-1, -1] for _ in range(G)]            # slots[rank][k] = the scan whose states it holds
        joined = [0] * G                                # all-gathers a rank has joined
        bad = []

        def rank_prog(r):
            for t in range(T):
                slots[r][t & 1] = t; yield              # k_shard_pack
                joined[r] = t + 1; yield                # enqueue the all-gather of scan t
                while min(joined) < t + 1:              # ... which completes only when every rank has joined it
                    yield
                for d in rnd.sample(range(G), G):       # k_shard_resample: peer reads, any order, not atomic as a group
                    if slots[d][t & 1] != t:
                        bad.append((t, r, d, slots[d][t & 1]))
                    yield

        progs = [rank_prog(r) for r in range(G)]
        live = list(range(G))
        while live:
            r = rnd.choice(live)
            try:
                next(progs[r])
            except StopIteration:
                live.remove(r)
        assert not bad, bad[:3]
