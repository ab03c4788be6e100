"""Run under torchrun (N >= 2 GPUs): the sharded filter INSIDE the library (tdr_shard_step / tdr_shard_pose,
csrc/shard.cu) against (a) the same update on ONE GPU over the concatenated particle set and (b) the torch.distributed
harness of round 1 — resampled states and indices bit for bit, pose bit for bit, over several consecutive scans (the
export slots alternate, so at least three).  Prints one JSON line on rank 0; exit code 1 on any mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/shard_check.py [--particles 20000] [--scans 4] [--workload global|tracking]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from top_down_renderer_b200 import sharded, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=20000)
    ap.add_argument("--scans", type=int, default=4)
    ap.add_argument("--workload", default="global")
    ap.add_argument("--side", type=int, default=1024)
    a = ap.parse_args()
    rank, world, local = bench.init_dist()
    wl = dict(bench.WORKLOADS[a.workload], n=a.particles, side=a.side)
    inp = bench.make_inputs(wl, rank)
    n, res = wl["n"], wl["res"]
    us = [float(x) for x in np.random.default_rng(7).random(a.scans, dtype=np.float32)]

    def run(kind):
        ctx = bench.setup_ctx(wl, inp, local)
        ctx.scan_set_points(inp["pts"])
        stream = torch.cuda.ExternalStream(ctx.stream, device=local)
        flt = (sharded.LibraryShardedFilter(ctx, rank, world, n) if kind == "library"
               else sharded.ShardedFilter(ctx, stream, rank, world))
        states, poses = [], []
        for u in us:
            with torch.cuda.stream(stream):
                flt.step(res, float(bench.ANG_RES), bench.N_THETA, bench.N_R, u, n * world)
            with torch.cuda.stream(stream):
                poses.append(np.concatenate([np.ravel(x) for x in flt.pose(True)]))
            ctx.sync()
            states.append(ctx.pf_get_states().copy())
        if hasattr(flt, "close"):
            flt.close()
        ctx.close()
        return states, poses

    lib_states, lib_poses = run("library")
    dist.barrier()
    tor_states, tor_poses = run("torch")
    dist.barrier()
    mine = {"lib": [bench.states_digest(s) for s in lib_states], "torch": [bench.states_digest(s) for s in tor_states],
            "lib_pose": [p.tobytes().hex() for p in lib_poses], "torch_pose": [p.tobytes().hex() for p in tor_poses]}
    got = [None] * world
    dist.all_gather_object(got, mine)
    ok = True
    if rank == 0:
        parts = [(synth.particles_global(n, inp["cm"], seed=bench.SEED + 101 * r) if wl["shifts"] > 1 else
                  synth.particles_tracking(n, inp["pose"], inp["heading"], seed=bench.SEED + 101 * r)) for r in range(world)]
        big = dict(inp, st=np.concatenate([p[0] for p in parts]), ld=np.concatenate([p[1] for p in parts]))
        ref = bench.setup_ctx(dict(wl, n=n * world), big, local)
        ref.scan_set_points(inp["pts"])
        want, want_pose = [], []
        for u in us:
            ref.step(res, float(bench.ANG_RES), bench.N_THETA, bench.N_R, u, n * world)
            want_pose.append(np.concatenate([np.ravel(x) for x in ref.pf_pose(True)]).tobytes().hex())
            st = ref.pf_get_states()
            want.append([bench.states_digest(st[r * n:(r + 1) * n]) for r in range(world)])
        ref.close()
        rec = {"ranks": world, "particles_per_rank": n, "scans": a.scans, "workload": a.workload}
        rec["library_states_equal_single_gpu"] = all(got[r]["lib"][k] == want[k][r] for r in range(world) for k in range(a.scans))
        rec["torch_states_equal_single_gpu"] = all(got[r]["torch"][k] == want[k][r] for r in range(world) for k in range(a.scans))
        rec["library_pose_equal_single_gpu"] = all(got[r]["lib_pose"][k] == want_pose[k] for r in range(world) for k in range(a.scans))
        rec["torch_pose_equal_single_gpu"] = all(got[r]["torch_pose"][k] == want_pose[k] for r in range(world) for k in range(a.scans))
        ok = all(v for k, v in rec.items() if k.endswith("single_gpu"))
        print(json.dumps(rec), flush=True)
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
