"""Worker for the multi-rank relocalization tests (launched with torch.distributed.run).

mode "gloo": CPU only - the hypothesis slices are scored by the oracle and the allreduce-argmin protocol of
             b200_reloc_argmin (all-gather of the ranks' (score key, index) winners, same reduction everywhere) runs over gloo.
mode "nccl": the product path - map replicated with b200_ndt_set_target_bcast, slices scored on each GPU, NCCL collectives.
Rank 0 writes a JSON result to argv[2].
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    mode, out_path = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    from pointcloud_slam_b200 import api, synth
    world_geo = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(150_000, synth.SEED, world=world_geo)
    p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
    T = synth.pose_vec_to_matrix(p_true)
    scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], synth.livox_dirs(3000, synth.SEED), world_geo, seed=synth.SEED)[:2000])
    poses = synth.hypothesis_grid(p_true, nx=5, ny=5, nyaw=3, pitch=1.0)   # 75 hypotheses: ragged over 2 ranks
    b, e = api.shard_range(len(poses), world, rank)
    if mode == "gloo":
        from oracle import binding as ob
        dist.init_process_group("gloo")
        o = ob.OracleNdt(resolution=1.0)
        o.set_target(mp)
        o.set_source(scan)
        scores = o.score_batch(poses[b:e]) if e > b else np.zeros(0)
        def gather(t):
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return torch.stack(out)
        best, score = api.argmin_protocol_host(scores, b, gather)
        full = o.score_batch(poses)
        res = dict(best=int(best), score=float(score), expect=int(np.argmax(full)), expect_score=float(full.max()), world=world,
                   slice=[b, e])
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ident = [api.Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        comm = api.Communicator(world, rank, ident[0], device=local)
        g = api.NormalDistributionsTransform(device=local)
        g.setInputTargetReplicated(comm, mp if rank == 0 else None, len(mp))
        g.setInputSource(scan)
        best, score, ms = api.relocalize(g, poses[b:e], comm, h_begin=b)
        nvox = g.numVoxels()
        # every rank must hold the same replica and the same answer
        t = torch.tensor([best, nvox], device="cuda", dtype=torch.int64)
        tmin, tmax = t.clone(), t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        full = g.calculateScore(poses)
        res = dict(best=int(best), score=float(score), expect=int(np.argmax(full)), expect_score=float(full.max()), world=world,
                   same_on_all_ranks=bool((tmin == tmax).all().item()), voxels=int(nvox), ms=float(ms), slice=[b, e])
        comm.close()
    if mode == "nccl":
        # construct_full_map over the ranks: contiguous keyframe blocks, exchange of partial sums, disjoint ownership
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_gpu_voxel import make_frames
        frames, poses = make_frames(synth, k=9, n=4000)
        fb, fe = api.shard_range(len(frames), world, rank)
        ident2 = [api.Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident2, src=0)
        comm2 = api.Communicator(world, rank, ident2[0], device=local)
        b = api.FullMapBuilder(leaf=0.1, capacity_voxels=400_000, device=local)
        for f, p in zip(frames[fb:fe], poses[fb:fe]):
            b.add_keyframe(f, p)
        b.merge(comm2)
        c, n = b.extract()
        single = api.FullMapBuilder(leaf=0.1, capacity_voxels=400_000, device=local)
        for f, p in zip(frames, poses):
            single.add_keyframe(f, p)
        cs, ns = single.extract()
        tot = torch.tensor([len(c), int(n.sum())], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        # every voxel this rank owns must equal the single-GPU voxel with the same centroid
        key = lambda a: np.floor(a[:, :3] / np.float32(0.1)).astype(np.int64)
        lut = {tuple(k): i for i, k in enumerate(key(cs))}
        idx = np.array([lut.get(tuple(k), -1) for k in key(c)])
        ok = bool((idx >= 0).all() and np.array_equal(ns[idx], n) and np.abs(cs[idx, :3] - c[:, :3]).max() < 4e-6
                  and np.abs(cs[idx, 3] - c[:, 3]).max() < 1e-3)   # intensity is a plain fp32 sum (order of the atomics)
        okt = torch.tensor([int(ok)], device="cuda", dtype=torch.int64)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        res.update(fullmap_voxels=int(tot[0]), fullmap_points=int(tot[1]), single_voxels=len(cs), single_points=int(ns.sum()),
                   fullmap_match=bool(okt.item()), exchange_ms=b.exchange_ms())
        comm2.close()
    if rank == 0:
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
