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
        # construct_full_map over the ranks, protocol only (numpy + gloo): contiguous keyframe blocks, per-rank voxel sums, records
        # routed to their owner (hash of the voxel key), owners add them up.  The union must be the oracle's single-process map.
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_gpu_voxel import make_frames
        frames, fposes = make_frames(synth, k=6, n=1500)
        leaf = np.float32(0.1)
        fb, fe = api.shard_range(len(frames), world, rank)
        cells, pts = [], []
        for f, p in zip(frames[fb:fe], fposes[fb:fe]):
            # pose7 -> matrix with the builder's arithmetic: fp64 quaternion products, narrowed to an fp32 3x4
            w, x, y, z = p[3], p[4], p[5], p[6]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]).astype(np.float32)
            t = p[:3].astype(np.float32)
            q = np.stack([((R[r, 0] * f[:, 0] + R[r, 1] * f[:, 1]) + R[r, 2] * f[:, 2]) + t[r] for r in range(3)], 1).astype(np.float32)
            cells.append(np.floor(q * (np.float32(1.0) / leaf)).astype(np.int64))
            pts.append(np.concatenate([q, f[:, 3:4]], 1).astype(np.float64))
        cells, pts = np.concatenate(cells), np.concatenate(pts)
        keys = api.voxel_key(cells)
        uk, inv = np.unique(keys, return_inverse=True)
        cnt = np.bincount(inv, minlength=len(uk))
        sums = np.zeros((len(uk), 4))
        np.add.at(sums, inv, pts)

        def exchange(per_dest):
            sizes = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([len(a) for a in per_dest], dtype=torch.int64))
            out = []
            for src in range(world):           # every rank broadcasts what it has for each destination; keep what is addressed to us
                for dst in range(world):
                    n = int(sizes[src][dst])
                    buf = torch.from_numpy(np.ascontiguousarray(per_dest[dst])) if src == rank else torch.zeros((n, 6), dtype=torch.float64)
                    dist.broadcast(buf, src=src)
                    if dst == rank:
                        out.append(buf.numpy().copy())
            return out
        ok_, cn_, sm_ = api.mapbuild_merge_protocol_host(uk, cnt, sums, world, rank, exchange)
        owned = [None] * world
        dist.all_gather_object(owned, (ok_, cn_, sm_))
        if rank == 0:
            c0, n0 = ob.full_map(frames, fposes, 0.1)
            allk = np.concatenate([o[0] for o in owned])
            alln = np.concatenate([o[1] for o in owned])
            alls = np.concatenate([o[2] for o in owned])
            order = np.argsort(allk)       # pack_key(cz, cy, cx): ascending key = the builder's (z, y, x) extraction order
            cent = (alls[order] / alln[order][:, None]).astype(np.float32)
            res.update(fullmap_disjoint=bool(len(np.unique(allk)) == len(allk)), fullmap_voxels=int(len(allk)), oracle_voxels=int(len(c0)),
                       fullmap_counts_equal=bool(len(allk) == len(c0) and np.array_equal(alln[order], n0)),
                       fullmap_max_centroid_diff=float(np.abs(cent - c0).max()) if len(allk) == len(c0) else None,
                       fullmap_owner_rule=bool(all((api.voxel_owner(o[0], world) == r).all() for r, o in enumerate(owned))))
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
