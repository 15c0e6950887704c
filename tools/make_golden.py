"""Generates tests/golden/r01_small.npz: a small seeded scene with the ORACLE's outputs for every stage of the path.

The reference ships no vectors for this path and cannot be built here, so these are not pins
of the reference: they freeze the oracle (and, through the GPU tests, the CUDA path) against accidental drift.
Regenerate with `python tools/make_golden.py` only when a deliberate change of the restated arithmetic is made.
`python tools/make_golden.py gicp` writes tests/golden/r02_gicp_small.npz (the GICP stages) the same way."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import binding as ob
from pointcloud_slam_b200 import synth


def main():
    world = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(30_000, synth.SEED, world=world)
    x_true = synth.make_state([3.0, -2.0, 1.2], [0.01, -0.02, 0.6])
    o_l, Rl = synth.lidar_pose(x_true)
    scan = np.ascontiguousarray(synth.raycast(o_l, Rl, synth.livox_dirs(900, synth.SEED), world, seed=synth.SEED)[:600])
    x_prop = synth.perturb_state(x_true, synth.SEED)
    P = synth.init_cov(synth.SEED)
    out = dict(map=mp, scan=scan, x_prop=x_prop, P=P)
    # local map + IEKF (P-horizon parameters)
    lio = ob.OracleLio(resolution=0.5, nearby=18)
    lio.insert(mp)
    qw = (scan.astype(np.float64) @ Rl.T + o_l).astype(np.float32)
    idx, d2, cnt = lio.knn5(qw)
    rc, x_post, P_post, st = lio.update(scan, x_prop, P)
    out.update(knn_query=qw, knn_idx=idx, knn_d2=d2, knn_cnt=cnt, iekf_rc=rc, x_post=x_post, P_post=P_post,
               n_eff=np.array(list(st.n_eff)), passes=st.passes, HtH0=np.array(st.HtH[0]).reshape(12, 12), Hth0=np.array(st.Hth[0]))
    # NDT
    ndt = ob.OracleNdt(resolution=2.0, trans_eps=0.01)
    ndt.set_target(mp)
    ndt.set_source(scan)
    L = ndt.leaves()
    p6 = np.array([3.05, -1.97, 1.22, 0.002, -0.003, 0.61])
    s, g, H = ndt.derivatives(p6)
    guess = synth.pose_vec_to_matrix(p6).astype(np.float32)
    rc, T, r = ndt.align(guess)
    poses = synth.hypothesis_grid(np.array([3.0, -2.0, 1.2, 0, 0, 0.6]), 3, 3, 2, 1.0)
    out.update(ndt_ids=L["ids"], ndt_npts=L["npts"], ndt_mean=L["mean"], ndt_icov=L["icov"], ndt_p6=p6, ndt_score=s, ndt_g=g, ndt_H=H,
               ndt_guess=guess, ndt_final=T, ndt_iters=r.iters, ndt_evals=r.evals, ndt_p_final=np.array(r.p_final), reloc_poses=poses,
               reloc_scores=ndt.score_batch(poses), fitness=np.array(ndt.fitness(T)))
    # VoxelGrid
    c, n = ob.voxel_grid(scan, 0.5)
    out.update(vg_centroids=c, vg_counts=n)
    path = os.path.join(ROOT, "tests", "golden", "r01_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main_gicp():
    """tests/golden/r02_gicp_small.npz: the oracle's GICP outputs on a small seeded scene (drift pin, like r01_small.npz)."""
    world = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(40_000, synth.SEED + 1, world=world)
    p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
    T = synth.pose_vec_to_matrix(p_true)
    scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], synth.livox_dirs(3000, synth.SEED + 1), world, seed=synth.SEED + 1)[:2400])
    guess = synth.pose_vec_to_matrix(p_true + np.array([0.15, -0.1, 0.05, 0.01, -0.01, 0.03])).astype(np.float32)
    g = ob.OracleGicp()
    g.set_target(mp)
    g.set_source(scan)
    cov_src, cov_tgt = g.covariances("source"), g.covariances("target")
    knn_src, _ = ob.exact_knn(scan, scan, 20)
    m, idx, maha, d2 = g.correspondences(np.eye(4), guess)
    x = np.array([0.01, -0.02, 0.005, 0.002, -0.001, 0.003])
    f_op, f_fdf, g_df, g_fdf = g.cost(x)
    rc, fin, r = g.align(guess)
    out = dict(map=mp, scan=scan, guess=guess, cov_src=cov_src, cov_tgt_every_16=cov_tgt[::16], knn_src=knn_src, corr_m=m, corr_idx=idx, corr_maha=maha,
               corr_d2=d2, cost_x=x, cost_f_op=f_op, cost_f_fdf=f_fdf, cost_g=g_fdf, align_rc=rc, align_final=fin, align_iterations=r.iterations,
               align_last_m=r.last_m, align_inner_total=r.inner_total, align_calls=np.array([r.n_f, r.n_df, r.n_fdf]))
    path = os.path.join(ROOT, "tests", "golden", "r02_gicp_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", "passes", r.iterations, "matches", r.last_m)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gicp":
        main_gicp()
    else:
        main()
