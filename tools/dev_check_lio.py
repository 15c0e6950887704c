import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from pointcloud_slam_b200 import synth, api
from oracle import binding as ob
NM=int(sys.argv[1]) if len(sys.argv)>1 else 200000
NS=int(sys.argv[2]) if len(sys.argv)>2 else 5000
c=synth.config1(n_map=NM,n_scan=NS)
for (res,nearby) in [(0.5,18),(0.2,26)]:
    o=ob.OracleLio(resolution=res,nearby=nearby)
    o.insert(c['map'])
    g=api.IVox(resolution=res,nearby=nearby)
    t=time.time(); g.AddPoints(c['map']); print('gpu insert s',time.time()-t)
    print('voxels',o.num_voxels,g.NumValidGrids(),'points',o.num_points,g.NumPoints())
    o_l,Rl=synth.lidar_pose(c['x_true'])
    qw=(c['scan'].astype(np.float64)@Rl.T+o_l).astype(np.float32)
    i0,d0,c0=o.knn5(qw)
    i1,d1,c1=g.GetClosestPoint(qw)
    print('knn idx equal',np.array_equal(i0,i1),'dist equal',np.array_equal(d0,d1),'cnt equal',np.array_equal(c0,c1), 'mismatch rows',int((i0!=i1).any(1).sum()))
    # incremental insert
    extra=synth.sample_map(30000,seed=777)
    o.insert(extra); g.AddPoints(extra)
    i0,d0,c0=o.knn5(qw); i1,d1,c1=g.GetClosestPoint(qw)
    print('after incr: voxels',o.num_voxels,g.NumValidGrids(),'idx equal',np.array_equal(i0,i1),np.array_equal(d0,d1))
    # IEKF
    o2=ob.OracleLio(resolution=res,nearby=nearby,extrinsic_est_en=(nearby==18)); o2.insert(c['map'])
    g2=api.IVox(resolution=res,nearby=nearby); g2.AddPoints(c['map'])
    kf=api.Esekf(g2,extrinsic_est_en=(nearby==18))
    rc0,HtH0,Hth0,ne0=o2.obs_model(c['scan'],c['x_prop'],True)
    rc1,HtH1,Hth1,ne1=kf.ObsModel(c['scan'],c['x_prop'],True)
    print('obs n_eff',ne0,ne1,'HtH relerr',np.abs(HtH0-HtH1).max()/np.abs(HtH0).max(),'Hth relerr',np.abs(Hth0-Hth1).max()/np.abs(Hth0).max())
    ps0=o2.point_state(NS if NS<=len(c['scan']) else len(c['scan'])); ps1=kf.point_state()
    n=len(c['scan'])
    print('plane equal',np.array_equal(ps0['plane'][:n],ps1['plane']),'res equal',np.array_equal(ps0['residual'][:n],ps1['residual']),'sel equal',np.array_equal(ps0['selected'][:n],ps1['selected']),'nn eq',np.array_equal(ps0['nn_idx'][:n],ps1['nn_idx']))
    # full update (fresh objects for clean per-point state)
    o3=ob.OracleLio(resolution=res,nearby=nearby,extrinsic_est_en=(nearby==18)); o3.insert(c['map'])
    g3=api.IVox(resolution=res,nearby=nearby); g3.AddPoints(c['map'])
    kf3=api.Esekf(g3,extrinsic_est_en=(nearby==18))
    rc,x0,P0,st0=o3.update(c['scan'],c['x_prop'],c['P'])
    kf3.change_x(c['x_prop']); kf3.change_P(c['P'])
    t=time.time(); rc1=kf3.update_iterated_dyn_share_modified(c['scan']); dt=time.time()-t
    st1=kf3.stats
    print('passes',st0.passes,st1.passes,'knn',st0.knn_passes,st1.knn_passes,'conv',st0.converged,st1.converged,'neff',list(st0.n_eff)[:4],list(st1.n_eff)[:4])
    print('x diff',np.abs(x0-kf3.x).max(),'P relerr',np.abs(P0-kf3.P).max()/np.abs(P0).max(),'gpu_ms',st1.gpu_ms,'wall ms',dt*1e3)
    for p in range(st1.passes):
        H,h,xin=kf3.last_HtH(p)
        H0=np.array(st0.HtH[p]).reshape(12,12); h0=np.array(st0.Hth[p])
        print(' pass',p,'HtH rel',np.abs(H-H0).max()/np.abs(H0).max(),'Hth rel',np.abs(h-h0).max()/max(np.abs(h0).max(),1e-30),'xin diff',np.abs(xin-np.array(st0.x_in[p])).max())
    for r in range(3):
        kf3.change_x(c['x_prop']); kf3.change_P(c['P'])
        t=time.time(); kf3.update_iterated_dyn_share_modified(c['scan']); dt=time.time()-t
        print('  rerun gpu_ms',kf3.stats.gpu_ms,'wall ms',dt*1e3)
    # map incremental
    tot,na,nd=o3.map_incremental(c['scan'],x0,True)
    na1,nd1=kf3.MapIncremental(x0,True)
    print('mapinc',na,nd,na1,nd1,'voxels',o3.num_voxels,g3.NumValidGrids())
    i0,d0,c0=o3.knn5(qw); i1,d1,c1=g3.GetClosestPoint(qw)
    print('after mapinc knn equal',np.array_equal(i0,i1),np.array_equal(d0,d1))
print('launches',api.kernel_launches())
