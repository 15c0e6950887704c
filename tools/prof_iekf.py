"""Dev helper: one full-size IEKF update workload (config 1) for ncu launch lists."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from pointcloud_slam_b200 import synth, api
params = sys.argv[1] if len(sys.argv) > 1 else 'livox'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
res, nearby, ext = (0.2, 26, False) if params == 'livox' else (0.5, 18, True)
c = synth.config1()
g = api.IVox(resolution=res, nearby=nearby)
t = time.time(); g.AddPoints(c['map']); print('insert 2M: %.1f ms, voxels %d' % ((time.time() - t) * 1e3, g.NumValidGrids()))
kf = api.Esekf(g, extrinsic_est_en=ext)
for r in range(reps):
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    t = time.time(); kf.update_iterated_dyn_share_modified(c['scan']); dt = time.time() - t
    st = kf.stats
    print('update n=%d gpu_ms %.4f wall_ms %.3f passes %d knn %d neff %s' % (len(c['scan']), st.gpu_ms, dt * 1e3, st.passes, st.knn_passes, list(st.n_eff)[:4]))
print('pos err', kf.x[:3] - c['x_true'][:3])
import ctypes as C
st = (C.c_longlong * (8 * 16))()
api.lib().b200_iekf_debug_stamps(kf.h, st)
for p in range(kf.stats.passes):
    v = [st[p * 16 + i] for i in range(8)]
    print('pass', p, 'post', [v[i + 1] - v[i] for i in range(7)], 'tot', v[7] - v[0], '| pre', st[p*16+8], 'wait', st[p*16+9], '| search/measure/accum', st[p*16+10], st[p*16+11], st[p*16+12], '| pre stages', st[p*16+13], st[p*16+14], st[p*16+15])
# per-kernel event timing (no graph)
api.lib().b200_iekf_set_profiling(kf.h, 1)
for r in range(3):
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    kf.update_iterated_dyn_share_modified(c['scan'])
    ms = (C.c_float * 17)()
    k = api.lib().b200_iekf_kernel_times(kf.h, ms, 17)
    print('profiled gpu_ms %.4f kernels(us):' % kf.stats.gpu_ms, ['%.1f' % (ms[i] * 1e3) for i in range(k)])
api.lib().b200_iekf_set_profiling(kf.h, 0)
api.lib().b200_iekf_set_graph(kf.h, 0)
for r in range(2):
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    t = time.time(); kf.update_iterated_dyn_share_modified(c['scan']); dt = time.time() - t
    print('no-graph gpu_ms %.4f wall %.3f' % (kf.stats.gpu_ms, dt * 1e3))
