import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from pointcloud_slam_b200 import synth, api
c = synth.config1()
g = api.IVox(resolution=0.2, nearby=26)
g.AddPoints(c['map'])
kf = api.Esekf(g)
ts = []
for r in range(12):
    api.flush_l2(0)
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    t = time.perf_counter(); kf.update_iterated_dyn_share_modified(c['scan']); ts.append((time.perf_counter() - t) * 1e6)
print('python wall us', ['%.0f' % t for t in ts])
pin = api.PinnedCloud(len(c['scan']), 3); pin.array[:] = c['scan']
ts = []
for r in range(12):
    api.flush_l2(0)
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    t = time.perf_counter(); kf.update_iterated_dyn_share_modified(pin.array); ts.append((time.perf_counter() - t) * 1e6)
print('pinned wall us', ['%.0f' % t for t in ts], kf.get_x()[:3])

api.lib().b200_iekf_set_graph(kf.h, 0)
ts = []; dev = []
for r in range(12):
    api.flush_l2(0)
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    t = time.perf_counter(); kf.update_iterated_dyn_share_modified(pin.array); ts.append((time.perf_counter() - t) * 1e6); dev.append(kf.stats.gpu_ms * 1e3)
print('pinned, plain launches: wall us', ['%.0f' % t for t in ts], 'device us', ['%.0f' % t for t in dev])
