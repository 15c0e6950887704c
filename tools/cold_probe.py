"""Dev helper: what the L2 flush costs the first k_obs of an update - kernel code or data?  After the flush a 64-point
update on another filter warms the code only; then the 20k-point update is timed."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from pointcloud_slam_b200 import synth, api
c = synth.config1()
g = api.IVox(resolution=0.2, nearby=26)
g.AddPoints(c['map'])
kf = api.Esekf(g)
kf2 = api.Esekf(g)
small = np.ascontiguousarray(c['scan'][:64])
def run(warm_code):
    out = []
    for r in range(8):
        api.flush_l2(0)
        if warm_code:
            kf2.change_x(c['x_prop']); kf2.change_P(c['P'])
            kf2.update_iterated_dyn_share_modified(small)
        kf.change_x(c['x_prop']); kf.change_P(c['P'])
        kf.update_iterated_dyn_share_modified(c['scan'])
        out.append(kf.stats.gpu_ms * 1e3)
    return np.median(out[2:])
print('cold (flush only): %.1f us; flush + code warmed by a 64-point update: %.1f us' % (run(False), run(True)))
w = []
for r in range(8):
    kf.change_x(c['x_prop']); kf.change_P(c['P'])
    kf.update_iterated_dyn_share_modified(c['scan']); w.append(kf.stats.gpu_ms * 1e3)
print('warm: %.1f us' % np.median(w[2:]))
