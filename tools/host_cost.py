import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sycl_points_b200 as spx
q = spx.DeviceQueue(0)
rs = np.random.RandomState(0)
pts = np.c_[rs.uniform(-20, 20, (3000, 3)), np.ones(3000)].astype(np.float32)
raw = spx.PointCloudShared(q, pts)
vg = spx.VoxelGrid(q, 0.25)
nn = spx.KNNResult()
params = spx.RegistrationParams(); params.robust.type = spx.RobustLossType.HUBER
reg = spx.Registration(q, params)
def chain():
    c = vg.downsampling(raw)
    t = spx.KDTree.build(q, c)
    t.knn_search_async(c, 10, nn)
    spx.covariance.estimate(nn, c)
    return c, t
for _ in range(20): c, t = chain(); t.close()
q.wait()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    c, t = chain(); t.close()
q.wait()
print("chain host+gpu per call (3000 pts): %.1f us" % ((time.perf_counter() - t0) / N * 1e6))
c, t = chain()
t0 = time.perf_counter()
for _ in range(N):
    r = reg.align(c, c, t)
print("align per call (3000 pts, %d its): %.1f us" % (r.iterations + 1, (time.perf_counter() - t0) / N * 1e6))
# individual stages
for name, fn in (("voxel", lambda: vg.downsampling(raw)), ("build", lambda: spx.KDTree.build(q, c).close()),
                 ("knn", lambda: t.knn_search_async(c, 10, nn)), ("cov", lambda: spx.covariance.estimate(nn, c))):
    q.wait(); t0 = time.perf_counter()
    for _ in range(N): fn()
    q.wait()
    print("  %s: %.1f us" % (name, (time.perf_counter() - t0) / N * 1e6))
