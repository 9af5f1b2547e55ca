"""GPU parity: linearise / error / robust weights / align through the C-ABI vs the oracle.

Tolerances (stated by the north star and SURVEY §8(c)): H, b, error within 1e-5 relative of the
oracle's fp64-accumulated sums for every factor (the transcendentals of the plane regularisation,
the robust rho and se3_exp are evaluated correctly rounded on both sides — spx_math.cuh cr_*,
orc_math.hpp cr_* — so the per-point terms are bit-identical and only the summation order differs);
poses within 1e-5 m / 1e-5 rad at equal iteration counts."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

LOSSES = ["NONE", "HUBER", "TUKEY", "CAUCHY", "GEMAN_MCCLURE"]
REGS = ["POINT_TO_POINT", "POINT_TO_PLANE", "GICP", "POINT_TO_DISTRIBUTION"]


@pytest.fixture(scope="module")
def q(spx):
    return spx.DeviceQueue(0)


@pytest.fixture(scope="module")
def pair(spx, q, bundled):
    """bundled scan pair with k=10 covariances + normals on the device and on the host."""
    src_h, tgt_h = bundled["source_ds"], bundled["target_ds"]
    src, tgt = spx.PointCloudShared(q, src_h), spx.PointCloudShared(q, tgt_h)
    ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
    ns, nt = ts.knn_search(src, 10), tt.knn_search(tgt, 10)
    spx.covariance.estimate(ns, src)
    spx.covariance.estimate(nt, tgt)
    spx.covariance.estimate_normals(nt, tgt)
    otree = oracle.KDTree(tgt_h)
    return dict(src=src, tgt=tgt, tree=tt, src_h=src_h, tgt_h=tgt_h, cov_s=src.covs_host(), cov_t=tgt.covs_host(),
                nrm_t=tgt.normals_host(), otree=otree)


def clouds_for(spx, q, pair, reg):
    """(target cloud, its host covariances) to use with factor `reg`.  Point-to-distribution inverts
    the RAW target covariance (factor.hpp:311-317); the 10-neighbour covariances of a planar LiDAR
    patch are near-singular in fp32, the inverse is then indefinite and the reference's own formula
    yields NaN residual norms (reproduced identically by oracle and GPU).  The parity comparison
    uses what a P2D user would: covariances with a 1 cm^2 isotropic floor."""
    if reg != "POINT_TO_DISTRIBUTION":
        return pair["tgt"], pair["cov_t"]
    if "tgt_p2d" not in pair:
        cov = pair["cov_t"].copy()
        cov[:, :3, :3] += np.float32(1e-2) * np.eye(3, dtype=np.float32)
        pair["tgt_p2d"] = spx.PointCloudShared(q, pair["tgt_h"], cov, pair["nrm_t"])
        pair["cov_t_p2d"] = cov
    return pair["tgt_p2d"], pair["cov_t_p2d"]


def pose_delta(Ta, Tb):
    d = np.linalg.inv(Ta.astype(np.float64)) @ Tb.astype(np.float64)
    # sin(angle) from the skew part: accurate for the tiny angles compared here (arccos of the
    # trace loses half the digits near identity)
    w = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return np.linalg.norm(d[:3, 3]), float(np.arcsin(min(1.0, np.linalg.norm(w))))


def rel(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("reg", REGS)
@pytest.mark.parametrize("loss", LOSSES)
def test_linearize_and_error_vs_oracle(spx, q, pair, reg, loss):
    T = oracle.se3_exp(np.array([0.004, -0.003, 0.012, 0.4, 0.1, -0.02], np.float32))
    nn_idx, nn_dist = pair["otree"].knn(pair["src_h"], 1, T)
    params = spx.RegistrationParams(reg_type=spx.RegType[reg])
    params.robust.type = spx.RobustLossType[loss]
    params.robust.default_scale = 0.7
    reg_obj = spx.Registration(q, params)
    tgt, cov_t = clouds_for(spx, q, pair, reg)
    lin = reg_obj.compute_linearized_result(pair["src"], tgt, pair["tree"], T)
    H, b, e, inl = oracle.linearize(oracle.REG[reg], oracle.LOSS[loss], pair["src_h"], pair["cov_s"], pair["tgt_h"],
                                    cov_t, pair["nrm_t"], nn_idx, nn_dist, T, 4.0, 0.7, mode=1)
    tol = 1e-5
    assert lin.inlier == inl
    assert rel(lin.H, H) <= tol and rel(lin.b, b) <= tol and abs(lin.error - e) <= tol * abs(e)
    assert np.array_equal(lin.H, lin.H.T)
    # frozen-neighbour error at a trial pose (registration.hpp:350-359)
    T2 = (T @ oracle.se3_exp(np.array([1e-3, 2e-3, -1e-3, 0.01, -0.02, 0.005], np.float32))).astype(np.float32)
    ge, gi = reg_obj.compute_error_frozen(pair["src"], tgt, T2)
    oe, oi = oracle.error(oracle.REG[reg], oracle.LOSS[loss], pair["src_h"], pair["cov_s"], pair["tgt_h"],
                          cov_t, pair["nrm_t"], nn_idx, nn_dist, T2, 4.0, 0.7, mode=1)
    assert gi == oi and abs(ge - oe) <= tol * abs(oe)


def test_linearize_without_covs_uses_identity(spx, q, pair):
    # registration.hpp:589-592: missing covariances -> identity (plane-regularised like any other)
    src = spx.PointCloudShared(q, pair["src_h"])
    tgt = spx.PointCloudShared(q, pair["tgt_h"])
    reg_obj = spx.Registration(q, spx.RegistrationParams())
    nn = spx.KNNResult()
    pair["tree"].nearest_neighbor_search_async(src, nn, None, np.eye(4))
    lin = reg_obj._linearize(src, tgt, nn, np.eye(4, dtype=np.float32), 10.0)
    H, b, e, inl = oracle.linearize(3, 0, pair["src_h"], None, pair["tgt_h"], None, None, nn.indices_host(),
                                    nn.distances_host(), np.eye(4), 4.0, 10.0, mode=1)
    assert lin.inlier == inl and rel(lin.H, H) <= 1e-5 and rel(lin.b, b) <= 1e-5


# ---- the reference's own solver tests (T/test_registration_pipeline.cpp:16-61, 411-508)
def make_counting_knn(spx):
    class CountingNearestKNN(spx.KNNBase):
        """host brute-force NN that counts calls (T/test_registration_pipeline.cpp:25-61)"""

        def __init__(self, queue):
            self.queue = queue
            self.call_count = 0
            self.target = None

        def set_target(self, target):
            self.target = target

        def knn_search_async(self, queries, k, result, depends=None, transT=None):
            self.call_count += 1
            T = np.eye(4, dtype=np.float32) if transT is None else np.asarray(transT, np.float32)
            qh, th = queries.points_host(), self.target.points_host()
            tq = (T @ qh.T).T
            d = ((tq[:, None, :3] - th[None, :, :3]) ** 2).sum(-1)
            idx = d.argmin(1).astype(np.int32)
            result.allocate(self.queue, len(qh), k)
            result.indices.upload(idx.reshape(-1, 1))
            result.distances.upload(d[np.arange(len(qh)), idx].astype(np.float32).reshape(-1, 1))

    return CountingNearestKNN


def test_compute_weights_zero_one_for_none_loss(spx, q):
    # RegistrationComputeWeightsUseZeroOneForNoneLoss, T/test_registration_pipeline.cpp:411-436
    src = spx.PointCloudShared(q, np.array([[0, 0, 0, 1], [1, 0, 0, 1], [5, 0, 0, 1]], np.float32))
    tgt = spx.PointCloudShared(q, np.array([[0, 0, 0, 1], [1, 0, 0, 1]], np.float32))
    params = spx.RegistrationParams(reg_type=spx.RegType.POINT_TO_POINT, max_iterations=1,
                                    max_correspondence_distance=1.5)
    reg_obj = spx.Registration(q, params)
    knn = make_counting_knn(spx)(q)
    knn.set_target(tgt)
    reg_obj.align(src, tgt, knn)
    calls = knn.call_count
    w = reg_obj.compute_icp_robust_weights(src, tgt, knn, np.eye(4), params.robust.default_scale)
    assert w.tolist() == [1.0, 1.0, 0.0]
    assert knn.call_count == calls + 1
    # ...FollowsProvidedSource, :438-475
    src2 = spx.PointCloudShared(q, np.array([[0, 0, 0, 1], [1, 0, 0, 1]], np.float32))
    reg_obj.align(src2, tgt, knn)
    w2 = reg_obj.compute_icp_robust_weights(src2, tgt, knn, np.eye(4), params.robust.default_scale)
    assert w2.tolist() == [1.0, 1.0]


def test_compute_weights_uses_provided_robust_scale(spx, q):
    # RegistrationComputeWeightsUsesProvidedRobustScale, T/test_registration_pipeline.cpp:477-508
    src = spx.PointCloudShared(q, np.array([[3, 0, 0, 1]], np.float32))
    tgt = spx.PointCloudShared(q, np.array([[0, 0, 0, 1]], np.float32))
    params = spx.RegistrationParams(reg_type=spx.RegType.POINT_TO_POINT, max_iterations=1,
                                    max_correspondence_distance=10.0)
    params.robust.type = spx.RobustLossType.HUBER
    reg_obj = spx.Registration(q, params)
    knn = make_counting_knn(spx)(q)
    knn.set_target(tgt)
    w = reg_obj.compute_icp_robust_weights(src, tgt, knn, np.eye(4), 1.0)
    assert abs(w[0] - 1 / 3) <= 1e-5
    reg_obj.align(src, tgt, knn)
    w = reg_obj.compute_icp_robust_weights(src, tgt, knn, np.eye(4), 2.0)
    assert abs(w[0] - 2 / 3) <= 1e-5


def test_validate_params_errors(spx, q, pair):
    # registration.hpp:129-193
    src = spx.PointCloudShared(q, pair["src_h"])
    tgt = spx.PointCloudShared(q, pair["tgt_h"])
    with pytest.raises(RuntimeError, match="Covariance matrices of source and target must be pre-computed"):
        spx.Registration(q, spx.RegistrationParams()).align(src, tgt, pair["tree"])
    with pytest.raises(RuntimeError, match="Normal vector or covariance matrices of target"):
        spx.Registration(q, spx.RegistrationParams(reg_type=spx.RegType.POINT_TO_PLANE)).align(src, tgt, pair["tree"])
    # empty source: result is the initial guess (registration.hpp:209-211)
    r = spx.Registration(q, spx.RegistrationParams()).align(
        spx.PointCloudShared(q, np.zeros((0, 4), np.float32)), tgt, pair["tree"], np.eye(4))
    assert np.array_equal(r.T, np.eye(4, dtype=np.float32)) and not r.converged


@pytest.mark.parametrize("reg", REGS)
@pytest.mark.parametrize("opt", ["GN", "LM", "DOGLEG"])
def test_align_matches_oracle_iteration_by_iteration(spx, q, pair, reg, opt):
    """Fixed iteration count (convergence criteria disabled) so discrete events cannot shift the
    comparison; pose after EVERY iteration within 1e-5 m / 1e-5 rad of the oracle."""
    iters = 6
    params = spx.RegistrationParams(reg_type=spx.RegType[reg], max_iterations=iters)
    params.robust.type = spx.RobustLossType.HUBER
    params.robust.default_scale = 1.0
    params.optimization_method = spx.OptimizationMethod({"GN": 0, "LM": 1, "DOGLEG": 2}[opt])
    params.criteria.translation = 0.0
    params.criteria.rotation = 0.0
    tgt, cov_t = clouds_for(spx, q, pair, reg)
    res = spx.Registration(q, params).align(pair["src"], tgt, pair["tree"], trace=True)
    P = oracle.default_params(reg_type=oracle.REG[reg], loss=1, opt_method=oracle.OPT[opt], max_iterations=iters,
                              robust_default_scale=1.0, crit_translation=0.0, crit_rotation=0.0)
    ores = oracle.align(P, pair["src_h"], pair["cov_s"], pair["tgt_h"], cov_t, pair["nrm_t"], pair["otree"],
                        trace=True)
    tol = 1e-5
    for it in range(iters):
        dt, da = pose_delta(ores["trace"][it], res.trace[it])
        assert dt < tol and da < tol, f"iteration {it}: dt={dt:.2e} da={da:.2e}"
    assert res.iterations == ores["iterations"] == iters - 1
    assert res.inlier == ores["inlier"]
    assert abs(res.error - ores["error"]) <= 1e-5 * abs(ores["error"])
    # b is a sum of terms that cancel near the optimum: its error is measured against H's scale
    # times the step the poses are compared at (1e-5), i.e. what b's error can move the solution by
    assert rel(res.H, ores["H"]) <= 1e-5
    assert np.abs(res.b - ores["b"]).max() <= 1e-5 * max(np.abs(ores["b"]).max(), 1e-2 * np.abs(ores["H"]).max())


@pytest.mark.parametrize("opt", ["GN", "LM", "DOGLEG"])
def test_align_converges_like_oracle(spx, q, pair, bundled, opt):
    """Default criteria: same number of iterations, same converged flag, same pose, and the pose
    lands near cpp/data/T_target_source.txt."""
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    params.optimization_method = spx.OptimizationMethod({"GN": 0, "LM": 1, "DOGLEG": 2}[opt])
    res = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"])
    P = oracle.default_params(reg_type=3, loss=1, opt_method=oracle.OPT[opt])
    ores = oracle.align(P, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], None, pair["otree"])
    assert res.converged and ores["converged"] and res.iterations == ores["iterations"]
    dt, da = pose_delta(ores["T"], res.T)
    assert dt < 1e-5 and da < 1e-5
    dt, da = pose_delta(bundled["T_target_source"], res.T)
    assert dt < 0.05 and np.degrees(da) < 0.3


def test_example_registration_pipeline(spx, q, pair, bundled):
    """E/example_registration.cpp:32-45,121 without the random sampling (SURVEY §8(d) cfg 1 (ii)):
    GICP + LM + GEMAN_MCCLURE, max_iter 10, robust auto-scale 10 -> 2.5 in 3 levels."""
    pp = spx.RegistrationPipelineParams()
    pp.registration.max_iterations = 10
    pp.registration.optimization_method = spx.OptimizationMethod.LEVENBERG_MARQUARDT
    pp.registration.robust.type = spx.RobustLossType.GEMAN_MCCLURE
    pp.random_sampling.enable = False
    pp.robust.auto_scale = True
    pp.robust.init_scale, pp.robust.min_scale, pp.robust.auto_scaling_iter = 10.0, 2.5, 3
    res = spx.RegistrationPipeline(q, pp).align(pair["src"], pair["tgt"], pair["tree"], np.eye(4))
    P = oracle.default_params(reg_type=3, loss=4, opt_method=1, max_iterations=10)
    ores = oracle.align_robust(P, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], None, pair["otree"],
                               np.eye(4), 10.0, 2.5, 3)
    dt, da = pose_delta(ores["T"], res.T)
    assert dt < 1e-5 and da < 1e-5 and res.iterations == ores["iterations"]
    dt, da = pose_delta(bundled["T_target_source"], res.T)
    assert dt < 0.05 and np.degrees(da) < 0.3


def test_injected_knn_matches_index_path(spx, q, pair):
    """A user KNNBase (host loop + GPU linearise) and the fused device loop agree."""
    params = spx.RegistrationParams(max_iterations=5)
    params.criteria.translation = params.criteria.rotation = 0.0

    class Wrap(spx.KNNBase):
        def __init__(self, tree):
            self.tree = tree

        def knn_search_async(self, queries, k, result, depends=None, transT=None):
            self.tree.knn_search_async(queries, k, result, depends, transT)

    a = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"], trace=True)
    b = spx.Registration(q, params).align(pair["src"], pair["tgt"], Wrap(pair["tree"]), trace=True)
    for it in range(5):
        dt, da = pose_delta(a.trace[it], b.trace[it])
        assert dt < 1e-5 and da < 1e-5


def test_sharded_sums_match_single(spx, q, pair):
    """Multi-GPU building blocks on one GPU: two source shards linearised separately, their sums
    added (what the NCCL all-reduce does), the update applied -> same poses as the fused loop."""
    import ctypes as C
    L = spx.lib()
    iters = 4
    params = spx.RegistrationParams(max_iterations=iters)
    params.robust.type = spx.RobustLossType.HUBER
    params.criteria.translation = params.criteria.rotation = 0.0
    ref = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"], trace=True)
    ns = pair["src"].size()
    cut = ns // 2 + 7
    shards = []
    for lo, hi in ((0, cut), (cut, ns)):
        cl = spx.PointCloudShared(q, pair["src_h"][lo:hi], pair["cov_s"][lo:hi])
        r = spx.Registration(q, params)
        shards.append((cl, r))
    t16 = np.ascontiguousarray(np.eye(4, dtype=np.float32).T).reshape(16)
    tgt = pair["tgt"]
    for cl, r in shards:
        spx._lib.check(L.spx_registration_shard_begin(r._h, cl.points.ptr, cl.covs.ptr, cl.size(), tgt.points.ptr,
                                                      tgt.covs.ptr, None, tgt.size(), pair["tree"].handle,
                                                      t16.ctypes.data_as(C.POINTER(C.c_float)), -1.0))
    sums = [spx.DeviceArray(q, (32,), np.float64) for _ in shards]
    total = spx.DeviceArray(q, (32,), np.float64)
    for _ in range(iters):
        for (cl, r), s in zip(shards, sums):
            spx._lib.check(L.spx_registration_shard_linearize(r._h, s.ptr))
        total.upload(sums[0].download() + sums[1].download())
        for cl, r in shards:
            spx._lib.check(L.spx_registration_shard_update(r._h, total.ptr))
    outs = []
    for cl, r in shards:
        R = spx._lib.RegistrationResultC()
        spx._lib.check(L.spx_registration_shard_finish(r._h, C.byref(R)))
        outs.append(spx.RegistrationResult.from_c(R))
    assert np.array_equal(outs[0].T, outs[1].T)  # every rank computes the identical update
    dt, da = pose_delta(ref.T, outs[0].T)
    assert dt < 1e-6 and da < 1e-6
    assert outs[0].inlier == ref.inlier


def _sharded_launch(spx, reg, comm_handle, cl, tgt, tree, t16):
    import ctypes as C
    spx._lib.check(spx.lib().spx_registration_align_sharded_launch(
        reg._h, comm_handle, cl.points.ptr if cl.size() else None, cl.covs.ptr if cl.size() else None, cl.size(),
        tgt.points.ptr, tgt.covs.ptr, None, tgt.size(), tree.handle, t16.ctypes.data_as(C.POINTER(C.c_float)), -1.0))


def _sharded_finish(spx, reg):
    import ctypes as C
    R = spx._lib.RegistrationResultC()
    spx._lib.check(spx.lib().spx_registration_align_sharded_finish(reg._h, C.byref(R)))
    return spx.RegistrationResult.from_c(R)


def test_fused_exchange_world1_equals_plain_align(spx, q, pair):
    """The NVLink-mailbox align kernel with a single rank: the self-exchange must leave the sums, and
    therefore every pose, bit-identical to the plain cooperative align."""
    from sycl_points_b200.multi_gpu import LocalCommunicators
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    ref = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"])
    comms = LocalCommunicators([q])
    reg = spx.Registration(q, params)
    spx._lib.check(spx.lib().spx_registration_set_params(reg._h, params.to_c()))
    t16 = np.ascontiguousarray(np.eye(4, dtype=np.float32).T).reshape(16)
    for _ in range(2):  # twice: the sequence numbers keep advancing across aligns
        _sharded_launch(spx, reg, comms.handles[0], pair["src"], pair["tgt"], pair["tree"], t16)
        out = _sharded_finish(spx, reg)
        assert np.array_equal(out.T, ref.T) and out.iterations == ref.iterations and out.converged == ref.converged
        assert out.inlier == ref.inlier and np.array_equal(out.H, ref.H) and np.array_equal(out.b, ref.b)
    comms.close()


def test_fused_exchange_keeps_correspondences_bit_exactly(spx, q, pair):
    """The sharded one-launch kernel carries correspondences over from one iteration to the next when their certified
    margin outlasts the query's motion (nn_search_grid_keep).  15 forced iterations from an offset start, single rank:
    H, b, the pose and the inlier count equal the plain cooperative align's (which searches every query every time),
    and a good part of the correspondences was in fact kept."""
    from sycl_points_b200.multi_gpu import LocalCommunicators
    params = spx.RegistrationParams(max_iterations=15)
    params.robust.type = spx.RobustLossType.HUBER
    params.criteria.translation = params.criteria.rotation = 0.0
    T0 = np.eye(4, dtype=np.float32)
    T0[:3, 3] = [0.3, -0.2, 0.05]
    ref = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"], T0)
    comms = LocalCommunicators([q])
    reg = spx.Registration(q, params)
    spx._lib.check(spx.lib().spx_registration_set_params(reg._h, params.to_c()))
    t16 = np.ascontiguousarray(T0.T).reshape(16)
    _sharded_launch(spx, reg, comms.handles[0], pair["src"], pair["tgt"], pair["tree"], t16)
    out = _sharded_finish(spx, reg)
    assert np.array_equal(out.T, ref.T) and out.iterations == ref.iterations == 14
    assert out.inlier == ref.inlier and np.array_equal(out.H, ref.H) and np.array_equal(out.b, ref.b)
    kept = reg.kept_correspondences()
    assert kept > 5 * pair["src"].size(), f"only {kept} of {15 * pair['src'].size()} correspondences kept"
    comms.close()


@pytest.mark.parametrize("split", ["half", "empty_tail"])
def test_fused_exchange_two_ranks_on_one_gpu(spx, pair, split):
    """Two ranks of the exchange protocol co-resident on ONE device (two queues, the persistent
    grids capped so that both fit): both end with the identical pose, equal to the unsharded align."""
    from sycl_points_b200.multi_gpu import LocalCommunicators
    qa, qb = spx.DeviceQueue(0), spx.DeviceQueue(0)
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    ref = spx.Registration(qa, params).align(pair["src"], pair["tgt"], pair["tree"])
    params.max_blocks = 64
    ns = pair["src"].size()
    cut = ns // 2 + 7 if split == "half" else ns
    tgt_a, tgt_b = pair["tgt"], spx.PointCloudShared(qb, pair["tgt_h"], pair["cov_t"])
    tree_b = spx.KDTree.build(qb, tgt_b)
    shards = [spx.PointCloudShared(qa, pair["src_h"][:cut], pair["cov_s"][:cut]),
              spx.PointCloudShared(qb, pair["src_h"][cut:], pair["cov_s"][cut:]) if cut < ns else spx.PointCloudShared(qb)]
    regs = [spx.Registration(qa, params), spx.Registration(qb, params)]
    for r in regs:
        spx._lib.check(spx.lib().spx_registration_set_params(r._h, params.to_c()))
    comms = LocalCommunicators([qa, qb])
    qa.wait()
    qb.wait()
    t16 = np.ascontiguousarray(np.eye(4, dtype=np.float32).T).reshape(16)
    _sharded_launch(spx, regs[0], comms.handles[0], shards[0], tgt_a, pair["tree"], t16)
    _sharded_launch(spx, regs[1], comms.handles[1], shards[1], tgt_b, tree_b, t16)
    outs = [_sharded_finish(spx, r) for r in regs]
    assert np.array_equal(outs[0].T, outs[1].T)
    assert outs[0].iterations == outs[1].iterations == ref.iterations and outs[0].converged == ref.converged
    dt, da = pose_delta(ref.T, outs[0].T)
    assert dt < 1e-6 and da < 1e-6
    assert outs[0].inlier == ref.inlier
    comms.close()


def test_fused_exchange_across_two_gpus(spx, pair):
    """The same protocol with the ranks on two devices (peer access over NVLink); skipped on a
    single-GPU box (the two-ranks-on-one-GPU test covers the protocol there)."""
    from sycl_points_b200.multi_gpu import LocalCommunicators, shard_bounds
    if spx.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    qs = [spx.DeviceQueue(0), spx.DeviceQueue(1)]
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    ref = spx.Registration(qs[0], params).align(pair["src"], pair["tgt"], pair["tree"])
    bounds = shard_bounds(pair["src"].size(), 2)
    tgts = [spx.PointCloudShared(qq, pair["tgt_h"], pair["cov_t"]) for qq in qs]
    trees = [spx.KDTree.build(qq, t) for qq, t in zip(qs, tgts)]
    shards = [spx.PointCloudShared(qq, pair["src_h"][lo:hi], pair["cov_s"][lo:hi]) for qq, (lo, hi) in zip(qs, bounds)]
    regs = [spx.Registration(qq, params) for qq in qs]
    for r in regs:
        spx._lib.check(spx.lib().spx_registration_set_params(r._h, params.to_c()))
    comms = LocalCommunicators(qs)
    for qq in qs:
        qq.wait()
    t16 = np.ascontiguousarray(np.eye(4, dtype=np.float32).T).reshape(16)
    for r, h, s, t, tr in zip(regs, comms.handles, shards, tgts, trees):
        _sharded_launch(spx, r, h, s, t, tr, t16)
    outs = [_sharded_finish(spx, r) for r in regs]
    assert np.array_equal(outs[0].T, outs[1].T)
    dt, da = pose_delta(ref.T, outs[0].T)
    assert dt < 1e-6 and da < 1e-6 and outs[0].inlier == ref.inlier
    comms.close()


def test_example_registration_pipeline_with_random_sampling(spx, q, pair, bundled):
    """E/example_registration.cpp as shipped (SURVEY §8(d) cfg 1 (i)): the pipeline's default random
    sampling draws 1000 source points from mt19937(1234); same subset as the oracle, same pose."""
    pp = spx.RegistrationPipelineParams()
    pp.registration.max_iterations = 10
    pp.registration.optimization_method = spx.OptimizationMethod.LEVENBERG_MARQUARDT
    pp.registration.robust.type = spx.RobustLossType.GEMAN_MCCLURE
    pp.robust.auto_scale = True
    pp.robust.init_scale, pp.robust.min_scale, pp.robust.auto_scaling_iter = 10.0, 2.5, 3
    assert pp.random_sampling.enable and pp.random_sampling.num == 1000  # reference defaults
    pipe = spx.RegistrationPipeline(q, pp)
    res = pipe.align(pair["src"], pair["tgt"], pair["tree"], np.eye(4))
    keep = oracle.Rng(1234).random_sampling_flags(len(pair["src_h"]), 1000).astype(bool)
    inp = pipe.get_registration_input_point_cloud()
    assert inp.size() == 1000 and np.array_equal(inp.points_host(), pair["src_h"][keep])
    assert np.array_equal(inp.covs_host(), pair["cov_s"][keep])
    P = oracle.default_params(reg_type=3, loss=4, opt_method=1, max_iterations=10)
    ores = oracle.align_robust(P, pair["src_h"][keep], pair["cov_s"][keep], pair["tgt_h"], pair["cov_t"], None,
                               pair["otree"], np.eye(4), 10.0, 2.5, 3)
    dt, da = pose_delta(ores["T"], res.T)
    assert dt < 1e-5 and da < 1e-5 and res.iterations == ores["iterations"]
    dt, da = pose_delta(bundled["T_target_source"], res.T)
    assert dt < 0.10 and np.degrees(da) < 0.5


@pytest.mark.parametrize("max_corr,shift", [(2.0, 0.0), (2.0, 1.5), (0.5, 0.3), (6.0, 4.0)])
@pytest.mark.parametrize("opt", ["GN", "LM"])
def test_align_correspondences_exact_with_warm_start(spx, q, pair, max_corr, shift, opt):
    """The fused nearest-neighbour search of the iteration kernels (warm start from the previous
    iteration, pruned first pass, warp-cooperative continuation): after 3 iterations the cached
    correspondences must be the exact brute-force answer at the pose the last iteration linearised
    at — bit-exact index and distance for everything within max_correspondence_distance."""
    import ctypes as C
    params = spx.RegistrationParams(max_iterations=3)
    params.max_correspondence_distance = max_corr
    params.optimization_method = spx.OptimizationMethod.GAUSS_NEWTON if opt == "GN" else \
        spx.OptimizationMethod.LEVENBERG_MARQUARDT
    params.criteria.translation = params.criteria.rotation = 0.0
    T0 = np.eye(4, dtype=np.float32)
    T0[:3, 3] = [shift, -0.5 * shift, 0.1 * shift]
    reg = spx.Registration(q, params)
    res = reg.align(pair["src"], pair["tgt"], pair["tree"], T0, trace=True)
    ip, dp, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
    spx._lib.check(spx.lib().spx_registration_neighbors(reg._h, C.byref(ip), C.byref(dp), C.byref(n)))
    assert n.value == pair["src"].size()
    idx = np.empty(n.value, np.int32)
    dist = np.empty(n.value, np.float32)
    spx._lib.check(spx.lib().spx_memcpy_d2h(q.handle, idx.ctypes.data_as(C.c_void_p), ip, idx.nbytes))
    spx._lib.check(spx.lib().spx_memcpy_d2h(q.handle, dist.ctypes.data_as(C.c_void_p), dp, dist.nbytes))
    q.wait()
    T_lin = res.trace[1]  # pose after 2 updates = the pose the 3rd iteration searched at
    oi, od = oracle.knn_bruteforce(pair["src_h"], pair["tgt_h"], 1, T_lin)
    oi, od = oi.reshape(-1), od.reshape(-1)
    within = od <= np.float32(max_corr) ** 2
    assert within.sum() > 100
    assert np.array_equal(idx[within], oi[within]) and np.array_equal(dist[within], od[within])
    assert (dist[~within] > np.float32(max_corr) ** 2).all()


@pytest.mark.parametrize("reg", REGS)
def test_split_kernel_path_equals_fused(spx, q, pair, reg, monkeypatch):
    """Large clouds run the Gauss-Newton loop as three launches per iteration (search kernels with
    their own register budget) instead of one cooperative launch; forced here on the small pair:
    identical correspondences, identical sums, hence bit-identical poses, iterations and inliers."""
    params = spx.RegistrationParams(reg_type=spx.RegType[reg], max_iterations=7)
    params.robust.type = spx.RobustLossType.HUBER
    T0 = np.eye(4, dtype=np.float32)
    T0[:3, 3] = [0.3, -0.2, 0.05]
    tgt, _ = clouds_for(spx, q, pair, reg)
    fused = spx.Registration(q, params).align(pair["src"], tgt, pair["tree"], T0, trace=True)
    monkeypatch.setenv("SPX_SPLIT_MIN", "0")
    r = spx.Registration(q, params)
    split = r.align(pair["src"], tgt, pair["tree"], T0, trace=True)
    assert r.last_timing()["launches"] >= 3
    monkeypatch.delenv("SPX_SPLIT_MIN")
    assert split.iterations == fused.iterations and split.converged == fused.converged
    assert split.inlier == fused.inlier
    assert np.array_equal(split.trace, fused.trace) and np.array_equal(split.T, fused.T)


@pytest.mark.parametrize("reg", ["GICP", "POINT_TO_POINT"])
def test_kept_correspondences_equal_a_search(spx, q, pair, reg, monkeypatch):
    """The split-kernel loop keeps a query's correspondence while the query has moved less than half the margin its
    last search certified (icp_keep, spx_registration.cu).  15 forced iterations from an offset start: the pose after
    EVERY iteration, the final neighbour indices and their distances must equal, bit for bit, those of a loop that
    searches every query every time — and most late-iteration correspondences must actually have been kept."""
    import ctypes as C
    params = spx.RegistrationParams(reg_type=spx.RegType[reg], max_iterations=15)
    params.robust.type = spx.RobustLossType.HUBER
    params.criteria.translation = params.criteria.rotation = 0.0
    T0 = np.eye(4, dtype=np.float32)
    T0[:3, 3] = [0.3, -0.2, 0.05]
    tgt, _ = clouds_for(spx, q, pair, reg)
    monkeypatch.setenv("SPX_SPLIT_MIN", "0")

    def run():
        r = spx.Registration(q, params)
        res = r.align(pair["src"], tgt, pair["tree"], T0, trace=True)
        ip, dp, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        spx._lib.check(spx.lib().spx_registration_neighbors(r._h, C.byref(ip), C.byref(dp), C.byref(n)))
        idx, dist = np.empty(n.value, np.int32), np.empty(n.value, np.float32)
        spx._lib.check(spx.lib().spx_memcpy_d2h(q.handle, idx.ctypes.data_as(C.c_void_p), ip, idx.nbytes))
        spx._lib.check(spx.lib().spx_memcpy_d2h(q.handle, dist.ctypes.data_as(C.c_void_p), dp, dist.nbytes))
        q.wait()
        return res, idx, dist, r.kept_correspondences()

    monkeypatch.setenv("SPX_KEEP_FRAC", "-1")
    every, idx_e, dist_e, kept_e = run()
    assert kept_e == 0
    monkeypatch.delenv("SPX_KEEP_FRAC")
    keep, idx_k, dist_k, kept_k = run()
    n = pair["src"].size()
    assert kept_k > 5 * n, f"only {kept_k} of {15 * n} correspondences kept"
    assert np.array_equal(keep.trace, every.trace) and np.array_equal(keep.T, every.T)
    assert keep.inlier == every.inlier and keep.error == every.error
    assert np.array_equal(idx_k, idx_e) and np.array_equal(dist_k, dist_e)


def test_config2_full_size_pipeline_matches_oracle(spx, q):
    """BASELINE config 2 at FULL size (2.0 M raw points per cloud -> ~120 k after the 0.25 m voxel grid):
    every stage of the hot path against the oracle on the same synthetic pair — voxel output and k = 10
    neighbour indices bit-exact, covariances bit-exact, GICP pose within 1e-5 m / 1e-5 rad (the north
    star's tolerance) with the same iteration count and inlier count, and the recovered motion within
    2 cm of the generator's ground truth."""
    import synthetic
    tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
    vg = spx.VoxelGrid(q, 0.25)
    src, tgt = vg.downsampling(spx.PointCloudShared(q, src_raw)), vg.downsampling(spx.PointCloudShared(q, tgt_raw))
    o_src, o_tgt = oracle.voxel_downsample(src_raw, 0.25), oracle.voxel_downsample(tgt_raw, 0.25)
    assert 110_000 < len(o_src) < 130_000
    assert np.array_equal(src.points_host(), o_src) and np.array_equal(tgt.points_host(), o_tgt)
    ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
    nn_s, nn_t = ts.knn_search(src, 10), tt.knn_search(tgt, 10)
    ots, ott = oracle.KDTree(o_src), oracle.KDTree(o_tgt)
    oi_s, od_s = ots.knn(o_src, 10)
    oi_t, od_t = ott.knn(o_tgt, 10)
    assert np.array_equal(nn_s.indices_host(), oi_s) and np.array_equal(nn_s.distances_host(), od_s)
    assert np.array_equal(nn_t.indices_host(), oi_t) and np.array_equal(nn_t.distances_host(), od_t)
    spx.covariance.estimate(nn_s, src)
    spx.covariance.estimate(nn_t, tgt)
    oc_s, oc_t = oracle.covariance(o_src, oi_s), oracle.covariance(o_tgt, oi_t)
    assert np.array_equal(src.covs_host(), oc_s) and np.array_equal(tgt.covs_host(), oc_t)
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    res = spx.Registration(q, params).align(src, tgt, tt)
    P = oracle.default_params(reg_type=3, loss=1)
    ores = oracle.align(P, o_src, oc_s, o_tgt, oc_t, None, ott)
    dt, da = pose_delta(ores["T"], res.T)
    assert dt < 1e-5 and da < 1e-5, (dt, da)
    assert res.iterations == ores["iterations"] and res.converged == ores["converged"]
    assert res.inlier == ores["inlier"]
    dt, da = pose_delta(T_gt, res.T)
    assert dt < 0.02 and np.degrees(da) < 0.05
    params.reg_type = spx.RegType.POINT_TO_POINT
    res = spx.Registration(q, params).align(src, tgt, tt)
    P = oracle.default_params(reg_type=0, loss=1)
    ores = oracle.align(P, o_src, None, o_tgt, None, None, ott)
    dt, da = pose_delta(ores["T"], res.T)
    assert dt < 1e-5 and da < 1e-5, (dt, da)
    assert res.iterations == ores["iterations"] and res.inlier == ores["inlier"]


# ---- batched align (BASELINE config 5 building block): one launch for P pairs
@pytest.mark.parametrize("reg", ["GICP", "POINT_TO_PLANE", "POINT_TO_POINT", "POINT_TO_DISTRIBUTION"])
def test_align_batch_equals_single_bit_for_bit(spx, q, pair, reg):
    """spx_registration_align_batch: every pair's result is bit for bit the single-pair result (the
    per-pair sums are folded in chunk order, independent of the batch and of the grid), for pairs of
    different sizes, different initial guesses and different iteration counts, incl. an empty source."""
    rs = np.random.RandomState(11)
    tgt, cov_t = clouds_for(spx, q, pair, reg)
    src_h, cov_s = pair["src_h"], pair["cov_s"]
    params = spx.RegistrationParams(reg_type=spx.RegType[reg])
    params.robust.type = spx.RobustLossType.HUBER
    params.robust.default_scale = 1.0
    reg_obj = spx.Registration(q, params)
    pairs = []
    for j, n in enumerate([len(src_h), 3000, 257, 256, 1, 0, 4500, 1000]):
        sel = np.sort(rs.choice(len(src_h), n, replace=False)) if n else np.zeros(0, np.int64)
        s = spx.PointCloudShared(q, src_h[sel], cov_s[sel]) if n else spx.PointCloudShared(q, np.zeros((0, 4), np.float32))
        T0 = oracle.se3_exp(rs.normal(0, [0.005, 0.005, 0.005, 0.05, 0.05, 0.02]).astype(np.float32)) if j % 2 else None
        pairs.append((s, tgt, pair["tree"], T0))
    batch = reg_obj.align_batch(pairs)
    singles = [reg_obj.align(s, t, k, T0) for s, t, k, T0 in pairs]
    its = set()
    for b, s1 in zip(batch, singles):
        assert np.array_equal(b.T, s1.T) and b.iterations == s1.iterations and b.converged == s1.converged
        assert np.array_equal(b.H, s1.H) and np.array_equal(b.b, s1.b) and b.error == s1.error and b.inlier == s1.inlier
        its.add(b.iterations)
    assert len(its) > 1  # the batch really mixes pairs that stop at different iterations
    # and the first pair (the whole cloud) against the oracle
    P = oracle.default_params(reg_type=oracle.REG[reg], loss=1, robust_default_scale=1.0)
    ores = oracle.align(P, src_h, cov_s, pair["tgt_h"], cov_t, pair["nrm_t"], pair["otree"])
    dt, da = pose_delta(ores["T"], batch[0].T)
    assert dt < 1e-5 and da < 1e-5 and batch[0].iterations == ores["iterations"]


def test_align_batch_64_pairs_of_60k(spx, q):
    """config-5 shape: 64 pairs of ~60 k points (voxelised synthetic scans, 4 distinct scenes x 16 random
    motions), GICP: the batch equals the single-pair path bit for bit and three of them are checked against the
    oracle at 1e-5."""
    import synthetic
    rs = np.random.RandomState(5)
    vg = spx.VoxelGrid(q, 0.25)
    scenes = []
    for seed in range(4):
        tgt_raw, _, _ = synthetic.kitti_pair(100 + seed, sweeps=8)
        tgt = vg.downsampling(spx.PointCloudShared(q, tgt_raw))
        tree = spx.KDTree.build(q, tgt)
        spx.covariance.estimate(tree.knn_search(tgt, 10), tgt)
        scenes.append((tgt, tree, tgt_raw))
    pairs, keep = [], []
    for j in range(64):
        tgt, tree, tgt_raw = scenes[j % 4]
        T_gt = synthetic.random_pose(rs, 0.6, 1.0).astype(np.float32)
        src_raw = oracle.transform_points(np.linalg.inv(T_gt), tgt_raw[rs.rand(len(tgt_raw)) < 0.9])
        src = vg.downsampling(spx.PointCloudShared(q, src_raw))
        ts = spx.KDTree.build(q, src)
        spx.covariance.estimate(ts.knn_search(src, 10), src)
        ts.close()
        pairs.append((src, tgt, tree, None))
        keep.append(T_gt)
    assert 40_000 < pairs[0][0].size() < 80_000
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    reg_obj = spx.Registration(q, params)
    batch = reg_obj.align_batch(pairs)
    for j in (0, 17, 63):
        s1 = reg_obj.align(*pairs[j])
        assert np.array_equal(batch[j].T, s1.T) and batch[j].iterations == s1.iterations and batch[j].inlier == s1.inlier
        src, tgt, _, _ = pairs[j]
        P = oracle.default_params(reg_type=3, loss=1)
        ores = oracle.align(P, src.points_host(), src.covs_host(), tgt.points_host(), tgt.covs_host(), None,
                            oracle.KDTree(tgt.points_host()))
        dt, da = pose_delta(ores["T"], batch[j].T)
        assert dt < 1e-5 and da < 1e-5 and batch[j].iterations == ores["iterations"], (j, dt, da)
    for j, r in enumerate(batch):
        dt, da = pose_delta(keep[j], r.T)
        assert r.converged and dt < 0.05, (j, dt)


def test_spx_align_batch_raw_pairs_equal_the_single_pair_chain(spx, q):
    """spx_align_batch (raw clouds in, results out; lanes + one batched align) == the single-pair chain
    voxel -> index -> KNN -> covariance -> align, bit for bit, incl. a target shared by several pairs."""
    import synthetic
    rs = np.random.RandomState(9)
    raws = []
    for seed in range(2):
        tgt_raw, src_raw, _ = synthetic.kitti_pair(200 + seed, sweeps=4)
        raws.append((spx.PointCloudShared(q, src_raw), spx.PointCloudShared(q, tgt_raw)))
    pairs = []
    for j in range(6):
        s, t = raws[j % 2]
        T0 = oracle.se3_exp(rs.normal(0, [0.002, 0.002, 0.004, 0.05, 0.05, 0.01]).astype(np.float32)) if j >= 2 else None
        pairs.append((s, t, T0))
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    ba = spx.BatchAligner(q, params, 0.25, 10, lanes=3)
    res, ns, nt = ba.align(pairs)
    vg = spx.VoxelGrid(q, 0.25)
    reg_obj = spx.Registration(q, params)
    for j, (s_raw, t_raw, T0) in enumerate(pairs):
        src, tgt = vg.downsampling(s_raw), vg.downsampling(t_raw)
        assert ns[j] == src.size() and nt[j] == tgt.size()
        ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
        spx.covariance.estimate(ts.knn_search(src, 10), src)
        spx.covariance.estimate(tt.knn_search(tgt, 10), tgt)
        one = reg_obj.align(src, tgt, tt, T0)
        assert np.array_equal(res[j].T, one.T) and res[j].iterations == one.iterations and res[j].inlier == one.inlier
        assert res[j].converged == one.converged and res[j].error == one.error
    assert ba.last_timing()["iterations"] >= 1
    ba.close()


# ---- GenZ (factor.hpp:378-449, registration.hpp:464-511)
@pytest.mark.parametrize("thr", [0.2, 0.02, 0.005])
@pytest.mark.parametrize("loss", ["NONE", "HUBER"])
def test_genz_linearize_error_weights_vs_oracle(spx, q, pair, thr, loss):
    """GenZ: plane factor x alpha for planar correspondences, point factor x (1 - alpha) for the others, alpha =
    planar inliers / inliers recounted per linearisation.  Thresholds chosen so that the planar share of the
    bundled pair is ~100 %, ~60 % and ~10 %."""
    oracle.set_genz_planarity_threshold(thr)
    try:
        T = oracle.se3_exp(np.array([0.004, -0.003, 0.012, 0.4, 0.1, -0.02], np.float32))
        nn_idx, nn_dist = pair["otree"].knn(pair["src_h"], 1, T)
        alpha = oracle.genz_alpha(pair["cov_t"], nn_idx, nn_dist, 4.0)
        params = spx.RegistrationParams(reg_type=spx.RegType.GENZ)
        params.genz.planarity_threshold = thr
        params.robust.type = spx.RobustLossType[loss]
        params.robust.default_scale = 0.7
        reg_obj = spx.Registration(q, params)
        lin = reg_obj.compute_linearized_result(pair["src"], pair["tgt"], pair["tree"], T)
        H, b, e, inl = oracle.linearize(4, oracle.LOSS[loss], pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"],
                                        pair["nrm_t"], nn_idx, nn_dist, T, 4.0, 0.7, mode=1)
        assert lin.inlier == inl
        assert rel(lin.H, H) <= 1e-5 and rel(lin.b, b) <= 1e-5 and abs(lin.error - e) <= 1e-5 * abs(e), (alpha, rel(lin.H, H))
        T2 = (T @ oracle.se3_exp(np.array([1e-3, 2e-3, -1e-3, 0.01, -0.02, 0.005], np.float32))).astype(np.float32)
        ge, gi = reg_obj.compute_error_frozen(pair["src"], pair["tgt"], T2)
        oe, oi = oracle.error(4, oracle.LOSS[loss], pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], pair["nrm_t"],
                              nn_idx, nn_dist, T2, 4.0, 0.7, mode=1)
        assert gi == oi and abs(ge - oe) <= 1e-5 * abs(oe)
        w = reg_obj.compute_icp_robust_weights(pair["src"], pair["tgt"], pair["tree"], T, 0.7)
        ow = oracle.robust_weights(4, oracle.LOSS[loss], pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"],
                                   pair["nrm_t"], nn_idx, nn_dist, T, 4.0, 0.7)
        assert np.abs(w - ow).max() <= 1e-6
        if thr == 0.02:
            assert 0.2 < alpha < 0.9, alpha  # a real mix of both factors
    finally:
        oracle.set_genz_planarity_threshold(0.2)


@pytest.mark.parametrize("opt", ["GN", "LM", "DOGLEG"])
def test_genz_align_matches_oracle(spx, q, pair, opt):
    thr = 0.02
    oracle.set_genz_planarity_threshold(thr)
    try:
        iters = 5
        params = spx.RegistrationParams(reg_type=spx.RegType.GENZ, max_iterations=iters)
        params.genz.planarity_threshold = thr
        params.robust.type = spx.RobustLossType.HUBER
        params.robust.default_scale = 1.0
        params.optimization_method = spx.OptimizationMethod({"GN": 0, "LM": 1, "DOGLEG": 2}[opt])
        params.criteria.translation = 0.0
        params.criteria.rotation = 0.0
        res = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"], trace=True)
        P = oracle.default_params(reg_type=4, loss=1, opt_method=oracle.OPT[opt], max_iterations=iters,
                                  robust_default_scale=1.0, crit_translation=0.0, crit_rotation=0.0)
        ores = oracle.align(P, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], pair["nrm_t"], pair["otree"],
                            trace=True)
        for it in range(iters):
            dt, da = pose_delta(ores["trace"][it], res.trace[it])
            assert dt < 1e-5 and da < 1e-5, f"iteration {it}: dt={dt:.2e} da={da:.2e}"
        assert res.inlier == ores["inlier"]
        # batched entry point with GenZ: one pair after the other, same results
        reg_obj = spx.Registration(q, params)
        b = reg_obj.align_batch([(pair["src"], pair["tgt"], pair["tree"], None)] * 2)
        assert np.array_equal(b[0].T, res.T) and np.array_equal(b[1].T, res.T)
    finally:
        oracle.set_genz_planarity_threshold(0.2)


# ---- rotation constraint (rotation_constraint.hpp:15-121; registration.hpp:629-649,757-764)
@pytest.mark.parametrize("reg", ["GICP", "POINT_TO_PLANE", "GENZ"])
def test_rotation_constraint_vs_oracle(spx, q, pair, reg):
    """the Jensen-Bregman LogDet term on the raw covariances, added to every correspondence of any factor:
    linearise / frozen error <= 1e-5 vs the oracle, and the GN / LM pose traces <= 1e-5."""
    oracle.set_rotation_constraint(True, 0.5, 3.0)
    try:
        T = oracle.se3_exp(np.array([0.02, -0.015, 0.03, 0.4, 0.1, -0.02], np.float32))
        nn_idx, nn_dist = pair["otree"].knn(pair["src_h"], 1, T)
        params = spx.RegistrationParams(reg_type=spx.RegType[reg])
        params.robust.type = spx.RobustLossType.HUBER
        params.robust.default_scale = 0.7
        params.rotation_constraint.enable = True
        params.rotation_constraint.weight = 0.5
        params.rotation_constraint.robust.default_scale = 3.0
        reg_obj = spx.Registration(q, params)
        lin = reg_obj.compute_linearized_result(pair["src"], pair["tgt"], pair["tree"], T)
        H, b, e, inl = oracle.linearize(oracle.REG[reg], 1, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"],
                                        pair["nrm_t"], nn_idx, nn_dist, T, 4.0, 0.7, mode=1)
        assert lin.inlier == inl
        assert rel(lin.H, H) <= 1e-5 and rel(lin.b, b) <= 1e-5 and abs(lin.error - e) <= 1e-5 * abs(e)
        # the term is really there
        oracle.set_rotation_constraint(False)
        H0, _, e0, _ = oracle.linearize(oracle.REG[reg], 1, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"],
                                        pair["nrm_t"], nn_idx, nn_dist, T, 4.0, 0.7, mode=1)
        oracle.set_rotation_constraint(True, 0.5, 3.0)
        assert rel(H0, H) > 1e-3 and e > e0
        ge, gi = reg_obj.compute_error_frozen(pair["src"], pair["tgt"], T)
        oe, oi = oracle.error(oracle.REG[reg], 1, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], pair["nrm_t"],
                              nn_idx, nn_dist, T, 4.0, 0.7, mode=1)
        assert gi == oi and abs(ge - oe) <= 1e-5 * abs(oe)
        for opt in ("GN", "LM"):
            iters = 4
            params.max_iterations = iters
            params.optimization_method = spx.OptimizationMethod({"GN": 0, "LM": 1}[opt])
            params.criteria.translation = params.criteria.rotation = 0.0
            res = spx.Registration(q, params).align(pair["src"], pair["tgt"], pair["tree"], trace=True)
            P = oracle.default_params(reg_type=oracle.REG[reg], loss=1, opt_method=oracle.OPT[opt], max_iterations=iters,
                                      robust_default_scale=0.7, crit_translation=0.0, crit_rotation=0.0)
            ores = oracle.align(P, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], pair["nrm_t"], pair["otree"],
                                trace=True)
            for it in range(iters):
                dt, da = pose_delta(ores["trace"][it], res.trace[it])
                assert dt < 1e-5 and da < 1e-5, f"{opt} iteration {it}: dt={dt:.2e} da={da:.2e}"
    finally:
        oracle.set_rotation_constraint(False)
    src_nocov = spx.PointCloudShared(q, pair["src_h"])
    with pytest.raises(RuntimeError, match="Covariance matrices of source are required"):
        reg_obj.align(src_nocov, pair["tgt"], pair["tree"])


# ------------------------------------------------------------------ solver add-ons: nl-reg and MAP prior
@pytest.mark.parametrize("opt", ["GN", "LM", "DOGLEG"])
@pytest.mark.parametrize("addon", ["nl_reg", "map_prior", "both"])
def test_align_with_addons_matches_oracle(spx, q, pair, opt, addon):
    """Degenerate regularisation (degenerate_regularization.hpp:58-112) and the MAP prior (map_prior.hpp:30-146)
    between the linearisation and the step (registration.hpp:248-253): pose after every iteration within 1e-5 of
    the oracle's restatement, H/b/error as returned (regularised) and H_raw/b_raw (raw) both checked.
    Thresholds are set so that nl-reg really fires (every rotation eigen-direction of this pair is 'degenerate'
    under a threshold of 1e9) and the prior comes from a real previous align."""
    iters = 5
    reg = "GICP"
    params = spx.RegistrationParams(reg_type=spx.RegType[reg], max_iterations=iters)
    params.robust.type = spx.RobustLossType.HUBER
    params.robust.default_scale = 1.0
    params.optimization_method = spx.OptimizationMethod({"GN": 0, "LM": 1, "DOGLEG": 2}[opt])
    params.criteria.translation = params.criteria.rotation = 0.0
    nl, mp = addon in ("nl_reg", "both"), addon in ("map_prior", "both")
    if nl:
        params.degenerate_reg.type = spx.DegenerateRegularizationType.nl_reg
        params.degenerate_reg.rot_eigenvalue_threshold = 1e9
        params.degenerate_reg.trans_eigenvalue_threshold = 0.5
        params.degenerate_reg.base_factor = 20.0
    params.map_prior.enabled = mp
    robj = spx.Registration(q, params)
    P = oracle.default_params(reg_type=oracle.REG[reg], loss=1, opt_method=oracle.OPT[opt], max_iterations=iters,
                              robust_default_scale=1.0, crit_translation=0.0, crit_rotation=0.0)
    oracle.set_addons(oracle.make_addons(nl_reg=nl, rot_thr=1e9, trans_thr=0.5, base_factor=20.0, map_prior=mp))
    try:
        T0 = oracle.se3_exp(np.array([0.002, -0.004, 0.006, 0.1, -0.05, 0.02], np.float32))
        if mp:
            # the "previous frame": a plain GN align without add-ons on both sides, then a predicted pose
            oracle.set_addons(oracle.make_addons())
            Pp = oracle.default_params(reg_type=oracle.REG[reg], loss=1, opt_method=0, max_iterations=4,
                                       robust_default_scale=1.0)
            oprev = oracle.align(Pp, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], pair["nrm_t"],
                                 pair["otree"])
            pp = spx.RegistrationParams(reg_type=spx.RegType[reg], max_iterations=4)
            pp.robust.type = spx.RobustLossType.HUBER
            pp.robust.default_scale = 1.0
            gprev = spx.Registration(q, pp).align(pair["src"], pair["tgt"], pair["tree"])
            oracle.set_addons(oracle.make_addons(nl_reg=nl, rot_thr=1e9, trans_thr=0.5, base_factor=20.0, map_prior=True))
            T_pred = (gprev.T @ oracle.se3_exp(np.array([0.01, 0.0, -0.02, 0.3, 0.1, 0.0], np.float32))).astype(np.float32)
            act_o, om_o = oracle.set_map_prior_state(oprev, T_pred)
            assert robj.set_map_prior_state(gprev, T_pred) and act_o
            om_g = robj._prior_omega.reshape(6, 6)
            assert rel(om_g, om_o) <= 1e-4  # built from each side's own H_raw (1e-5 apart) through a 6x6 inverse
            T0 = T_pred
        res = robj.align(pair["src"], pair["tgt"], pair["tree"], T0, trace=True)
        ores = oracle.align(P, pair["src_h"], pair["cov_s"], pair["tgt_h"], pair["cov_t"], pair["nrm_t"], pair["otree"],
                            T_init=T0, trace=True)
    finally:
        oracle.set_addons(oracle.make_addons())
    tol = 2e-5 if mp else 1e-5  # the prior's information matrix differs by the two previous aligns' rounding
    for it in range(iters):
        dt, da = pose_delta(ores["trace"][it], res.trace[it])
        assert dt < tol and da < tol, f"iteration {it}: dt={dt:.2e} da={da:.2e}"
    assert res.iterations == ores["iterations"] == iters - 1
    assert res.inlier == ores["inlier"]
    assert rel(res.H_raw, ores["H_raw"]) <= 1e-5
    assert rel(res.H, ores["H"]) <= (1e-4 if mp else 1e-5)
    assert rel(res.H, res.H_raw) > 1e-5  # the add-on really changed the system (H of GICP is ~1e6-1e7 here, the prior ~1e2-1e4)
    assert abs(res.error - ores["error"]) <= 2e-5 * abs(ores["error"])
    # with the add-ons cleared the same handle is back on the one-launch path and equals a fresh plain align
    params.degenerate_reg.type = spx.DegenerateRegularizationType.none
    params.map_prior.enabled = False
    plain = robj.align(pair["src"], pair["tgt"], pair["tree"], T0)
    params2 = spx.RegistrationParams(reg_type=spx.RegType[reg], max_iterations=iters)
    params2.robust.type = spx.RobustLossType.HUBER
    params2.robust.default_scale = 1.0
    params2.optimization_method = params.optimization_method
    params2.criteria.translation = params2.criteria.rotation = 0.0
    fresh = spx.Registration(q, params2).align(pair["src"], pair["tgt"], pair["tree"], T0)
    assert np.array_equal(plain.T, fresh.T)


def test_degenerate_regularize_entry_point_vs_oracle(spx, q, pair):
    """compute_linearized_result(..., initial_pose) (registration.hpp:312-323) and the host entry point
    spx_degenerate_regularize against the oracle, including the no-op cases (type none, zero inliers)."""
    T = oracle.se3_exp(np.array([0.004, -0.003, 0.012, 0.4, 0.1, -0.02], np.float32))
    T0 = np.eye(4, dtype=np.float32)
    params = spx.RegistrationParams(reg_type=spx.RegType.GICP)
    params.degenerate_reg.type = spx.DegenerateRegularizationType.nl_reg
    params.degenerate_reg.rot_eigenvalue_threshold = 1e9
    params.degenerate_reg.trans_eigenvalue_threshold = 1e9
    robj = spx.Registration(q, params)
    raw = robj.compute_linearized_result(pair["src"], pair["tgt"], pair["tree"], T)
    regd = robj.compute_linearized_result(pair["src"], pair["tgt"], pair["tree"], T, initial_pose=T0)
    Ho, bo = oracle.degenerate_regularize(oracle.make_addons(nl_reg=True, rot_thr=1e9, trans_thr=1e9), raw.H, raw.b,
                                          raw.inlier, T, T0)
    assert rel(regd.H, Ho) <= 1e-6 and rel(regd.b, bo) <= 1e-5
    # every direction penalised with lambda = inlier: H grows by lambda * I on both diagonal blocks
    lam = float(raw.inlier)
    np.testing.assert_allclose(regd.H - raw.H, lam * np.eye(6), atol=2e-4 * lam)
    params.degenerate_reg.type = spx.DegenerateRegularizationType.none
    same = robj.compute_linearized_result(pair["src"], pair["tgt"], pair["tree"], T, initial_pose=T0)
    assert np.array_equal(same.H, raw.H) and np.array_equal(same.b, raw.b)


def test_pipeline_intensity_weighted_sampling(spx, q, pair):
    """RegistrationPipeline with random_sampling.use_intensities (registration_pipeline.hpp:131-134): the source is
    drawn by mixed_random_sampling with the intensities as weights — the registration input equals the oracle's
    selection from the same mt19937(1234) stream, and the align runs on it."""
    n = len(pair["src_h"])
    inten = np.random.default_rng(12).uniform(0.0, 1.0, n).astype(np.float32)
    src = spx.PointCloudShared(q, pair["src_h"], pair["cov_s"])
    src.set_intensities(inten)
    pp = spx.RegistrationPipelineParams()
    pp.random_sampling.num = 1500
    pp.random_sampling.use_intensities = True
    pp.random_sampling.weighted_ratio = 0.6
    pipe = spx.RegistrationPipeline(q, pp)
    res = pipe.align(src, pair["tgt"], pair["tree"])
    used = pipe.get_registration_input_point_cloud()
    keep = oracle.Rng(1234).mixed_random_sampling_flags(inten, 1500, 0.6).astype(bool)
    assert used.size() == 1500 and np.array_equal(used.points_host(), pair["src_h"][keep])
    assert res.inlier > 1000
