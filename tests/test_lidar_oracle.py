"""CPU checks of the oracle's lidar tight-coupling pass (oracle/refba.cpp: lidarAssociate, uError, uLinearize), the
restatement of g2oOptimizer.cc:979-1117 + types_six_dof_expmap.h:206-262 + base_unary_edge.hpp:58-123 that the GPU
path is compared with in test_gpu_lidar.py.  The reference has no test or golden vector for this block (it needs PCL,
OpenCV and lidar data), so the restatement is checked for internal consistency: the central-difference Jacobians agree
with the closed form, the association agrees with an independent numpy nearest-neighbour search, absent edges leave
the third pass untouched, and the edges do pull the current keyframe onto the lidar structure."""
import numpy as np
import pytest

from oracle import refba


def world_points(pose7, pts):
    """Twc applied to points of a keyframe (float32 boundary as in the reference: float Tcw, double product, float out)."""
    t, q = pose7[:3], pose7[3:]
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    Rf = R.astype(np.float32).astype(np.float64)
    tf = t.astype(np.float32).astype(np.float64)
    Ow = (-(Rf.T @ tf)).astype(np.float32).astype(np.float64)
    return (pts.astype(np.float64) @ Rf + Ow).astype(np.float32)  # p_w = R^T p_c + Ow


@pytest.fixture(scope="module")
def window(synth):
    prob = synth.small_window(seed=1, n_free=6, n_fixed=3, n_points=300)
    return prob, synth.lidar_data(prob, seed=1, n_flat=400, n_corner=100)


def test_numeric_and_closed_form_jacobians_agree(window):
    prob, ld = window
    sol = {}
    for numeric in (True, False):
        r = refba.RefBA(prob)
        r.set_lidar(ld, numeric_jacobian=numeric)
        r.solve_local(third_pass_iters=20)
        sol[numeric] = (r.poses(), r.trace(), r.lidar_matches(), r.num_lidar_edges(), r.outliers())
    assert sol[True][3] == sol[False][3] > 100
    assert np.array_equal(sol[True][2], sol[False][2])
    assert len(sol[True][1]) == len(sol[False][1])
    # central differences with delta = 1e-9 carry ~1e-7 of rounding noise per entry: the iterates agree to that level
    np.testing.assert_allclose(sol[True][1][:, 5], sol[False][1][:, 5], rtol=1e-8)
    assert np.abs(sol[True][0] - sol[False][0]).max() < 1e-7
    assert np.array_equal(sol[True][4], sol[False][4])


def test_association_is_the_nearest_neighbour(window):
    prob, ld = window
    # two visual passes first: the association uses the pass-2 estimates
    r2 = refba.RefBA(prob)
    r2.solve_local(third_pass_iters=0)
    P2 = r2.poses()
    r = refba.RefBA(prob)
    r.set_lidar(ld)
    r.solve_local(third_pass_iters=20)
    m = r.lidar_matches()
    nf = len(ld.flat_xyz)
    for cur, mapx, mapp, sl in ((ld.flat_xyz, ld.map_flat_xyz, ld.map_flat_pose, slice(0, nf)),
                                (ld.corner_xyz, ld.map_corner_xyz, ld.map_corner_pose, slice(nf, None))):
        mw = np.zeros_like(mapx)
        for k in np.unique(mapp):
            mw[mapp == k] = world_points(P2[k], mapx[mapp == k])
        cw = world_points(P2[ld.cur_pose], cur)
        d2 = ((cw[:, None, :].astype(np.float64) - mw[None, :, :]) ** 2).sum(-1)
        nn, dmin = d2.argmin(1), d2.min(1)
        got = m[sl]
        clear = np.abs(dmin - ld.distance_sq_threshold) > 1e-4  # away from the threshold (float vs double distances)
        assert np.array_equal(got[clear] >= 0, (dmin < ld.distance_sq_threshold)[clear])
        hit = got >= 0
        # the matched point is (one of) the nearest: equal distance up to float rounding
        np.testing.assert_allclose(d2[np.flatnonzero(hit), got[hit]], dmin[hit], rtol=1e-4, atol=1e-7)
        assert (got[hit] == nn[hit]).mean() > 0.99
    assert r.num_lidar_edges() == int((m >= 0).sum())


def test_no_edges_means_plain_third_pass(window):
    prob, ld = window
    a = refba.RefBA(prob)
    a.solve_local(third_pass_iters=20)
    b = refba.RefBA(prob)
    n = 10
    b.set_lidar_edges(ld.cur_pose, np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), n_flat=6)
    b.solve_local(third_pass_iters=20)
    assert np.array_equal(a.poses(), b.poses()) and np.array_equal(a.trace(), b.trace())
    # a far-away map gives no correspondence at all
    far = type(ld)(**{**ld.__dict__, "map_flat_xyz": ld.map_flat_xyz + 500.0, "map_corner_xyz": ld.map_corner_xyz + 500.0})
    c = refba.RefBA(prob)
    c.set_lidar(far)
    c.solve_local(third_pass_iters=20)
    assert c.num_lidar_edges() == 0 and (c.lidar_matches() == -1).all()
    assert np.array_equal(a.poses(), c.poses())


def test_edges_pull_the_keyframe_onto_the_planes(synth):
    # explicit point-to-plane edges that say "the ground is 5 cm higher than the vision-only solution thinks"
    prob = synth.small_window(seed=4, n_free=5, n_fixed=3, n_points=250)
    base = refba.RefBA(prob)
    base.solve_local(third_pass_iters=20)
    cur = int(np.flatnonzero(prob.pose_fixed == 0)[-1])
    P = base.poses()[cur]
    rng = np.random.default_rng(0)
    n = 300
    pc = np.stack([rng.uniform(-8, 8, n), np.full(n, 1.65), rng.uniform(3, 25, n)], -1)  # in the keyframe's frame
    nrm = np.tile([0.0, -1.0, 0.0], (n, 1))
    # world points = Twc * (pc shifted by 5 cm along the normal)
    t, q = P[:3], P[3:]
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    qw = (pc + 0.05 * nrm - t) @ R  # R^T (p - t)
    r = refba.RefBA(prob)
    r.set_lidar_edges(cur, pc, qw, nrm, np.full(n, 5000.0), n_flat=n)
    r.solve_local(third_pass_iters=20)
    tr = r.trace()
    p3 = tr[tr[:, 0] == 2]
    assert p3[0, 4] > p3[-1, 5]  # cost of the third pass went down
    # residual of the edges at the final pose: well below the 5 cm they started with
    Pn = r.poses()[cur]
    t2, (x, y, z, w) = Pn[:3], Pn[3:]
    R2 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    res = ((qw @ R2.T + t2 - pc) * nrm).sum(1)
    assert np.abs(res).mean() < 0.02
