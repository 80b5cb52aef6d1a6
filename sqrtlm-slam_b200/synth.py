"""Seeded synthetic KITTI-00-shaped bundle-adjustment problems.

The reference ships no data and no tests, so every workload in BASELINE.json is
generated here.  Constants come from the reference's own KITTI config
(cfg/KITTI00-02.yaml:8-11,18-19,25,42,45) and the pyramid sigma table is computed
in float32 exactly as the ORB extractor does (src/frontend/ORBextractor.cc:482-505).

All state crosses the BA boundary the way the reference's does: poses as float32
4x4 Tcw matrices converted with the Converter::toSE3Quat recipe
(src/utils/Converter.cc:55-68), points as float32 3-vectors, measurements as
float32 (cv::KeyPoint::pt, mvuRight, mvInvLevelSigma2 are float containers,
include/data_structure/KeyFrame.h:394,402-403,427).

Layout produced (identical to what include/sqrtba.h consumes):
  pose_qt   (n_pose,7) f64  tx,ty,tz,qx,qy,qz,qw   (SE3Quat::toVector order, se3quat.h:138-148)
  pose_fixed(n_pose,)  u8
  cam       (n_pose,5) f64  fx,fy,cx,cy,bf
  point_xyz (n_point,3) f64
  obs_pose  (n_obs,) i32 ; obs_point (n_obs,) i32  -- grouped by landmark, pose-sorted inside
  obs_meas  (n_obs,4) f32  u,v,ur(<0 = monocular),invSigma2
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

# cfg/KITTI00-02.yaml -- the reference stores these as float (KeyFrame.h:394)
FX = float(np.float32(718.856))
FY = float(np.float32(718.856))
CX = float(np.float32(607.1928))
CY = float(np.float32(185.2157))
BF = float(np.float32(386.1448))
IMG_W, IMG_H = 1241.0, 376.0
N_LEVELS = 8


def inv_level_sigma2() -> np.ndarray:
    """mvInvLevelSigma2 in float32, ORBextractor.cc:482-505 (scaleFactor_ is a double member
    initialised from the float 1.2f, include/frontend/ORBextractor.h:187)."""
    scale = float(np.float32(1.2))
    sf = np.zeros(N_LEVELS, dtype=np.float32)
    s2 = np.zeros(N_LEVELS, dtype=np.float32)
    sf[0] = 1.0
    s2[0] = 1.0
    for i in range(1, N_LEVELS):
        sf[i] = np.float32(float(sf[i - 1]) * scale)
        s2[i] = np.float32(sf[i] * sf[i])
    return (np.float32(1.0) / s2).astype(np.float32)


# ----------------------------------------------------------------------------- small SO3/SE3 helpers (vectorised)

def _skew(w):
    z = np.zeros(w.shape[:-1])
    return np.stack([np.stack([z, -w[..., 2], w[..., 1]], -1),
                     np.stack([w[..., 2], z, -w[..., 0]], -1),
                     np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def so3_exp(w):
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = _skew(w)
    th_s = np.where(th < 1e-12, 1.0, th)
    a = np.where(th < 1e-12, 1.0, np.sin(th_s) / th_s)
    b = np.where(th < 1e-12, 0.5, (1 - np.cos(th_s)) / (th_s * th_s))
    return np.eye(3) + a * K + b * (K @ K)


def rotmat_to_quat_eigen(R, normalize=True):
    """Eigen's Quaterniond(Matrix3d) branches (what SE3Quat(R,t) calls, se3quat.h:58),
    followed by SE3Quat::normalizeRotation (se3quat.h:280-285) unless normalize=False (g2o::Sim3(R,t,s) keeps the
    quaternion as Eigen returns it, sim3.h:63-66). Returns (...,4) x,y,z,w."""
    R = np.asarray(R, dtype=np.float64)
    flat = R.reshape(-1, 3, 3)
    out = np.zeros((flat.shape[0], 4))
    for n, m in enumerate(flat):
        t = m[0, 0] + m[1, 1] + m[2, 2]
        if t > 0:
            t = np.sqrt(t + 1.0)
            w = 0.5 * t
            t = 0.5 / t
            q = [(m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t, w]
        else:
            i = 0
            if m[1, 1] > m[0, 0]:
                i = 1
            if m[2, 2] > m[i, i]:
                i = 2
            j = (i + 1) % 3
            k = (j + 1) % 3
            t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
            q = [0.0, 0.0, 0.0, 0.0]
            q[i] = 0.5 * t
            t = 0.5 / t
            q[3] = (m[k, j] - m[j, k]) * t
            q[j] = (m[j, i] + m[i, j]) * t
            q[k] = (m[k, i] + m[i, k]) * t
        q = np.array(q)
        if not normalize:
            out[n] = q
            continue
        if q[3] < 0:
            q = -q
        out[n] = q / np.sqrt(np.dot(q, q))
    return out.reshape(R.shape[:-2] + (4,))


def quat_to_rotmat(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


# ----------------------------------------------------------------------------- problem container

@dataclass
class Problem:
    pose_qt: np.ndarray
    pose_fixed: np.ndarray
    cam: np.ndarray
    point_xyz: np.ndarray
    obs_pose: np.ndarray
    obs_point: np.ndarray
    obs_meas: np.ndarray
    name: str = ""
    truth: dict = field(default_factory=dict)

    @property
    def n_pose(self):
        return int(self.pose_qt.shape[0])

    @property
    def n_point(self):
        return int(self.point_xyz.shape[0])

    @property
    def n_obs(self):
        return int(self.obs_pose.shape[0])

    @property
    def n_free(self):
        return int((self.pose_fixed == 0).sum())

    def lm_ptr(self) -> np.ndarray:
        cnt = np.bincount(self.obs_point, minlength=self.n_point)
        return np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)

    def copy(self) -> "Problem":
        return Problem(self.pose_qt.copy(), self.pose_fixed.copy(), self.cam.copy(), self.point_xyz.copy(),
                       self.obs_pose.copy(), self.obs_point.copy(), self.obs_meas.copy(), self.name, dict(self.truth))

    def save(self, path):
        np.savez_compressed(path, pose_qt=self.pose_qt, pose_fixed=self.pose_fixed, cam=self.cam,
                            point_xyz=self.point_xyz, obs_pose=self.obs_pose, obs_point=self.obs_point,
                            obs_meas=self.obs_meas, name=np.array(self.name))

    @staticmethod
    def load(path) -> "Problem":
        z = np.load(path)
        return Problem(z["pose_qt"], z["pose_fixed"], z["cam"], z["point_xyz"], z["obs_pose"], z["obs_point"],
                       z["obs_meas"], str(z["name"]))


# ----------------------------------------------------------------------------- generator

def _trajectory(rng, n_kf, loop=False):
    """KITTI-like forward motion: camera z forward, x right, y down. Returns R_wc (n,3,3), c_w (n,3)."""
    step = rng.uniform(0.8, 1.5, n_kf)
    dyaw = np.deg2rad(rng.normal(0.0, 1.5, n_kf))
    if loop:
        # a bit more than one full turn so the tail re-observes the head (loop-closure overlap)
        dyaw = dyaw * 0.3 + 2.0 * np.pi * 1.04 / n_kf
    yaw = np.cumsum(dyaw) - dyaw[0]
    pitch = np.deg2rad(rng.normal(0.0, 0.2, n_kf))
    roll = np.deg2rad(rng.normal(0.0, 0.2, n_kf))
    # yaw about camera y (down) axis, pitch about x, roll about z
    Ry = so3_exp(np.stack([np.zeros(n_kf), yaw, np.zeros(n_kf)], -1))
    Rx = so3_exp(np.stack([pitch, np.zeros(n_kf), np.zeros(n_kf)], -1))
    Rz = so3_exp(np.stack([np.zeros(n_kf), np.zeros(n_kf), roll], -1))
    R_wc = Ry @ Rx @ Rz
    fwd = R_wc[:, :, 2]
    c = np.zeros((n_kf, 3))
    c[1:] = np.cumsum(fwd[:-1] * step[1:, None], axis=0)
    return R_wc, c


def make_problem(seed: int, n_kf: int, n_fixed_head: int, n_points: int, mean_track: float,
                 stereo: bool = True, outlier_frac: float = 0.05, loop: bool = False,
                 cand_halfwidth: int | None = None, extra_fixed=(), name: str = "",
                 max_track: int | None = None) -> Problem:
    """n_kf keyframes in trajectory order; the first n_fixed_head are fixed (plus `extra_fixed` indices)."""
    rng = np.random.default_rng(seed)
    R_wc, c_w = _trajectory(rng, n_kf, loop=loop)
    R_cw = np.swapaxes(R_wc, -1, -2)
    t_cw = -np.einsum("nij,nj->ni", R_cw, c_w)

    inv_s2 = inv_level_sigma2()

    # ---- points: back-project a random pixel at a random depth from a random non-fixed "home" keyframe
    lo_home = min(n_fixed_head, n_kf - 1)
    home = rng.integers(lo_home, n_kf, n_points)
    pu = rng.uniform(0, IMG_W, n_points)
    pv = rng.uniform(0, IMG_H, n_points)
    depth = rng.uniform(4.0, 60.0, n_points)
    xc = np.stack([(pu - CX) / FX * depth, (pv - CY) / FY * depth, depth], -1)
    Xw = np.einsum("nij,nj->ni", R_wc[home], xc) + c_w[home]

    # ---- candidate keyframes per point
    if cand_halfwidth is None:
        cand = np.broadcast_to(np.arange(n_kf)[None, :], (n_points, n_kf))
        valid = np.ones_like(cand, dtype=bool)
    else:
        off = np.arange(-cand_halfwidth, cand_halfwidth + 1)
        if loop:
            per = int(round(n_kf / 1.04))
            off = np.concatenate([off, off + per, off - per])
        cand = home[:, None] + off[None, :]
        valid = (cand >= 0) & (cand < n_kf)
        cand = np.clip(cand, 0, n_kf - 1)

    # ---- true projections into candidates
    Xc = np.einsum("pkij,pj->pki", R_cw[cand], Xw) + t_cw[cand]
    z = Xc[..., 2]
    zs = np.where(z > 0.5, z, 1.0)
    u = FX * Xc[..., 0] / zs + CX
    v = FY * Xc[..., 1] / zs + CY
    vis = valid & (z > 0.5) & (z < 90.0) & (u >= 0) & (u < IMG_W) & (v >= 0) & (v < IMG_H)
    vis[np.arange(n_points), np.argmax(cand == home[:, None], axis=1)] = True  # home always sees it

    # ---- contiguous-in-time tracks: keep the L visible candidates nearest (by index) to the home keyframe
    L = 2 + rng.poisson(max(mean_track - 2.0, 0.1), n_points)
    if max_track is not None:
        L = np.minimum(L, max_track)
    dist = np.abs(cand - home[:, None]).astype(np.float64)
    dist = dist + rng.uniform(0, 0.5, dist.shape)  # break ties between the two sides
    dist[~vis] = np.inf
    order = np.argsort(dist, axis=1)
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.broadcast_to(np.arange(order.shape[1])[None, :], order.shape), axis=1)
    keep = vis & (rank < L[:, None])

    # drop points with fewer than 2 observations (cannot be triangulated; the front-end never creates them)
    nobs_pt = keep.sum(1)
    good = nobs_pt >= 2
    keep = keep[good]
    cand_g, Xw, depth_g = cand[good], Xw[good], depth[good]
    u, v, z = u[good], v[good], z[good]
    n_pt = int(good.sum())

    pt_idx, col = np.nonzero(keep)  # row-major => grouped by landmark
    kf_idx = cand_g[pt_idx, col]
    # sort observations inside each landmark by keyframe index
    o = np.lexsort((kf_idx, pt_idx))
    pt_idx, col, kf_idx = pt_idx[o], col[o], kf_idx[o]
    n_obs = pt_idx.size
    ut, vt, zt = u[pt_idx, col], v[pt_idx, col], z[pt_idx, col]

    # ---- measurements
    level = np.minimum(rng.geometric(0.35, n_obs) - 1, N_LEVELS - 1)
    sigma = 1.2 ** level
    mu = ut + rng.normal(0, 1, n_obs) * sigma
    mv = vt + rng.normal(0, 1, n_obs) * sigma
    mur = ut - BF / zt + rng.normal(0, 1, n_obs) * sigma
    is_out = rng.uniform(0, 1, n_obs) < outlier_frac
    n_out = int(is_out.sum())
    sgn = lambda n: rng.choice([-1.0, 1.0], n)
    mu[is_out] += sgn(n_out) * rng.uniform(10, 50, n_out)
    mv[is_out] += sgn(n_out) * rng.uniform(10, 50, n_out)
    mur[is_out] += sgn(n_out) * rng.uniform(10, 50, n_out)
    meas = np.zeros((n_obs, 4), dtype=np.float32)
    meas[:, 0] = mu.astype(np.float32)
    meas[:, 1] = mv.astype(np.float32)
    if stereo:
        # a real right-image coordinate is never negative; the adapter uses ur<0 as "monocular"
        meas[:, 2] = np.maximum(mur, 0.0).astype(np.float32)
    else:
        meas[:, 2] = -1.0
    meas[:, 3] = inv_s2[level]

    # ---- initial estimates: perturbed truth, then rounded through the float32 map boundary
    fixed = np.zeros(n_kf, dtype=np.uint8)
    fixed[:n_fixed_head] = 1
    for i in extra_fixed:
        fixed[i] = 1
    xi_r = np.deg2rad(rng.normal(0, 0.3, (n_kf, 3)))
    xi_t = rng.normal(0, 0.05, (n_kf, 3))
    free = fixed == 0
    dR = so3_exp(xi_r)
    R0 = np.where(free[:, None, None], dR @ R_cw, R_cw)
    t0 = np.where(free[:, None], np.einsum("nij,nj->ni", dR, t_cw) + xi_t, t_cw)
    R0f = R0.astype(np.float32).astype(np.float64)
    t0f = t0.astype(np.float32).astype(np.float64)
    q0 = rotmat_to_quat_eigen(R0f)
    pose_qt = np.concatenate([t0f, q0], axis=1)

    X0 = Xw + rng.normal(0, 1, Xw.shape) * (0.02 * depth_g)[:, None]
    X0 = X0.astype(np.float32).astype(np.float64)

    cam = np.tile(np.array([FX, FY, CX, CY, BF]), (n_kf, 1))
    truth = dict(R_cw=R_cw, t_cw=t_cw, Xw=Xw, is_outlier=is_out)
    return Problem(np.ascontiguousarray(pose_qt), fixed, np.ascontiguousarray(cam), np.ascontiguousarray(X0),
                   kf_idx.astype(np.int32), pt_idx.astype(np.int32), meas, name, truth)


# ----------------------------------------------------------------------------- the BASELINE.json configs

def config_c0(seed=0, scale=1.0):
    """C0: stereo local window, 20 free + 10 fixed KFs, ~6k points, ~70k observations."""
    return make_problem(seed, 30, 10, int(6000 * scale), 11.7, stereo=True, name=f"C0-seed{seed}")


def config_c1(seed=0, scale=1.0):
    """C1: the same window in monocular mode; first window KF additionally fixed (gauge)."""
    return make_problem(seed, 30, 10, int(6000 * scale), 11.7, stereo=False, extra_fixed=(10,),
                        name=f"C1-seed{seed}")


def config_c2(seed=0, scale=1.0):
    """C2: large stereo window, 100 KFs (first fixed), 50k points, ~600k observations."""
    return make_problem(seed, 100, 1, int(50000 * scale), 12.0, stereo=True, cand_halfwidth=40,
                        name=f"C2-seed{seed}")


def config_c3(seed=0, scale=1.0, n_kf=1500):
    """C3: global BA, 1500 KFs on a loop (first fixed), 300k points, ~3M observations."""
    return make_problem(seed, int(n_kf), 1, int(300000 * scale), 10.0, stereo=True, loop=True,
                        cand_halfwidth=20, name=f"C3-seed{seed}")


def small_window(seed=0, n_free=6, n_fixed=3, n_points=150, mean_track=5.0, stereo=True, outlier_frac=0.05,
                 extra_fixed=()):
    """Tiny window for parity tests."""
    return make_problem(seed, n_free + n_fixed, n_fixed, n_points, mean_track, stereo=stereo,
                        outlier_frac=outlier_frac, extra_fixed=extra_fixed, name=f"small-seed{seed}")


def concat_windows(windows):
    """Batch of independent windows (config C4) as one block-diagonal problem + CSR window offsets."""
    pose_ptr = np.zeros(len(windows) + 1, dtype=np.int64)
    point_ptr = np.zeros(len(windows) + 1, dtype=np.int64)
    obs_ptr = np.zeros(len(windows) + 1, dtype=np.int64)
    for i, w in enumerate(windows):
        pose_ptr[i + 1] = pose_ptr[i] + w.n_pose
        point_ptr[i + 1] = point_ptr[i] + w.n_point
        obs_ptr[i + 1] = obs_ptr[i] + w.n_obs
    p = Problem(
        np.concatenate([w.pose_qt for w in windows]),
        np.concatenate([w.pose_fixed for w in windows]),
        np.concatenate([w.cam for w in windows]),
        np.concatenate([w.point_xyz for w in windows]),
        np.concatenate([w.obs_pose + pose_ptr[i] for i, w in enumerate(windows)]).astype(np.int32),
        np.concatenate([w.obs_point + point_ptr[i] for i, w in enumerate(windows)]).astype(np.int32),
        np.concatenate([w.obs_meas for w in windows]),
        name=f"batch{len(windows)}")
    return p, pose_ptr, point_ptr, obs_ptr


def frame_problem(seed=0, n_points=1500, stereo=False, outlier_frac=0.1, pose_noise=(0.05, 0.3)):
    """One tracked frame for pose-only optimisation (g2oOptimizer::PoseOptimization): map points in front of the
    camera, their measurements in the frame, a perturbed initial pose.  Returns (pose7, cam5, xyz[n,3], meas[n,4])
    plus the true pose and the injected-outlier mask.  Map points and the pose cross the boundary as float32, the way
    Frame::mTcw (cv::Mat CV_32F) and MapPoint::GetWorldPos() do (g2oOptimizer.cc:407, 472-475)."""
    rng = np.random.default_rng(seed)
    R_wc, c_w = _trajectory(rng, 3)
    R_cw = np.swapaxes(R_wc, -1, -2)[1]
    t_cw = -R_cw @ c_w[1]
    pu = rng.uniform(0, IMG_W, n_points)
    pv = rng.uniform(0, IMG_H, n_points)
    depth = rng.uniform(4.0, 60.0, n_points)
    xc = np.stack([(pu - CX) / FX * depth, (pv - CY) / FY * depth, depth], -1)
    Xw = (xc - t_cw) @ R_cw  # R_cw^T (xc - t)
    inv_s2 = inv_level_sigma2()
    level = np.minimum(rng.geometric(0.35, n_points) - 1, N_LEVELS - 1)
    sigma = 1.2 ** level
    mu = pu + rng.normal(0, 1, n_points) * sigma
    mv = pv + rng.normal(0, 1, n_points) * sigma
    mur = pu - BF / depth + rng.normal(0, 1, n_points) * sigma
    is_out = rng.uniform(0, 1, n_points) < outlier_frac
    k = int(is_out.sum())
    mu[is_out] += rng.choice([-1.0, 1.0], k) * rng.uniform(10, 50, k)
    mv[is_out] += rng.choice([-1.0, 1.0], k) * rng.uniform(10, 50, k)
    mur[is_out] += rng.choice([-1.0, 1.0], k) * rng.uniform(10, 50, k)
    meas = np.zeros((n_points, 4), dtype=np.float32)
    meas[:, 0], meas[:, 1] = mu, mv
    meas[:, 2] = np.maximum(mur, 0.0) if stereo else -1.0
    meas[:, 3] = inv_s2[level]
    dR = so3_exp(np.deg2rad(rng.normal(0, pose_noise[1], 3)))
    R0 = (dR @ R_cw).astype(np.float32).astype(np.float64)
    t0 = (dR @ t_cw + rng.normal(0, pose_noise[0], 3)).astype(np.float32).astype(np.float64)
    pose7 = np.concatenate([t0, rotmat_to_quat_eigen(R0)])
    cam = np.array([FX, FY, CX, CY, BF])
    truth = dict(R_cw=R_cw, t_cw=t_cw, is_outlier=is_out)
    return pose7, cam, Xw.astype(np.float32).astype(np.float64), meas, truth


# ----------------------------------------------------------------------------- lidar feature clouds (row N4)

@dataclass
class LidarData:
    """Inputs of the lidar tight-coupling pass (include/sqrtba.h: sqrtba_lidar), all float32 like pcl::PointXYZI.

    Defaults of the thresholds / weights: include/utils/lidarconfig.h:53-56 (cfg/lidar_slam.yaml:52-61 only uses the
    flat points; both kinds are generated so both edge types are exercised)."""
    cur_pose: int
    flat_xyz: np.ndarray
    flat_normal: np.ndarray
    corner_xyz: np.ndarray
    map_flat_xyz: np.ndarray
    map_flat_pose: np.ndarray
    map_corner_xyz: np.ndarray
    map_corner_pose: np.ndarray
    distance_sq_threshold: float = 0.2
    flat_weight: float = 50.0
    corner_weight: float = 30.0
    use_flat: bool = True
    use_corner: bool = True


def lidar_data(prob: Problem, seed: int = 0, cur_pose: int | None = None, n_flat: int = 1200, n_corner: int = 300,
               sigma: float = 0.02) -> LidarData:
    """Synthetic lidar features for a window made by make_problem: a ground plane 1.65 m below the cameras, two walls
    and a row of vertical poles, seen from every FREE keyframe (the local keyframes) at its TRUE pose, expressed in
    that keyframe's frame with `sigma` metres of noise.  The current keyframe (default: the newest free one) gets the
    feature + normal clouds of KeyFrame.h:438-442; the others form the local map."""
    rng = np.random.default_rng(10_000 + seed)
    R_cw, t_cw = prob.truth["R_cw"], prob.truth["t_cw"]
    free = np.flatnonzero(prob.pose_fixed == 0)
    if cur_pose is None:
        cur_pose = int(free[-1])
    c_w = -np.einsum("nji,nj->ni", R_cw, t_cw)  # camera centres
    x_mid = float(np.mean(c_w[free, 0]))
    z_lo, z_hi = float(c_w[free, 2].min()) - 10.0, float(c_w[free, 2].max()) + 30.0
    poles = np.stack([x_mid + rng.choice([-6.0, 6.0], 40), np.zeros(40), rng.uniform(z_lo, z_hi, 40)], -1)

    def sample_flat(k, n):
        which = rng.integers(0, 3, n)  # 0 ground, 1 left wall, 2 right wall
        z = c_w[k, 2] + rng.uniform(-5.0, 30.0, n)
        x = np.where(which == 0, c_w[k, 0] + rng.uniform(-12.0, 12.0, n), np.where(which == 1, x_mid - 8.0, x_mid + 8.0))
        y = np.where(which == 0, 1.65, rng.uniform(-3.0, 1.65, n))
        nw = np.where((which == 0)[:, None], np.array([0.0, -1.0, 0.0]),
                      np.where((which == 1)[:, None], np.array([1.0, 0.0, 0.0]), np.array([-1.0, 0.0, 0.0])))
        return np.stack([x, y, z], -1), nw

    def sample_corner(k, n):
        near = np.argsort(np.abs(poles[:, 2] - (c_w[k, 2] + 10.0)))[:12]
        p = poles[rng.choice(near, n)]
        p = p.copy()
        p[:, 1] = rng.uniform(-2.5, 1.5, n)
        return p

    def to_kf(k, pw):
        return np.einsum("ij,nj->ni", R_cw[k], pw) + t_cw[k] + rng.normal(0, sigma, pw.shape)

    fw, nw = sample_flat(cur_pose, n_flat)
    flat = to_kf(cur_pose, fw)
    nrm = np.einsum("ij,nj->ni", R_cw[cur_pose], nw) + rng.normal(0, 0.01, nw.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    corner = to_kf(cur_pose, sample_corner(cur_pose, n_corner))
    mf, mfp, mc, mcp = [], [], [], []
    for k in free:
        if k == cur_pose:
            continue
        mf.append(to_kf(k, sample_flat(k, n_flat)[0]))
        mfp.append(np.full(n_flat, k, np.int32))
        mc.append(to_kf(k, sample_corner(k, n_corner)))
        mcp.append(np.full(n_corner, k, np.int32))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return LidarData(cur_pose, f32(flat), f32(nrm), f32(corner), f32(np.concatenate(mf)), np.concatenate(mfp),
                     f32(np.concatenate(mc)), np.concatenate(mcp))


@dataclass
class FrameLidar:
    """Inputs of the lidar block of PoseOptimization (include/sqrtba.h: sqrtba_frame_lidar; g2oOptimizer.cc:560-640): the
    frame's flat / sharp feature points in its own frame (Frame.h:273-277) and the local lidar map in the world frame
    (Tracking's local_lidarmap_cloud_ptr_), float32 like the pcl point types."""
    flat_xyz: np.ndarray
    flat_normal: np.ndarray
    corner_xyz: np.ndarray
    map_xyz: np.ndarray
    distance_sq_threshold: float = 0.2
    flat_weight: float = 50.0
    corner_weight: float = 30.0
    use_flat: bool = True
    use_corner: bool = True


def frame_lidar(truth, seed: int = 0, n_flat: int = 800, n_corner: int = 200, n_map: int = 20000, sigma: float = 0.02,
                miss_frac: float = 0.1) -> FrameLidar:
    """Lidar features for a frame made by frame_problem (truth = its last return value): a ground plane 1.65 m below
    the camera, two walls and vertical poles.  The map cloud samples the scene in the world frame, the frame's feature
    points sample it again, seen from the TRUE pose with `sigma` metres of noise; `miss_frac` of them are moved away
    from every map point (no correspondence)."""
    rng = np.random.default_rng(20_000 + seed)
    R_cw, t_cw = truth["R_cw"], truth["t_cw"]
    c_w = -R_cw.T @ t_cw
    poles = np.stack([c_w[0] + rng.choice([-6.0, 6.0], 24), np.zeros(24), c_w[2] + rng.uniform(-5.0, 30.0, 24)], -1)

    def flat(n):
        which = rng.integers(0, 3, n)
        z = c_w[2] + rng.uniform(-5.0, 30.0, n)
        x = np.where(which == 0, c_w[0] + rng.uniform(-12.0, 12.0, n), np.where(which == 1, c_w[0] - 8.0, c_w[0] + 8.0))
        y = np.where(which == 0, c_w[1] + 1.65, c_w[1] + rng.uniform(-3.0, 1.65, n))
        nw = np.where((which == 0)[:, None], np.array([0.0, -1.0, 0.0]),
                      np.where((which == 1)[:, None], np.array([1.0, 0.0, 0.0]), np.array([-1.0, 0.0, 0.0])))
        return np.stack([x, y, z], -1), nw

    def corner(n):
        p = poles[rng.integers(0, len(poles), n)].copy()
        p[:, 1] = c_w[1] + rng.uniform(-2.5, 1.5, n)
        return p

    n_mc = n_map // 5
    mapw = np.concatenate([flat(n_map - n_mc)[0], corner(n_mc)])
    fw, nw = flat(n_flat)
    cw = corner(n_corner)
    miss_f, miss_c = rng.uniform(0, 1, n_flat) < miss_frac, rng.uniform(0, 1, n_corner) < miss_frac
    fw[miss_f, 1] -= 6.0                                  # above everything
    cw[miss_c, 0] += 30.0
    to_cam = lambda pw: pw @ R_cw.T + t_cw + rng.normal(0, sigma, pw.shape)
    nrm = nw @ R_cw.T + rng.normal(0, 0.01, nw.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return FrameLidar(f32(to_cam(fw)), f32(nrm), f32(to_cam(cw)), f32(mapw))


# ----------------------------------------------------------------------------- essential graph (Sim3 pose graph, row N3)
# Sim3 as 8 numbers in g2o's operator[] order: qx qy qz qw | tx ty tz | s   (types/sim3.h)

def _q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _q_rot(q, v):
    return quat_to_rotmat(q) @ v


def sim3_mul(a, b):
    return np.concatenate([_q_mul(a[:4], b[:4]), a[7] * _q_rot(a[:4], b[4:7]) + a[4:7], [a[7] * b[7]]])


def sim3_inv(a):
    qi = np.array([-a[0], -a[1], -a[2], a[3]])
    return np.concatenate([qi, _q_rot(qi, (-1.0 / a[7]) * a[4:7]), [1.0 / a[7]]])


def sim3_small(rng, s_rot, s_t, s_scale):
    """A small random similarity: rotation vector ~ N(0, s_rot), translation ~ N(0, s_t), log-scale ~ N(0, s_scale)."""
    w = rng.normal(0, s_rot, 3)
    th = np.linalg.norm(w)
    q = np.concatenate([np.sin(th / 2) * w / th, [np.cos(th / 2)]]) if th > 0 else np.array([0, 0, 0, 1.0])
    return np.concatenate([q, rng.normal(0, s_t, 3), [float(np.exp(rng.normal(0, s_scale))) if s_scale > 0 else 1.0]])


def sim3_pair(seed=0, n_matches=150, fix_scale=False, outlier_frac=0.15, point_noise=0.01, init_noise=(0.01, 0.05, 0.02)):
    """A loop-candidate keyframe pair for g2oOptimizer::OptimizeSim3 (g2oOptimizer.cc:1560-1796): matched map points
    expressed in the frame of keyframe 1 (P1c) and of keyframe 2 (P2c), their keypoints in both images, and an initial
    S12 as a RANSAC Sim3Solver would hand it over (truth times a small similarity).  Map points cross the boundary as
    float32 (cv::Mat CV_32F products, :1650-1662).  Returns (s12_init 8, cam8, p1c[n,3], p2c[n,3], meas6[n,6] float32,
    truth dict)."""
    rng = np.random.default_rng(seed)
    S12 = sim3_small(rng, 0.08, 0.8, 0.0 if fix_scale else 0.15)
    S21 = sim3_inv(S12)
    n = n_matches
    pu, pv = rng.uniform(40, IMG_W - 40, n), rng.uniform(40, IMG_H - 40, n)
    depth = rng.uniform(5.0, 40.0, n)
    p1 = np.stack([(pu - CX) / FX * depth, (pv - CY) / FY * depth, depth], -1)
    p2 = np.array([S21[7] * _q_rot(S21[:4], x) + S21[4:7] for x in p1])
    p2 *= 1.0 + rng.normal(0, point_noise, (n, 1))  # the two maps disagree a little about every point
    p2[:, 2] = np.maximum(p2[:, 2], 1.0)
    inv_s2 = inv_level_sigma2()
    l1 = np.minimum(rng.geometric(0.35, n) - 1, N_LEVELS - 1)
    l2 = np.minimum(rng.geometric(0.35, n) - 1, N_LEVELS - 1)
    meas = np.zeros((n, 6), np.float32)
    meas[:, 0] = pu + rng.normal(0, 1, n) * 1.2 ** l1
    meas[:, 1] = pv + rng.normal(0, 1, n) * 1.2 ** l1
    meas[:, 2] = inv_s2[l1]
    meas[:, 3] = p2[:, 0] / p2[:, 2] * FX + CX + rng.normal(0, 1, n) * 1.2 ** l2
    meas[:, 4] = p2[:, 1] / p2[:, 2] * FY + CY + rng.normal(0, 1, n) * 1.2 ** l2
    meas[:, 5] = inv_s2[l2]
    is_out = rng.uniform(0, 1, n) < outlier_frac
    k = int(is_out.sum())
    side = rng.integers(0, 2, k) * 3
    rows = np.nonzero(is_out)[0]
    meas[rows, side] += (rng.choice([-1.0, 1.0], k) * rng.uniform(8, 60, k)).astype(np.float32)
    meas[rows, side + 1] += (rng.choice([-1.0, 1.0], k) * rng.uniform(8, 60, k)).astype(np.float32)
    s0 = sim3_mul(sim3_small(rng, init_noise[0], init_noise[1], 0.0 if fix_scale else init_noise[2]), S12)
    cam8 = np.array([FX, FY, CX, CY, FX, FY, CX, CY])
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    return s0, cam8, f32(p1), f32(p2), meas, dict(S12=S12, is_outlier=is_out)


def pose_graph(seed=0, n_kf=200, fix_scale=True, covis_span=4, covis_prob=0.5, n_loop=6, radius=60.0):
    """An essential graph the way OptimizeEssentialGraph sees it after a loop closure (g2oOptimizer.cc:1212-1460):
    keyframes on a closed loop, dead-reckoned estimates that drift (in scale too when it is free), spanning-tree edges
    (parent = previous keyframe), covisibility edges to older neighbours, and `n_loop` loop edges between the end of the
    trajectory and its start, measured from the true geometry.  Vertex 0 (the loop keyframe) is fixed.
    Returns (vert0 n x 8 [S_iw], fixed n, edge_ij m x 2 [vertex 0 = i, vertex 1 = j], meas m x 8 [S_ji])."""
    rng = np.random.default_rng(seed)
    truth = []
    for k in range(n_kf):
        ang = 2 * np.pi * k / n_kf
        q = np.array([0, np.sin(ang / 2), 0, np.cos(ang / 2)])
        c = np.array([radius * np.sin(ang), 0.3 * np.sin(3 * ang), radius * (1 - np.cos(ang))])
        truth.append(np.concatenate([q, -_q_rot(q, c), [1.0]]))  # S_kw
    edges, meas = [], []
    est = [truth[0].copy()]
    sc = 0.0 if fix_scale else 0.004
    for k in range(1, n_kf):  # spanning tree: child k -> parent k-1, measurement S_parent,child (vertex 0 = child)
        rel = sim3_mul(truth[k], sim3_inv(truth[k - 1]))  # S_k,k-1
        m = sim3_mul(sim3_small(rng, 0.004, 0.03, sc), rel)
        est.append(sim3_mul(m, est[k - 1]))                # dead reckoning
        edges.append((k, k - 1))
        meas.append(sim3_inv(m))                           # S_ji with i = k, j = k-1
    for k in range(2, n_kf):  # covisibility edges to older neighbours (pKFn->mnId < pKF->mnId), from the estimates' own
        for d in range(2, covis_span + 1):  # relative pose as the reference does (Sji = Sjw * Swi of the SAME estimates)
            if k - d >= 0 and rng.random() < covis_prob:
                edges.append((k, k - d))
                meas.append(sim3_mul(est[k - d], sim3_inv(est[k])))
    for t in range(n_loop):   # loop edges: the last keyframes see the first ones again; measured by the loop detector
        i, j = n_kf - 1 - t, t % 3
        edges.append((i, j))
        meas.append(sim3_mul(sim3_mul(sim3_small(rng, 0.001, 0.01, 0.0), truth[j]), sim3_inv(truth[i])))
    fixed = np.zeros(n_kf, np.uint8)
    fixed[0] = 1
    return np.stack(est), fixed, np.array(edges, np.int32), np.stack(meas)
