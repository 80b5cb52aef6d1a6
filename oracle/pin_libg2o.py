"""Pins the oracle's per-edge arithmetic against the reference's OWN compiled code.

The reference cannot be built in this image (Eigen, OpenCV, PCL, Ceres and ROS are missing), but its tree ships a
prebuilt `Thirdparty/g2o/lib/libg2o.so` whose only dependencies are libstdc++/libm/libc.  The functions below are exported
by that binary and can be called through ctypes with hand-made argument blocks (Itanium C++ ABI on x86-64: a class
returned by value comes back through a hidden first pointer; `this` follows it):

  g2o::SE3Quat::exp(Vector6d const&)                                   se3quat.h:223-257
  g2o::project2d(Vector3d const&)                                      types_six_dof_expmap.cpp:37-42
  g2o::EdgeSE3ProjectXYZ::cam_project(Vector3d const&) const           types_six_dof_expmap.cpp:141-147
  g2o::EdgeStereoSE3ProjectXYZ::cam_project(Vector3d const&, float const&) const      .cpp:150-157 (float invz quirk)
  g2o::RobustKernelHuber::robustify(double, Vector3d&) const           core/robust_kernel_impl.cpp:78-91
  g2o::VertexSE3Expmap::VertexSE3Expmap / setToOriginImpl / oplusImpl  types_six_dof_expmap.h:59-76 (a REAL vertex object)

The edge / kernel objects are never constructed: the methods are const and only read a few scalar members (fx, fy, cx,
cy), whose byte offsets inside `this` are found by probing (set one candidate slot of a zeroed block to 1.0 and watch
the output); the Huber kernel's members are written by the binary's own RobustKernelHuber::setDelta.  TEST INFRASTRUCTURE ONLY; runs where /root/reference exists.  `make_golden()` stores inputs and
the binary's outputs in tests/golden/libg2o_vectors.npz so that the same check runs on machines without the reference.

Caveat: the provenance of the prebuilt binary relative to the fork's modified sources cannot be
verified; the functions pinned here are untouched upstream g2o / ORB-SLAM2 code."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIBG2O = "/root/reference/Thirdparty/g2o/lib/libg2o.so"
SYM = {
    "exp": "_ZN3g2o7SE3Quat3expERKN5Eigen6MatrixIdLi6ELi1ELi0ELi6ELi1EEE",
    "project2d": "_ZN3g2o9project2dERKN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEE",
    "cam_mono": "_ZNK3g2o17EdgeSE3ProjectXYZ11cam_projectERKN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEE",
    "cam_stereo": "_ZNK3g2o23EdgeStereoSE3ProjectXYZ11cam_projectERKN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEERKf",
    "huber": "_ZNK3g2o17RobustKernelHuber9robustifyEdRN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEE",
    "set_delta": "_ZN3g2o17RobustKernelHuber8setDeltaEd",
    "vtx_ctor": "_ZN3g2o15VertexSE3ExpmapC1Ev",
    "vtx_origin": "_ZN3g2o15VertexSE3Expmap15setToOriginImplEv",
    "vtx_oplus": "_ZN3g2o15VertexSE3Expmap9oplusImplEPKd",
}
THIS_DOUBLES = 256  # probe window: 2 KB of `this`


def available() -> bool:
    return os.path.exists(LIBG2O)


def _aligned(n_doubles: int):
    """64-byte aligned block of doubles: Eigen's fixed-size vectorisable members are 16-byte aligned, 32 with AVX
    (Quaterniond inside the returned SE3Quat is written with an aligned 256-bit store)."""
    raw = np.zeros(n_doubles + 8, np.float64)
    off = ((64 - raw.ctypes.data % 64) % 64) // 8
    return raw[off:off + n_doubles]


class LibG2O:
    def __init__(self):
        self.L = C.CDLL(LIBG2O)
        self.f = {k: getattr(self.L, v) for k, v in SYM.items()}
        for k in ("exp", "project2d", "cam_mono", "cam_stereo"):
            self.f[k].restype = C.c_void_p
        self.f["exp"].argtypes = [C.c_void_p, C.c_void_p]
        self.f["project2d"].argtypes = [C.c_void_p, C.c_void_p]
        self.f["cam_mono"].argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.f["cam_stereo"].argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.f["huber"].restype = None
        self.f["huber"].argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        self.f["set_delta"].restype = None
        self.f["set_delta"].argtypes = [C.c_void_p, C.c_double]
        for k in ("vtx_ctor", "vtx_origin"):
            self.f[k].restype = None
            self.f[k].argtypes = [C.c_void_p]
        self.f["vtx_oplus"].restype = None
        self.f["vtx_oplus"].argtypes = [C.c_void_p, C.c_void_p]
        self._mono_off = self._probe_cam(stereo=False)
        self._stereo_off = self._probe_cam(stereo=True)

    # ---- raw calls
    def _cam_raw(self, this, xyz, bf, stereo):
        out = _aligned(4)
        v = _aligned(4)
        v[:3] = xyz
        if stereo:
            b = C.c_float(bf)
            self.f["cam_stereo"](out.ctypes.data, this.ctypes.data, v.ctypes.data, C.addressof(b))
            return out[:3].copy()
        self.f["cam_mono"](out.ctypes.data, this.ctypes.data, v.ctypes.data)
        return out[:2].copy()

    def _probe_cam(self, stereo):
        """offsets (in doubles) of fx, fy, cx, cy inside `this`: res0 = x/z*fx + cx, res1 = y/z*fy + cy."""
        xyz = np.array([2.0, 3.0, 1.0])
        found = {}
        for i in range(THIS_DOUBLES):
            this = _aligned(THIS_DOUBLES)
            this[i] = 1.0
            r = self._cam_raw(this, xyz, 0.0, stereo)
            if r[0] == 2.0 and r[1] == 0.0:
                found["fx"] = i
            elif r[0] == 1.0 and r[1] == 0.0:
                found["cx"] = i
            elif r[1] == 3.0 and r[0] == 0.0:
                found["fy"] = i
            elif r[1] == 1.0 and r[0] == 0.0:
                found["cy"] = i
        assert set(found) == {"fx", "fy", "cx", "cy"}, found
        return found

    # ---- the reference's functions
    def se3_exp(self, upd6):
        out = _aligned(8)
        u = _aligned(6)
        u[:] = upd6
        self.f["exp"](out.ctypes.data, u.ctypes.data)
        return np.array([out[4], out[5], out[6], out[0], out[1], out[2], out[3]])  # (t, q) like SE3Quat::toVector

    def oplus_chain(self, updates):
        """VertexSE3Expmap::oplusImpl applied in sequence from the origin (types_six_dof_expmap.h:73-76: estimate <-
        exp(update) * estimate, i.e. SE3Quat::operator* + normalizeRotation, se3quat.h:104-110,280-285).  A real vertex
        object is built with the binary's own constructor; its `_estimate` (an SE3Quat: q.x q.y q.z q.w | t) is located
        by looking for the identity quaternion that setToOriginImpl writes.  Returns the estimate as (t, q)."""
        this = _aligned(512)                       # 4 KB: room for the whole BaseVertex<6, SE3Quat>
        self.f["vtx_ctor"](this.ctypes.data)
        self.f["vtx_origin"](this.ctypes.data)
        off = None
        for i in range(0, 500, 2):                 # SE3Quat is at least 16-byte aligned
            if this[i] == 0 and this[i + 1] == 0 and this[i + 2] == 0 and this[i + 3] == 1.0 and not this[i + 4:i + 7].any():
                u = _aligned(6)
                u[:] = [0.1, -0.2, 0.3, 1.0, 2.0, 3.0]
                probe = this.copy()                # the copy keeps the vptr etc.; only used to confirm the slot moves
                before = this[i:i + 7].copy()
                self.f["vtx_oplus"](this.ctypes.data, u.ctypes.data)
                if not np.array_equal(this[i:i + 7], before):
                    off = i
                    break
                del probe
        assert off is not None, "VertexSE3Expmap::_estimate not found"
        self.f["vtx_origin"](this.ctypes.data)
        for upd in updates:
            u = _aligned(6)
            u[:] = upd
            self.f["vtx_oplus"](this.ctypes.data, u.ctypes.data)
        e = this[off:off + 7]
        return np.array([e[4], e[5], e[6], e[0], e[1], e[2], e[3]])

    def project2d(self, xyz):
        out = _aligned(2)
        v = _aligned(4)
        v[:3] = xyz
        self.f["project2d"](out.ctypes.data, v.ctypes.data)
        return out.copy()

    def cam_project(self, xyz, cam5, stereo):
        this = _aligned(THIS_DOUBLES)
        off = self._stereo_off if stereo else self._mono_off
        for k, v in zip(("fx", "fy", "cx", "cy"), cam5[:4]):
            this[off[k]] = v
        return self._cam_raw(this, xyz, float(cam5[4]), stereo)

    def huber(self, delta, e2):
        this = _aligned(16)                          # vptr (unused: direct calls) + the kernel's scalar members
        self.f["set_delta"](this.ctypes.data, float(delta))  # RobustKernelHuber::setDelta writes _delta and dsqr itself
        rho = _aligned(4)
        self.f["huber"](this.ctypes.data, float(e2), rho.ctypes.data)
        return rho[:3].copy()


def make_golden(path: str, n: int = 300, seed: int = 0):
    """Inputs + outputs of the reference binary, committed as a fixture."""
    g = LibG2O()
    rng = np.random.default_rng(seed)
    upd = rng.normal(0, 1, (n, 6)) * np.array([0.3, 0.3, 0.3, 1, 1, 1])
    upd[:20, :3] *= 1e-6   # the small-angle branch (theta < 1e-5) of SE3Quat::exp
    upd[0] = 0.0
    xyz = np.stack([rng.uniform(-30, 30, n), rng.uniform(-10, 10, n), rng.uniform(0.6, 80, n)], 1)
    cam = np.array([718.856, 718.856, 607.1928, 185.2157, 386.1448])
    cam = cam.astype(np.float32).astype(np.float64)   # the reference stores them as float (KeyFrame.h:394)
    e2 = np.concatenate([rng.uniform(0, 5.9, n // 2), rng.uniform(6, 5000, n - n // 2)])
    delta = float(np.float32(np.sqrt(5.991)))
    out = dict(upd=upd, xyz=xyz, cam=cam, e2=e2, delta=np.array(delta),
               exp=np.stack([g.se3_exp(u) for u in upd]),
               project2d=np.stack([g.project2d(x) for x in xyz]),
               cam_mono=np.stack([g.cam_project(x, cam, False) for x in xyz]),
               cam_stereo=np.stack([g.cam_project(x, cam, True) for x in xyz]),
               huber=np.stack([g.huber(delta, e) for e in e2]),
               # chains of three oplus updates from the origin: composition + re-normalisation of the reference
               oplus=np.stack([g.oplus_chain(upd[3 * k:3 * k + 3]) for k in range(n // 3)]))
    np.savez(path, **out)
    return out


if __name__ == "__main__":
    # Run as a script (a clean process): the prebuilt binary must not share a process with torch & co, whose bundled
    # libstdc++/OpenMP runtimes it was not linked against.
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "tests", "golden", "libg2o_vectors.npz")
    make_golden(p)
    print("wrote", p)
    sys.stdout.flush()
    os._exit(0)  # skip the binary's static destructors (its factory singletons crash at interpreter shutdown)
