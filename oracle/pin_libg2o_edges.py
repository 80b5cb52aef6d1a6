"""Edge-level pinning: REAL EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ objects of the reference's prebuilt libg2o.so.

Builds one pose vertex, one point vertex and one edge with the binary's own constructors, wires them, calls the
binary's computeError() and linearizeOplus(JacobianWorkspace&) (types_six_dof_expmap.h:90-95,122-127,
.cpp:103-139,188-234; base_binary_edge.hpp:40-53) and reads `_error` and both Jacobians back.  Member offsets inside the
objects (the `_vertices` vector, `_measurement`, `_error`, fx..bf, the workspace pointers) are not taken from headers --
Eigen is not available here -- but located by probing (write a marker, call, watch what moves), see pin_libg2o.py.
Run as a script in a clean interpreter; appends `edge_*` arrays to tests/golden/libg2o_vectors.npz."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pin_libg2o as P  # noqa: E402

SYM = {
    "stereo_ctor": "_ZN3g2o23EdgeStereoSE3ProjectXYZC1Ev",
    "mono_ctor": "_ZN3g2o17EdgeSE3ProjectXYZC1Ev",
    "stereo_err": "_ZN3g2o23EdgeStereoSE3ProjectXYZ12computeErrorEv",
    "mono_err": "_ZN3g2o17EdgeSE3ProjectXYZ12computeErrorEv",
    "stereo_lin": "_ZN3g2o14BaseBinaryEdgeILi3EN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEENS_17VertexSBAPointXYZENS_15VertexSE3ExpmapEE14linearizeOplusERNS_17JacobianWorkspaceE",
    "mono_lin": "_ZN3g2o14BaseBinaryEdgeILi2EN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEENS_17VertexSBAPointXYZENS_15VertexSE3ExpmapEE14linearizeOplusERNS_17JacobianWorkspaceE",
    "pt_ctor": "_ZN3g2o17VertexSBAPointXYZC1Ev",
    "pt_origin": "_ZN3g2o17VertexSBAPointXYZ15setToOriginImplEv",
    "pt_oplus": "_ZN3g2o17VertexSBAPointXYZ9oplusImplEPKd",
    "jw_ctor": "_ZN3g2o17JacobianWorkspaceC1Ev",
    "jw_size": "_ZN3g2o17JacobianWorkspace10updateSizeEPKNS_10HyperGraph4EdgeE",
    "jw_alloc": "_ZN3g2o17JacobianWorkspace8allocateEv",
}
OBJ = 1024  # doubles per object block (8 KB)


def _u64(block):
    return block.view(np.uint64)


def _vector_slots(block, elem_bytes):
    """(index, begin) of std::vector members {begin, end, cap} inside an object block whose size is elem_bytes."""
    u = _u64(block)
    out = []
    for i in range(len(u) - 2):
        b, e, c = int(u[i]), int(u[i + 1]), int(u[i + 2])
        if b > 0x10000 and e - b == elem_bytes and c >= e and c - b <= 4096:
            out.append((i, b))
    return out


class Edges:
    def __init__(self):
        self.g = P.LibG2O()
        L = self.g.L
        self.f = {k: getattr(L, v) for k, v in SYM.items()}
        for k, fn in self.f.items():
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.c_void_p] if k in ("stereo_lin", "mono_lin", "pt_oplus", "jw_size") else [C.c_void_p]
        self.f["jw_alloc"].restype = C.c_bool

    def _pose_vertex(self, upd):
        v = P._aligned(OBJ)
        self.g.f["vtx_ctor"](v.ctypes.data)
        self.g.f["vtx_origin"](v.ctypes.data)
        u = P._aligned(6)
        u[:] = upd
        self.g.f["vtx_oplus"](v.ctypes.data, u.ctypes.data)   # estimate = exp(upd) * identity
        return v

    def _point_vertex(self, X):
        v = P._aligned(OBJ)
        self.f["pt_ctor"](v.ctypes.data)
        self.f["pt_origin"](v.ctypes.data)
        u = P._aligned(4)
        u[:3] = X
        self.f["pt_oplus"](v.ctypes.data, u.ctypes.data)      # estimate = 0 + X
        return v

    def _edge(self, stereo):
        e = P._aligned(OBJ)
        self.f["stereo_ctor" if stereo else "mono_ctor"](e.ctypes.data)
        vec = _vector_slots(e, 16)                             # _vertices: two Vertex* (resize(2) in BaseBinaryEdge())
        assert len(vec) >= 1, "edge._vertices not found"
        return e, vec[0][1]

    def layout(self, stereo):
        """offsets (doubles) of _measurement[0], _error[0] and bf inside the edge, by differential probing."""
        d = 3 if stereo else 2
        upd = np.array([0.02, -0.01, 0.03, 0.1, -0.2, 0.3])
        X = np.array([1.0, -0.5, 12.0])
        cam = np.array([700.0, 710.0, 600.0, 180.0, 380.0])
        off = self.g._stereo_off if stereo else self.g._mono_off

        def run(mutate):
            vp, vx = self._pose_vertex(upd), self._point_vertex(X)
            e, vbeg = self._edge(stereo)
            ptrs = (C.c_uint64 * 2).from_address(vbeg)
            ptrs[0], ptrs[1] = vx.ctypes.data, vp.ctypes.data
            for k, val in zip(("fx", "fy", "cx", "cy"), cam[:4]):
                e[off[k]] = val
            mutate(e)
            before = e.copy()
            self.f["stereo_err" if stereo else "mono_err"](e.ctypes.data)
            return before, e.copy(), (vp, vx)

        b0, a0, _ = run(lambda e: None)
        changed = np.nonzero(b0 != a0)[0]
        assert len(changed) >= 2, changed
        err_off = int(changed[0])                              # _error: d consecutive doubles
        # measurement: the slot whose +1 shifts _error[0] by exactly +1
        meas_off = None
        zero = np.nonzero(_u64(b0) == 0)[0]                    # members the constructor left alone (Eigen does not
        for i in zero:                                         # initialise _measurement); never touch pointer slots
            if i < 2 or i > 200 or i in changed or i in off.values():
                continue
            try_b, try_a, _ = run(lambda e, i=i: e.__setitem__(i, 1.0))
            if abs((try_a[err_off] - a0[err_off]) - 1.0) < 1e-9 and abs(try_a[err_off + 1] - a0[err_off + 1]) < 1e-12:
                meas_off = i
                break
        assert meas_off is not None, "_measurement not found"
        bf_off = None
        if stereo:                                             # bf: the slot that moves only _error[2]
            for i in range(off["cy"] + 1, off["cy"] + 8):
                if _u64(b0)[i] != 0:
                    continue
                try_b, try_a, _ = run(lambda e, i=i: e.__setitem__(i, 100.0))
                if try_a[err_off + 2] != a0[err_off + 2] and try_a[err_off] == a0[err_off]:
                    bf_off = i
                    break
            assert bf_off is not None, "bf not found"
        return dict(err=err_off, meas=meas_off, bf=bf_off, d=d)

    def evaluate(self, stereo, lay, upd, X, cam, meas):
        """(error d, J_point d x 3, J_pose d x 6) from the binary for one observation."""
        d = lay["d"]
        off = self.g._stereo_off if stereo else self.g._mono_off
        vp, vx = self._pose_vertex(upd), self._point_vertex(X)
        e, vbeg = self._edge(stereo)
        ptrs = (C.c_uint64 * 2).from_address(vbeg)
        ptrs[0], ptrs[1] = vx.ctypes.data, vp.ctypes.data
        for k, val in zip(("fx", "fy", "cx", "cy"), cam[:4]):
            e[off[k]] = val
        if stereo:
            e[lay["bf"]] = cam[4]
        e[lay["meas"]:lay["meas"] + d] = meas[:d]
        self.f["stereo_err" if stereo else "mono_err"](e.ctypes.data)
        err = e[lay["err"]:lay["err"] + d].copy()
        jw = P._aligned(64)
        self.f["jw_ctor"](jw.ctypes.data)
        self.f["jw_size"](jw.ctypes.data, e.ctypes.data)
        assert self.f["jw_alloc"](jw.ctypes.data)
        self.f["stereo_lin" if stereo else "mono_lin"](e.ctypes.data, jw.ctypes.data)
        ws = _vector_slots(jw, 32)                             # _workspace: one Eigen::VectorXd {data, size} per vertex
        assert ws, "JacobianWorkspace::_workspace not found"
        p = (C.c_uint64 * 4).from_address(ws[0][1])
        Ji = np.ctypeslib.as_array((C.c_double * (d * 3)).from_address(p[0])).copy().reshape(3, d).T   # column-major
        Jj = np.ctypeslib.as_array((C.c_double * (d * 6)).from_address(p[2])).copy().reshape(6, d).T
        return err, Ji, Jj, (vp, vx, e, jw)


def make(path, n=120, seed=1):
    E = Edges()
    rng = np.random.default_rng(seed)
    cam = np.array([718.856, 718.856, 607.1928, 185.2157, 386.1448]).astype(np.float32).astype(np.float64)
    out = dict(np.load(path)) if os.path.exists(path) else {}
    keep = []
    for stereo in (False, True):
        lay = E.layout(stereo)
        tag = "stereo" if stereo else "mono"
        upd = rng.normal(0, 1, (n, 6)) * np.array([0.2, 0.2, 0.2, 2, 2, 2])
        X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-8, 8, n), rng.uniform(4, 70, n)], 1)
        meas = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n), rng.uniform(0, 1241, n)], 1)
        meas = meas.astype(np.float32).astype(np.float64)
        errs, Jis, Jjs = [], [], []
        for k in range(n):
            err, Ji, Jj, objs = E.evaluate(stereo, lay, upd[k], X[k], cam, meas[k])
            keep.append(objs)
            errs.append(err); Jis.append(Ji); Jjs.append(Jj)
        pose = np.stack([E.g.se3_exp(u) for u in upd])         # the vertex estimate the edge saw: exp(upd)
        out.update({f"edge_{tag}_pose": pose, f"edge_{tag}_X": X, f"edge_{tag}_meas": meas, f"edge_{tag}_err": np.stack(errs),
                    f"edge_{tag}_Jl": np.stack(Jis), f"edge_{tag}_Jp": np.stack(Jjs)})
    out["edge_cam"] = cam
    np.savez(path, **out)
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "tests", "golden", "libg2o_vectors.npz")
    make(p)
    print("wrote", p)
    sys.stdout.flush()
    os._exit(0)
