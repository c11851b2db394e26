"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the two CPU checkers:

* ``kind="port"``  -> oracle/lib/libogjk_oracle_{f32,f64}.so : the C restatement (ogjk_oracle.c)
* ``kind="ref"``   -> oracle/_ref/libogjk_ref_{f32,f64}.so   : the reference's own CPU sources
  (GJK/cpu/openGJK.c, GJK/cpu/EPA.c) compiled unmodified by oracle/build_ref.sh

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  The product (opengjk-gpu_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def simplex_dtype(dtype) -> np.dtype:
    """numpy mirror of gkSimplex (reference GJK/common.h:84-89)."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return np.dtype(
            {
                "names": ["nvrtx", "vrtx", "vrtx_idx", "witnesses"],
                "formats": ["<i4", ("<f4", (4, 3)), ("<i4", (4, 2)), ("<f4", (2, 3))],
                "offsets": [0, 4, 52, 84],
                "itemsize": 108,
            }
        )
    if dtype == np.float64:
        return np.dtype(
            {
                "names": ["nvrtx", "vrtx", "vrtx_idx", "witnesses"],
                "formats": ["<i4", ("<f8", (4, 3)), ("<i4", (4, 2)), ("<f8", (2, 3))],
                "offsets": [0, 8, 104, 136],
                "itemsize": 184,
            }
        )
    raise TypeError(dtype)


def _lib_path(kind: str, dtype: np.dtype) -> str:
    tag = "f32" if dtype == np.float32 else "f64"
    if kind == "port":
        return os.path.join(_HERE, "lib", f"libogjk_oracle_{tag}.so")
    if kind == "ref":
        return os.path.join(_HERE, "_ref", f"libogjk_ref_{tag}.so")
    raise ValueError(kind)


def available(kind: str, dtype=np.float32) -> bool:
    return os.path.exists(_lib_path(kind, np.dtype(dtype)))


def _flat(c, off, dtype):
    """-> (contiguous coords [total,3], offsets int64 or None, uniform nv, n)"""
    c = np.ascontiguousarray(c, dtype=dtype)
    if off is None:
        assert c.ndim == 3 and c.shape[2] == 3, "uniform input must be [n, V, 3]"
        return c, None, int(c.shape[1]), int(c.shape[0])
    off = np.ascontiguousarray(off, dtype=np.int64)
    return c, off, 0, int(off.shape[0] - 1)


class Oracle:
    def __init__(self, kind: str = "port", dtype=np.float32):
        self.kind = kind
        self.dtype = np.dtype(dtype)
        self.sdtype = simplex_dtype(self.dtype)
        path = _lib_path(kind, self.dtype)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle`")
        self.lib = ctypes.CDLL(path)
        self.prefix = "ogjk_oracle_" if kind == "port" else "ogjk_ref_"
        f = lambda name: getattr(self.lib, self.prefix + name)
        assert f("sizeof_real")() == self.dtype.itemsize
        assert f("sizeof_simplex")() == self.sdtype.itemsize
        self._gjk = f("gjk_batch")
        self._epa = f("epa_batch")
        self._idx = f("gjk_epa_indexed")
        for fn in (self._gjk, self._epa, self._idx):
            fn.restype = None

    @staticmethod
    def _p(a):
        return ctypes.c_void_p(0 if a is None else a.ctypes.data)

    def gjk(self, c1, c2, off1=None, off2=None, nthreads: int = 1, want_iters: bool = False):
        c1, off1, nv1, n = _flat(c1, off1, self.dtype)
        c2, off2, nv2, n2 = _flat(c2, off2, self.dtype)
        assert n == n2
        simp = np.zeros(n, dtype=self.sdtype)
        dist = np.zeros(n, dtype=self.dtype)
        iters = np.zeros(n, dtype=np.int32) if (want_iters and self.kind == "port") else None
        args = [ctypes.c_long(n), self._p(c1), self._p(off1), ctypes.c_int(nv1), self._p(c2), self._p(off2),
                ctypes.c_int(nv2), self._p(simp), self._p(dist)]
        if self.kind == "port":
            args.append(self._p(iters))
        args.append(ctypes.c_int(nthreads))
        self._gjk(*args)
        return (simp, dist, iters) if want_iters else (simp, dist)

    def epa(self, c1, c2, simplices, distances, off1=None, off2=None, nthreads: int = 1,
            want_iters: bool = False, normals_init: float = 0.0):
        """Returns (simplices, distances, normals) -- copies; the inputs are not modified."""
        c1, off1, nv1, n = _flat(c1, off1, self.dtype)
        c2, off2, nv2, _ = _flat(c2, off2, self.dtype)
        simp = np.array(simplices, dtype=self.sdtype, copy=True)
        dist = np.array(distances, dtype=self.dtype, copy=True)
        nrm = np.full((n, 3), normals_init, dtype=self.dtype)
        iters = np.zeros(n, dtype=np.int32) if (want_iters and self.kind == "port") else None
        args = [ctypes.c_long(n), self._p(c1), self._p(off1), ctypes.c_int(nv1), self._p(c2), self._p(off2),
                ctypes.c_int(nv2), self._p(simp), self._p(dist), self._p(nrm)]
        if self.kind == "port":
            args.append(self._p(iters))
        args.append(ctypes.c_int(nthreads))
        self._epa(*args)
        return (simp, dist, nrm, iters) if want_iters else (simp, dist, nrm)

    def gjk_epa_indexed(self, pool, pairs, off=None, do_gjk=True, do_epa=True, simplices=None,
                        distances=None, nthreads: int = 1):
        pool, off, nv, _ = _flat(pool, off, self.dtype)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        n = int(pairs.shape[0])
        simp = np.zeros(n, dtype=self.sdtype) if simplices is None else np.array(simplices, dtype=self.sdtype, copy=True)
        dist = np.zeros(n, dtype=self.dtype) if distances is None else np.array(distances, dtype=self.dtype, copy=True)
        nrm = np.zeros((n, 3), dtype=self.dtype)
        self._idx(ctypes.c_long(n), self._p(pool), self._p(off), ctypes.c_int(nv), self._p(pairs), self._p(simp),
                  self._p(dist), self._p(nrm), ctypes.c_int(int(do_gjk)), ctypes.c_int(int(do_epa)),
                  ctypes.c_int(nthreads))
        return simp, dist, nrm


class RefGpu:
    """The reference's own GPU library (GJK/gpu/openGJK.cu compiled unmodified for sm_100 by
    oracle/build_ref_gpu.sh): second baseline ("kernel to beat") and second oracle.  fp32 only.  Needs a GPU."""

    PATH = os.path.join(_HERE, "_ref_gpu", "libogjk_refgpu_f32.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    def __init__(self):
        if not self.available():
            raise FileNotFoundError(f"{self.PATH} missing: run oracle/build_ref_gpu.sh where /root/reference exists")
        self.lib = ctypes.CDLL(self.PATH)
        self.dtype = np.dtype(np.float32)
        self.sdtype = simplex_dtype(self.dtype)
        assert self.lib.ogjk_refgpu_sizeof_real() == 4
        assert self.lib.ogjk_refgpu_sizeof_simplex() == self.sdtype.itemsize

    def gjk_epa(self, c1, c2, do_epa=True, reps=1):
        """-> (simplices, distances, normals, gjk_ms, epa_ms); timing = cudaEvent pairs around the reference's
        compute_minimum_distance_device / compute_epa_device, mean over `reps`."""
        c1 = np.ascontiguousarray(c1, dtype=np.float32)
        c2 = np.ascontiguousarray(c2, dtype=np.float32)
        n = c1.shape[0]
        simp = np.zeros(n, dtype=self.sdtype)
        dist = np.zeros(n, dtype=np.float32)
        nrm = np.zeros((n, 3), dtype=np.float32)
        ms = (ctypes.c_float * 2)(0, 0)
        rc = self.lib.ogjk_refgpu_gjk_epa(ctypes.c_long(n), Oracle._p(c1), ctypes.c_int(c1.shape[1]), Oracle._p(c2),
                                          ctypes.c_int(c2.shape[1]), Oracle._p(simp), Oracle._p(dist), Oracle._p(nrm),
                                          ctypes.c_int(int(do_epa)), ctypes.c_int(reps), ms)
        if rc != 0:
            raise RuntimeError(f"reference GPU library failed: cudaError {rc}")
        return simp, dist, nrm, float(ms[0]), float(ms[1])

    def gjk_epa_indexed(self, pool, pairs, do_epa=True, reps=1):
        pool = np.ascontiguousarray(pool, dtype=np.float32)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pairs.shape[0]
        simp = np.zeros(n, dtype=self.sdtype)
        dist = np.zeros(n, dtype=np.float32)
        nrm = np.zeros((n, 3), dtype=np.float32)
        ms = (ctypes.c_float * 2)(0, 0)
        rc = self.lib.ogjk_refgpu_gjk_epa_indexed(ctypes.c_int(pool.shape[0]), ctypes.c_int(pool.shape[1]), Oracle._p(pool),
                                                  ctypes.c_long(n), Oracle._p(pairs), Oracle._p(simp), Oracle._p(dist),
                                                  Oracle._p(nrm), ctypes.c_int(int(do_epa)), ctypes.c_int(reps), ms)
        if rc != 0:
            raise RuntimeError(f"reference GPU library failed: cudaError {rc}")
        return simp, dist, nrm, float(ms[0]), float(ms[1])


class RefVis:
    """The reference visualiser's own kernels either side of the hot path (broad phase, world transform, contact
    response, descriptor upkeep), compiled unmodified by oracle/build_ref_vis.sh.  Needs a GPU."""

    PATH = os.path.join(_HERE, "_ref_gpu", "libogjk_refvis_f32.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    def __init__(self):
        if not self.available():
            raise FileNotFoundError(f"{self.PATH} missing: run oracle/build_ref_vis.sh where /root/reference exists")
        self.lib = ctypes.CDLL(self.PATH)
        self.lib.ogjk_refvis_broadphase.restype = ctypes.c_long

    def broadphase(self, pos_radius, cell_size, boundary, grid_size, max_pairs):
        """-> int32 [m, 2] in the reference's emission order (grouped by idx1; order inside a group follows the
        atomic slot order of insert_objects_kernel, i.e. is not deterministic)"""
        p = np.ascontiguousarray(pos_radius, np.float32)
        out = np.zeros((max_pairs, 2), np.int32)
        total = self.lib.ogjk_refvis_broadphase(ctypes.c_int(p.shape[0]), Oracle._p(p), ctypes.c_float(cell_size),
                                                ctypes.c_float(boundary), ctypes.c_int(grid_size), Oracle._p(out),
                                                ctypes.c_int(max_pairs))
        if total < 0:
            raise RuntimeError("reference broad phase failed" if total == -1 else
                               "reference broad phase dropped objects (more than 512 per cell)")
        return out[: min(total, max_pairs)], int(total)

    def transform(self, positions, quats, scales, verts_local, offsets, counts, sub_body):
        pos = np.ascontiguousarray(positions, np.float32)
        q = np.ascontiguousarray(quats, np.float32)
        sc = np.ascontiguousarray(scales, np.float32)
        lv = np.ascontiguousarray(verts_local, np.float32).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, np.int32)
        cnt = np.ascontiguousarray(counts, np.int32)
        sb = np.ascontiguousarray(sub_body, np.int32)
        out = np.zeros_like(lv)
        rc = self.lib.ogjk_refvis_transform(ctypes.c_int(len(off)), ctypes.c_int(pos.shape[0]), ctypes.c_int(lv.shape[0]),
                                            Oracle._p(pos), Oracle._p(q), Oracle._p(sc), Oracle._p(lv), Oracle._p(out),
                                            Oracle._p(off), Oracle._p(cnt), Oracle._p(sb))
        if rc:
            raise RuntimeError(f"reference transform failed: cudaError {rc}")
        return out

    def response(self, pairs, distances, simplices, normals, sub_mesh_body, positions, vel_ping, ang_ping, quats,
                 inv_inertia, epsilon):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        dist = np.ascontiguousarray(distances, np.float32)
        simp = np.ascontiguousarray(simplices)
        nrm = np.ascontiguousarray(normals, np.float32)
        smb = np.ascontiguousarray(sub_mesh_body, np.int32)
        pos = np.array(positions, np.float32, copy=True)
        vp = np.ascontiguousarray(vel_ping, np.float32)
        ap = np.ascontiguousarray(ang_ping, np.float32)
        q = np.ascontiguousarray(quats, np.float32)
        ii = np.ascontiguousarray(inv_inertia, np.float32)
        vq, aq = np.zeros_like(vp), np.zeros_like(ap)
        rc = self.lib.ogjk_refvis_response(ctypes.c_int(pairs.shape[0]), Oracle._p(pairs), Oracle._p(dist), Oracle._p(simp),
                                           Oracle._p(nrm), Oracle._p(smb), ctypes.c_int(len(smb)), ctypes.c_int(pos.shape[0]),
                                           Oracle._p(pos), Oracle._p(vp), Oracle._p(vq), Oracle._p(ap), Oracle._p(aq),
                                           Oracle._p(q), Oracle._p(ii), ctypes.c_float(epsilon))
        if rc:
            raise RuntimeError(f"reference response failed: cudaError {rc}")
        return pos, vq, aq

    def init_polytopes(self, offsets, counts):
        off = np.ascontiguousarray(offsets, np.int32)
        cnt = np.ascontiguousarray(counts, np.int32)
        npts = np.zeros(len(off), np.int32)
        coff = np.zeros(len(off), np.int64)
        rc = self.lib.ogjk_refvis_init_polytopes(ctypes.c_int(len(off)), Oracle._p(off), Oracle._p(cnt), Oracle._p(npts),
                                                 Oracle._p(coff))
        if rc:
            raise RuntimeError(f"reference init_polytopes failed: cudaError {rc}")
        return npts, coff
