"""opengjk-gpu_b200 -- Python mirror of the reference's GPU API over the C-ABI library.

The product is ``lib/libopengjk_b200.so`` (hand-written sm_100a kernels behind ``include/opengjk_b200.h``).
This module only *binds* it with ctypes so that tests and bench.py can drive the same entry points a C++
user of the reference calls (``compute_minimum_distance``, ``computeCollisionInformation``,
``compute_gjk_epa``, the ``*_device`` and ``*_indexed`` families; reference GJK/gpu/openGJK.h:91-505).

There is no CPU fallback: if the library is missing or no CUDA device is usable, calls raise.
The directory name contains a hyphen, so import it with ``load_package()`` from ``_pkgpath.py`` at the repo
root (tests/conftest.py and bench.py do) -- it registers the package as ``opengjk_gpu_b200``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopengjk_b200.so")


class OgjkError(RuntimeError):
    pass


def polytope_dtype(dtype) -> np.dtype:
    """numpy mirror of gkPolytope (reference GJK/common.h:68-78)."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return np.dtype({"names": ["numpoints", "s", "s_idx", "coord"],
                         "formats": ["<i4", ("<f4", (3,)), "<i4", "<u8"],
                         "offsets": [0, 4, 16, 24], "itemsize": 32})
    if dtype == np.float64:
        return np.dtype({"names": ["numpoints", "s", "s_idx", "coord"],
                         "formats": ["<i4", ("<f8", (3,)), "<i4", "<u8"],
                         "offsets": [0, 8, 32, 40], "itemsize": 48})
    raise TypeError(dtype)


def simplex_dtype(dtype) -> np.dtype:
    """numpy mirror of gkSimplex (reference GJK/common.h:84-89)."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return np.dtype({"names": ["nvrtx", "vrtx", "vrtx_idx", "witnesses"],
                         "formats": ["<i4", ("<f4", (4, 3)), ("<i4", (4, 2)), ("<f4", (2, 3))],
                         "offsets": [0, 4, 52, 84], "itemsize": 108})
    if dtype == np.float64:
        return np.dtype({"names": ["nvrtx", "vrtx", "vrtx_idx", "witnesses"],
                         "formats": ["<i4", ("<f8", (4, 3)), ("<i4", (4, 2)), ("<f8", (2, 3))],
                         "offsets": [0, 8, 104, 136], "itemsize": 184})
    raise TypeError(dtype)


PAIR_DTYPE = np.dtype([("idx1", "<i4"), ("idx2", "<i4")])  # gkCollisionPair, GJK/gpu/openGJK.h:309-312

_lib = None


def load_library() -> ctypes.CDLL:
    """Loads the CUDA library; raises loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OgjkError(f"{LIB_PATH} not found: run `python opengjk-gpu_b200/build.py` (needs nvcc)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.ogjk_last_error.restype = ctypes.c_char_p
        lib.ogjk_version.restype = ctypes.c_char_p
        lib.ogjk_launch_count.restype = ctypes.c_longlong
        lib.ogjk_last_kernel.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _ptr(x) -> ctypes.c_void_p:
    """numpy array / torch tensor / int address / None -> void*"""
    if x is None:
        return ctypes.c_void_p(0)
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, np.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError(type(x))


def make_polytopes(coords, dtype=None):
    """Builds a host gkPolytope array over `coords`.

    coords: ndarray [n, V, 3] (uniform) or a sequence of [Vi, 3] arrays (ragged).
    Returns (descriptors, keepalive): descriptors' `coord` fields point into keepalive buffers, which the
    caller owns, exactly as in the reference (GJK/common.h:75-77)."""
    if isinstance(coords, np.ndarray) and coords.ndim == 3:
        dtype = np.dtype(dtype or coords.dtype)
        buf = np.ascontiguousarray(coords, dtype=dtype)
        n, nv = buf.shape[0], buf.shape[1]
        desc = np.zeros(n, dtype=polytope_dtype(dtype))
        desc["numpoints"] = nv
        desc["coord"] = buf.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(nv * 3 * dtype.itemsize)
        return desc, buf
    arrs = [np.ascontiguousarray(c, dtype=dtype or np.asarray(c).dtype) for c in coords]
    dtype = np.dtype(dtype or arrs[0].dtype)
    desc = np.zeros(len(arrs), dtype=polytope_dtype(dtype))
    for i, a in enumerate(arrs):
        desc["numpoints"][i] = a.shape[0]
        desc["coord"][i] = a.ctypes.data
    return desc, arrs


class Engine:
    """One precision of the C ABI (``ogjk_f32_*`` or ``ogjk_f64_*``)."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.tag = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[self.dtype]
        self.lib = load_library()
        self.sdtype = simplex_dtype(self.dtype)
        self.pdtype = polytope_dtype(self.dtype)

    # -- plumbing ------------------------------------------------------------------------------------------
    def _call(self, name, *args):
        rc = getattr(self.lib, f"ogjk_{self.tag}_{name}")(*args)
        if rc != 0:
            raise OgjkError(f"ogjk_{self.tag}_{name} failed ({rc}): {self.lib.ogjk_last_error().decode()}")

    def set_device(self, device: int):
        if self.lib.ogjk_set_device(ctypes.c_int(device)) != 0:
            raise OgjkError(self.lib.ogjk_last_error().decode())

    def set_stream(self, stream_ptr: int):
        self.lib.ogjk_set_stream(ctypes.c_void_p(stream_ptr))

    def set_sync(self, enabled: bool):
        self.lib.ogjk_set_sync(ctypes.c_int(int(enabled)))

    def set_timing(self, enabled: bool):
        """per-stage CUDA-event timing of gjk_epa_uniform_device calls (see ogjk_stage_times)"""
        self.lib.ogjk_set_timing(ctypes.c_int(int(enabled)))

    def stage_times(self):
        """-> (gjk_ms, epa_ms, calls) summed over the timed calls since the last query"""
        g, e, c = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
        rc = self.lib.ogjk_stage_times(ctypes.byref(g), ctypes.byref(e), ctypes.byref(c))
        if rc != 0:
            raise OgjkError(f"ogjk_stage_times failed ({rc}): {self.lib.ogjk_last_error().decode()}")
        return g.value, e.value, c.value

    def broadphase_pairs_device(self, n, d_pos_radius, cell_size, boundary, grid_size, d_pairs, max_pairs) -> int:
        """uniform-grid broad phase on device memory; returns the number of pairs found (may exceed max_pairs)"""
        total = ctypes.c_longlong(0)
        rc = self.lib.ogjk_broadphase_pairs_device(ctypes.c_int(n), _ptr(d_pos_radius), ctypes.c_float(cell_size),
                                                   ctypes.c_float(boundary), ctypes.c_int(grid_size), _ptr(d_pairs),
                                                   ctypes.c_int(max_pairs), ctypes.byref(total))
        if rc != 0:
            raise OgjkError(f"ogjk_broadphase_pairs_device failed ({rc}): {self.lib.ogjk_last_error().decode()}")
        return int(total.value)

    def transform_to_world_device(self, n_sub, d_positions, d_quats, d_scales, d_local, d_world, d_offsets=None,
                                  d_counts=None, d_sub_body=None, uniform_count=0):
        rc = self.lib.ogjk_transform_to_world_device(ctypes.c_int(n_sub), _ptr(d_positions), _ptr(d_quats), _ptr(d_scales),
                                                     _ptr(d_local), _ptr(d_world), _ptr(d_offsets), _ptr(d_counts),
                                                     _ptr(d_sub_body), ctypes.c_int(uniform_count))
        if rc != 0:
            raise OgjkError(f"ogjk_transform_to_world_device failed ({rc}): {self.lib.ogjk_last_error().decode()}")

    def init_polytopes_device(self, d_polytopes, d_world, n_sub, d_offsets=None, d_counts=None, uniform_count=0):
        self._call("init_polytopes_device", _ptr(d_polytopes), _ptr(d_world), _ptr(d_offsets), _ptr(d_counts),
                   ctypes.c_int(uniform_count), ctypes.c_int(n_sub))

    def contact_response_device(self, num_pairs, d_pairs, d_distances, d_simplices, d_normals, num_objects, d_positions,
                                d_vel_ping, d_vel_pong, d_ang_ping, d_ang_pong, d_quats, d_inv_inertia,
                                d_sub_mesh_body=None, epsilon=0.0, restitution=0.7, restitution_threshold=2.0,
                                baumgarte_beta=0.2):
        """contact response over GJK/EPA outputs (reference collision_response_kernel); corrects d_positions in place
        and writes ping + impulses to the pong buffers; defaults = reference visualization/sim_config.h:60-64"""
        prm = (ctypes.c_float * 4)(epsilon, restitution, restitution_threshold, baumgarte_beta)
        self._call("contact_response_device", ctypes.c_int(num_pairs), _ptr(d_pairs), _ptr(d_distances), _ptr(d_simplices),
                   _ptr(d_normals), _ptr(d_sub_mesh_body), ctypes.c_int(num_objects), _ptr(d_positions), _ptr(d_vel_ping),
                   _ptr(d_vel_pong), _ptr(d_ang_ping), _ptr(d_ang_pong), _ptr(d_quats), _ptr(d_inv_inertia), prm)

    def set_devices(self, devices=None):
        """devices the host-pointer calls fan out over (None / [] / one entry: single device); see ogjk_set_devices"""
        devices = list(devices or [])
        arr = (ctypes.c_int * max(len(devices), 1))(*devices)
        if self.lib.ogjk_set_devices(ctypes.c_int(len(devices)), arr) != 0:
            raise OgjkError(self.lib.ogjk_last_error().decode())

    def release_cached_buffers(self):
        if self.lib.ogjk_release_cached_buffers() != 0:
            raise OgjkError(self.lib.ogjk_last_error().decode())

    def release_pool(self, d_polytopes):
        self.lib.ogjk_release_pool(_ptr(d_polytopes))

    def last_kernel(self) -> str:
        return self.lib.ogjk_last_kernel().decode()

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.ogjk_launch_count(ctypes.c_int(int(reset))))

    # -- high level (host arrays), reference GJK/gpu/openGJK.h:91-141 ----------------------------------------
    def compute_minimum_distance(self, bd1, bd2, simplices=None, distances=None):
        n = len(bd1)
        simplices = np.zeros(n, self.sdtype) if simplices is None else simplices
        distances = np.zeros(n, self.dtype) if distances is None else distances
        self._call("compute_minimum_distance", ctypes.c_int(n), _ptr(bd1), _ptr(bd2), _ptr(simplices), _ptr(distances))
        return simplices, distances

    def compute_collision_information(self, bd1, bd2, simplices, distances, contact_normals=None):
        n = len(bd1)
        contact_normals = np.zeros((n, 3), self.dtype) if contact_normals is None else contact_normals
        self._call("compute_collision_information", ctypes.c_int(n), _ptr(bd1), _ptr(bd2), _ptr(simplices),
                   _ptr(distances), _ptr(contact_normals))
        return simplices, distances, contact_normals

    def compute_gjk_epa(self, bd1, bd2, simplices=None, distances=None, contact_normals=None):
        n = len(bd1)
        simplices = np.zeros(n, self.sdtype) if simplices is None else simplices
        distances = np.zeros(n, self.dtype) if distances is None else distances
        contact_normals = np.zeros((n, 3), self.dtype) if contact_normals is None else contact_normals
        self._call("compute_gjk_epa", ctypes.c_int(n), _ptr(bd1), _ptr(bd2), _ptr(simplices), _ptr(distances),
                   _ptr(contact_normals))
        return simplices, distances, contact_normals

    def compute_collision_information_witness(self, bd1, bd2):
        """README spelling (reference README.md:38-47): GJK+EPA with separate witness arrays."""
        n = len(bd1)
        simplices = np.zeros(n, self.sdtype)
        distances = np.zeros(n, self.dtype)
        w1 = np.zeros((n, 3), self.dtype)
        w2 = np.zeros((n, 3), self.dtype)
        nrm = np.zeros((n, 3), self.dtype)
        self._call("compute_collision_information_witness", ctypes.c_int(n), _ptr(bd1), _ptr(bd2), _ptr(simplices),
                   _ptr(distances), _ptr(w1), _ptr(w2), _ptr(nrm))
        return simplices, distances, w1, w2, nrm

    # -- indexed (host arrays), reference GJK/gpu/openGJK.h:398-505 ----------------------------------------
    def compute_minimum_distance_indexed(self, polytopes, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pairs.shape[0]
        simplices = np.zeros(n, self.sdtype)
        distances = np.zeros(n, self.dtype)
        self._call("compute_minimum_distance_indexed", ctypes.c_int(len(polytopes)), ctypes.c_int(n), _ptr(polytopes),
                   _ptr(pairs), _ptr(simplices), _ptr(distances))
        return simplices, distances

    def compute_epa_indexed(self, polytopes, pairs, simplices, distances):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pairs.shape[0]
        nrm = np.zeros((n, 3), self.dtype)
        self._call("compute_epa_indexed", ctypes.c_int(len(polytopes)), ctypes.c_int(n), _ptr(polytopes), _ptr(pairs),
                   _ptr(simplices), _ptr(distances), _ptr(nrm))
        return simplices, distances, nrm

    def compute_gjk_epa_indexed(self, polytopes, pairs, out=None):
        """`out` = (simplices, distances, normals) arrays to write into (e.g. views of pinned memory); allocated when None"""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pairs.shape[0]
        if out is not None:
            simplices, distances, nrm = out
            assert simplices.shape == (n,) and simplices.dtype == self.sdtype and distances.shape == (n,)
            assert nrm.shape == (n, 3) and distances.dtype == self.dtype and nrm.dtype == self.dtype
        else:
            simplices = np.zeros(n, self.sdtype)
            distances = np.zeros(n, self.dtype)
            nrm = np.zeros((n, 3), self.dtype)
        self._call("compute_gjk_epa_indexed", ctypes.c_int(len(polytopes)), ctypes.c_int(n), _ptr(polytopes),
                   _ptr(pairs), _ptr(simplices), _ptr(distances), _ptr(nrm))
        return simplices, distances, nrm

    # -- mid level (explicit device memory), reference GJK/gpu/openGJK.h:155-299 ------------------------------
    def allocate_and_copy_device_arrays(self, bd1, bd2):
        n = len(bd1)
        out = [ctypes.c_void_p() for _ in range(6)]
        self._call("allocate_and_copy_device_arrays", ctypes.c_int(n), _ptr(bd1), _ptr(bd2), *[ctypes.byref(o) for o in out])
        return tuple(o.value for o in out)  # d_bd1, d_bd2, d_coord1, d_coord2, d_simplices, d_distances

    def compute_minimum_distance_device(self, n, d_bd1, d_bd2, d_simplices, d_distances):
        self._call("compute_minimum_distance_device", ctypes.c_int(n), _ptr(d_bd1), _ptr(d_bd2), _ptr(d_simplices),
                   _ptr(d_distances))

    def compute_epa_device(self, n, d_bd1, d_bd2, d_simplices, d_distances, d_normals):
        self._call("compute_epa_device", ctypes.c_int(n), _ptr(d_bd1), _ptr(d_bd2), _ptr(d_simplices), _ptr(d_distances),
                   _ptr(d_normals))

    def copy_results_from_device(self, n, d_simplices, d_distances):
        simplices = np.zeros(n, self.sdtype)
        distances = np.zeros(n, self.dtype)
        self._call("copy_results_from_device", ctypes.c_int(n), _ptr(d_simplices), _ptr(d_distances), _ptr(simplices),
                   _ptr(distances))
        return simplices, distances

    def free_device_arrays(self, d_bd1, d_bd2, d_coord1, d_coord2, d_simplices, d_distances):
        self._call("free_device_arrays", _ptr(d_bd1), _ptr(d_bd2), _ptr(d_coord1), _ptr(d_coord2), _ptr(d_simplices),
                   _ptr(d_distances))

    def allocate_epa_device_arrays(self, n):
        out = [ctypes.c_void_p() for _ in range(3)]
        self._call("allocate_epa_device_arrays", ctypes.c_int(n), *[ctypes.byref(o) for o in out])
        return tuple(o.value for o in out)

    def copy_epa_results_from_device(self, n, d_w1, d_w2, d_nrm):
        w1 = np.zeros((n, 3), self.dtype)
        w2 = np.zeros((n, 3), self.dtype)
        nrm = np.zeros((n, 3), self.dtype)
        self._call("copy_epa_results_from_device", ctypes.c_int(n), _ptr(d_w1), _ptr(d_w2), _ptr(d_nrm), _ptr(w1), _ptr(w2),
                   _ptr(nrm))
        return w1, w2, nrm

    def free_epa_device_arrays(self, d_w1, d_w2, d_nrm):
        self._call("free_epa_device_arrays", _ptr(d_w1), _ptr(d_w2), _ptr(d_nrm))

    # -- indexed device level, reference GJK/gpu/openGJK.h:330-457 ---------------------------------------------
    def allocate_indexed_device(self, polytopes, max_pairs, want_normals=True):
        out = [ctypes.c_void_p() for _ in range(6)]
        refs = [ctypes.byref(o) for o in out]
        if not want_normals:
            refs[5] = ctypes.c_void_p(0)
        self._call("allocate_indexed_device", ctypes.c_int(len(polytopes)), ctypes.c_int(max_pairs), _ptr(polytopes), *refs)
        return tuple(o.value for o in out)  # d_polytopes, d_coords, d_pairs, d_simplices, d_distances, d_normals

    def upload_pairs_device(self, pairs, d_pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self._call("upload_pairs_device", ctypes.c_int(pairs.shape[0]), _ptr(pairs), _ptr(d_pairs))

    def compute_minimum_distance_indexed_device(self, n, d_polytopes, d_pairs, d_simplices, d_distances):
        self._call("compute_minimum_distance_indexed_device", ctypes.c_int(n), _ptr(d_polytopes), _ptr(d_pairs),
                   _ptr(d_simplices), _ptr(d_distances))

    def compute_epa_indexed_device(self, n, d_polytopes, d_pairs, d_simplices, d_distances, d_normals):
        self._call("compute_epa_indexed_device", ctypes.c_int(n), _ptr(d_polytopes), _ptr(d_pairs), _ptr(d_simplices),
                   _ptr(d_distances), _ptr(d_normals))

    def gjk_epa_indexed_device(self, n, d_polytopes, d_pairs, d_simplices, d_distances, d_normals):
        """GJK then EPA over an indexed batch in one call (fused EPA gate; feeds stage_times)."""
        self._call("gjk_epa_indexed_device", ctypes.c_int(n), _ptr(d_polytopes), _ptr(d_pairs), _ptr(d_simplices),
                   _ptr(d_distances), _ptr(d_normals))

    def free_indexed_device(self, d_polytopes, d_coords, d_pairs, d_simplices, d_distances, d_normals):
        self._call("free_indexed_device", _ptr(d_polytopes), _ptr(d_coords), _ptr(d_pairs), _ptr(d_simplices),
                   _ptr(d_distances), _ptr(d_normals))

    # -- flat uniform batches, device resident ---------------------------------------------------------------
    def gjk_uniform_device(self, n, nv1, d_coord1, nv2, d_coord2, d_simplices, d_distances):
        self._call("gjk_uniform_device", ctypes.c_int(n), ctypes.c_int(nv1), _ptr(d_coord1), ctypes.c_int(nv2),
                   _ptr(d_coord2), _ptr(d_simplices), _ptr(d_distances))

    def gjk_epa_uniform_device(self, n, nv1, d_coord1, nv2, d_coord2, d_simplices, d_distances, d_normals):
        """GJK then EPA in one call (lets the library fuse the EPA gate into the GJK kernel)."""
        self._call("gjk_epa_uniform_device", ctypes.c_int(n), ctypes.c_int(nv1), _ptr(d_coord1), ctypes.c_int(nv2),
                   _ptr(d_coord2), _ptr(d_simplices), _ptr(d_distances), _ptr(d_normals))

    def epa_uniform_device(self, n, nv1, d_coord1, nv2, d_coord2, d_simplices, d_distances, d_normals):
        self._call("epa_uniform_device", ctypes.c_int(n), ctypes.c_int(nv1), _ptr(d_coord1), ctypes.c_int(nv2),
                   _ptr(d_coord2), _ptr(d_simplices), _ptr(d_distances), _ptr(d_normals))


from . import sharding, workloads  # noqa: E402,F401
