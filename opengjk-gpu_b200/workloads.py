"""Seeded synthetic workloads for the GJK/EPA hot path.

The shapes follow the reference's own generators so the numbers are comparable:
  * random_polytopes   -- ``generatePolytope`` (reference examples/gpu/example.cu:122-138):
    per vertex theta=U[0,2pi), phi=U[0,pi), r=1+0.5*U[0,1), point = r*(sin phi cos theta,
    sin phi sin theta, cos phi) + offset, body offsets (U[0,1)-0.5)*S per axis (:303-309).
  * cube_grid          -- ``generateCubeWithGrid`` (:145-226)
  * sphere_surface     -- ``generateSphereSurface`` (:233-256)
  * unit_sphere_hulls  -- ``gen_gjk_random_hull`` (visualization/rendering/mesh_builder.cpp:379-398)

A counter-based generator (numpy Philox) replaces the reference's wall-clock seeded rand()
(examples/main.cpp:102) so the CPU oracle and the GPU see identical bytes on every box.
"""
from __future__ import annotations

import numpy as np

_CHUNK = 1 << 15


def _rng(seed: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed), int(stream)]))


def random_polytopes(n: int, nverts: int, spread: float, seed: int, dtype=np.float32, stream: int = 0):
    """n random point clouds of `nverts` vertices, each translated by (U-0.5)*spread per axis.
    Returns [n, nverts, 3] of `dtype`."""
    out = np.empty((n, nverts, 3), dtype=dtype)
    rng = _rng(seed, stream)
    for lo in range(0, n, _CHUNK):
        m = min(_CHUNK, n - lo)
        u = rng.random((m, nverts, 3))
        off = (rng.random((m, 1, 3)) - 0.5) * spread
        theta = u[..., 0] * (2.0 * np.pi)
        phi = u[..., 1] * np.pi
        r = 1.0 + 0.5 * u[..., 2]
        sp = np.sin(phi)
        blk = np.stack([r * sp * np.cos(theta), r * sp * np.sin(theta), r * np.cos(phi)], axis=-1) + off
        out[lo:lo + m] = blk.astype(dtype)
    return out


def random_pairs(n: int, nverts: int, spread: float, seed: int = 12345, dtype=np.float32):
    """Body-1 and body-2 clouds of one batch of pairs (BASELINE configs 2, 3, 4)."""
    return (random_polytopes(n, nverts, spread, seed, dtype, stream=1),
            random_polytopes(n, nverts, spread, seed, dtype, stream=2))


def cube_grid(grid: int, half: float, offset=(0.0, 0.0, 0.0), dtype=np.float32):
    """6*grid*grid vertices on the faces of an axis-aligned cube (reference generateCubeWithGrid)."""
    t = -half + (2.0 * half * np.arange(grid)) / (grid - 1)
    a, b = np.meshgrid(t, t, indexing="ij")
    a, b = a.ravel(), b.ravel()
    s = np.full_like(a, half)
    faces = [np.stack([s, a, b], -1), np.stack([-s, a, b], -1), np.stack([a, s, b], -1),
             np.stack([a, -s, b], -1), np.stack([a, b, s], -1), np.stack([a, b, -s], -1)]
    return (np.concatenate(faces, 0) + np.asarray(offset)).astype(dtype)


def unit_cube(offset=(0.0, 0.0, 0.0), dtype=np.float32):
    """The 8 corners (+-1) in the x,y,z nested-loop order of examples/usage/EPAUsage.cpp:35-45."""
    pts = [(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]
    return (np.asarray(pts, dtype=np.float64) + np.asarray(offset)).astype(dtype)


def rotated_cube_readme(dtype=np.float32):
    """Cube rotated 45 deg about x, y, z then shifted +1 in x (examples/usage/EPAUsage.cpp:24-83),
    evaluated in `dtype` arithmetic like the reference does in gkFloat."""
    T = np.dtype(dtype).type
    angle = T(T(45.0) * T(3.14159265358979323846) / T(180.0))
    ca, sa = T(np.cos(angle)), T(np.sin(angle))
    out = []
    for x in (-1, 1):
        for y in (-1, 1):
            for z in (-1, 1):
                px, py, pz = T(x), T(y), T(z)
                ty = T(py * ca - pz * sa); tz = T(py * sa + pz * ca); py, pz = ty, tz
                tx = T(px * ca + pz * sa); tz = T(-px * sa + pz * ca); px, pz = tx, tz
                tx = T(px * ca - py * sa); ty = T(px * sa + py * ca); px, py = tx, ty
                out.append((T(px + T(1.0)), py, pz))
    return np.asarray(out, dtype=dtype)


def sphere_surface(npts: int, radius: float, offset=(0.0, 0.0, 0.0), seed: int = 7, dtype=np.float32):
    rng = _rng(seed, 3)
    u, v = rng.random(npts), rng.random(npts)
    theta = 2.0 * np.pi * u
    phi = np.arccos(2.0 * v - 1.0)
    pts = radius * np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], -1)
    return (pts + np.asarray(offset)).astype(dtype)


def unit_sphere_hulls(npoly: int, nverts: int, seed: int = 2024, dtype=np.float32):
    """Point sets on the unit sphere, one per pool entry (local coordinates)."""
    rng = _rng(seed, 4)
    p = rng.standard_normal((npoly, nverts, 3))
    p /= np.linalg.norm(p, axis=-1, keepdims=True)
    return p.astype(dtype)


def broadphase_pool(npoly: int, nverts: int, npairs: int, seed: int = 2024, dtype=np.float32,
                    scale_range=(0.3, 2.5), return_spheres: bool = False):
    """BASELINE config 5: a pool of `npoly` world-space hulls (unit-sphere hulls scaled by
    U[0.3,2.5], reference visualization/sim_config.h:41-44) and `npairs` candidate pairs whose
    bounding spheres overlap (what the reference's grid broad phase emits,
    visualization/integrate_final_gjk.cu:529-570).  Returns (pool [npoly,nverts,3], pairs [npairs,2] int32).
    """
    rng = _rng(seed, 5)
    local = unit_sphere_hulls(npoly, nverts, seed, np.float64)
    radius = scale_range[0] + (scale_range[1] - scale_range[0]) * rng.random(npoly)
    # box edge chosen so that the number of overlapping sphere pairs is ~npairs.  The centres are confined to the box,
    # so hulls near its faces have fewer neighbours than the unbounded-medium estimate assumes (at 20 000 hulls and
    # 16 M pairs the box is only ~2 interaction diameters wide); the edge is therefore solved for by bisection on the
    # exact pair count of a fixed random subsample of the hulls, scaled by (npoly / subsample)^2.
    unit = rng.random((npoly, 3))
    m = min(npoly, 3000)
    sub = _rng(seed, 6).choice(npoly, size=m, replace=False) if m < npoly else np.arange(npoly)
    rsum = radius[sub][:, None] + radius[sub][None, :]
    iu = np.triu_indices(m, 1)
    rsum2 = (rsum[iu] ** 2)
    diff = unit[sub][:, None, :] - unit[sub][None, :, :]
    d2_unit = np.einsum("ijk,ijk->ij", diff, diff)[iu]
    scale_up = (npoly * (npoly - 1.0)) / max(m * (m - 1.0), 1.0)

    def expected_pairs(edge_len):
        return float(np.count_nonzero(d2_unit * (edge_len * edge_len) < rsum2)) * scale_up

    lo_e, hi_e = 2.0 * scale_range[1], 2.0 * scale_range[1]
    while expected_pairs(hi_e) > npairs and hi_e < 1e6:
        hi_e *= 2.0
    for _ in range(40):
        mid = 0.5 * (lo_e + hi_e)
        if expected_pairs(mid) > npairs:
            lo_e = mid
        else:
            hi_e = mid
    edge = lo_e  # slightly more than npairs; the list is truncated below
    centre = unit * edge
    pool = (local * radius[:, None, None] + centre[:, None, :]).astype(dtype)
    # candidate search on the host (the generator is not the thing measured): blocked all-pairs test of
    # |ci - cj|^2 < (ri + rj)^2, i < j, in float64 -- at ~800 neighbours per hull the box is only two interaction
    # diameters wide, so a cell list would visit nearly every pair anyway
    out = []
    c2 = np.einsum("ij,ij->i", centre, centre)
    blk = max(1, (1 << 24) // max(npoly, 1))
    for lo in range(0, npoly, blk):
        hi = min(npoly, lo + blk)
        d2 = c2[lo:hi, None] + c2[None, :] - 2.0 * (centre[lo:hi] @ centre.T)
        rs = radius[lo:hi, None] + radius[None, :]
        ii, jj = np.nonzero(d2 < rs * rs)
        keep = jj > ii + lo
        out.append(np.stack([ii[keep] + lo, jj[keep]], 1))
    pairs = np.concatenate(out, 0) if out else np.zeros((0, 2), np.int64)
    if pairs.shape[0] > npairs:
        pairs = pairs[:npairs]
    if return_spheres:  # (centre, bounding radius) per hull + the box edge: the input of the device broad phase
        spheres = np.concatenate([centre, radius[:, None]], 1).astype(np.float32)
        return pool, pairs.astype(np.int32), spheres, float(edge)
    return pool, pairs.astype(np.int32)
