"""Scenes for the ray-tracing hot path: the two shader-constant scenes of the reference and the
seeded synthetic sphere scenes named by BASELINE.json's configs.

All arrays use the byte layouts of include/vkrt.h, so the same Scene feeds libvkrt_cuda and
(in the tests) the CPU oracle.
"""
import hashlib
from dataclasses import dataclass

import numpy as np

from .renderer import pack_materials

MAT_DIFFUSE, MAT_DIELECTRIC = 0, 1


@dataclass
class Scene:
    name: str
    materials: np.ndarray   # (n, 12) float32, vkrt_material
    spheres: np.ndarray     # (n, 4) float32 {cx, cy, cz, r}
    sphere_mat: np.ndarray  # (n,) uint32
    planes: np.ndarray      # (n, 4) float32 {nx, ny, nz, len}
    plane_mat: np.ndarray   # (n,) uint32
    triangles: np.ndarray   # (n, 12) float32, 3 x vec3 padded to 16 B
    tri_mat: int
    tri_mats: np.ndarray = None   # optional (n,) uint32: per-triangle material ids (meshes); None = every triangle uses tri_mat

    def digest(self):
        """sha256 over every array: the scene hash quoted next to benchmark numbers."""
        h = hashlib.sha256()
        for a in (self.materials, self.spheres, self.sphere_mat, self.planes, self.plane_mat, self.triangles):
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(str(self.tri_mat).encode())
        if self.tri_mats is not None:
            h.update(np.ascontiguousarray(self.tri_mats).tobytes())
        return h.hexdigest()[:16]


# the single triangle the host uploads: Source/GraphicsDevice.cpp:798-803 == Raytracer.comp:116
_HOST_TRIANGLE = np.array([[10, 10, 0, 0, 0, 20, 0, 0, -10, 10, 0, 0]], dtype=np.float32)

# Tracer.comp:186-194 -- (albedo, emissive, roughness, metalness, type)
_TRACER_MATS = [
    ((1.0, 1.0, 1.0), 0.0, 0.3, 0.7, MAT_DIFFUSE),       # 0 matte_white
    ((0.75, 0.25, 0.25), 0.0, 0.4, 0.0, MAT_DIFFUSE),    # 1 matte_red
    ((0.25, 0.75, 0.25), 0.0, 0.4, 0.0, MAT_DIFFUSE),    # 2 matte_green
    ((0.25, 0.25, 0.75), 0.0, 0.4, 0.0, MAT_DIFFUSE),    # 3 matte_blue
    ((0.25, 0.25, 0.75), 0.0, 0.3, 0.6, MAT_DIFFUSE),    # 4 plastic
    ((1.0, 0.5, 0.5), 0.0, 0.0, 1.0, MAT_DIFFUSE),       # 5 mirror
    ((1.0, 1.0, 1.0), 0.0, 0.42, 0.0, MAT_DIELECTRIC),   # 6 glass
    ((1.0, 1.0, 1.0), 128.0, 0.6, 0.0, MAT_DIFFUSE),     # 7 light
]
# Tracer.comp:204-211
_TRACER_PLANES = np.array([[0, 1, 0, 0], [0, -1, 0, 128], [1, 0, 0, 64], [0, 0, -1, 64], [-1, 0, 0, 64]], dtype=np.float32)
_TRACER_PLANE_MAT = np.array([0, 0, 1, 2, 3], dtype=np.uint32)
_LIGHT_SPHERE = (0.0, 96.0, 0.0, 12.0)   # Tracer.comp:199


def _mats(rows):
    return pack_materials([r[0] for r in rows], [r[2] for r in rows], [[r[1]] * 3 for r in rows],
                          [r[3] for r in rows], [r[4] for r in rows])


def tracer_default():
    """The shader-constant scene of Assets/Tracer.comp:186-211."""
    spheres = np.array([[42, 16, 12, 16], _LIGHT_SPHERE, [-32, 24, 24, 24], [-24, 11, -48, 11]], dtype=np.float32)
    return Scene("tracer_default", _mats(_TRACER_MATS), spheres, np.array([6, 7, 5, 4], dtype=np.uint32),
                 _TRACER_PLANES.copy(), _TRACER_PLANE_MAT.copy(), _HOST_TRIANGLE.copy(), 5)


def raytracer_default(with_emitter=False):
    """Assets/Raytracer.comp:98-127.  Materials {reflective, diffuse} map to metalness 1 / roughness 0
    when reflective, else metalness 0 / roughness 0.4 (the reference defines no
    such mapping -- the whitted integrator only reads albedo and metalness >= 0.5).
    with_emitter adds Tracer.comp's `light` sphere so that the PATH integrator has a light (config 2)."""
    rows = [((1, 1, 1), True), ((1, 0, 0), False), ((0, 1, 0), True), ((0, 0, 1), False), ((1, 1, 0), False), ((1, 0, 1), False)]
    mats = [(c, 0.0, 0.0 if refl else 0.4, 1.0 if refl else 0.0, MAT_DIFFUSE) for c, refl in rows]
    spheres = [[-14, 12, 32, 5], [32, 24, 25, 12]]
    sphere_mat = [5, 4]
    if with_emitter:
        mats.append(_TRACER_MATS[7])
        spheres.append(list(_LIGHT_SPHERE))
        sphere_mat.append(6)
    planes = np.array([[0, 1, 0, 0], [0, -1, 0, 128], [0, 0, -1, 64], [1, 0, 0, 64], [-1, 0, 0, 64]], dtype=np.float32)
    return Scene("raytracer_default" + ("+light" if with_emitter else ""), _mats(mats), np.array(spheres, dtype=np.float32),
                 np.array(sphere_mat, dtype=np.uint32), planes, np.array([0, 0, 2, 1, 3], dtype=np.uint32),
                 _HOST_TRIANGLE.copy(), 1)


# ---- seeded synthetic scenes ----------------------------------------------------------------
def _pcg(v):
    v = np.asarray(v, dtype=np.uint64) & 0xFFFFFFFF
    state = (v * 747796405 + 2891336453) & 0xFFFFFFFF
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    return ((word >> 22) ^ word) & 0xFFFFFFFF


def _u01(seed, i, k):
    """uniform [0,1) float32 for (sphere i, stream k): pcg(pcg(seed ^ pcg(i)) + k) >> 8 * 2^-24."""
    h = _pcg(_pcg(np.uint64(seed) ^ _pcg(i)) + np.uint64(k))
    return (h >> 8).astype(np.float32) * np.float32(2.0 ** -24)


def _synthetic_materials(seed, n):
    """hash(i) mod 10: 0-5 lambertian, 6-8 metal, 9 dielectric."""
    i = np.arange(n, dtype=np.uint64)
    kind = (_pcg(np.uint64(seed) + 977 * i + 13) % 10).astype(np.int64)
    lam, met, die = kind < 6, (kind >= 6) & (kind < 9), kind == 9
    albedo = np.stack([0.2 + 0.7 * _u01(seed, i, 10 + c) for c in range(3)], axis=1).astype(np.float32)
    albedo[met] = (0.5 + 0.5 * ((albedo[met] - 0.2) / 0.7)).astype(np.float32)
    albedo[die] = 1.0
    rough = np.full(n, 0.4, dtype=np.float32)
    rough[met] = (0.3 * _u01(seed, i, 13))[met]
    rough[die] = (0.42 + 0.25 * _u01(seed, i, 14))[die]
    metal = np.where(met, 1.0, 0.0).astype(np.float32)
    mtype = np.where(die, MAT_DIELECTRIC, MAT_DIFFUSE).astype(np.uint32)
    return albedo, rough, metal, mtype


def _assemble(name, seed, centres, radii):
    n = centres.shape[0]
    albedo, rough, metal, mtype = _synthetic_materials(seed, n)
    base = _mats(_TRACER_MATS)                                   # ids 0..7: planes + light keep Tracer.comp's materials
    procedural = pack_materials(albedo, rough, np.zeros((n, 3), np.float32), metal, mtype)
    spheres = np.concatenate([np.array([_LIGHT_SPHERE], dtype=np.float32),
                              np.concatenate([centres, radii[:, None]], axis=1).astype(np.float32)], axis=0)
    sphere_mat = np.concatenate([np.array([7], dtype=np.uint32), 8 + np.arange(n, dtype=np.uint32)])
    return Scene(name, np.concatenate([base, procedural], axis=0), spheres, sphere_mat, _TRACER_PLANES.copy(),
                 _TRACER_PLANE_MAT.copy(), np.zeros((0, 12), dtype=np.float32), 5)


def random_spheres(n=1024, seed=None):
    """Config 3: n spheres, centres uniform in x,z in [-60,60], y in [2,120], radii uniform [0.5,3],
    mixed lambertian / metal / dielectric, + Tracer.comp's light sphere and 5 planes."""
    seed = n if seed is None else seed
    i = np.arange(n, dtype=np.uint64)
    cx = -60.0 + 120.0 * _u01(seed, i, 0)
    cy = 2.0 + 118.0 * _u01(seed, i, 1)
    cz = -60.0 + 120.0 * _u01(seed, i, 2)
    r = 0.5 + 2.5 * _u01(seed, i, 3)
    return _assemble("random_spheres_%d" % n, seed, np.stack([cx, cy, cz], axis=1).astype(np.float32), r.astype(np.float32))


def grid_spheres(nx=50, ny=40, nz=50, seed=None):
    """Config 4/5: jittered nx*ny*nz grid filling the room box, radius = 0.35 x the smallest cell edge,
    jitter keeps every sphere inside its own cell (no overlaps).  50x40x50 = 100,000 spheres."""
    n = nx * ny * nz
    seed = n if seed is None else seed
    i = np.arange(n, dtype=np.uint64)
    ix, iy, iz = (i % nx).astype(np.float32), ((i // nx) % ny).astype(np.float32), (i // (nx * ny)).astype(np.float32)
    cell = np.array([120.0 / nx, 118.0 / ny, 120.0 / nz], dtype=np.float32)
    r = np.float32(0.35) * cell.min()
    amp = cell * 0.5 - r
    cx = -60.0 + (ix + 0.5) * cell[0] + (2.0 * _u01(seed, i, 0) - 1.0) * amp[0]
    cy = 2.0 + (iy + 0.5) * cell[1] + (2.0 * _u01(seed, i, 1) - 1.0) * amp[1]
    cz = -60.0 + (iz + 0.5) * cell[2] + (2.0 * _u01(seed, i, 2) - 1.0) * amp[2]
    return _assemble("grid_spheres_%d" % n, seed, np.stack([cx, cy, cz], axis=1).astype(np.float32),
                     np.full(n, r, dtype=np.float32))


# ---- triangle meshes (ref: the TODO at Assets/Raytracer.comp:10) ---------------------
def torus_triangles(centre=(0.0, 40.0, 10.0), major=20.0, minor=7.0, n_major=36, n_minor=18, tilt=0.6):
    """A closed torus as (n, 12) float32 triangles in the reference's SSBO layout, counter-clockwise seen from outside
    (tri_intersect culls back faces, Tracer.comp:348), tilted about the x axis."""
    u = np.arange(n_major + 1, dtype=np.float64) * (2.0 * np.pi / n_major)
    v = np.arange(n_minor + 1, dtype=np.float64) * (2.0 * np.pi / n_minor)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    x = (major + minor * np.cos(vv)) * np.cos(uu)
    z = (major + minor * np.cos(vv)) * np.sin(uu)
    y = minor * np.sin(vv)
    ct, st = np.cos(tilt), np.sin(tilt)
    y, z = ct * y - st * z, st * y + ct * z
    p = np.stack([x + centre[0], y + centre[1], z + centre[2]], axis=-1)
    tris = []
    for i in range(n_major):
        for j in range(n_minor):
            a, b, c, d = p[i, j], p[i + 1, j], p[i + 1, j + 1], p[i, j + 1]
            tris.append((a, c, b))
            tris.append((a, d, c))
    t = np.zeros((len(tris), 12), dtype=np.float32)
    for k, (a, b, c) in enumerate(tris):
        t[k, 0:3], t[k, 4:7], t[k, 8:11] = a, b, c
    return t


def write_obj(path, triangles):
    """Writes a triangle list as a Wavefront OBJ file (one `v` per corner, %.9g: float32 round-trips exactly)."""
    with open(path, "w") as f:
        f.write("# vk-renderer_b200 triangle list\n")
        for t in triangles:
            for k in range(3):
                f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in t[4 * k:4 * k + 3]))
        for i in range(len(triangles)):
            f.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))


def mesh_scene(triangles, n_spheres=300, seed=77):
    """A triangle mesh with per-triangle materials (diffuse / mirror / glass / plastic by triangle index) among
    `n_spheres` random spheres, Tracer.comp's light and planes."""
    s = random_spheres(n_spheres, seed=seed)
    k = np.arange(triangles.shape[0], dtype=np.uint32)
    # materials 0..7 are Tracer.comp's: 1 matte_red, 5 mirror, 6 glass, 4 plastic
    tri_mats = np.array([1, 5, 4, 6, 2, 5], dtype=np.uint32)[(k // 7) % 6]
    return Scene("mesh_%d+%s" % (triangles.shape[0], s.name), s.materials, s.spheres, s.sphere_mat, s.planes, s.plane_mat,
                 np.ascontiguousarray(triangles, dtype=np.float32), 5, tri_mats)
