"""Builds b200pt_scene_desc structures in Python (ctypes mirrors of include/b200pt.h) for synthetic parity scenes.

The five BASELINE scenes never touch plastic / thin-dielectric / rough-diffuse BSDFs, anisotropic conductors, bump maps,
cylinders, spot / point / sun / constant emitters or isotropic media; the scenes assembled here do, so that those branches of
the path are pinned against the reference build and compared with the CUDA path like everything else.
"""
import ctypes

import numpy as np

u32, i32, u64, f32 = ctypes.c_uint32, ctypes.c_int32, ctypes.c_uint64, ctypes.c_float
INVALID = 0xFFFFFFFF
NO_OFFSET = 0xFFFFFFFFFFFFFFFF

TEX_CONSTANT, TEX_CHECKERBOARD, TEX_BITMAP = 1, 2, 3
BSDF_AREA_LIGHT, BSDF_DIFFUSE, BSDF_ROUGH_DIFFUSE, BSDF_CONDUCTOR, BSDF_DIELECTRIC, BSDF_THIN_DIELECTRIC, BSDF_PLASTIC = 1, 2, 3, 4, 5, 6, 7
INST_CUBE, INST_RECTANGLE, INST_MESHES, INST_SPHERE, INST_DISK, INST_CYLINDER = 1, 2, 3, 4, 5, 6
EMIT_POINT, EMIT_SPOT, EMIT_DIRECTIONAL, EMIT_SUN, EMIT_ENVMAP, EMIT_CONSTANT = 1, 2, 3, 4, 5, 6


class Camera(ctypes.Structure):
    _fields_ = [("spp", u32), ("width", i32), ("height", i32), ("fov_x", f32), ("eye", f32 * 3), ("look_at", f32 * 3), ("up", f32 * 3)]


class Integrator(ctypes.Structure):
    _fields_ = [("type", u32), ("hide_emitters", u32), ("pdf_rr", f32), ("depth_rr", u32), ("depth_max", u32)]


class Texture(ctypes.Structure):
    _fields_ = [("type", u32), ("color0", f32 * 3), ("color1", f32 * 3), ("to_uv", f32 * 16), ("width", i32), ("height", i32),
                ("channels", i32), ("reserved", u32), ("pixel_offset", u64)]


class Bsdf(ctypes.Structure):
    _fields_ = [("type", u32), ("twosided", u32), ("id_opacity", u32), ("id_bump_map", u32), ("id_radiance", u32),
                ("id_diffuse_reflectance", u32), ("id_roughness_u", u32), ("id_roughness_v", u32), ("id_specular_reflectance", u32),
                ("id_specular_transmittance", u32), ("eta", f32), ("reflectivity", f32 * 3), ("edgetint", f32 * 3),
                ("area_light_weight", f32), ("use_fast_approx", u32)]


class Medium(ctypes.Structure):
    _fields_ = [("type", u32), ("sigma_a", f32 * 3), ("sigma_s", f32 * 3), ("phase_type", u32), ("g", f32 * 3)]


class Instance(ctypes.Structure):
    _fields_ = [("type", u32), ("id_bsdf", u32), ("id_medium_int", u32), ("id_medium_ext", u32), ("flip_normals", u32),
                ("to_world", f32 * 16), ("sphere_radius", f32), ("sphere_center", f32 * 3), ("cylinder_radius", f32),
                ("cylinder_p0", f32 * 3), ("cylinder_p1", f32 * 3), ("reserved", u32), ("num_vertices", u64), ("num_triangles", u64),
                ("position_offset", u64), ("normal_offset", u64), ("texcoord_offset", u64), ("tangent_offset", u64),
                ("bitangent_offset", u64), ("index_offset", u64)]


class Emitter(ctypes.Structure):
    _fields_ = [("type", u32), ("position", f32 * 3), ("direction", f32 * 3), ("radiance", f32 * 3), ("cutoff_angle", f32),
                ("beam_width", f32), ("cos_cutoff_angle", f32), ("id_texture", u32), ("to_world", f32 * 16)]


class SceneDesc(ctypes.Structure):
    _fields_ = [("abi_version", u32), ("reserved", u32), ("camera", Camera), ("integrator", Integrator),
                ("num_textures", u64), ("textures", ctypes.c_void_p), ("num_pixels", u64), ("pixels", ctypes.c_void_p),
                ("num_bsdfs", u64), ("bsdfs", ctypes.c_void_p), ("num_media", u64), ("media", ctypes.c_void_p),
                ("num_instances", u64), ("instances", ctypes.c_void_p), ("num_emitters", u64), ("emitters", ctypes.c_void_p),
                ("num_positions", u64), ("positions", ctypes.c_void_p), ("num_normals", u64), ("normals", ctypes.c_void_p),
                ("num_texcoords", u64), ("texcoords", ctypes.c_void_p), ("num_tangents", u64), ("tangents", ctypes.c_void_p),
                ("num_bitangents", u64), ("bitangents", ctypes.c_void_p), ("num_triangles", u64), ("indices", ctypes.c_void_p)]


IDENTITY = np.eye(4, dtype=np.float32)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def scale(x, y, z):
    return np.diag([x, y, z, 1.0]).astype(np.float32)


def rotate_x(deg):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def rotate_y(deg):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def _vec(dst, values):
    for k, v in enumerate(values):
        dst[k] = float(v)


class SceneBuilder:
    """Mirrors what src/parser/parser.cpp appends to a csrt::RendererConfig, one call per XML element."""

    def __init__(self, width=64, height=64, spp=64, fov_x=40.0, eye=(0, 1, 5), look_at=(0, 1, 0), up=(0, 1, 0),
                 volpath=False, depth_max=8, depth_rr=5, pdf_rr=0.95):
        self.textures, self.bsdfs, self.media, self.instances, self.emitters = [], [], [], [], []
        self.pixels = []
        self.positions, self.normals, self.texcoords, self.indices = [], [], [], []
        self.camera = Camera(spp, width, height, fov_x)
        _vec(self.camera.eye, eye), _vec(self.camera.look_at, look_at), _vec(self.camera.up, up)
        self.integrator = Integrator(1 if volpath else 0, 0, pdf_rr, depth_rr, depth_max)
        self._keep = []

    # ---- textures ----
    def constant(self, r, g=None, b=None):
        t = Texture()
        t.type = TEX_CONSTANT
        _vec(t.color0, (r, r if g is None else g, r if b is None else b))
        _vec(t.to_uv, IDENTITY.reshape(-1))
        self.textures.append(t)
        return len(self.textures) - 1

    def checkerboard(self, c0, c1, to_uv=IDENTITY):
        t = Texture()
        t.type = TEX_CHECKERBOARD
        _vec(t.color0, c0), _vec(t.color1, c1), _vec(t.to_uv, np.asarray(to_uv, dtype=np.float32).reshape(-1))
        self.textures.append(t)
        return len(self.textures) - 1

    def bitmap(self, array):
        """array: [h, w, channels] float32 (already linear, as image_io::Read returns it)."""
        array = np.ascontiguousarray(array, dtype=np.float32)
        t = Texture()
        t.type = TEX_BITMAP
        t.height, t.width, t.channels = array.shape
        t.pixel_offset = sum(len(p) for p in self.pixels)
        _vec(t.to_uv, IDENTITY.reshape(-1))
        self.pixels.append(array.reshape(-1))
        self.textures.append(t)
        return len(self.textures) - 1

    # ---- BSDFs ----
    def _bsdf(self, kind, twosided=False, bump=INVALID, opacity=INVALID):
        b = Bsdf()
        b.type, b.twosided, b.id_opacity, b.id_bump_map = kind, int(twosided), opacity, bump
        b.id_radiance = b.id_diffuse_reflectance = b.id_roughness_u = b.id_roughness_v = INVALID
        b.id_specular_reflectance = b.id_specular_transmittance = INVALID
        b.eta, b.area_light_weight = 1.0, 1.0
        self.bsdfs.append(b)
        return b, len(self.bsdfs) - 1

    def diffuse(self, reflectance_tex, twosided=True, bump=INVALID, opacity=INVALID):
        b, i = self._bsdf(BSDF_DIFFUSE, twosided, bump, opacity)
        b.id_diffuse_reflectance = reflectance_tex
        return i

    def rough_diffuse(self, reflectance_tex, alpha, twosided=True):
        b, i = self._bsdf(BSDF_ROUGH_DIFFUSE, twosided)
        b.id_diffuse_reflectance = reflectance_tex
        b.id_roughness_u = b.id_roughness_v = self.constant(alpha)
        return i

    def conductor(self, alpha_u, alpha_v, eta, k, twosided=True):
        # parser.cpp:944-951: reflectivity / edgetint from eta, k
        eta, k = np.asarray(eta, dtype=np.float32), np.asarray(k, dtype=np.float32)
        refl = ((eta - 1) ** 2 + k ** 2) / ((eta + 1) ** 2 + k ** 2)
        t1, t2, t3 = 1 + np.sqrt(refl), 1 - np.sqrt(refl), (1 - refl) / (1 + refl)
        edge = (t1 - eta * t2) / (t1 - t3 * t2)
        b, i = self._bsdf(BSDF_CONDUCTOR, twosided)
        b.id_roughness_u, b.id_roughness_v = self.constant(alpha_u), self.constant(alpha_v)
        b.id_specular_reflectance = self.constant(1.0)
        _vec(b.reflectivity, refl), _vec(b.edgetint, edge)
        return i

    def dielectric(self, eta, alpha=0.001, thin=False):
        b, i = self._bsdf(BSDF_THIN_DIELECTRIC if thin else BSDF_DIELECTRIC, True)
        b.id_roughness_u = b.id_roughness_v = self.constant(alpha)
        b.id_specular_reflectance, b.id_specular_transmittance = self.constant(1.0), self.constant(1.0)
        b.eta = eta
        return i

    def plastic(self, diffuse_tex, eta=1.5, alpha=0.001, twosided=True):
        b, i = self._bsdf(BSDF_PLASTIC, twosided)
        b.id_roughness_u = b.id_roughness_v = self.constant(alpha)
        b.id_diffuse_reflectance, b.id_specular_reflectance = diffuse_tex, self.constant(1.0)
        b.eta = eta
        return i

    def area_light(self, radiance):
        b, i = self._bsdf(BSDF_AREA_LIGHT, False)
        b.id_radiance = self.constant(*radiance)
        return i

    # ---- media ----
    def medium(self, sigma_a, sigma_s, g=None):
        m = Medium()
        _vec(m.sigma_a, sigma_a), _vec(m.sigma_s, sigma_s)
        m.phase_type = 0 if g is None else 1
        _vec(m.g, (g or 0.0,) * 3)
        self.media.append(m)
        return len(self.media) - 1

    # ---- shapes ----
    def _instance(self, kind, bsdf, to_world, medium_int=INVALID, medium_ext=INVALID):
        s = Instance()
        s.type, s.id_bsdf, s.id_medium_int, s.id_medium_ext = kind, bsdf, medium_int, medium_ext
        _vec(s.to_world, np.asarray(to_world, dtype=np.float32).reshape(-1))
        s.sphere_radius, s.cylinder_radius = 1.0, 1.0
        s.position_offset = s.normal_offset = s.texcoord_offset = s.tangent_offset = s.bitangent_offset = NO_OFFSET
        self.instances.append(s)
        return s

    def rectangle(self, bsdf, to_world=IDENTITY, **kw):
        self._instance(INST_RECTANGLE, bsdf, to_world, **kw)

    def cube(self, bsdf, to_world=IDENTITY, **kw):
        self._instance(INST_CUBE, bsdf, to_world, **kw)

    def sphere(self, bsdf, center, radius, to_world=IDENTITY, **kw):
        s = self._instance(INST_SPHERE, bsdf, to_world, **kw)
        s.sphere_radius = radius
        _vec(s.sphere_center, center)

    def disk(self, bsdf, to_world=IDENTITY, **kw):
        self._instance(INST_DISK, bsdf, to_world, **kw)

    def cylinder(self, bsdf, p0, p1, radius, to_world=IDENTITY, **kw):
        s = self._instance(INST_CYLINDER, bsdf, to_world, **kw)
        s.cylinder_radius = radius
        _vec(s.cylinder_p0, p0), _vec(s.cylinder_p1, p1)

    def mesh(self, bsdf, positions, indices, normals=None, texcoords=None, to_world=IDENTITY, **kw):
        s = self._instance(INST_MESHES, bsdf, to_world, **kw)
        positions = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
        indices = np.asarray(indices, dtype=np.uint32).reshape(-1, 3)
        s.num_vertices, s.num_triangles = len(positions), len(indices)
        s.position_offset = sum(len(p) for p in self.positions) // 3
        self.positions.append(positions.reshape(-1))
        s.index_offset = sum(len(p) for p in self.indices) // 3
        self.indices.append(indices.reshape(-1))
        if normals is not None:
            s.normal_offset = sum(len(p) for p in self.normals) // 3
            self.normals.append(np.asarray(normals, dtype=np.float32).reshape(-1))
        if texcoords is not None:
            s.texcoord_offset = sum(len(p) for p in self.texcoords) // 2
            self.texcoords.append(np.asarray(texcoords, dtype=np.float32).reshape(-1))

    # ---- emitters ----
    def _emitter(self, kind):
        e = Emitter()
        e.type, e.id_texture = kind, INVALID
        _vec(e.to_world, IDENTITY.reshape(-1))
        self.emitters.append(e)
        return e

    def directional(self, direction, radiance):
        e = self._emitter(EMIT_DIRECTIONAL)
        d = np.asarray(direction, dtype=np.float32)
        _vec(e.direction, d / np.linalg.norm(d)), _vec(e.radiance, radiance)

    def point(self, position, intensity):
        e = self._emitter(EMIT_POINT)
        _vec(e.position, position), _vec(e.radiance, intensity)

    def spot(self, to_world, intensity, cutoff_deg=25.0, beam_deg=18.0, texture=INVALID):
        e = self._emitter(EMIT_SPOT)
        _vec(e.to_world, np.asarray(to_world, dtype=np.float32).reshape(-1)), _vec(e.radiance, intensity)
        e.cutoff_angle, e.beam_width, e.id_texture = np.radians(cutoff_deg), np.radians(beam_deg), texture
        return e

    def constant_env(self, radiance):
        _vec(self._emitter(EMIT_CONSTANT).radiance, radiance)

    def envmap(self, texture, to_world=IDENTITY):
        e = self._emitter(EMIT_ENVMAP)
        e.id_texture = texture
        _vec(e.to_world, np.asarray(to_world, dtype=np.float32).reshape(-1))

    def sun(self, direction, radiance, texture, cos_cutoff=0.9999):
        e = self._emitter(EMIT_SUN)
        d = np.asarray(direction, dtype=np.float32)
        _vec(e.direction, d / np.linalg.norm(d)), _vec(e.radiance, radiance)
        e.id_texture, e.cos_cutoff_angle = texture, cos_cutoff

    # ---- finish ----
    def desc(self):
        """Returns a SceneDesc whose pointers stay valid as long as this builder lives."""
        d = SceneDesc()
        d.abi_version = 1
        d.camera, d.integrator = self.camera, self.integrator

        def table(kind, items):
            arr = (kind * max(1, len(items)))(*items)
            self._keep.append(arr)
            return len(items), ctypes.cast(arr, ctypes.c_void_p)

        def pool(chunks, dtype):
            arr = np.concatenate(chunks).astype(dtype) if chunks else np.zeros(1, dtype=dtype)
            self._keep.append(arr)
            return arr

        d.num_textures, d.textures = table(Texture, self.textures)
        d.num_bsdfs, d.bsdfs = table(Bsdf, self.bsdfs)
        d.num_media, d.media = table(Medium, self.media)
        d.num_instances, d.instances = table(Instance, self.instances)
        d.num_emitters, d.emitters = table(Emitter, self.emitters)
        px = pool(self.pixels, np.float32)
        d.num_pixels, d.pixels = (len(px) if self.pixels else 0), px.ctypes.data
        pos, nrm, uv, idx = pool(self.positions, np.float32), pool(self.normals, np.float32), pool(self.texcoords, np.float32), pool(self.indices, np.uint32)
        d.num_positions, d.positions = (len(pos) // 3 if self.positions else 0), pos.ctypes.data
        d.num_normals, d.normals = (len(nrm) // 3 if self.normals else 0), nrm.ctypes.data
        d.num_texcoords, d.texcoords = (len(uv) // 2 if self.texcoords else 0), uv.ctypes.data
        d.num_tangents, d.tangents, d.num_bitangents, d.bitangents = 0, None, 0, None
        d.num_triangles, d.indices = (len(idx) // 3 if self.indices else 0), idx.ctypes.data
        self._keep.append(d)
        return d


def stage(b):
    """Floor + back wall + ceiling area light shared by the synthetic scenes."""
    floor = b.diffuse(b.checkerboard((0.7, 0.7, 0.7), (0.25, 0.3, 0.35), scale(4, 4, 1)))
    wall = b.diffuse(b.constant(0.6, 0.55, 0.5))
    b.rectangle(floor, translate(0, 0, 0) @ rotate_x(-90) @ scale(4, 4, 1))
    b.rectangle(wall, translate(0, 2, -3) @ scale(4, 2, 1))
    b.rectangle(b.area_light((18, 17, 15)), translate(0, 3.6, 0.5) @ rotate_x(90) @ scale(0.7, 0.7, 1))


def uv_sphere(n_lat=12, n_lon=24, radius=1.0):
    """A tessellated sphere with normals and uvs (exercises the kMeshes path with texcoords)."""
    pos, nrm, uv, idx = [], [], [], []
    for a in range(n_lat + 1):
        theta = np.pi * a / n_lat
        for o in range(n_lon + 1):
            phi = 2 * np.pi * o / n_lon
            n = (np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi))
            nrm.append(n), pos.append([radius * c for c in n]), uv.append((o / n_lon, a / n_lat))
    for a in range(n_lat):
        for o in range(n_lon):
            i0 = a * (n_lon + 1) + o
            i1, i2, i3 = i0 + 1, i0 + n_lon + 1, i0 + n_lon + 2
            if a > 0:
                idx.append((i0, i2, i1))
            if a < n_lat - 1:
                idx.append((i1, i2, i3))
    return pos, idx, nrm, uv


def synthetic_scenes():
    """name -> SceneBuilder; every branch of bsdfs/*, emitters/*, medium/* and primitives/* not covered by the BASELINE scenes."""
    rng = np.random.RandomState(7)
    scenes = {}

    b = SceneBuilder(depth_max=8)
    stage(b)
    b.sphere(b.plastic(b.constant(0.2, 0.4, 0.7), eta=1.5, alpha=0.15), (-1.3, 0.6, 0.3), 0.6)
    b.sphere(b.plastic(b.constant(0.7, 0.3, 0.2), eta=1.49), (0.0, 0.6, 0.6), 0.6)
    b.cube(b.rough_diffuse(b.constant(0.6, 0.6, 0.3), 0.4), translate(1.4, 0.5, 0.0) @ rotate_y(30) @ scale(0.5, 0.5, 0.5))
    b.constant_env((0.15, 0.17, 0.2))
    scenes["plastic_roughdiffuse"] = b

    b = SceneBuilder(depth_max=10)
    stage(b)
    b.rectangle(b.dielectric(1.5, thin=True), translate(-1.2, 1.0, 1.0) @ rotate_y(20) @ scale(0.6, 0.9, 1))
    b.sphere(b.dielectric(1.33, alpha=0.2), (0.9, 0.7, 0.5), 0.7)
    b.cylinder(b.conductor(0.05, 0.3, (0.2, 0.92, 1.1), (3.9, 2.45, 2.14)), (-0.2, 0.0, -1.0), (-0.2, 1.6, -1.0), 0.45)
    b.point((2.5, 3.0, 2.0), (30, 30, 30))
    b.spot(translate(-2.5, 3.0, 2.5) @ rotate_y(-45) @ rotate_x(215), (60, 55, 50), 30.0, 20.0)
    scenes["dielectrics_conductor_cylinder"] = b

    b = SceneBuilder(depth_max=6)
    stage(b)
    bump = b.bitmap(rng.rand(32, 32, 3).astype(np.float32))
    albedo = b.bitmap((0.2 + 0.6 * rng.rand(16, 16, 3)).astype(np.float32))
    pos, idx, nrm, uv = uv_sphere()
    b.mesh(b.diffuse(albedo, bump=bump), pos, idx, nrm, uv, to_world=translate(-0.9, 0.9, 0.3) @ scale(0.9, 0.9, 0.9))
    b.mesh(b.diffuse(b.checkerboard((0.8, 0.2, 0.2), (0.9, 0.9, 0.9), scale(6, 3, 1))), pos, idx, nrm, uv,
           to_world=translate(1.1, 0.7, 0.6) @ scale(0.7, 0.7, 0.7))
    b.disk(b.diffuse(b.constant(0.3, 0.6, 0.3)), translate(0.2, 0.02, 1.8) @ rotate_x(-90) @ scale(1.2, 1.2, 1))
    b.directional((0.4, -1.0, -0.5), (2.5, 2.4, 2.2))
    scenes["bump_bitmap_mesh_disk"] = b

    b = SceneBuilder(depth_max=8)
    sky = np.zeros((16, 32, 3), dtype=np.float32)
    sky[:8] = (0.35, 0.5, 0.9)
    sky[8:] = (0.25, 0.22, 0.2)
    sky[2:4, 4:6] = (40.0, 38.0, 30.0)
    b.envmap(b.bitmap(sky), rotate_y(35))
    b.sun((0.3, -0.8, -0.4), (6.0, 5.5, 5.0), b.bitmap(sky), cos_cutoff=0.995)
    floor = b.diffuse(b.constant(0.5, 0.5, 0.5))
    b.rectangle(floor, rotate_x(-90) @ scale(5, 5, 1))
    b.sphere(b.conductor(0.2, 0.2, (0.14, 0.37, 1.44), (3.98, 2.38, 1.6)), (-0.8, 0.7, 0), 0.7)
    b.cube(b.diffuse(b.constant(0.7, 0.2, 0.2), twosided=False), translate(1.0, 0.5, 0.2) @ rotate_y(20) @ scale(0.5, 0.5, 0.5))
    scenes["envmap_sun_onesided"] = b

    b = SceneBuilder(volpath=True, depth_max=8)
    stage(b)
    fog = b.medium((0.05, 0.05, 0.05), (0.5, 0.6, 0.7))
    b.cube(INVALID, translate(0, 1.0, 0.3) @ scale(1.2, 0.9, 0.9), medium_int=fog)  # a BSDF-less box filled with an isotropic medium
    b.sphere(b.diffuse(b.constant(0.8, 0.8, 0.2)), (0.0, 0.8, 0.3), 0.4, medium_ext=fog)
    scenes["isotropic_medium_null_surface"] = b

    # Opacity masks: Bsdf::IsTransparent inside every primitive test (triangle.cpp:116, sphere.cpp:42, disk.cpp:41,
    # cylinder.cpp:50) with constant (constant_texture.cpp:18) and 4-channel bitmap (bitmap.cpp:70) alpha.
    b = SceneBuilder(depth_max=6)
    stage(b)
    rgba = np.zeros((8, 8, 4), dtype=np.float32)
    rgba[..., :3] = 0.5
    rgba[..., 3] = (np.add.outer(np.arange(8), np.arange(8)) % 2) * 0.8 + 0.1   # alpha 0.1 / 0.9 checker
    mask = b.bitmap(rgba)
    pos, idx, nrm, uv = uv_sphere()
    b.mesh(b.diffuse(b.constant(0.8, 0.3, 0.2), opacity=mask), pos, idx, nrm, uv, to_world=translate(-1.2, 0.8, 0.4) @ scale(0.8, 0.8, 0.8))
    b.rectangle(b.diffuse(b.constant(0.2, 0.3, 0.8), opacity=b.constant(0.4)), translate(0.2, 1.0, 1.4) @ rotate_y(-15) @ scale(0.7, 0.8, 1))
    b.sphere(b.diffuse(b.constant(0.3, 0.7, 0.3), opacity=b.constant(0.6)), (1.3, 0.6, 0.2), 0.6)
    b.disk(b.diffuse(b.constant(0.7, 0.7, 0.2), opacity=mask), translate(0.0, 0.05, 2.2) @ rotate_x(-90) @ scale(1.6, 1.6, 1))
    b.cylinder(b.diffuse(b.constant(0.6, 0.3, 0.6), opacity=b.constant(0.5)), (0.3, 0.0, -0.8), (0.3, 1.8, -0.8), 0.35)
    b.point((2.0, 3.0, 3.0), (25, 25, 25))
    scenes["opacity_masks"] = b

    # Depth limits of the path loop `depth < depth_rr || (depth < depth_max && rand < pdf_rr)` (path.cpp:57-60):
    # a max_depth below rr_depth still walks rr_depth - 1 vertices; Russian roulette from the first vertex on.
    b = SceneBuilder(depth_max=2, depth_rr=5)
    stage(b)
    b.sphere(b.diffuse(b.constant(0.75, 0.7, 0.3)), (-0.8, 0.7, 0.4), 0.7)
    b.cube(b.conductor(0.15, 0.15, (0.2, 0.92, 1.1), (3.9, 2.45, 2.14)), translate(1.1, 0.5, 0.3) @ rotate_y(25) @ scale(0.5, 0.5, 0.5))
    b.constant_env((0.3, 0.35, 0.45))
    scenes["depth_max_below_rr"] = b

    b = SceneBuilder(depth_max=3, depth_rr=1, pdf_rr=0.4)
    stage(b)
    b.sphere(b.diffuse(b.constant(0.3, 0.7, 0.75)), (0.0, 0.8, 0.2), 0.8)
    b.directional((0.3, -1.0, -0.6), (1.5, 1.4, 1.3))
    b.constant_env((0.2, 0.2, 0.25))
    scenes["early_rr"] = b
    return scenes
