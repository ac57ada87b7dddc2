"""TEST INFRASTRUCTURE ONLY — ctypes access to the reference build (oracle/_ref/libcsrt_ref_*.so),
the C restatement (oracle/_ref/liboracle.so).  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference arm; never by the product package."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class RefLib:
    """The unmodified reference CPU renderer (csrt::Renderer, src/renderer/renderer.cpp:259,678)."""

    def __init__(self, variant="woop"):
        path = os.path.join(REF_DIR, f"libcsrt_ref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        os.environ.setdefault("B200PT_EXR_SIDECAR_DIR", os.path.join(REF_DIR, "exr"))
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_pack_from_xml.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        L.ref_render.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.b200pt_scene_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.b200pt_scene_get_desc.argtypes = [ctypes.c_void_p]
        L.b200pt_scene_get_desc.restype = ctypes.c_void_p
        L.b200pt_scene_free.argtypes = [ctypes.c_void_p]
        L.ref_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.ref_create.restype = ctypes.c_void_p
        L.ref_draw.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_draw.restype = ctypes.c_double
        L.ref_destroy.argtypes = [ctypes.c_void_p]
        L.ref_destroy.restype = None
        L.ref_tea4.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.ref_tea4.restype = ctypes.c_uint32
        L.ref_random_float.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.ref_random_float.restype = ctypes.c_float
        L.ref_van_der_corput2.argtypes = [ctypes.c_uint32]
        L.ref_van_der_corput2.restype = ctypes.c_float
        if hasattr(L, "ref_van_der_corput3"):
            L.ref_van_der_corput3.argtypes = [ctypes.c_uint32]
            L.ref_van_der_corput3.restype = ctypes.c_float
        L.ref_mis_weight.argtypes = [ctypes.c_float, ctypes.c_float]
        L.ref_mis_weight.restype = ctypes.c_float

    def error(self):
        return self.lib.ref_last_error().decode()

    def pack_from_xml(self, xml, out_pack, width=0, height=0, spp=0):
        rc = self.lib.ref_pack_from_xml(xml.encode(), width, height, spp, out_pack.encode())
        if rc != 0:
            raise RuntimeError(self.error())

    def render_pack(self, pack_path, width=0, height=0, spp=0):
        """Returns (frame[h,w,3] float32, build_seconds, render_seconds)."""
        scene = ctypes.c_void_p()
        if self.lib.b200pt_scene_load(pack_path.encode(), ctypes.byref(scene)) != 0:
            raise RuntimeError(f"cannot load {pack_path}")
        try:
            desc = self.lib.b200pt_scene_get_desc(scene)
            cam = np.ctypeslib.as_array(ctypes.cast(desc + 8, ctypes.POINTER(ctypes.c_int32)), shape=(3,))
            w = width or int(cam[1])
            h = height or int(cam[2])
            frame = np.zeros((h, w, 3), dtype=np.float32)
            b, r = ctypes.c_double(), ctypes.c_double()
            rc = self.lib.ref_render(desc, width, height, spp, frame.ctypes.data, ctypes.byref(b), ctypes.byref(r))
            if rc != 0:
                raise RuntimeError(self.error())
            return frame, b.value, r.value
        finally:
            self.lib.b200pt_scene_free(scene)


class RefHit(ctypes.Structure):
    """struct ref_hit of oracle/ref_glue.cpp."""
    _fields_ = [("t", ctypes.c_float), ("valid", ctypes.c_uint32), ("inside", ctypes.c_uint32), ("id_instance", ctypes.c_uint32),
                ("id_primitive", ctypes.c_uint32), ("position", ctypes.c_float * 3), ("normal", ctypes.c_float * 3),
                ("texcoord", ctypes.c_float * 2), ("tangent", ctypes.c_float * 3), ("bitangent", ctypes.c_float * 3)]


class RefTracer:
    """csrt::Scene + TLAS::Intersect / IntersectAny (tlas.cpp:13-76) of the reference build on caller-supplied rays."""

    def __init__(self, ref, pack_path=None, desc=None):
        self.ref = ref
        L = ref.lib
        L.ref_scene_create.argtypes = [ctypes.c_void_p]
        L.ref_scene_create.restype = ctypes.c_void_p
        L.ref_scene_destroy.argtypes = [ctypes.c_void_p]
        L.ref_scene_destroy.restype = None
        L.ref_trace.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        self.scene = ctypes.c_void_p()
        if pack_path is not None:
            if L.b200pt_scene_load(pack_path.encode(), ctypes.byref(self.scene)) != 0:
                raise RuntimeError(f"cannot load {pack_path}")
            desc = L.b200pt_scene_get_desc(self.scene)
        self.handle = L.ref_scene_create(desc)
        if not self.handle:
            raise RuntimeError(ref.error())

    def trace(self, rays, any_hit=False):
        """rays: float32 [n, 8] (origin, direction, t_min, t_max) -> numpy record array of RefHit."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = (RefHit * len(rays))()
        if self.ref.lib.ref_trace(self.handle, len(rays), rays.ctypes.data, 1 if any_hit else 0, out) != 0:
            raise RuntimeError(self.ref.error())
        return np.ctypeslib.as_array(out).copy() if len(rays) else np.zeros(0, dtype=RefHit)

    def close(self):
        if self.handle:
            self.ref.lib.ref_scene_destroy(self.handle)
            self.handle = None
        if self.scene:
            self.ref.lib.b200pt_scene_free(self.scene)
            self.scene = ctypes.c_void_p()


class OracleLib:
    """The plain-C restatement (oracle/pt_oracle.c -> oracle/_ref/liboracle.so)."""

    def __init__(self):
        path = os.path.join(REF_DIR, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.oracle_render.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.oracle_tea4.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.oracle_tea4.restype = ctypes.c_uint32
        L.oracle_random_float.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_random_float.restype = ctypes.c_float
        L.oracle_van_der_corput2.argtypes = [ctypes.c_uint32]
        L.oracle_van_der_corput2.restype = ctypes.c_float
        L.oracle_van_der_corput3.argtypes = [ctypes.c_uint32]
        L.oracle_van_der_corput3.restype = ctypes.c_float
        L.oracle_render_progressive.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_mis_weight.argtypes = [ctypes.c_float, ctypes.c_float]
        L.oracle_mis_weight.restype = ctypes.c_float
        L.oracle_sample_hemis_cos.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        L.oracle_kulla_conty.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_build_bvh.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
        L.oracle_build_bvh.restype = ctypes.c_uint32

    def scene(self, pack_path, watertight=True):
        """A committed oracle scene for the pointwise entries (oracle_eval / oracle_trace)."""
        return OracleScene(self, pack_path, watertight)

    def render_pack(self, pack_path, width=0, height=0, spp=0, watertight=True, threads=None):
        """Loads the pack with the product's pack reader (data I/O only) and renders it with the C restatement."""
        loader = _pack_loader()
        scene = ctypes.c_void_p()
        if loader.b200pt_scene_load(pack_path.encode(), ctypes.byref(scene)) != 0:
            raise RuntimeError(f"cannot load {pack_path}")
        try:
            desc = loader.b200pt_scene_get_desc(scene)
            cam = np.ctypeslib.as_array(ctypes.cast(desc + 8, ctypes.POINTER(ctypes.c_int32)), shape=(3,))
            w, h = width or int(cam[1]), height or int(cam[2])
            frame = np.zeros((h, w, 3), dtype=np.float32)
            rc = self.lib.oracle_render(desc, width, height, spp, 1 if watertight else 0, threads or (os.cpu_count() or 1), frame.ctypes.data)
            if rc != 0:
                raise RuntimeError("oracle_render failed")
            return frame
        finally:
            loader.b200pt_scene_free(scene)


    def render_progressive(self, pack_path, width, height, num_frames, watertight=True):
        """num_frames calls of the preview path (renderer.cpp:97-138 restated): returns (running-mean frame, sRGB bottom-up copy)."""
        loader = _pack_loader()
        scene = ctypes.c_void_p()
        if loader.b200pt_scene_load(pack_path.encode(), ctypes.byref(scene)) != 0:
            raise RuntimeError(f"cannot load {pack_path}")
        try:
            desc = loader.b200pt_scene_get_desc(scene)
            frame = np.zeros((height, width, 3), dtype=np.float32)
            srgb = np.zeros((height, width, 3), dtype=np.float32)
            for k in range(num_frames):
                if self.lib.oracle_render_progressive(desc, width, height, 1 if watertight else 0, k, frame.ctypes.data, srgb.ctypes.data) != 0:
                    raise RuntimeError("oracle_render_progressive failed")
            return frame, srgb
        finally:
            loader.b200pt_scene_free(scene)


class OracleScene:
    """oracle_scene_create + oracle_eval / oracle_trace: the C restatement's leaf functions at caller-supplied inputs."""

    def __init__(self, oracle, pack_path, watertight=True):
        self.lib = oracle.lib
        L = self.lib
        L.oracle_scene_create.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_scene_create.restype = ctypes.c_void_p
        L.oracle_scene_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_scene_destroy.restype = None
        L.oracle_eval.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_trace.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        self.loader = _pack_loader()
        self.pack = ctypes.c_void_p()
        if self.loader.b200pt_scene_load(pack_path.encode(), ctypes.byref(self.pack)) != 0:
            raise RuntimeError(f"cannot load {pack_path}")
        self.handle = L.oracle_scene_create(self.loader.b200pt_scene_get_desc(self.pack), 1 if watertight else 0)
        if not self.handle:
            raise RuntimeError("oracle_scene_create failed")

    def eval(self, what, index, inputs):
        inputs = np.ascontiguousarray(inputs, dtype=np.float32).reshape(-1, 32)
        out = np.zeros((len(inputs), 16), dtype=np.float32)
        if self.lib.oracle_eval(self.handle, what, index, len(inputs), inputs.ctypes.data, out.ctypes.data) != 0:
            raise RuntimeError("oracle_eval failed")
        return out

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = (RefHit * len(rays))()
        if self.lib.oracle_trace(self.handle, len(rays), rays.ctypes.data, 1 if any_hit else 0, out) != 0:
            raise RuntimeError("oracle_trace failed")
        return np.ctypeslib.as_array(out).copy()

    def close(self):
        if self.handle:
            self.lib.oracle_scene_destroy(self.handle)
            self.handle = None
        if self.pack:
            self.loader.b200pt_scene_free(self.pack)
            self.pack = ctypes.c_void_p()


_PACK_LOADER = None


def _pack_loader():
    """Any library that exports the scene-pack reader: the product library, else a reference build."""
    global _PACK_LOADER
    if _PACK_LOADER is None:
        candidates = [os.path.join(os.path.dirname(HERE), "monte-carlo-path-tracing_b200", "libb200pt.so"),
                      os.path.join(REF_DIR, "libcsrt_ref_woop.so"), os.path.join(REF_DIR, "libcsrt_ref_mt.so")]
        for path in candidates:
            if os.path.exists(path):
                L = ctypes.CDLL(path)
                L.b200pt_scene_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
                L.b200pt_scene_get_desc.argtypes = [ctypes.c_void_p]
                L.b200pt_scene_get_desc.restype = ctypes.c_void_p
                L.b200pt_scene_free.argtypes = [ctypes.c_void_p]
                _PACK_LOADER = L
                break
        else:
            raise FileNotFoundError("no library with b200pt_scene_load found")
    return _PACK_LOADER


class RefRenderer:
    """A persistent csrt::Renderer over a scene pack: build once, draw() many times (for timing)."""

    def __init__(self, ref, pack_path, width=0, height=0, spp=0):
        self.ref = ref
        self.scene = ctypes.c_void_p()
        if ref.lib.b200pt_scene_load(pack_path.encode(), ctypes.byref(self.scene)) != 0:
            raise RuntimeError(f"cannot load {pack_path}")
        desc = ref.lib.b200pt_scene_get_desc(self.scene)
        cam = np.ctypeslib.as_array(ctypes.cast(desc + 8, ctypes.POINTER(ctypes.c_int32)), shape=(3,))
        self.width, self.height, self.spp = width or int(cam[1]), height or int(cam[2]), spp or int(cam[0])
        b = ctypes.c_double()
        self.handle = ref.lib.ref_create(desc, width, height, spp, ctypes.byref(b))
        if not self.handle:
            raise RuntimeError(ref.error())
        self.build_seconds = b.value
        self.frame = np.zeros((self.height, self.width, 3), dtype=np.float32)

    def eval(self, what, index, inputs):
        """ref_eval (oracle/ref_glue.cpp): the renderer's own Bsdf / Emitter / Medium / Texture objects at fixed inputs;
        inputs float32 [n, 32] -> float32 [n, 16], layouts as b200pt_debug_eval (include/b200pt.h)."""
        L = self.ref.lib
        L.ref_eval.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        inputs = np.ascontiguousarray(inputs, dtype=np.float32).reshape(-1, 32)
        out = np.zeros((len(inputs), 16), dtype=np.float32)
        if L.ref_eval(self.handle, what, index, len(inputs), inputs.ctypes.data, out.ctypes.data) != 0:
            raise RuntimeError(self.ref.error())
        return out

    def trace(self, rays):
        """Closest hits through the renderer's scene WITH its BSDFs (bump-mapped hit frames): record array of RefHit."""
        L = self.ref.lib
        L.ref_trace_renderer.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = (RefHit * len(rays))()
        if L.ref_trace_renderer(self.handle, len(rays), rays.ctypes.data, out) != 0:
            raise RuntimeError(self.ref.error())
        return np.ctypeslib.as_array(out).copy()

    def draw(self):
        seconds = self.ref.lib.ref_draw(self.handle, self.frame.ctypes.data)
        if seconds < 0:
            raise RuntimeError(self.ref.error())
        return seconds

    def close(self):
        if self.handle:
            self.ref.lib.ref_destroy(self.handle)
            self.handle = None
        if self.scene:
            self.ref.lib.b200pt_scene_free(self.scene)
            self.scene = ctypes.c_void_p()


_REF_CACHE = {}


def ref_lib(variant="woop"):
    if variant not in _REF_CACHE:
        _REF_CACHE[variant] = RefLib(variant)
    return _REF_CACHE[variant]


def render_checker(pack_path, width=0, height=0, spp=0):
    """CPU frame for parity checks: the reference build (Woop triangles, like the CUDA path) when
    oracle/_ref holds it, else the C restatement.  Returns (frame, kind)."""
    try:
        frame, _, _ = ref_lib("woop").render_pack(pack_path, width, height, spp)
        return frame, "reference"
    except FileNotFoundError:
        pass
    return OracleLib().render_pack(pack_path, width, height, spp), "port"


def metrics(a, b, box=8):
    """Parity metrics of SURVEY.md §8d: per-pixel rel-L2, mean ratio, box-filtered rel-L2 (b = oracle)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    h, w = a.shape[:2]
    hb, wb = (h // box) * box, (w // box) * box
    fa = a[:hb, :wb].reshape(hb // box, box, wb // box, box, 3).mean(axis=(1, 3))
    fb = b[:hb, :wb].reshape(hb // box, box, wb // box, box, 3).mean(axis=(1, 3))
    return {
        "rel_l2": float(np.linalg.norm(a - b) / np.linalg.norm(b)),
        "mean_ratio": float(a.mean() / b.mean()),
        "box_rel_l2": float(np.linalg.norm(fa - fb) / np.linalg.norm(fb)),
        "mean_a": float(a.mean()),
        "mean_b": float(b.mean()),
    }
