"""b200pt — host-side mirror of the reference's renderer interface over the C ABI (include/b200pt.h).

The reference exposes, behind its XML parser, exactly two things:

    csrt::Renderer(const RendererConfig&)     src/renderer/renderer.cpp:259
    csrt::Renderer::Draw(float *frame)        src/renderer/renderer.cpp:678
    csrt::RayTracer(config).Draw(filename)    src/ray_tracer.cpp:124,155   (Draw + sRGB PNG)

`Renderer` / `RayTracer` below have the same names, argument meaning and error behaviour (a failing call
raises `MyException`, the reference's only exception type, include/csrt/utils/misc.hpp:54-63), but every
sample is computed by the CUDA library `libb200pt.so`.  There is no CPU fallback: importing works anywhere,
constructing a Renderer without the built library or without a GPU raises.

PyTorch is used only as plumbing (device buffers, streams, torch.distributed for the tile gather).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200PT_LIB", os.path.join(_HERE, "libb200pt.so"))  # override only for A/B experiments
INVALID_ID = 0xFFFFFFFF


class MyException(RuntimeError):
    """Mirror of csrt::MyException (include/csrt/utils/misc.hpp:54-63)."""


class RenderOpts(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("spp", ctypes.c_uint32),
                ("seed", ctypes.c_uint64), ("tile_rank", ctypes.c_uint32), ("tile_world", ctypes.c_uint32),
                ("collect_stats", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


class CreateOpts(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("max_leaf_size", ctypes.c_uint32),
                ("max_paths_in_flight", ctypes.c_uint64), ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class KernelStats(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double), ("launches", ctypes.c_uint64), ("rays", ctypes.c_uint64),
                ("node_visits", ctypes.c_uint64), ("prim_tests", ctypes.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


STATS_COUNTERS = 1  # B200PT_STATS_COUNTERS
STATS_TIMING = 2    # B200PT_STATS_TIMING
RENDER_NO_TILE_CULL = 1  # B200PT_RENDER_NO_TILE_CULL
CREATE_GPU_LBVH = 1      # B200PT_CREATE_GPU_LBVH
CREATE_BVH8 = 2          # B200PT_CREATE_BVH8
DEBUG_ANY_HIT = 1        # B200PT_DEBUG_ANY_HIT
DEBUG_PER_LANE_LOOP = 2  # B200PT_DEBUG_PER_LANE_LOOP
DEBUG_RAW_PRIM = 4       # B200PT_DEBUG_RAW_PRIM
DEBUG_PACKET_LOOP = 8    # B200PT_DEBUG_PACKET_LOOP
(EVAL_BSDF_EVALUATE, EVAL_BSDF_SAMPLE, EVAL_EMITTER_SAMPLE, EVAL_EMITTER_DIR, EVAL_MEDIUM_SAMPLE, EVAL_MEDIUM_EVALUATE,
 EVAL_PHASE_SAMPLE, EVAL_PHASE_EVALUATE, EVAL_TEXTURE, EVAL_SURFACE) = range(10)
EVAL_IN, EVAL_OUT = 32, 16


class Stats(ctypes.Structure):
    _fields_ = [("render_ms", ctypes.c_double), ("upload_ms", ctypes.c_double), ("bvh_build_ms", ctypes.c_double),
                ("bvh_gpu_ms", ctypes.c_double), ("samples", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64),
                ("num_bvh_nodes", ctypes.c_uint64), ("num_triangles", ctypes.c_uint64), ("num_prims", ctypes.c_uint64),
                ("bvh_width", ctypes.c_uint32), ("bvh_depth", ctypes.c_uint32),
                ("local_tiles", ctypes.c_uint64), ("active_tiles", ctypes.c_uint64), ("active_pixels", ctypes.c_uint64),
                ("primary", KernelStats), ("extend", KernelStats), ("shadow", KernelStats), ("shade", KernelStats),
                ("other", KernelStats), ("tail", KernelStats)]

    def as_dict(self):
        out = {}
        for name, kind in self._fields_:
            value = getattr(self, name)
            out[name] = value.as_dict() if kind is KernelStats else value
        return out


# Every symbol include/b200pt.h declares; tests check the library exports all of them.
EXPORTED_SYMBOLS = [
    "b200pt_create", "b200pt_destroy", "b200pt_render", "b200pt_render_device", "b200pt_render_progressive_device", "b200pt_tile_buffer_floats",
    "b200pt_render_tiles_device", "b200pt_assemble_tiles_device", "b200pt_get_stats", "b200pt_last_error",
    "b200pt_get_kulla_conty", "b200pt_get_envmap_tables", "b200pt_scene_load", "b200pt_scene_save",
    "b200pt_scene_get_desc", "b200pt_scene_free", "b200pt_debug_trace", "b200pt_debug_eval",
    "b200pt_debug_render_replay", "b200pt_debug_light_order",
]

_lib = None


def lib():
    """Loads libb200pt.so (built in-tree by __graft_entry__.build()).  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MyException(f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64
    L.b200pt_create.argtypes = [vp, ctypes.POINTER(CreateOpts), ctypes.POINTER(vp)]
    L.b200pt_destroy.argtypes = [vp]
    L.b200pt_destroy.restype = None
    L.b200pt_render.argtypes = [vp, ctypes.POINTER(RenderOpts), vp]
    L.b200pt_render_device.argtypes = [vp, ctypes.POINTER(RenderOpts), vp, vp]
    L.b200pt_render_progressive_device.argtypes = [vp, ctypes.POINTER(RenderOpts), u32, vp, vp, vp]
    L.b200pt_tile_buffer_floats.argtypes = [u32, u32, u32]
    L.b200pt_tile_buffer_floats.restype = u64
    L.b200pt_render_tiles_device.argtypes = [vp, ctypes.POINTER(RenderOpts), vp, vp]
    L.b200pt_assemble_tiles_device.argtypes = [vp, u32, u32, u32, vp, vp, vp]
    L.b200pt_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.b200pt_last_error.argtypes = [vp]
    L.b200pt_last_error.restype = ctypes.c_char_p
    L.b200pt_get_kulla_conty.argtypes = [vp, vp, vp]
    L.b200pt_get_envmap_tables.argtypes = [vp, vp, u64, ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_float)]
    L.b200pt_scene_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    L.b200pt_scene_save.argtypes = [vp, ctypes.c_char_p]
    L.b200pt_scene_get_desc.argtypes = [vp]
    L.b200pt_scene_get_desc.restype = vp
    L.b200pt_scene_free.argtypes = [vp]
    L.b200pt_scene_free.restype = None
    L.b200pt_debug_trace.argtypes = [vp, vp, u64, u32, vp]
    L.b200pt_debug_eval.argtypes = [vp, u32, u32, u64, vp, vp]
    L.b200pt_debug_render_replay.argtypes = [vp, u32, u32, u32, vp]
    L.b200pt_debug_light_order.argtypes = [vp, u32, vp, vp, u64, ctypes.POINTER(u64)]
    _lib = L
    return L


def _check(rc, handle=None):
    if rc != 0:
        raise MyException(lib().b200pt_last_error(handle).decode(errors="replace"))


def light_order(desc, light):
    """Test hook b200pt_debug_light_order (host only): (triangle numbers within the instance, cdf) of mesh area light `light`
    in the order of its sampling distribution.  `desc`: address of / ctypes reference to a b200pt_scene_desc."""
    n = ctypes.c_uint64()
    _check(lib().b200pt_debug_light_order(desc, light, None, None, 0, ctypes.byref(n)))
    ids, cdf = np.zeros(n.value, dtype=np.uint32), np.zeros(n.value, dtype=np.float32)
    if n.value:
        _check(lib().b200pt_debug_light_order(desc, light, ids.ctypes.data, cdf.ctypes.data, n.value, ctypes.byref(n)))
    return ids, cdf


class Scene:
    """A parsed scene (the flattened csrt::RendererConfig, renderer.hpp:18-28) loaded from a scene pack."""

    def __init__(self, path):
        self._ptr = ctypes.c_void_p()
        self.path = path
        _check(lib().b200pt_scene_load(os.fsencode(path), ctypes.byref(self._ptr)))
        self.desc = lib().b200pt_scene_get_desc(self._ptr)
        # b200pt_scene_desc: {u32 abi, u32 reserved, camera{u32 spp, i32 w, i32 h, f32 fov_x, ...}}
        head = np.ctypeslib.as_array(ctypes.cast(self.desc + 8, ctypes.POINTER(ctypes.c_int32)), shape=(3,))
        self.spp, self.width, self.height = int(head[0]), int(head[1]), int(head[2])

    def save(self, path):
        _check(lib().b200pt_scene_save(self.desc, os.fsencode(path)))

    def close(self):
        if self._ptr:
            lib().b200pt_scene_free(self._ptr)
            self._ptr = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """csrt::Renderer (renderer.hpp:30-82): construct from a config, Draw(frame) fills w*h*3 linear floats."""

    def __init__(self, scene, device=-1, max_paths_in_flight=0, max_leaf_size=0, flags=0):
        self.scene = scene
        self._h = ctypes.c_void_p()
        opts = CreateOpts(device, max_leaf_size, max_paths_in_flight, flags, 0)
        _check(lib().b200pt_create(scene.desc, ctypes.byref(opts), ctypes.byref(self._h)))

    def _opts(self, width, height, spp, seed, tile_rank=0, tile_world=1, stats=0, flags=0):
        return RenderOpts(width or 0, height or 0, spp or 0, seed, tile_rank, tile_world, int(stats), int(flags))

    def Draw(self, frame=None, width=0, height=0, spp=0, seed=0, stats=0, flags=0):
        """Host-buffer render (the reference's Draw(float*)): H2D/D2H copies happen inside the call."""
        w, h = width or self.scene.width, height or self.scene.height
        if frame is None:
            frame = np.zeros((h, w, 3), dtype=np.float32)
        if frame.dtype != np.float32 or frame.size != w * h * 3 or not frame.flags["C_CONTIGUOUS"]:
            raise MyException("frame must be a C-contiguous float32 array of width*height*3 elements.")
        opts = self._opts(width, height, spp, seed, stats=stats, flags=flags)
        _check(lib().b200pt_render(self._h, ctypes.byref(opts), frame.ctypes.data), self._h)
        return frame

    def draw_device(self, frame_tensor, width=0, height=0, spp=0, seed=0, stream=None, stats=0, flags=0):
        """Frame stays in HBM: `frame_tensor` is a CUDA float32 tensor with width*height*3 elements."""
        opts = self._opts(width, height, spp, seed, stats=stats, flags=flags)
        _check(lib().b200pt_render_device(self._h, ctypes.byref(opts), frame_tensor.data_ptr(), stream), self._h)

    def draw_progressive_device(self, frame_tensor, srgb_tensor, frame_index, width=0, height=0, seed=0, stream=None, flags=0):
        """csrt::Renderer::Draw(index_frame, frame, frame_srgb) (renderer.cpp:723-746): one more sample per pixel into the running
        mean `frame_tensor` (CUDA float32, width*height*3, updated in place); `srgb_tensor` (or None) gets the bottom-up sRGB copy."""
        opts = self._opts(width, height, 1, seed, flags=flags)
        _check(lib().b200pt_render_progressive_device(self._h, ctypes.byref(opts), int(frame_index), frame_tensor.data_ptr(),
                                                      srgb_tensor.data_ptr() if srgb_tensor is not None else None, stream), self._h)

    def draw_tiles_device(self, tiles_tensor, tile_rank, tile_world, width=0, height=0, spp=0, seed=0, stream=None, stats=0, flags=0):
        opts = self._opts(width, height, spp, seed, tile_rank, tile_world, stats, flags)
        _check(lib().b200pt_render_tiles_device(self._h, ctypes.byref(opts), tiles_tensor.data_ptr(), stream), self._h)

    def assemble_tiles_device(self, gathered_tensor, frame_tensor, width, height, tile_world, stream=None):
        _check(lib().b200pt_assemble_tiles_device(self._h, width, height, tile_world, gathered_tensor.data_ptr(),
                                                  frame_tensor.data_ptr(), stream), self._h)

    def debug_trace(self, rays, any_hit=False, per_lane_loop=False, raw_prim=False, packet_loop=False):
        """Test hook (b200pt_debug_trace): rays = float32 [n, 8] (origin, direction, tmin, tmax) through the traversal kernels.
        Returns (t [n] float32, prim [n] uint32, uv [n, 2] float32)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.zeros((len(rays), 4), dtype=np.float32)
        flags = (DEBUG_ANY_HIT if any_hit else 0) | (DEBUG_PER_LANE_LOOP if per_lane_loop else 0) | (DEBUG_RAW_PRIM if raw_prim else 0) | (DEBUG_PACKET_LOOP if packet_loop else 0)
        _check(lib().b200pt_debug_trace(self._h, rays.ctypes.data, len(rays), flags, out.ctypes.data), self._h)
        return out[:, 0].copy(), out[:, 1].copy().view(np.uint32), out[:, 2:4].copy()

    def debug_eval(self, what, index, inputs):
        """Test hook (b200pt_debug_eval): inputs float32 [n, 32] -> float32 [n, 16]; layouts in include/b200pt.h."""
        inputs = np.ascontiguousarray(inputs, dtype=np.float32).reshape(-1, EVAL_IN)
        out = np.zeros((len(inputs), EVAL_OUT), dtype=np.float32)
        _check(lib().b200pt_debug_eval(self._h, what, index, len(inputs), inputs.ctypes.data, out.ctypes.data), self._h)
        return out

    def render_replay(self, width=0, height=0, spp=0):
        """Test hook (b200pt_debug_render_replay): the frame rendered with the reference's loop shape and per-pixel LCG
        stream (Renderer::DrawPixel, renderer.cpp:62-85) around the product's device functions -> float32 [h, w, 3]."""
        width, height, spp = width or self.scene.width, height or self.scene.height, spp or self.scene.spp
        frame = np.zeros((height, width, 3), dtype=np.float32)
        _check(lib().b200pt_debug_render_replay(self._h, width, height, spp, frame.ctypes.data), self._h)
        return frame

    def stats(self):
        s = Stats()
        _check(lib().b200pt_get_stats(self._h, ctypes.byref(s)), self._h)
        return s.as_dict()

    def kulla_conty(self):
        brdf = np.zeros((128, 128), dtype=np.float32)
        albedo = np.zeros(128, dtype=np.float32)
        _check(lib().b200pt_get_kulla_conty(self._h, brdf.ctypes.data, albedo.ctypes.data), self._h)
        return brdf, albedo

    def envmap_tables(self):
        n, norm = ctypes.c_uint64(), ctypes.c_float()
        _check(lib().b200pt_get_envmap_tables(self._h, None, 0, ctypes.byref(n), ctypes.byref(norm)), self._h)
        out = np.zeros(n.value, dtype=np.float32)
        if n.value:
            _check(lib().b200pt_get_envmap_tables(self._h, out.ctypes.data, n.value, ctypes.byref(n), ctypes.byref(norm)), self._h)
        return out, norm.value

    def close(self):
        if self._h:
            lib().b200pt_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tile_buffer_floats(width, height, tile_world):
    return int(lib().b200pt_tile_buffer_floats(width, height, tile_world))


def linear_to_srgb8(frame):
    """image_io::Write's transfer function (src/utils/image_io.cpp:25-38): sRGB OETF, truncation to 8 bit."""
    f = np.asarray(frame, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        v = np.where(f <= 0.0031308, 12.92 * f, 1.055 * np.power(np.maximum(f, 0.0), 1.0 / 2.4) - 0.055)
    return np.where(v > 1.0, 255.0, v * 255.0).astype(np.int32).astype(np.uint8)


class RayTracer:
    """csrt::RayTracer (include/csrt/ray_tracer.hpp:12-36): owns a Renderer and the frame; Draw(filename) writes a PNG."""

    def __init__(self, scene, width=0, height=0, spp=0, device=-1):
        self.width, self.height, self.spp = width or scene.width, height or scene.height, spp or scene.spp
        self.renderer = Renderer(scene, device=device)
        self.frame = np.zeros((self.height, self.width, 3), dtype=np.float32)

    def Draw(self, output_filename):
        """csrt::RayTracer::Draw (ray_tracer.cpp:155-159): render, then image_io::Write.  A name ending in .pfm writes the
        linear float frame losslessly instead of the reference's 8-bit sRGB PNG."""
        self.renderer.Draw(self.frame, self.width, self.height, self.spp)
        if output_filename.lower().endswith(".pfm"):
            write_pfm(output_filename, self.frame)
        else:
            from PIL import Image
            Image.fromarray(linear_to_srgb8(self.frame)).save(output_filename)
        return self.frame


def write_pfm(path, frame):
    """Portable float map: "PF", size, negative scale = little endian, rows bottom-up, 3 x float32 per pixel."""
    frame = np.ascontiguousarray(frame, dtype="<f4")
    with open(path, "wb") as f:
        f.write(f"PF\n{frame.shape[1]} {frame.shape[0]}\n-1.0\n".encode())
        f.write(frame[::-1].tobytes())


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    return data[::-1].astype(np.float32)
