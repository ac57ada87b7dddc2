"""Pointwise parity of the device leaf functions (-m gpu): the very functions k_shade inlines (shading.cuh), evaluated at
fixed inputs through the test hook b200pt_debug_eval, against the UNMODIFIED reference's Bsdf / Emitter / Medium / Texture
objects called through ref_eval (oracle/ref_glue.cpp) on a csrt::Renderer built from the same scene pack.

Image-level tests cannot see a bias of a few tenths of a percent in a rarely taken branch; these do.  Sampling routines
draw from the reference's LCG on both sides (debug build of the Rng, vecmath.cuh), started from the same seed, and must
leave it in the same state — the NUMBER and ORDER of draws is part of the contract.

Tolerance: `valid` flags equal except within 1e-4 of a threshold the functions cut at (pdf < 0.01, cos < 1e-6 ...:
at most 0.2 % of the probes may disagree); values within 2e-4 relative of the record's magnitude (sinf / cosf / powf / expf
differ by a few ulp between glibc and CUDA, and GGX terms amplify that near the specular peak), directions within 2e-4."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, pack

import refcheck

pytestmark = pytest.mark.gpu

RTOL = 2e-4
BSDF_NAMES = {2: "diffuse", 3: "rough_diffuse", 4: "conductor", 5: "dielectric", 6: "thin_dielectric", 7: "plastic"}
SCENES = ["matpreview", "volumetric-caustic", "mercury", "synthetic_plastic_roughdiffuse", "synthetic_dielectrics_conductor_cylinder",
          "synthetic_envmap_sun_onesided", "synthetic_isotropic_medium_null_surface", "synthetic_bump_bitmap_mesh_disk"]
_CACHE = {}


def setup(pkg, name):
    if name not in _CACHE:
        path = os.path.join(GOLDEN, name + ".b200scene") if name.startswith("synthetic_") else pack(name)
        scene = pkg.Scene(path)
        ours = pkg.Renderer(scene, device=0)
        ref = refcheck.RefRenderer(refcheck.ref_lib("woop"), path, 8, 8, 1)
        _CACHE[name] = (scene, ours, ref)
    return _CACHE[name]


def counts(scene):
    import ctypes
    import scene_builder as sb
    d = ctypes.cast(scene.desc, ctypes.POINTER(sb.SceneDesc)).contents
    bsdfs = ctypes.cast(d.bsdfs, ctypes.POINTER(sb.Bsdf))
    return d, [int(bsdfs[i].type) for i in range(d.num_bsdfs)]


def unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def frames(n, rng):
    normal = unit(rng.randn(n, 3))
    helper = unit(rng.randn(n, 3))
    tangent = unit(np.cross(helper, normal))
    bitangent = np.cross(normal, tangent)
    return normal.astype(np.float32), tangent.astype(np.float32), bitangent.astype(np.float32)


def hemisphere(normal, tangent, bitangent, rng, grazing=0.15):
    """Directions on the normal's side; a share of them at grazing angles (cos in [0, 0.05])."""
    n = len(normal)
    cos = np.where(rng.rand(n) < grazing, rng.rand(n) * 0.05, rng.rand(n))
    phi = rng.rand(n) * 2 * np.pi
    sin = np.sqrt(1 - cos * cos)
    v = (sin * np.cos(phi))[:, None] * tangent + (sin * np.sin(phi))[:, None] * bitangent + cos[:, None] * normal
    return unit(v).astype(np.float32)


def bsdf_inputs(n, rng, both_sides):
    inp = np.zeros((n, 32), dtype=np.float32)
    normal, tangent, bitangent = frames(n, rng)
    wo = hemisphere(normal, tangent, bitangent, rng)
    arriving = hemisphere(normal, tangent, bitangent, rng)  # -wi: from where the light comes
    if both_sides:  # transmission: light arriving from below the surface on half of the probes
        flip = rng.rand(n) < 0.5
        arriving[flip] = hemisphere(-normal[flip], tangent[flip], bitangent[flip], rng)
    inp[:, 0:3], inp[:, 3:6], inp[:, 6:9], inp[:, 9:12], inp[:, 12:15] = -arriving, wo, normal, tangent, bitangent
    inp[:, 15:17] = rng.rand(n, 2) * 3.0
    inp[:, 17] = rng.rand(n) < 0.5
    inp[:, 18] = rng.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    return inp


def compare(a, b, value_cols, what, dir_cols=None, flag_cols=(0,), scale_cols=None, max_share=0.002):
    """a = ours, b = reference: [n, 16] outputs.  max_share: probes that may differ (float-conditioning cases of the reference itself)."""
    flags_equal = np.all(a[:, list(flag_cols)] == b[:, list(flag_cols)], axis=1)
    assert flags_equal.mean() >= 1.0 - max_share, f"{what}: valid flags differ on {(~flags_equal).sum()} of {len(a)} probes"
    both = flags_equal & (b[:, 0] != 0) if 0 in flag_cols else flags_equal
    if not both.any():
        return
    va, vb = a[both][:, value_cols].astype(np.float64), b[both][:, value_cols].astype(np.float64)
    assert np.isfinite(va).all() or not np.isfinite(vb).all(), f"{what}: non-finite values"
    scale = np.maximum(np.abs(vb).max(axis=1, keepdims=True), 1e-3)
    err = np.abs(va - vb) / scale
    bad = err.max(axis=1) > RTOL
    assert bad.mean() <= max_share, f"{what}: {bad.sum()} of {both.sum()} probes off by more than {RTOL} (worst {err.max():.3e}: ours {va[err.max(axis=1).argmax()]}, reference {vb[err.max(axis=1).argmax()]})"
    if dir_cols is not None:
        da, db = a[both][:, dir_cols].astype(np.float64), b[both][:, dir_cols].astype(np.float64)
        off = np.abs(da - db).max(axis=1) > RTOL
        assert off.mean() <= max_share, f"{what}: {off.sum()} sampled directions differ (worst {np.abs(da - db).max():.3e})"


@pytest.mark.parametrize("scene_name", SCENES)
def test_bsdf_evaluate_and_sample(pkg, scene_name):
    """Bsdf::Evaluate / Bsdf::Sample (bsdfs/*.cpp) of every BSDF of the scene: attenuation, pdf, validity, sampled direction,
    and the LCG state after sampling."""
    scene, ours, ref = setup(pkg, scene_name)
    _, types = counts(scene)
    rng = np.random.RandomState(17)
    tested = 0
    for index, kind in enumerate(types):
        if kind not in BSDF_NAMES:
            continue
        inp = bsdf_inputs(6000, rng, both_sides=kind in (5, 6))
        what = f"{scene_name}: {BSDF_NAMES[kind]} #{index}"
        # 200 000 probes per BSDF (tools/pointwise_report.py, profiles/r02_pointwise_report.log): no probe with another number of draws,
        # another validity or another direction, 2 values off by more than 2e-4 -> at most 3 of these 6 000 may differ
        compare(ours.debug_eval(pkg.EVAL_BSDF_EVALUATE, index, inp), ref.eval(pkg.EVAL_BSDF_EVALUATE, index, inp), [1, 2, 3, 4], what + " Evaluate", max_share=0.0005)
        a, b = ours.debug_eval(pkg.EVAL_BSDF_SAMPLE, index, inp), ref.eval(pkg.EVAL_BSDF_SAMPLE, index, inp)
        assert np.array_equal(a[:, 15].view(np.uint32), b[:, 15].view(np.uint32)), f"{what} Sample: number of random draws differs"
        compare(a, b, [1, 2, 3, 4], what + " Sample", dir_cols=[5, 6, 7], max_share=0.0005)
        tested += 1
    if tested == 0:
        pytest.skip("no scattering BSDF in this scene")


@pytest.mark.parametrize("scene_name", SCENES)
def test_emitters(pkg, scene_name):
    """Emitter::Sample / Evaluate / Pdf (emitters/*.cpp), env-map tables with the reference's wiring (Q9) included."""
    scene, ours, ref = setup(pkg, scene_name)
    d, _ = counts(scene)
    if d.num_emitters == 0:
        pytest.skip("no emitter list in this scene (area lights only)")
    rng = np.random.RandomState(23)
    for index in range(d.num_emitters):
        inp = np.zeros((8000, 32), dtype=np.float32)
        inp[:, 0:3] = rng.randn(8000, 3) * 2.0
        inp[:, 3:5] = rng.rand(8000, 2)
        a, b = ours.debug_eval(pkg.EVAL_EMITTER_SAMPLE, index, inp), ref.eval(pkg.EVAL_EMITTER_SAMPLE, index, inp)
        what = f"{scene_name}: emitter #{index}"
        compare(a, b, [2, 6, 7, 8, 9], what + " Sample", dir_cols=[3, 4, 5], flag_cols=(0, 1))
        inp = np.zeros((8000, 32), dtype=np.float32)
        inp[:, 0:3] = unit(rng.randn(8000, 3))
        a, b = ours.debug_eval(pkg.EVAL_EMITTER_DIR, index, inp), ref.eval(pkg.EVAL_EMITTER_DIR, index, inp)
        compare(a, b, [0, 1, 2, 3], what + " Evaluate(dir) / Pdf(dir)", flag_cols=())


@pytest.mark.parametrize("scene_name", ["volumetric-caustic", "synthetic_isotropic_medium_null_surface"])
def test_media_and_phase_functions(pkg, scene_name):
    """Medium::Sample / Evaluate (homogeneous.cpp, Q10) and SamplePhase / EvaluatePhase (isotropic.cpp, henyey_greenstein.cpp)."""
    scene, ours, ref = setup(pkg, scene_name)
    d, _ = counts(scene)
    assert d.num_media > 0
    rng = np.random.RandomState(29)
    n = 8000
    for index in range(d.num_media):
        what = f"{scene_name}: medium #{index}"
        inp = np.zeros((n, 32), dtype=np.float32)
        inp[:, 0] = np.where(rng.rand(n) < 0.1, 3.0e38, rng.rand(n) * 6.0)   # max_distance, FLT_MAX-like for escaping segments
        inp[:, 18] = rng.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        a, b = ours.debug_eval(pkg.EVAL_MEDIUM_SAMPLE, index, inp), ref.eval(pkg.EVAL_MEDIUM_SAMPLE, index, inp)
        assert np.array_equal(a[:, 15].view(np.uint32), b[:, 15].view(np.uint32)), f"{what} Sample: number of random draws differs"
        compare(a, b, [2, 3, 4, 5, 6], what + " Sample", flag_cols=(0, 1))
        a, b = ours.debug_eval(pkg.EVAL_MEDIUM_EVALUATE, index, inp), ref.eval(pkg.EVAL_MEDIUM_EVALUATE, index, inp)
        compare(a, b, [2, 4, 5, 6], what + " Evaluate", flag_cols=(0, 1))
        inp[:, 0:3], inp[:, 3:6] = unit(rng.randn(n, 3)), unit(rng.randn(n, 3))
        a, b = ours.debug_eval(pkg.EVAL_PHASE_SAMPLE, index, inp), ref.eval(pkg.EVAL_PHASE_SAMPLE, index, inp)
        assert np.array_equal(a[:, 15].view(np.uint32), b[:, 15].view(np.uint32)), f"{what} SamplePhase: number of random draws differs"
        compare(a, b, [1, 2, 3, 4], what + " SamplePhase", dir_cols=[5, 6, 7])
        a, b = ours.debug_eval(pkg.EVAL_PHASE_EVALUATE, index, inp), ref.eval(pkg.EVAL_PHASE_EVALUATE, index, inp)
        compare(a, b, [1, 2, 3, 4], what + " EvaluatePhase")


@pytest.mark.parametrize("scene_name", ["mercury", "matpreview", "synthetic_bump_bitmap_mesh_disk", "synthetic_opacity_masks"])
def test_textures(pkg, scene_name):
    """Texture::GetColor (bitmap.cpp:6-56 manual bilinear with wrap, checkboard.cpp:6-21, constant_texture.cpp:6)."""
    scene, ours, ref = setup(pkg, scene_name)
    d, _ = counts(scene)
    rng = np.random.RandomState(31)
    for index in range(d.num_textures):
        inp = np.zeros((6000, 32), dtype=np.float32)
        inp[:, 0:2] = rng.rand(6000, 2) * 3.0
        inp[:200, 0:2] = np.float32(np.round(inp[:200, 0:2] * 4) / 4)  # texel / tile borders
        a, b = ours.debug_eval(pkg.EVAL_TEXTURE, index, inp), ref.eval(pkg.EVAL_TEXTURE, index, inp)
        compare(a, b, [0, 1, 2], f"{scene_name}: texture #{index}", flag_cols=())


@pytest.mark.parametrize("scene_name", ["matpreview", "mercury", "synthetic_bump_bitmap_mesh_disk", "synthetic_dielectrics_conductor_cylinder", "dragon"])
def test_hit_attributes(pkg, scene_name):
    """The hit record the shading stage rebuilds from (primitive, u, v) — SurfTriangle / SurfAnalytic — against the csrt::Hit
    that Primitive::Intersect fills (triangle.cpp:115-148, sphere.cpp:46-88, disk.cpp:42-110, cylinder.cpp:51-90):
    position, shading normal (bump map applied, flipped towards the ray), tangent frame, texture coordinate, side."""
    scene, ours, ref = setup(pkg, scene_name)
    rng = np.random.RandomState(37)
    n = 30000
    origins = (rng.randn(n, 3) * np.repeat([3.0, 30.0, 300.0], n // 3)[:, None]).astype(np.float32)  # scenes come in very different scales
    far = np.zeros((n, 8), dtype=np.float32)
    far[:, 0:3] = origins
    far[:, 3:6] = unit(-origins + rng.randn(n, 3) * 0.5)
    far[:, 6], far[:, 7] = 1e-4, 3.0e38
    t, prim, uv = ours.debug_trace(far, raw_prim=True)
    hit = prim != 0xFFFFFFFF
    assert hit.sum() > 200, "probe rays miss the scene"
    rays = far[hit]
    href = ref.trace(rays)
    missed = href["valid"] == 0
    assert missed.mean() <= 0.001, f"the reference misses {missed.sum()} of {len(rays)} rays we hit (our prims {np.unique(prim[hit][missed] >> 28)})"
    keep = ~missed
    rays, href, hit = rays[keep], href[keep], np.flatnonzero(hit)[keep]
    rel = np.abs(t[hit] / href["t"] - 1)   # analytic primitives seen from far away: an ill-conditioned quadratic (test_gpu_traversal.py)
    assert np.median(rel) < 2e-6 and rel.max() < 5e-2, f"t off by median {np.median(rel):.3e}, max {rel.max():.3e}"
    near = rel < 1e-5                      # the attributes below are compared where both sides stand on the same point
    rays, href, hit = rays[near], href[near], hit[near]
    inp = np.zeros((len(rays), 32), dtype=np.float32)
    inp[:, 0] = prim[hit].view(np.float32)
    inp[:, 1:3] = uv[hit]
    inp[:, 3:9] = rays[:, 0:6]
    inp[:, 9] = t[hit]
    out = ours.debug_eval(pkg.EVAL_SURFACE, 0, inp)
    # ours is o + t d with the t both sides agree on to 1e-5 (the filter above): the bar scales with the distance travelled
    scale = np.maximum(np.abs(href["position"]).max(), 1.0)
    err = np.abs(out[:, 0:3] - href["position"]).max(axis=1)
    assert (err <= 2e-5 * scale + 1.5e-5 * t[hit]).all(), f"hit position: worst {err.max():.3e} (scene scale {scale:.1f}, t up to {t[hit].max():.1f})"
    import ctypes
    import scene_builder as sb
    d, _ = counts(scene)
    inst_type = np.array([ctypes.cast(d.instances, ctypes.POINTER(sb.Instance))[i].type for i in range(d.num_instances)])
    kind = inst_type[href["id_instance"]]
    # two surfaces at one distance (a rectangle standing on the stage): either answer is right, the rest is compared elsewhere
    other_instance = out[:, 15].view(np.uint32) != href["id_instance"]
    assert other_instance.mean() <= 0.002, f"hit instance differs on {other_instance.sum()} of {len(other_instance)} hits"
    # `inside` of a sphere / cylinder is the sign of c = |o|^2 - r^2 in float (sphere.cpp:49, cylinder.cpp:61): origins within
    # rounding of the surface may fall on either side of it with another contraction of the products
    side_differs = (out[:, 14] != 0) != (href["inside"] != 0)
    assert side_differs.mean() <= 0.002, f"hit side differs on {side_differs.sum()} of {len(side_differs)} hits (instance types {np.unique(kind[side_differs])})"
    # a bump map perturbs the normal by texture differences over 1e-4-wide steps (bsdf.cpp:238-254): float noise is amplified
    # spheres, disks and cylinders rebuild the frame from a hit point that both sides know to 1e-5 of the distance travelled
    # (hundreds of units here), and the sphere's bitangent is a finite difference over a 0.03-radian arc (sphere.cpp:58-67)
    # -> the frame of a unit-sized primitive is known to ~ 1e-5 t / radius
    analytic = np.isin(kind, [sb.INST_SPHERE, sb.INST_DISK, sb.INST_CYLINDER])
    tol = np.where(analytic, 3e-3 + 4e-5 * t[hit], 3e-3 if "bump" in scene_name else 2e-4)
    # The tangent frame of a DISK hangs on phi = atan2(z, x) of a point whose local z is rounding noise around 0
    # (disk.cpp:38-60 with the y-up CartesianToSpherical, math.cpp:102-119): phi jumps between 0 / pi / -pi with the sign of that
    # noise, and `flip_tangent = phi' > pi` with it.  The frame is compared up to that sign there; normals are not affected.
    disk = kind == sb.INST_DISK
    same_side = ~side_differs & ~other_instance
    for cols, field in (((3, 6), "normal"), ((6, 9), "tangent"), ((9, 12), "bitangent")):
        a, b = out[:, cols[0]:cols[1]], href[field]
        err = np.abs(a - b).max(axis=1)
        if field != "normal":
            err = np.where(disk, np.minimum(err, np.abs(a + b).max(axis=1)), err)
        bad = (err > tol)[same_side]
        assert bad.mean() <= 0.002, f"{field}: {bad.sum()} of {len(bad)} differ (worst {err[same_side].max():.3e}; instance types {np.unique(kind[same_side][bad])})"
    err = np.abs(out[:, 12:14] - href["texcoord"])
    err = np.where(analytic[:, None], np.minimum(err, np.abs(1.0 - err)), err).max(axis=1)[same_side]  # phi / 2 pi wraps at the seam
    uv_tol = np.where(analytic, 2e-5 + 1e-5 * t[hit], 2e-5)[same_side]
    assert (err > uv_tol).mean() <= 0.002, f"texcoord: {(err > uv_tol).sum()} of {len(err)} differ, worst {err.max():.3e}"
