"""The C restatement (oracle/pt_oracle.c) pinned POINTWISE on the reference's own functions — no GPU needed.

Frames (tests/test_oracle_pinning.py) pin whole paths bit for bit but visit rare branches rarely; here every BSDF /
emitter / medium / texture of the parity scenes is probed at thousands of fixed inputs, grazing angles and
transmission included, through oracle_eval (pt_oracle.c) and ref_eval (oracle/ref_glue.cpp, the csrt::Renderer's own
objects), and rays are traced through oracle_trace and the reference's TLAS::Intersect.  Same compiler, same libm, no
contraction on either side: the results must be BIT-EQUAL.  Needs oracle/_ref/libcsrt_ref_woop.so (built where
/root/reference exists; it travels to the GPU box)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, pack

import refcheck
from test_gpu_pointwise import BSDF_NAMES, bsdf_inputs, counts, unit

pytestmark = []   # CPU only
SCENES = ["matpreview", "volumetric-caustic", "cornell-box", "synthetic_plastic_roughdiffuse", "synthetic_dielectrics_conductor_cylinder",
          "synthetic_envmap_sun_onesided", "synthetic_isotropic_medium_null_surface", "synthetic_bump_bitmap_mesh_disk", "synthetic_opacity_masks"]
EVAL = dict(BSDF_EVALUATE=0, BSDF_SAMPLE=1, EMITTER_SAMPLE=2, EMITTER_DIR=3, MEDIUM_SAMPLE=4, MEDIUM_EVALUATE=5, PHASE_SAMPLE=6, PHASE_EVALUATE=7, TEXTURE=8)


class PackScene:
    def __init__(self, path):
        import __graft_entry__ as ge
        self.scene = ge.load_package().Scene(path)
        self.desc = self.scene.desc


@pytest.fixture(scope="module", params=SCENES)
def both(request, oracle):
    name = request.param
    path = os.path.join(GOLDEN, name + ".b200scene") if name.startswith("synthetic_") else pack(name)
    try:
        ref = refcheck.RefRenderer(refcheck.ref_lib("woop"), path, 8, 8, 1)
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libcsrt_ref_woop.so not built")
    ours = oracle.scene(path)
    yield name, PackScene(path), ours, ref
    ours.close()
    ref.close()


def same(a, b, what):
    bad = np.flatnonzero(np.any(a.view(np.uint32) != b.view(np.uint32), axis=1) & ~np.all(np.isnan(a) == np.isnan(b), axis=1) | np.any((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b)), axis=1))
    assert len(bad) == 0, f"{what}: {len(bad)} of {len(a)} probes differ, first #{bad[0]}: oracle {a[bad[0]]}, reference {b[bad[0]]}"


def test_bsdfs_bit_equal(both):
    name, scene, ours, ref = both
    _, types = counts(scene)
    rng = np.random.RandomState(41)
    for index, kind in enumerate(types):
        if kind not in BSDF_NAMES:
            continue
        inp = bsdf_inputs(3000, rng, both_sides=kind in (5, 6))
        for what in ("BSDF_EVALUATE", "BSDF_SAMPLE"):
            same(ours.eval(EVAL[what], index, inp), ref.eval(EVAL[what], index, inp), f"{name}: {BSDF_NAMES[kind]} #{index} {what}")


def test_emitters_media_textures_bit_equal(both):
    name, scene, ours, ref = both
    d, _ = counts(scene)
    rng = np.random.RandomState(43)
    n = 3000
    for index in range(d.num_emitters):
        inp = np.zeros((n, 32), dtype=np.float32)
        inp[:, 0:3], inp[:, 3:5] = rng.randn(n, 3) * 2.0, rng.rand(n, 2)
        same(ours.eval(EVAL["EMITTER_SAMPLE"], index, inp), ref.eval(EVAL["EMITTER_SAMPLE"], index, inp), f"{name}: emitter #{index} Sample")
        inp[:, 0:3] = unit(rng.randn(n, 3))
        same(ours.eval(EVAL["EMITTER_DIR"], index, inp), ref.eval(EVAL["EMITTER_DIR"], index, inp), f"{name}: emitter #{index} Evaluate / Pdf")
    for index in range(d.num_media):
        inp = np.zeros((n, 32), dtype=np.float32)
        inp[:, 0] = np.where(rng.rand(n) < 0.1, 3.0e38, rng.rand(n) * 6.0)
        inp[:, 18] = rng.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        for what in ("MEDIUM_SAMPLE", "MEDIUM_EVALUATE"):
            same(ours.eval(EVAL[what], index, inp), ref.eval(EVAL[what], index, inp), f"{name}: medium #{index} {what}")
        inp[:, 0:3], inp[:, 3:6] = unit(rng.randn(n, 3)), unit(rng.randn(n, 3))
        for what in ("PHASE_SAMPLE", "PHASE_EVALUATE"):
            same(ours.eval(EVAL[what], index, inp), ref.eval(EVAL[what], index, inp), f"{name}: medium #{index} {what}")
    for index in range(d.num_textures):
        inp = np.zeros((n, 32), dtype=np.float32)
        inp[:, 0:2] = rng.rand(n, 2) * 3.0
        same(ours.eval(EVAL["TEXTURE"], index, inp), ref.eval(EVAL["TEXTURE"], index, inp), f"{name}: texture #{index}")


def test_traversal_and_hit_records_bit_equal(both):
    """TLAS::Intersect / IntersectAny with the scene's BSDFs: distance, side, ids and the whole hit frame (bump maps applied)."""
    name, scene, ours, ref = both
    rng = np.random.RandomState(47)
    n = 6000
    origins = (rng.randn(n, 3) * np.repeat([3.0, 30.0, 300.0], n // 3)[:, None]).astype(np.float32)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = origins
    rays[:, 3:6] = unit(-origins + rng.randn(n, 3) * np.linalg.norm(origins, axis=1, keepdims=True) * 0.2)
    rays[:, 6], rays[:, 7] = 1e-4, 3.0e38
    a, b = ours.trace(rays), ref.trace(rays)
    assert (b["valid"] != 0).sum() > 100, "probe rays miss the scene"
    if name != "synthetic_opacity_masks":   # alpha tests consume the LCG inside traversal: ref.trace starts every ray at seed 0, so does the oracle
        pass
    for field in a.dtype.names:
        fa, fb = a[field], b[field]
        equal = fa.view(np.uint32) == fb.view(np.uint32) if fa.dtype == np.float32 else fa == fb
        assert equal.all(), f"{name}: {field} differs on {np.count_nonzero(~np.all(equal.reshape(n, -1), axis=1))} of {n} rays"
