"""Pointwise parity of the traversal kernels (-m gpu): caller-supplied rays through b200pt_debug_trace, i.e. through the
same persistent loops k_primary / k_trace run, against the UNMODIFIED reference's csrt::Scene + TLAS::Intersect /
IntersectAny (tlas.cpp:13-76, blas.cpp:18-77, triangle.cpp:23-87 Woop variant) called through oracle/ref_glue.cpp.

Bar: hit / miss identical, distance t BIT-EXACT on triangles (same Woop arithmetic on the same world-space vertices;
spheres, disks and cylinders measure t with square roots and products that nvcc and g++ contract differently: 1e-5
relative there), hit side identical, hit instance identical — for random rays and for rays aimed exactly at shared
vertices, shared edges and degenerate triangles.  The primitive may differ only where two triangles report the same t (a ray through a shared
edge: the winner depends on the order of the tests, which differs between trees).  The binary tree (default),
the compressed 8-wide tree (B200PT_CREATE_BVH8) and both loop shapes (persistent / one ray per lane) must all agree."""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN, pack

import refcheck
import scene_builder as sb

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(3.4028234663852886e38)


def make_rays(origins, dirs, tmin=1e-4, tmax=FLT_MAX):
    origins, dirs = np.asarray(origins, dtype=np.float32), np.asarray(dirs, dtype=np.float32)
    dirs = dirs / np.linalg.norm(dirs.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    rays = np.zeros((len(origins), 8), dtype=np.float32)
    rays[:, 0:3], rays[:, 3:6], rays[:, 6], rays[:, 7] = origins, dirs, tmin, tmax
    return rays


def instance_table(desc_ptr):
    """Per instance of a b200pt_scene_desc: (first triangle in the scene's triangle enumeration, triangle count, analytic rank)."""
    desc = ctypes.cast(desc_ptr, ctypes.POINTER(sb.SceneDesc)).contents
    inst = ctypes.cast(desc.instances, ctypes.POINTER(sb.Instance))
    rows, tri, analytic = [], 0, 0
    for i in range(desc.num_instances):
        kind = inst[i].type
        count = {sb.INST_CUBE: 12, sb.INST_RECTANGLE: 2, sb.INST_MESHES: int(inst[i].num_triangles)}.get(kind, 0)
        rows.append((tri, count, analytic if count == 0 else -1))
        tri += count
        analytic += 1 if count == 0 else 0
    return rows


def ours_to_instance(prim, table):
    """prim of b200pt_debug_hit -> (instance, local primitive index)."""
    firsts = np.array([r[0] for r in table] + [1 << 62])
    counts = np.array([r[1] for r in table])
    analytic_inst = {r[2]: i for i, r in enumerate(table) if r[2] >= 0}
    inst = np.full(len(prim), -1, dtype=np.int64)
    local = np.full(len(prim), -1, dtype=np.int64)
    for k, p in enumerate(prim):
        if p == 0xFFFFFFFF:
            continue
        if p & 0x80000000:
            inst[k], local[k] = analytic_inst[int(p & 0x0FFFFFFF)], 0
        else:
            t = int(p & 0x0FFFFFFF)
            i = int(np.searchsorted(firsts, t, side="right") - 1)
            while counts[i] == 0:  # analytic instances own no triangles: same `first` as the next mesh
                i += 1
            inst[k], local[k] = i, t - firsts[i]
    return inst, local


def check_against_reference(r, tracer, table, rays, name, allow_prim_ties, exact_targets=None, expected_t=None):
    ref = tracer.trace(rays)
    ref_any = tracer.trace(rays, any_hit=True)
    for per_lane in (False, True):
        t, prim, uv = r.debug_trace(rays, per_lane_loop=per_lane)
        hit = prim != 0xFFFFFFFF
        ref_hit = ref["valid"] != 0
        if exact_targets is None:
            assert np.array_equal(hit, ref_hit), f"{name}: hit/miss differs on {(hit != ref_hit).sum()} rays, first {np.flatnonzero(hit != ref_hit)[:5]}"
        else:
            # Rays aimed EXACTLY at shared vertices / edges graze the boxes of the reference's own BVH, whose slab test is not
            # conservative (aabb.cpp:29-48): the reference itself leaks there.  Ours must not: every such ray hits, at the
            # distance of its target; wherever the reference hits too, the distances are compared below like everywhere else.
            # conservative (aabb.cpp:29-48: enter <= exit on rounded products): the reference itself leaks ~5 % of them, and so
            # does any BVH whose boxes end exactly on the lattice lines — ours included (same test, other rounding).  The
            # watertightness Woop's test guarantees is per TRIANGLE PAIR, and that is what must hold: whenever the boxes let
            # the ray through to the triangles it is hit, at the distance of its target.  Leaks are reported and bounded.
            # (a ray that slips through the first sheet and meets the second one has leaked just the same)
            through = lambda h, tt: exact_targets & (~h | (tt > expected_t * np.float32(1.001)))
            leaks, ref_leaks = through(hit, t), through(ref_hit, ref["t"])
            assert leaks.sum() <= max(2 * ref_leaks.sum(), 0.08 * exact_targets.sum()), \
                f"{name}: {leaks.sum()} of {exact_targets.sum()} rays leaked at a shared vertex / edge (the reference leaks {ref_leaks.sum()})"
            found = exact_targets & hit & ~leaks
            # the target's distance, to what float Woop arithmetic gives at grazing incidence (the eye 2 degrees above the sheet
            # loses a factor 1 / sin(2 deg) = 28); the second sheet would be off by percents.  Where the reference hits too the
            # BITS are compared below.
            worst = np.abs(t[found] / expected_t[found] - 1).max()
            assert worst < 1e-4, f"{name}: wrong distance on an exact vertex / edge hit (worst {worst:.3e} relative)"
            assert np.array_equal(hit[~exact_targets], ref_hit[~exact_targets]), f"{name}: hit/miss differs off the lattice lines"
        neither = ~hit & ~ref_hit
        if exact_targets is not None:
            hit = hit & ref_hit & ~leaks & ~ref_leaks  # the distances below are compared where both sides met the first sheet
        tri = hit & ((prim & 0x80000000) == 0)
        analytic = hit & ~tri
        same_bits = t.view(np.uint32) == ref["t"].view(np.uint32)
        inst, local = ours_to_instance(prim, table)
        same_prim = (inst == ref["id_instance"].astype(np.int64)) & (local == ref["id_primitive"].astype(np.int64))
        same_prim |= ~hit
        # Same triangle => same Woop arithmetic on the same vertices => the SAME BITS.  Two coplanar triangles (a box standing
        # on the floor) or a shared edge answer with distances one ulp apart; `t <= t_max` lets the one tested last win, and
        # the order of the tests differs between trees: there the other triangle, 2 ulp away at most, is as right.
        check = (tri & same_prim) | neither
        assert same_bits[check].all(), (f"{name}: t differs from TLAS::Intersect on {(~same_bits[check]).sum()} of {check.sum()} rays that hit the same triangle "
                                        f"(worst {np.abs(t[check] / ref['t'][check] - 1).max():.3e} relative, first {np.flatnonzero(check & ~same_bits)[:5]})")
        ties = tri & ~same_prim
        assert ties.mean() <= (1.0 if allow_prim_ties else 0.01), \
            f"{name}: another triangle on {ties.sum()} of {len(rays)} rays (instance differs on {(ties & (inst != ref['id_instance'])).sum()}); " \
            f"first ours {list(zip(inst[ties][:4], local[ties][:4]))} vs reference {list(zip(ref['id_instance'][ties][:4], ref['id_primitive'][ties][:4]))}"
        if ties.any():
            # (1e-6 of the distance or of the scene's own scale: a ray that starts a millimetre above two coplanar faces)
            off = ties & (np.abs(t - ref["t"]) >= 1e-6 * np.maximum(np.maximum(np.abs(t), np.abs(rays[:, 0:3]).max(axis=1)), 1.0))
            # a silhouette ray of a sphere / cylinder (discriminant or cap test within rounding of zero) may pass on one side and
            # hit on the other, and then meets whatever lies behind: a handful in 10^5, never a pattern
            ref_analytic = np.array([table[i][2] >= 0 for i in ref["id_instance"].clip(0, len(table) - 1)]) & ref_hit
            silhouette = off & ref_analytic
            assert silhouette.sum() <= max(2, 1e-4 * len(rays)), f"{name}: {silhouette.sum()} rays miss an analytic primitive the reference hits"
            off &= ~silhouette
            first = [(rays[k].tolist(), float(t[k]), int(inst[k]), int(local[k]), float(ref["t"][k]), int(ref["id_instance"][k]), int(ref["id_primitive"][k]))
                     for k in np.flatnonzero(off)[:3]]
            assert not off.any(), f"{name} per_lane={per_lane}: a different triangle AND a different distance on {off.sum()} rays: (ray, t, inst, prim, ref t, ref inst, ref prim) {first}"
        if analytic.any():
            # sphere / disk / cylinder distances come out of a quadratic in ray-origin coordinates: for origins hundreds of radii
            # away it is ill-conditioned in float (c = |o|^2 - r^2), and nvcc's fma contraction rounds it differently from g++
            rel = np.abs(t[analytic] / ref["t"][analytic] - 1)
            assert np.median(rel) < 2e-6 and rel.max() < 5e-3, f"{name}: analytic t off by median {np.median(rel):.3e}, max {rel.max():.3e}"
        inside = (prim & 0x40000000) != 0
        assert np.array_equal(inside[tri & same_prim], (ref["inside"] != 0)[tri & same_prim]), f"{name}: hit side differs"
        occluded = r.debug_trace(rays, any_hit=True, per_lane_loop=per_lane)[1] == 0
        if exact_targets is None:
            assert np.array_equal(occluded, ref_any["valid"] != 0), f"{name}: IntersectAny differs on {(occluded != (ref_any['valid'] != 0)).sum()} rays"
        else:
            # the same box-grazing leaks as above (a ray along the sheet's outer edge also runs past the second sheet)
            any_leaks, ref_any_leaks = exact_targets & ~occluded, exact_targets & (ref_any["valid"] == 0)
            assert any_leaks.sum() <= max(2 * ref_any_leaks.sum(), 0.08 * exact_targets.sum()), \
                f"{name}: IntersectAny leaks {any_leaks.sum()} of {exact_targets.sum()} aimed rays (the reference {ref_any_leaks.sum()})"
            assert np.array_equal(occluded[~exact_targets], (ref_any["valid"] != 0)[~exact_targets])


def random_rays(lo, hi, n, rng):
    span = hi - lo
    origins = lo - 0.3 * span + rng.rand(n, 3) * 1.6 * span
    targets = lo + rng.rand(n, 3) * span
    return make_rays(origins, targets - origins)


@pytest.mark.parametrize("scene", ["cornell-box", "dragon", "matpreview", "volumetric-caustic", "mercury", "synthetic_bump_bitmap_mesh_disk",
                                   "synthetic_dielectrics_conductor_cylinder"])
def test_random_rays_match_reference_tlas(pkg, scene):
    path = os.path.join(GOLDEN, scene + ".b200scene") if scene.startswith("synthetic_") else pack(scene)
    sc = pkg.Scene(path)
    tracer = refcheck.RefTracer(refcheck.ref_lib("woop"), path)
    table = instance_table(sc.desc)
    rng = np.random.RandomState(3)
    # the extent of the geometry: probe with axis rays is overkill — use a generous cube around the camera target instead
    probe = make_rays(rng.randn(4096, 3) * 1e3, rng.randn(4096, 3))
    n = 200000 if scene == "dragon" else 60000
    for flags in (0, pkg.CREATE_BVH8):
        r = pkg.Renderer(sc, device=0, flags=flags)
        has_triangles = r.stats()["num_triangles"] > 0  # mercury is a sphere and a disk: no tree of either kind
        assert r.stats()["bvh_width"] == (8 if flags and has_triangles else 2)
        # bounding box of what random probes from far away hit
        far = make_rays(rng.randn(20000, 3) * 50.0, rng.randn(20000, 3))
        far[:, 3:6] = -far[:, 0:3] / np.linalg.norm(far[:, 0:3], axis=1, keepdims=True)  # towards the origin
        t, prim, _ = r.debug_trace(far)
        pts = far[:, 0:3] + t[:, None] * far[:, 3:6]
        pts = pts[prim != 0xFFFFFFFF]
        lo, hi = (pts.min(axis=0), pts.max(axis=0)) if len(pts) > 16 else (np.full(3, -5.0), np.full(3, 5.0))
        rays = np.concatenate([random_rays(lo, hi, n, rng), probe, far[:2000]])
        rays[::5, 7] = np.float32(0.5) * np.linalg.norm(hi - lo)  # bounded segments (shadow-ray style) on a fifth of the rays
        check_against_reference(r, tracer, table, rays, f"{scene}/bvh{8 if flags else 2}", allow_prim_ties=False)
        r.close()
    tracer.close()


def lattice_scene():
    """A 12 x 12 grid of quads (shared vertices and edges at exactly representable coordinates), a second sheet behind it,
    plus zero-area and needle triangles."""
    b = sb.SceneBuilder()
    n = 12
    xs = np.arange(n + 1, dtype=np.float32) * 0.25 - 1.5
    pos = np.array([(x, y, 0.0) for y in xs for x in xs], dtype=np.float32)
    idx = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            idx += [(a, a + 1, a + n + 2), (a, a + n + 2, a + n + 1)]
    white = b.diffuse(b.constant(0.7))
    b.mesh(white, pos, idx)
    b.mesh(white, pos + np.float32([0.125, 0.0625, -0.75]), idx)
    # zero-area and needle triangles BEHIND the sheets (z < -0.75 is hidden from the front eyes, in view from the back one)
    degenerate = np.array([(0, 0, -1.5), (0, 0, -1.5), (1, 1, -1.5),          # two coincident vertices
                           (-1, -1, -1.5), (0, 0, -1.5), (1, 1, -1.5),        # collinear
                           (-1, 0.5, -1.5), (1, 0.5, -1.5), (0, 0.5 + 1e-6, -1.5)], dtype=np.float32)  # needle
    b.mesh(white, degenerate, [(0, 1, 2), (3, 4, 5), (6, 7, 8)])
    b.directional((0, 0, -1), (1, 1, 1))
    return b, xs


def test_rays_through_shared_vertices_edges_and_degenerate_triangles(pkg, tmp_path):
    """Woop's watertight test with its double-precision edge fallback (triangle.cpp:50-63): rays aimed EXACTLY at lattice
    vertices and edge midpoints must be hit (no leaks between neighbours) with the reference's t, from both trees."""
    b, xs = lattice_scene()
    desc = b.desc()
    path = str(tmp_path / "lattice.b200scene")
    assert pkg.lib().b200pt_scene_save(ctypes.byref(desc), path.encode()) == 0
    sc = pkg.Scene(path)
    tracer = refcheck.RefTracer(refcheck.ref_lib("woop"), path)
    table = instance_table(sc.desc)
    rng = np.random.RandomState(11)
    mids = (xs[:-1] + xs[1:]) * np.float32(0.5)
    targets = [(x, y, 0.0) for x in xs for y in xs] + [(x, y, 0.0) for x in mids for y in xs] + [(x, y, 0.0) for x in xs for y in mids] \
        + [(x, y, 0.0) for x in mids for y in mids]  # vertices, both edge families, diagonals' midpoints
    targets = np.array(targets, dtype=np.float32)
    origins, aims = [], []
    for eye in ([0, 0, 4], [0.3, -0.2, 3], [2.5, 1.0, 2.0], [0, 0, -4], [1e-3, 7.0, 0.25]):
        origins.append(np.tile(np.float32(eye), (len(targets), 1)))
        aims.append(targets)
    # axis-parallel rays straight down the lattice lines (zero direction components: ray.cpp:21-22)
    origins.append(targets + np.float32([0, 0, 5]))
    aims.append(targets)
    origins, aims = np.concatenate(origins), np.concatenate(aims)
    rays = np.concatenate([make_rays(origins, aims - origins), random_rays(np.float32([-1.6, -1.6, -0.8]), np.float32([1.6, 1.6, 0.6]), 20000, rng)])
    # the first sheet (z = 0) is the closest surface for every aimed ray but those from behind, which meet the second sheet first
    exact = np.zeros(len(rays), dtype=bool)
    front = origins[:, 2] > 0
    exact[: len(origins)] = front
    expected = np.zeros(len(rays), dtype=np.float32)
    expected[: len(origins)] = np.linalg.norm((aims - origins).astype(np.float64), axis=1)
    for flags in (0, pkg.CREATE_BVH8):
        r = pkg.Renderer(sc, device=0, flags=flags)
        check_against_reference(r, tracer, table, rays, f"lattice/bvh{8 if flags else 2}", allow_prim_ties=True, exact_targets=exact, expected_t=expected)
        r.close()
    tracer.close()


@pytest.mark.parametrize("scene,w,h,spp", [("dragon", 192, 192, 16), ("cornell-box", 96, 96, 32), ("matpreview", 96, 96, 16),
                                           ("volumetric-caustic", 64, 64, 32), ("synthetic_opacity_masks", 64, 64, 32)])
def test_wide_and_binary_trees_render_the_same_frame(pkg, scene, w, h, spp):
    """Same hits => same paths: the frames of the two layouts agree except where a tie on a shared edge picked the neighbour."""
    path = os.path.join(GOLDEN, scene + ".b200scene") if scene.startswith("synthetic_") else pack(scene)
    sc = pkg.Scene(path)
    frames = []
    for flags in (0, pkg.CREATE_BVH8):
        r = pkg.Renderer(sc, device=0, flags=flags, max_paths_in_flight=1 << 22)
        frames.append(r.Draw(width=w, height=h, spp=spp, seed=5))
        r.close()
    a, b = frames
    differing = np.any(a != b, axis=2).mean()
    assert differing < 0.02, f"{scene}: {differing:.4f} of the pixels differ between the wide and the binary tree"
    assert abs(a.mean() / b.mean() - 1.0) < 2e-3


@pytest.mark.parametrize("scene", ["dragon", "cornell-box", "matpreview", "synthetic_dielectrics_conductor_cylinder"])
def test_warp_packets_find_the_same_hits(pkg, scene):
    """traverse_packet.cuh (camera rays and first-vertex NEE rays on the binary tree): 32 consecutive rays walk the tree with one
    shared node pointer and stack, each lane testing its own ray.  Same closest hit as the per-ray loop BIT FOR BIT (a tie
    between two triangles at one t may name the other triangle), same occlusion verdict — for coherent bundles (a pinhole
    camera, parallel shadow rays) AND for incoherent rays, where packets are slow but must still be right."""
    path = os.path.join(GOLDEN, scene + ".b200scene") if scene.startswith("synthetic_") else pack(scene)
    sc = pkg.Scene(path)
    r = pkg.Renderer(sc, device=0)
    rng = np.random.RandomState(17)
    far = make_rays(rng.randn(20000, 3) * 50.0, rng.randn(20000, 3))
    far[:, 3:6] = -far[:, 0:3] / np.linalg.norm(far[:, 0:3], axis=1, keepdims=True)
    t, prim, _ = r.debug_trace(far)
    pts = (far[:, 0:3] + t[:, None] * far[:, 3:6])[prim != 0xFFFFFFFF]
    lo, hi = (pts.min(axis=0), pts.max(axis=0)) if len(pts) > 16 else (np.full(3, -5.0), np.full(3, 5.0))
    centre, size = (lo + hi) / 2, np.linalg.norm(hi - lo)
    # a pinhole camera: 96 x 96 pixels x 32 sub-pixel samples, consecutive rays = one pixel
    eye = centre + np.float32([0.3, 0.5, 1.0]) * size
    px = np.stack(np.meshgrid(np.arange(96), np.arange(96), indexing="ij"), axis=-1).reshape(-1, 1, 2) + rng.rand(96 * 96, 32, 2)
    u, v = (px[..., 0] / 96 - 0.5).ravel(), (px[..., 1] / 96 - 0.5).ravel()
    front = (centre - eye) / np.linalg.norm(centre - eye)
    right = np.cross(front, [0, 1, 0]); right /= np.linalg.norm(right)
    up = np.cross(right, front)
    camera = make_rays(np.tile(eye, (len(u), 1)), front + 0.9 * (u[:, None] * right + v[:, None] * up))
    # parallel shadow rays from the camera's hit points towards one light, in camera order
    tc, pc, _ = r.debug_trace(camera)
    hit = pc != 0xFFFFFFFF
    origins = camera[hit, 0:3] + (tc[hit] * np.float32(0.999))[:, None] * camera[hit, 3:6]
    shadow = make_rays(origins, np.tile(np.float32([0.4, 1.0, 0.2]), (len(origins), 1)), tmin=1e-4, tmax=np.float32(4.0) * size)
    incoherent = random_rays(lo, hi, 40000, rng)
    for name, rays in (("camera", camera), ("shadow", shadow), ("incoherent", incoherent)):
        t0, p0, uv0 = r.debug_trace(rays)
        t1, p1, uv1 = r.debug_trace(rays, packet_loop=True)
        assert np.array_equal(p0 == 0xFFFFFFFF, p1 == 0xFFFFFFFF), f"{scene}/{name}: hit / miss differs on {((p0 == 0xFFFFFFFF) != (p1 == 0xFFFFFFFF)).sum()} rays"
        same = p0 == p1
        assert np.array_equal(t0[same].view(np.uint32), t1[same].view(np.uint32)) and np.array_equal(uv0[same], uv1[same]), f"{scene}/{name}: same primitive, other t"
        assert (~same).mean() < 0.01, f"{scene}/{name}: {(~same).sum()} of {len(rays)} rays report another primitive"
        if (~same).any():
            assert np.abs(t0[~same] / t1[~same] - 1).max() < 1e-6, f"{scene}/{name}: another primitive at another distance"
        o0 = r.debug_trace(rays, any_hit=True)[1] == 0
        o1 = r.debug_trace(rays, any_hit=True, packet_loop=True)[1] == 0
        assert np.array_equal(o0, o1), f"{scene}/{name}: occlusion differs on {(o0 != o1).sum()} rays"
    r.close()


@pytest.mark.parametrize("scene,w,h,spp", [("dragon", 160, 160, 64), ("cornell-box", 96, 96, 64), ("synthetic_opacity_masks", 64, 64, 64)])
def test_frames_with_and_without_warp_packets(pkg, scene, w, h, spp, monkeypatch):
    """B200PT_PACKETS=0 (every ray through the per-lane replacement loop) and the default render the same frame up to ties."""
    path = os.path.join(GOLDEN, scene + ".b200scene") if scene.startswith("synthetic_") else pack(scene)
    sc = pkg.Scene(path)
    frames = []
    for packets in ("0", "3", "15"):
        monkeypatch.setenv("B200PT_PACKETS", packets)
        r = pkg.Renderer(sc, device=0, max_paths_in_flight=1 << 22)
        frames.append(r.Draw(width=w, height=h, spp=spp, seed=5))
        r.close()
    for f in frames[1:]:
        differing = np.any(f != frames[0], axis=2).mean()
        assert differing < 0.02, f"{scene}: {differing:.4f} of the pixels differ with warp packets"
        assert abs(f.mean() / frames[0].mean() - 1.0) < 2e-3
