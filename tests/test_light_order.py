"""Mesh area lights: the same random number must pick the same triangle as the reference's BLAS::Sample.

BLAS::Sample (src/rtcore/accel/blas.cpp:79-98) walks the instance's BVH from the root with thresh = area * xi_0 and goes left
while thresh < area(left): a cumulative distribution over the LEAVES IN TREE ORDER.  The product samples a mesh light through
an explicit CDF (csrc/scene_build.cpp); for the exact-mode comparison (tests/test_gpu_replay.py) that CDF has to stand in the
leaf order of the reference's Morton-built tree (bvh_builder.cpp:92-141).  Checked here without a GPU: the order
b200pt_debug_light_order reports against the leaf order of the checker's LBVH (oracle_build_bvh, bit-equal to the reference's
builder: tests/test_oracle_pinning.py) over the same triangle boxes."""
import ctypes
import sys

import numpy as np
import pytest

import scene_builder as sb


def world_positions(to_world, positions):
    """TransformPoint (mat4.cpp:265-268) in float32, term by term as the reference evaluates it."""
    m = np.asarray(to_world, dtype=np.float32).reshape(4, 4)
    p = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
    out = np.zeros_like(p)
    for r in range(3):
        acc = m[r, 0] * p[:, 0]
        acc = (acc + m[r, 1] * p[:, 1]).astype(np.float32)
        acc = (acc + m[r, 2] * p[:, 2]).astype(np.float32)
        out[:, r] = (acc + m[r, 3] * np.float32(1.0)).astype(np.float32)
    return out


def reference_leaf_order(oracle, tri_boxes):
    n = len(tri_boxes)
    boxes = np.ascontiguousarray(tri_boxes, dtype=np.float32)
    areas = np.ones(n, dtype=np.float32)
    nodes = np.zeros((2 * n, 4), dtype=np.uint32)
    area = np.zeros(2 * n, dtype=np.float32)
    count = oracle.lib.oracle_build_bvh(n, boxes.ctypes.data, areas.ctypes.data, nodes.ctypes.data, area.ctypes.data, 2 * n)
    assert count == 2 * n - 1
    order, stack = [], [0]
    while stack:  # depth first, left before right
        k = stack.pop()
        if nodes[k, 0]:
            order.append(int(nodes[k, 3]))
        else:
            stack.append(int(nodes[k, 2]))
            stack.append(int(nodes[k, 1]))
    return np.array(order, dtype=np.uint32)


def test_mesh_light_cdf_stands_in_the_reference_leaf_order(pkg, oracle):
    rng = np.random.RandomState(3)
    b = sb.SceneBuilder(depth_max=6)
    sb.stage(b)                                                      # light 0: the built-in rectangle of the stage
    pos, idx, nrm, uv = sb.uv_sphere()
    sphere_to_world = sb.translate(0.3, 1.0, 0.2) @ sb.rotate_y(25) @ sb.scale(0.5, 0.7, 0.4)
    b.mesh(b.area_light((5, 4, 3)), pos, idx, nrm, uv, to_world=sphere_to_world)            # light 1: 528 triangles
    soup = rng.rand(300, 3, 3).astype(np.float32)                                            # light 2: a triangle soup
    soup_to_world = sb.translate(-1.5, 0.2, 0.4) @ sb.scale(0.8, 1.3, 0.6)
    b.mesh(b.area_light((2, 2, 2)), soup.reshape(-1, 3), np.arange(900, dtype=np.uint32).reshape(-1, 3), to_world=soup_to_world)
    flat = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 0, 0], [2, 1, 0]], dtype=np.float32)  # light 3: flat (one axis of extent 0)
    b.mesh(b.area_light((1, 1, 1)), flat, np.array([[0, 1, 2], [2, 3, 0], [1, 4, 5], [5, 2, 1]], dtype=np.uint32), to_world=sb.translate(0, 3.0, 0))
    desc = b.desc()
    meshes = [(sphere_to_world, pos, idx), (soup_to_world, soup.reshape(-1, 3), np.arange(900).reshape(-1, 3)),
              (sb.translate(0, 3.0, 0), flat, np.array([[0, 1, 2], [2, 3, 0], [1, 4, 5], [5, 2, 1]]))]
    ids, cdf = pkg.light_order(ctypes.byref(desc), 0)
    assert len(ids) == 2 and sorted(ids) == [0, 1] and cdf[-1] == 1.0
    for light, (to_world, positions, indices) in enumerate(meshes, start=1):
        ids, cdf = pkg.light_order(ctypes.byref(desc), light)
        tri = world_positions(to_world, positions)[np.asarray(indices).reshape(-1, 3)]       # [n, 3 corners, xyz]
        assert len(ids) == len(tri)
        boxes = np.concatenate([tri.min(axis=1), tri.max(axis=1)], axis=1)
        expected = reference_leaf_order(oracle, boxes)
        assert np.array_equal(ids, expected), (light, int(np.argmax(ids != expected)))
        # the CDF is the running sum of |e1 x e2| in that order, normalised
        e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
        w = np.linalg.norm(np.cross(e1.astype(np.float64), e2.astype(np.float64)), axis=1)[expected]
        assert np.allclose(cdf, np.cumsum(w) / w.sum(), rtol=0, atol=2e-6)
    with pytest.raises(pkg.MyException):
        pkg.light_order(ctypes.byref(desc), 9)
