"""CPU check of the DEFAULT tree: the product's own host builder (parallel binned SAH + flatten, csrc/scene_build.cpp, reached
through libb200pt.so) on real scene packs — structure invariants and closest hits equal to brute force (tests/sah_bvh_check.cpp).
No GPU needed; the CUDA traversal of the same tree is compared with the reference's TLAS::Intersect in tests/test_gpu_traversal.py."""
import os
import subprocess

import pytest

from conftest import GOLDEN, pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "monte-carlo-path-tracing_b200")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    if not os.path.exists(os.path.join(PKG, "libb200pt.so")):
        pytest.skip("libb200pt.so is not built")
    exe = str(tmp_path_factory.mktemp("sah") / "sah_bvh_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(PKG, "csrc"), "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "sah_bvh_check.cpp"), "-L" + PKG, "-lb200pt", "-Wl,-rpath," + PKG, "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("scene,rays", [("cornell-box", 2000), ("volumetric-caustic", 2000), ("matpreview", 600), ("box", 600),
                                        ("synthetic_bump_bitmap_mesh_disk", 2000), ("dragon", 40)])
def test_default_tree_structure_and_closest_hits(checker, scene, rays):
    path = os.path.join(GOLDEN, scene + ".b200scene") if scene.startswith("synthetic_") else pack(scene)
    out = subprocess.run([checker, path, str(rays), "7"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
