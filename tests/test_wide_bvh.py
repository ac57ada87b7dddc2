"""CPU check of the compressed 8-wide BVH builder and of the group-walk traversal scheme (tests/wide_bvh_check.cpp):
structure invariants + closest hits bit-identical to brute force.  No GPU needed; the CUDA traversal itself is compared
with the reference's TLAS::Intersect in tests/test_gpu_traversal.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "monte-carlo-path-tracing_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("wide") / "wide_bvh_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "wide_bvh_check.cpp"), os.path.join(CSRC, "bvh_wide.cpp"), "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("tris,rays,seed", [(0, 10, 1), (1, 200, 2), (2, 200, 3), (3, 200, 4), (9, 500, 5), (25, 500, 6),
                                            (1000, 3000, 7), (30000, 6000, 8), (120000, 6000, 9)])
def test_wide_bvh_matches_brute_force(checker, tris, rays, seed):
    out = subprocess.run([checker, str(tris), str(rays), str(seed)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
