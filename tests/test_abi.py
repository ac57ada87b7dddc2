"""The C-ABI shared library loads and exports every symbol include/b200pt.h declares (no compute without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT, pack


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200pt_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"libb200pt.so does not export {name}"
    assert sorted(pkg.EXPORTED_SYMBOLS) == names


def test_struct_layouts_match_the_header(pkg, tmp_path):
    """ctypes mirrors in the Python host layer must have the C compiler's sizes."""
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "b200pt.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(b200pt_render_opts),sizeof(b200pt_create_opts),sizeof(b200pt_stats),sizeof(b200pt_kernel_stats),"
                   "sizeof(b200pt_camera),sizeof(b200pt_texture),sizeof(b200pt_bsdf),sizeof(b200pt_instance),sizeof(b200pt_emitter),"
                   "sizeof(b200pt_medium));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes[0] == ctypes.sizeof(pkg.RenderOpts)
    assert sizes[1] == ctypes.sizeof(pkg.CreateOpts)
    assert sizes[2] == ctypes.sizeof(pkg.Stats)
    assert sizes[3] == ctypes.sizeof(pkg.KernelStats)
    assert sizes[4:] == [52, 120, 76, 200, 120, 44]  # the pack format stores these structs verbatim


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c89.c"
    src.write_text('#include "b200pt.h"\nint main(void){return B200PT_ABI_VERSION == 1u ? 0 : 1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "c.o")])


def test_no_cpu_fallback_without_a_gpu(pkg):
    """Constructing a Renderer where CUDA is unusable must fail loudly, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    scene = pkg.Scene(pack("cornell-box"))
    with pytest.raises(pkg.MyException, match="CUDA error"):
        pkg.Renderer(scene)


def test_product_never_touches_the_oracle():
    """libb200pt.so and the package sources must not link, load or import anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "monte-carlo-path-tracing_b200")
    for base, _, files in os.walk(pkg_dir):
        if os.path.basename(base) == "build":
            continue
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and name != "csrt_glue.hpp":
                text = open(os.path.join(base, name), errors="replace").read()
                assert "liboracle" not in text and "pt_oracle" not in text and "refcheck" not in text, name
                assert not re.search(r"#include\s+[\"<].*oracle/", text), name
    needed = subprocess.check_output(["readelf", "-d", os.path.join(pkg_dir, "libb200pt.so")], text=True)
    assert "oracle" not in needed and "csrt_ref" not in needed


def test_pfm_frame_dump_round_trips(pkg, tmp_path):
    """The float frame dump next to the reference's 8-bit PNG: lossless, top row first on read."""
    import numpy as np
    frame = np.random.RandomState(3).rand(5, 7, 3).astype(np.float32) * 4.0
    pkg.write_pfm(str(tmp_path / "f.pfm"), frame)
    assert np.array_equal(pkg.read_pfm(str(tmp_path / "f.pfm")), frame)


def test_kulla_conty_tables_equal_the_reference_tables_without_a_gpu(pkg):
    """The energy-compensation tables of the rough conductor / dielectric / plastic models (bsdf.cpp:112-186), which
    b200pt_create computes on the host: bit-equal to the tables the reference build produced (tests/golden/kulla_conty.npz).
    Without a handle the entry computes them right away, so this runs on a CPU box."""
    import numpy as np
    golden = np.load(os.path.join(ROOT, "tests", "golden", "kulla_conty.npz"))
    brdf, albedo = np.zeros((128, 128), dtype=np.float32), np.zeros(128, dtype=np.float32)
    assert pkg.lib().b200pt_get_kulla_conty(None, brdf.ctypes.data, albedo.ctypes.data) == 0
    assert np.array_equal(brdf, golden["brdf_avg"])
    assert np.array_equal(albedo, golden["albedo_avg"])
