"""Scene packs: the lossless serialisation of b200pt_scene_desc that carries parsed scenes to machines without the reference."""
import ctypes
import os

import numpy as np
import pytest

from conftest import pack


def desc_counts(pkg, scene):
    """(num_textures, num_pixels, num_bsdfs, num_media, num_instances, num_emitters, num_positions, ..., num_triangles)."""
    # header: abi u32, reserved u32, camera 52 B, integrator 20 B -> 80 bytes, then (count u64, pointer) pairs
    base = scene.desc + 80
    raw = np.ctypeslib.as_array(ctypes.cast(base, ctypes.POINTER(ctypes.c_uint64)), shape=(24,))
    return [int(raw[2 * k]) for k in range(12)]


def test_known_scene_contents(pkg):
    dragon = pkg.Scene(pack("dragon"))
    counts = desc_counts(pkg, dragon)
    assert counts[4] == 16            # instances (resources/scene/dragon/scene.xml)
    assert counts[11] == 831812       # triangles
    assert counts[5] == 1             # one directional emitter
    assert (dragon.width, dragon.height, dragon.spp) == (1024, 1024, 256)
    cornell = pkg.Scene(pack("cornell-box"))
    assert desc_counts(pkg, cornell)[4] == 8  # 5 walls + 2 boxes + 1 area light


def test_round_trip_is_lossless(pkg, tmp_path):
    for name in ("cornell-box", "matpreview", "volumetric-caustic"):
        a = pkg.Scene(pack(name))
        out = str(tmp_path / (name + ".b200scene"))
        a.save(out)
        b = pkg.Scene(out)
        ca, cb = desc_counts(pkg, a), desc_counts(pkg, b)
        assert ca == cb
        # compare every array byte for byte
        elem = [120, 4, 76, 44, 200, 120, 12, 12, 8, 12, 12, 12]
        pa = np.ctypeslib.as_array(ctypes.cast(a.desc + 80, ctypes.POINTER(ctypes.c_uint64)), shape=(24,))
        pb = np.ctypeslib.as_array(ctypes.cast(b.desc + 80, ctypes.POINTER(ctypes.c_uint64)), shape=(24,))
        for k in range(12):
            n = ca[k] * elem[k]
            if n:
                assert ctypes.string_at(int(pa[2 * k + 1]), n) == ctypes.string_at(int(pb[2 * k + 1]), n), (name, k)
        assert ctypes.string_at(a.desc, 80) == ctypes.string_at(b.desc, 80)


def test_corrupt_and_missing_packs_fail_loudly(pkg, tmp_path):
    with pytest.raises(pkg.MyException, match="cannot open"):
        pkg.Scene(str(tmp_path / "nope.b200scene"))
    bad = tmp_path / "bad.b200scene"
    bad.write_bytes(b"not a scene pack at all")
    with pytest.raises(pkg.MyException, match="bad magic"):
        pkg.Scene(str(bad))
    good = open(pack("cornell-box"), "rb").read()
    truncated = tmp_path / "trunc.b200scene"
    truncated.write_bytes(good[: len(good) // 2])
    with pytest.raises(pkg.MyException):
        pkg.Scene(str(truncated))


def test_palette_coded_bitmap_pool(pkg):
    """mercury's 8192x4096 JPEG texture (403 MB of floats) is palette-coded in the pack and must decode to <=256 levels."""
    path = os.path.join(os.path.dirname(pack("cornell-box")), "mercury.b200scene")
    if not os.path.exists(path):
        pytest.skip("mercury pack not present")
    scene = pkg.Scene(path)
    counts = desc_counts(pkg, scene)
    assert counts[1] == 8192 * 4096 * 3
    raw = np.ctypeslib.as_array(ctypes.cast(scene.desc + 80, ctypes.POINTER(ctypes.c_uint64)), shape=(24,))
    pixels = np.ctypeslib.as_array(ctypes.cast(int(raw[3]), ctypes.POINTER(ctypes.c_float)), shape=(1 << 22,))
    assert len(np.unique(pixels)) <= 256
    assert 0.0 <= pixels.min() and pixels.max() <= 1.0


def test_hostile_section_sizes_come_back_as_errors(pkg, tmp_path):
    """Round-1 advisor finding: the section headers carry 64-bit sizes.  A pack that claims petabytes, or a count whose product
    with the element size overflows, must be refused with B200PT_EIO before anything is allocated — no std::bad_alloc through
    the C boundary."""
    import struct
    good = bytearray(open(pack("cornell-box"), "rb").read())
    header = 8 + 4 + 52 + 20            # magic, version, camera, integrator
    tag, codec, count, raw_bytes, stored_bytes = struct.unpack_from("<IIQQQ", good, header)
    assert tag != 0 and stored_bytes < len(good)
    for field, value in ((24, 1 << 60), (16, 1 << 50), (8, (1 << 64) // 120 + 7)):  # stored_bytes, raw_bytes, count
        bad = bytearray(good)
        struct.pack_into("<Q", bad, header + field, value)
        path = tmp_path / f"hostile_{field}.b200scene"
        path.write_bytes(bytes(bad))
        with pytest.raises(pkg.MyException, match="scene pack"):
            pkg.Scene(str(path))
