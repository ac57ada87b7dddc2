#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — writes the RGBA float32 sidecar that monte-carlo-path-tracing_b200/host/shims/tinyexr.h reads.

Usage: make_exr_sidecar.py <in.exr> <out_dir>
Decodes with OpenCV's OpenEXR reader (BGR float32) and stores int32 w, int32 h, RGBA float32 (A=1),
top row first, which is what tinyexr's LoadEXR hands to src/utils/image_io.cpp:79-87.
"""
import os
import sys

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402


def main():
    src, out_dir = sys.argv[1], sys.argv[2]
    img = cv2.imread(src, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise SystemExit(f"cannot decode {src}")
    img = img.astype(np.float32)
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)
    h, w = img.shape[:2]
    rgba = np.ones((h, w, 4), dtype=np.float32)
    rgba[:, :, 0] = img[:, :, 2]
    rgba[:, :, 1] = img[:, :, 1]
    rgba[:, :, 2] = img[:, :, 0]
    if img.shape[2] == 4:
        rgba[:, :, 3] = img[:, :, 3]
    os.makedirs(out_dir, exist_ok=True)
    dst = os.path.join(out_dir, os.path.basename(src) + ".rgba32f")
    with open(dst, "wb") as f:
        f.write(np.array([w, h], dtype=np.int32).tobytes())
        f.write(rgba.tobytes())
    print(dst, w, h, float(rgba[:, :, :3].max()))


if __name__ == "__main__":
    main()
