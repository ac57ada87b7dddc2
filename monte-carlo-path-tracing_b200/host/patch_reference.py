#!/usr/bin/env python3
"""Applies the INTEGRATION.md binding to a BUILD-DIRECTORY COPY of the four reference files it touches.

    python patch_reference.py /root/reference <out_dir>

Nothing is written under the reference tree and nothing of it is committed: the patched copies live in the (git-ignored)
build directory and exist only to compile the drop-in host `bin/RayTracer` (host/Makefile).  Every edit is anchored on a
line of the reference; a reference that no longer has that line makes the script fail instead of producing a silent
mis-patch.  The hunks:

  include/csrt/utils/memory.hpp:18-24   BackendType gains kB200
  include/csrt/ray_tracer.hpp:12-36     RayTracer gains the library handle and the flattened scene
  src/ray_tracer.cpp:124-159            constructor / ReleaseData / Draw forward to b200pt_create / _destroy / _render;
                                        a .pfm output name writes the linear float frame instead of the 8-bit sRGB PNG
  apps/main.cpp:130-149, 188-198        `--b200` / `-b` selects the backend; `.pfm` is accepted as output suffix
"""
import os
import sys


def patch(text, anchor, replacement, path, count=1):
    if text.count(anchor) != count:
        sys.exit(f"patch_reference: anchor not found exactly {count}x in {path}:\n{anchor}")
    return text.replace(anchor, replacement)


def main():
    ref, out = sys.argv[1], sys.argv[2]

    def read(rel):
        with open(os.path.join(ref, rel), encoding="utf-8") as f:
            return f.read()

    def write(rel, text):
        path = os.path.join(out, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        if os.path.exists(path) and open(path, encoding="utf-8").read() == text:
            return  # keep the time stamp: make does not rebuild
        with open(path, "w", encoding="utf-8") as f:
            f.write(text)

    # ---- BackendType::kB200 ----
    rel = "include/csrt/utils/memory.hpp"
    t = read(rel)
    t = patch(t, "    kCpu,\n#ifdef ENABLE_CUDA\n    kCuda\n#endif\n};", "    kCpu,\n#ifdef ENABLE_CUDA\n    kCuda,\n#endif\n    kB200\n};", rel)
    write(rel, t)

    # ---- RayTracer members ----
    rel = "include/csrt/ray_tracer.hpp"
    t = read(rel)
    # the copy lives in another directory: its two relative includes become include-path ones
    t = patch(t, '#include "parser/parser.hpp"\n', '#include "csrt/parser/parser.hpp"\n', rel)
    t = patch(t, '#include "renderer/renderer.hpp"\n', '#include "csrt/renderer/renderer.hpp"\n\n#include "b200pt.h"\n', rel)
    t = patch(t, "    int width_;\n", "    int width_;\n    b200pt_handle b200_ = nullptr;       // BackendType::kB200: the library's renderer ...\n"
                                      "    b200pt_scene *scene_b200_ = nullptr; // ... and the flattened RendererConfig it was created from\n", rel)
    write(rel, t)

    # ---- RayTracer::RayTracer / ReleaseData / Draw ----
    rel = "src/ray_tracer.cpp"
    t = read(rel)
    t = patch(t, '#include "csrt/utils.hpp"\n', '#include "csrt/utils.hpp"\n\n#include "host/csrt_glue.hpp"\n#include "host/pfm.hpp"\n', rel)
    t = patch(t, "        renderer_ = new Renderer(config);\n",
              "        if (config.backend_type == BackendType::kB200)\n"
              "        {\n"
              "            scene_b200_ = b200pt_glue::FlattenConfig(config).release();\n"
              "            if (b200pt_create(&scene_b200_->desc, nullptr, &b200_) != B200PT_OK)\n"
              "                throw MyException(b200pt_last_error(nullptr));\n"
              "        }\n"
              "        else\n"
              "            renderer_ = new Renderer(config);\n", rel)
    t = patch(t, "    csrt::DeleteElement(BackendType::kCpu, renderer_);\n",
              "    b200pt_destroy(b200_);\n"
              "    b200_ = nullptr;\n"
              "    delete scene_b200_;\n"
              "    scene_b200_ = nullptr;\n"
              "    csrt::DeleteElement(BackendType::kCpu, renderer_);\n", rel)
    t = patch(t, "    renderer_->Draw(frame_);\n    csrt::image_io::Write(frame_, width_, height_, output_filename);\n",
              "    if (b200_ != nullptr)\n"
              "    {\n"
              "        if (b200pt_render(b200_, nullptr, frame_) != B200PT_OK)\n"
              "            throw MyException(b200pt_last_error(b200_));\n"
              "    }\n"
              "    else\n"
              "        renderer_->Draw(frame_);\n"
              "    if (GetSuffix(output_filename) == \"pfm\")\n"
              "        b200pt_host::WritePfm(frame_, width_, height_, output_filename);\n"
              "    else\n"
              "        csrt::image_io::Write(frame_, width_, height_, output_filename);\n", rel)
    write(rel, t)

    # ---- CLI ----
    rel = "apps/main.cpp"
    t = read(rel)
    t = patch(t, '        else if ((argv[i] == std::string("--width") ||\n',
              '        else if (argv[i] == std::string("--b200") ||\n'
              '                 argv[i] == std::string("-b"))\n'
              '        {\n'
              '            param.type = csrt::BackendType::kB200;\n'
              '            param.preview = false;\n'
              '        }\n'
              '        else if ((argv[i] == std::string("--width") ||\n', rel)
    t = patch(t, '    if (suffix != "png")\n', '    if (suffix != "png" && suffix != "pfm")\n', rel)
    write(rel, t)


if __name__ == "__main__":
    main()
