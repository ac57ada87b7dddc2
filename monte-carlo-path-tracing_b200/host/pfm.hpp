// pfm.hpp — lossless float frame dump for the drop-in host (image_io::Write, src/utils/image_io.cpp:25-53,
// only writes 8-bit sRGB PNGs).  Portable float map: "PF", size, negative scale = little endian, rows bottom-up.
#pragma once
#include <cstdio>
#include <string>

namespace b200pt_host {

inline bool WritePfm(const float *frame, int width, int height, const std::string &filename) {
    FILE *f = fopen(filename.c_str(), "wb");
    if (f == nullptr) {
        fprintf(stderr, "[error] write image failed.\n");
        return false;
    }
    fprintf(f, "PF\n%d %d\n-1.0\n", width, height);
    for (int row = height - 1; row >= 0; --row) fwrite(frame + static_cast<size_t>(row) * width * 3, sizeof(float), static_cast<size_t>(width) * 3, f);
    fclose(f);
    fprintf(stderr, "[info] save result as image \"%s\".\n", filename.c_str());
    return true;
}

} // namespace b200pt_host
