// tinyexr.h — TEST INFRASTRUCTURE ONLY (oracle build shim, never shipped).
//
// The reference reads OpenEXR through tinyexr (src/utils/image_io.cpp:9,79-87), which is
// neither vendored nor installed here.  Only LoadEXR / FreeEXRErrorMessage / TINYEXR_SUCCESS
// are used.  This shim does not decode EXR: it reads a sidecar produced once by
// oracle/make_exr_sidecar.py (OpenCV's OpenEXR reader): int32 width, int32 height, then
// width*height RGBA float32 texels, top row first — the layout tinyexr's LoadEXR returns.
// Sidecars are looked up as  $B200PT_EXR_SIDECAR_DIR/<basename>.rgba32f
// (default directory: oracle/_ref/exr).
#ifndef ORACLE_SHIM_TINYEXR_H
#define ORACLE_SHIM_TINYEXR_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#define TINYEXR_SUCCESS 0
#define TINYEXR_ERROR_CANT_OPEN_FILE (-7)

inline void FreeEXRErrorMessage(const char *msg) { free(const_cast<char *>(msg)); }

inline int LoadEXR(float **out_rgba, int *width, int *height, const char *filename, const char **err) {
    std::string name = filename;
    const size_t slash = name.find_last_of("/\\");
    if (slash != std::string::npos) name = name.substr(slash + 1);
    const char *dir = getenv("B200PT_EXR_SIDECAR_DIR");
    const std::string path = std::string(dir ? dir : "oracle/_ref/exr") + "/" + name + ".rgba32f";
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        if (err) *err = strdup(("EXR sidecar not found: " + path).c_str());
        return TINYEXR_ERROR_CANT_OPEN_FILE;
    }
    int32_t wh[2] = {0, 0};
    bool ok = fread(wh, sizeof(int32_t), 2, f) == 2 && wh[0] > 0 && wh[1] > 0;
    float *data = nullptr;
    if (ok) {
        const size_t n = static_cast<size_t>(wh[0]) * wh[1] * 4;
        data = new float[n]; // the reference frees image buffers with delete[] (parser.cpp SAFE_DELETE_ARRAY)
        ok = fread(data, sizeof(float), n, f) == n;
    }
    fclose(f);
    if (!ok) {
        delete[] data;
        if (err) *err = strdup(("corrupt EXR sidecar: " + path).c_str());
        return TINYEXR_ERROR_CANT_OPEN_FILE;
    }
    *out_rgba = data;
    *width = wh[0];
    *height = wh[1];
    return TINYEXR_SUCCESS;
}

#endif
