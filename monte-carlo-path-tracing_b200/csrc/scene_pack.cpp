// scene_pack.cpp — lossless binary serialisation of b200pt_scene_desc ("scene pack").
//
// The reference keeps its scene only as an in-memory csrt::RendererConfig
// (include/csrt/renderer/renderer.hpp:18-28) produced by its XML parser.  A pack
// is that structure written to disk so the same parsed scene can be consumed by
// the CUDA path, by the CPU oracle and by the reference build, on a machine where
// the reference tree (XML, OBJ, JPEG, EXR readers) is not present.
//
// Layout: magic, header (camera + integrator), then tagged sections.  Float/u32
// sections are byte-shuffled (stride 4) and zlib-deflated; the bitmap pool is
// palette-coded to u8 when it holds <= 256 distinct float values (true for any
// texture decoded from an 8-bit image), which is still lossless.
#include "b200pt.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <unordered_map>
#include <vector>

#include "host_util.hpp"
#include "scene_storage.hpp"

namespace {

constexpr char kMagic[8] = {'B', '2', 'P', 'T', 'S', 'C', 'N', '1'};

enum Tag : uint32_t {
    kTextures = 1, kPixels, kBsdfs, kMedia, kInstances, kEmitters,
    kPositions, kNormals, kTexcoords, kTangents, kBitangents, kIndices
};
enum Codec : uint32_t { kRaw = 0, kShuffleZ = 1, kPaletteZ = 2, kZ = 3 };

struct SectionHeader {
    uint32_t tag, codec;
    uint64_t count;        // elements (tag-specific unit)
    uint64_t raw_bytes;    // bytes after decoding
    uint64_t stored_bytes; // bytes in the file
};

bool Deflate(const std::vector<uint8_t> &in, std::vector<uint8_t> *out) {
    uLongf cap = compressBound(in.size());
    out->resize(cap);
    if (compress2(out->data(), &cap, in.data(), in.size(), 4) != Z_OK) return false;
    out->resize(cap);
    return true;
}

bool Inflate(const uint8_t *in, size_t n_in, size_t n_out, std::vector<uint8_t> *out) {
    out->resize(n_out);
    uLongf len = n_out;
    if (n_out == 0) return true;
    if (uncompress(out->data(), &len, in, n_in) != Z_OK) return false;
    return len == n_out;
}

void Shuffle4(const uint8_t *src, size_t bytes, std::vector<uint8_t> *dst) {
    dst->resize(bytes);
    const size_t n = bytes / 4;
    for (size_t i = 0; i < n; ++i)
        for (int b = 0; b < 4; ++b) (*dst)[b * n + i] = src[i * 4 + b];
    for (size_t i = n * 4; i < bytes; ++i) (*dst)[i] = src[i];
}

void Unshuffle4(const std::vector<uint8_t> &src, uint8_t *dst) {
    const size_t bytes = src.size(), n = bytes / 4;
    for (size_t i = 0; i < n; ++i)
        for (int b = 0; b < 4; ++b) dst[i * 4 + b] = src[b * n + i];
    for (size_t i = n * 4; i < bytes; ++i) dst[i] = src[i];
}

bool WriteSection(FILE *f, uint32_t tag, uint64_t count, const void *data, size_t bytes, bool numeric) {
    SectionHeader h{tag, kRaw, count, bytes, bytes};
    std::vector<uint8_t> tmp, z;
    const uint8_t *payload = static_cast<const uint8_t *>(data);
    if (bytes > 0) {
        if (numeric) {
            Shuffle4(payload, bytes, &tmp);
            h.codec = kShuffleZ;
        } else {
            tmp.assign(payload, payload + bytes);
            h.codec = kZ;
        }
        if (!Deflate(tmp, &z)) return false;
        h.stored_bytes = z.size();
        payload = z.data();
    }
    if (fwrite(&h, sizeof(h), 1, f) != 1) return false;
    if (h.stored_bytes && fwrite(payload, 1, h.stored_bytes, f) != h.stored_bytes) return false;
    return true;
}

// Palette coding of the bitmap pool: [u32 n_palette][n_palette floats][u8 index per float].
bool WritePixels(FILE *f, const float *pixels, uint64_t n) {
    std::unordered_map<uint32_t, uint8_t> palette;
    std::vector<uint32_t> order;
    bool ok = n > 0;
    const uint32_t *bits = reinterpret_cast<const uint32_t *>(pixels);
    for (uint64_t i = 0; ok && i < n; ++i) {
        if (palette.find(bits[i]) == palette.end()) {
            if (order.size() == 256) { ok = false; break; }
            palette.emplace(bits[i], static_cast<uint8_t>(order.size()));
            order.push_back(bits[i]);
        }
    }
    if (!ok) return WriteSection(f, kPixels, n, pixels, n * sizeof(float), true);

    std::vector<uint8_t> raw(4 + order.size() * 4 + n);
    const uint32_t np = static_cast<uint32_t>(order.size());
    memcpy(raw.data(), &np, 4);
    memcpy(raw.data() + 4, order.data(), order.size() * 4);
    uint8_t *idx = raw.data() + 4 + order.size() * 4;
    for (uint64_t i = 0; i < n; ++i) idx[i] = palette[bits[i]];
    std::vector<uint8_t> z;
    if (!Deflate(raw, &z)) return false;
    SectionHeader h{kPixels, kPaletteZ, n, raw.size(), z.size()};
    if (fwrite(&h, sizeof(h), 1, f) != 1) return false;
    return fwrite(z.data(), 1, z.size(), f) == z.size();
}

} // namespace

extern "C" int b200pt_scene_save(const b200pt_scene_desc *s, const char *path) {
    if (!s || !path) return b200pt::SetGlobalError(B200PT_EINVAL, "b200pt_scene_save: null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return b200pt::SetGlobalError(B200PT_EIO, std::string("cannot open '") + path + "' for writing");
    bool ok = fwrite(kMagic, 8, 1, f) == 1;
    const uint32_t version = B200PT_ABI_VERSION;
    ok = ok && fwrite(&version, 4, 1, f) == 1;
    ok = ok && fwrite(&s->camera, sizeof(s->camera), 1, f) == 1;
    ok = ok && fwrite(&s->integrator, sizeof(s->integrator), 1, f) == 1;
    ok = ok && WriteSection(f, kTextures, s->num_textures, s->textures, s->num_textures * sizeof(b200pt_texture), false);
    ok = ok && WritePixels(f, s->pixels, s->num_pixels);
    ok = ok && WriteSection(f, kBsdfs, s->num_bsdfs, s->bsdfs, s->num_bsdfs * sizeof(b200pt_bsdf), false);
    ok = ok && WriteSection(f, kMedia, s->num_media, s->media, s->num_media * sizeof(b200pt_medium), false);
    ok = ok && WriteSection(f, kInstances, s->num_instances, s->instances, s->num_instances * sizeof(b200pt_instance), false);
    ok = ok && WriteSection(f, kEmitters, s->num_emitters, s->emitters, s->num_emitters * sizeof(b200pt_emitter), false);
    ok = ok && WriteSection(f, kPositions, s->num_positions, s->positions, s->num_positions * 12, true);
    ok = ok && WriteSection(f, kNormals, s->num_normals, s->normals, s->num_normals * 12, true);
    ok = ok && WriteSection(f, kTexcoords, s->num_texcoords, s->texcoords, s->num_texcoords * 8, true);
    ok = ok && WriteSection(f, kTangents, s->num_tangents, s->tangents, s->num_tangents * 12, true);
    ok = ok && WriteSection(f, kBitangents, s->num_bitangents, s->bitangents, s->num_bitangents * 12, true);
    ok = ok && WriteSection(f, kIndices, s->num_triangles, s->indices, s->num_triangles * 12, true);
    const SectionHeader end{0, 0, 0, 0, 0};
    ok = ok && fwrite(&end, sizeof(end), 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return b200pt::SetGlobalError(B200PT_EIO, std::string("short write to '") + path + "'");
    return B200PT_OK;
}

namespace {

constexpr uint64_t kMaxSectionBytes = 1ull << 36; // 64 GiB: far above any scene, far below an overflow

template <typename T>
bool Decode(const SectionHeader &h, const std::vector<uint8_t> &stored, size_t elem_bytes, std::vector<T> *out) {
    if (elem_bytes == 0 || h.count > kMaxSectionBytes / elem_bytes || h.raw_bytes != h.count * elem_bytes) return false; // no overflow
    out->resize(h.raw_bytes / sizeof(T));
    uint8_t *dst = reinterpret_cast<uint8_t *>(out->data());
    std::vector<uint8_t> tmp;
    switch (h.codec) {
    case kRaw:
        if (stored.size() != h.raw_bytes) return false;
        if (h.raw_bytes) memcpy(dst, stored.data(), h.raw_bytes);
        return true;
    case kZ:
        if (!Inflate(stored.data(), stored.size(), h.raw_bytes, &tmp)) return false;
        if (h.raw_bytes) memcpy(dst, tmp.data(), h.raw_bytes);
        return true;
    case kShuffleZ:
        if (!Inflate(stored.data(), stored.size(), h.raw_bytes, &tmp)) return false;
        Unshuffle4(tmp, dst);
        return true;
    default:
        return false;
    }
}

bool DecodePixels(const SectionHeader &h, const std::vector<uint8_t> &stored, std::vector<float> *out) {
    if (h.codec != kPaletteZ) return Decode(h, stored, sizeof(float), out);
    std::vector<uint8_t> raw;
    if (!Inflate(stored.data(), stored.size(), h.raw_bytes, &raw) || raw.size() < 4) return false;
    uint32_t np;
    memcpy(&np, raw.data(), 4);
    if (np > 256 || raw.size() != 4 + size_t(np) * 4 + h.count) return false;
    const float *palette = reinterpret_cast<const float *>(raw.data() + 4);
    const uint8_t *idx = raw.data() + 4 + size_t(np) * 4;
    out->resize(h.count);
    for (uint64_t i = 0; i < h.count; ++i) {
        if (idx[i] >= np) return false;
        (*out)[i] = palette[idx[i]];
    }
    return true;
}

} // namespace

// A corrupt or hostile pack must come back as B200PT_EIO, not as std::bad_alloc through the C boundary: every size in a section
// header is checked against the bytes the file still holds and against a hard cap before anything is allocated.
static int SceneLoadChecked(const char *path, b200pt_scene **out);

extern "C" int b200pt_scene_load(const char *path, b200pt_scene **out) {
    if (!path || !out) return b200pt::SetGlobalError(B200PT_EINVAL, "b200pt_scene_load: null argument");
    *out = nullptr;
    try {
        return SceneLoadChecked(path, out);
    } catch (const std::exception &e) {
        return b200pt::SetGlobalError(B200PT_EIO, std::string("scene pack '") + path + "': " + e.what());
    } catch (...) {
        return b200pt::SetGlobalError(B200PT_EIO, std::string("scene pack '") + path + "': unknown failure");
    }
}

static int SceneLoadChecked(const char *path, b200pt_scene **out) {
    FILE *f = fopen(path, "rb");
    if (!f) return b200pt::SetGlobalError(B200PT_EIO, std::string("cannot open scene pack '") + path + "'");
    auto fail = [&](const std::string &why, b200pt_scene *s) {
        fclose(f);
        delete s;
        return b200pt::SetGlobalError(B200PT_EIO, "scene pack '" + std::string(path) + "': " + why);
    };
    fseek(f, 0, SEEK_END);
    const uint64_t file_bytes = static_cast<uint64_t>(ftell(f));
    fseek(f, 0, SEEK_SET);
    char magic[8];
    uint32_t version = 0;
    if (fread(magic, 8, 1, f) != 1 || memcmp(magic, kMagic, 8) != 0) return fail("bad magic", nullptr);
    if (fread(&version, 4, 1, f) != 1 || version != B200PT_ABI_VERSION) return fail("ABI version mismatch", nullptr);
    b200pt_scene *s = new b200pt_scene();
    if (fread(&s->desc.camera, sizeof(b200pt_camera), 1, f) != 1 ||
        fread(&s->desc.integrator, sizeof(b200pt_integrator), 1, f) != 1)
        return fail("truncated header", s);
    for (;;) {
        SectionHeader h;
        if (fread(&h, sizeof(h), 1, f) != 1) return fail("truncated section header", s);
        if (h.tag == 0) break;
        const uint64_t here = static_cast<uint64_t>(ftell(f));
        if (h.stored_bytes > file_bytes - std::min(here, file_bytes)) return fail("section larger than the file", s);
        if (h.raw_bytes > kMaxSectionBytes) return fail("section larger than the format allows", s);
        std::vector<uint8_t> stored(h.stored_bytes);
        if (h.stored_bytes && fread(stored.data(), 1, h.stored_bytes, f) != h.stored_bytes)
            return fail("truncated section payload", s);
        bool ok = false;
        switch (h.tag) {
        case kTextures: ok = Decode(h, stored, sizeof(b200pt_texture), &s->textures); break;
        case kPixels: ok = DecodePixels(h, stored, &s->pixels); break;
        case kBsdfs: ok = Decode(h, stored, sizeof(b200pt_bsdf), &s->bsdfs); break;
        case kMedia: ok = Decode(h, stored, sizeof(b200pt_medium), &s->media); break;
        case kInstances: ok = Decode(h, stored, sizeof(b200pt_instance), &s->instances); break;
        case kEmitters: ok = Decode(h, stored, sizeof(b200pt_emitter), &s->emitters); break;
        case kPositions: ok = Decode(h, stored, 12, &s->positions); break;
        case kNormals: ok = Decode(h, stored, 12, &s->normals); break;
        case kTexcoords: ok = Decode(h, stored, 8, &s->texcoords); break;
        case kTangents: ok = Decode(h, stored, 12, &s->tangents); break;
        case kBitangents: ok = Decode(h, stored, 12, &s->bitangents); break;
        case kIndices: ok = Decode(h, stored, 12, &s->indices); break;
        default: ok = true; break; // unknown section: skip (forward compatible)
        }
        if (!ok) return fail("corrupt section " + std::to_string(h.tag), s);
    }
    fclose(f);
    s->Finalize();
    *out = s;
    return B200PT_OK;
}

extern "C" const b200pt_scene_desc *b200pt_scene_get_desc(const b200pt_scene *s) { return s ? &s->desc : nullptr; }

extern "C" void b200pt_scene_free(b200pt_scene *s) { delete s; }
