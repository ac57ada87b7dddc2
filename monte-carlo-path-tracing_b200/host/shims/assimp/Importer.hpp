// assimp/Importer.hpp — TEST INFRASTRUCTURE ONLY (oracle build shim, never shipped).
//
// The reference loads OBJ / PLY meshes through Assimp (src/parser/model_loader.cpp:506-531),
// which is neither vendored nor installed here.  This is a from-scratch Wavefront-OBJ and
// ASCII-PLY reader exposing the tiny slice of Assimp::Importer that model_loader.cpp calls:
//   Importer::ReadFile(path, flags) -> const aiScene*,  Importer::GetErrorString().
// It honours the post-process flags the reference passes (model_loader.cpp:512-520):
//   Triangulate (fan), GenSmoothNormals (when the file has no `vn`), FlipUVs (v -> 1-v),
//   CalcTangentSpace (per-face UV-derivative tangents accumulated per vertex, Gram-Schmidt
//   against the normal; degenerate UVs fall back to an orthonormal frame built from the
//   normal, so no NaN tangents are ever produced — see SURVEY.md §7 "Oracle dependencies").
// Identical (v,vt,vn) corners are joined so a mesh stays indexed; the triangles produced
// are the same as with un-joined corners.
// The scene graph returned is a root node with no meshes and one child holding one mesh,
// which keeps the reference's child index-offset arithmetic (model_loader.cpp:395-396) at 0.
#ifndef ORACLE_SHIM_ASSIMP_IMPORTER_HPP
#define ORACLE_SHIM_ASSIMP_IMPORTER_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "postprocess.h"
#include "scene.h"

namespace Assimp {

class Importer {
public:
    Importer() { memset(&scene_, 0, sizeof(scene_)); }
    Importer(const Importer &) = delete;
    Importer &operator=(const Importer &) = delete;

    const char *GetErrorString() const { return error_.c_str(); }

    const aiScene *ReadFile(const std::string &path, unsigned int flags) {
        error_.clear();
        const size_t dot = path.find_last_of('.');
        const std::string suffix = dot == std::string::npos ? "" : path.substr(dot + 1);
        if (suffix == "ply" || suffix == "PLY") {
            if (!LoadPly(path, flags)) return nullptr;
        } else if (suffix == "obj" || suffix == "OBJ") {
            if (!LoadObj(path, flags)) return nullptr;
        } else {
            error_ = "assimp shim: only Wavefront OBJ and ASCII PLY are supported ('" + path + "')";
            return nullptr;
        }
        Publish();
        return &scene_;
    }

private:
    struct V3 { float x, y, z; };
    struct Key {
        int v, vt, vn;
        bool operator==(const Key &o) const { return v == o.v && vt == o.vt && vn == o.vn; }
    };
    struct KeyHash {
        size_t operator()(const Key &k) const {
            size_t h = static_cast<size_t>(k.v) * 73856093u;
            h ^= static_cast<size_t>(k.vt + 1) * 19349663u;
            h ^= static_cast<size_t>(k.vn + 1) * 83492791u;
            return h;
        }
    };

    static V3 Sub(const V3 &a, const V3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
    static V3 Cross(const V3 &a, const V3 &b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
    static float Dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    static V3 Scale(const V3 &a, float s) { return {a.x * s, a.y * s, a.z * s}; }
    static bool NormalizeSafe(V3 *v) {
        const float len = std::sqrt(Dot(*v, *v));
        if (!(len > 1e-20f) || !std::isfinite(len)) return false;
        *v = Scale(*v, 1.0f / len);
        return true;
    }
    static V3 AnyPerpendicular(const V3 &n) {
        V3 a = std::fabs(n.x) > 0.9f ? V3{0, 1, 0} : V3{1, 0, 0};
        V3 t = Sub(a, Scale(n, Dot(a, n)));
        NormalizeSafe(&t);
        return t;
    }

    static int ResolveIndex(long idx, size_t count) {
        if (idx > 0) return static_cast<int>(idx - 1);
        if (idx < 0) return static_cast<int>(static_cast<long>(count) + idx);
        return -1;
    }

    bool LoadObj(const std::string &path, unsigned int flags) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) {
            error_ = "cannot open '" + path + "'";
            return false;
        }
        std::vector<V3> file_v, file_vn;
        std::vector<V3> file_vt;
        std::vector<Key> corners; // 3 per triangle
        std::vector<char> line(1 << 16);
        while (fgets(line.data(), static_cast<int>(line.size()), f)) {
            char *s = line.data();
            while (*s == ' ' || *s == '\t') ++s;
            if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
                V3 v{0, 0, 0};
                char *e = s + 1;
                v.x = strtof(e, &e); v.y = strtof(e, &e); v.z = strtof(e, &e);
                file_v.push_back(v);
            } else if (s[0] == 'v' && s[1] == 'n') {
                V3 v{0, 0, 0};
                char *e = s + 2;
                v.x = strtof(e, &e); v.y = strtof(e, &e); v.z = strtof(e, &e);
                file_vn.push_back(v);
            } else if (s[0] == 'v' && s[1] == 't') {
                V3 v{0, 0, 0};
                char *e = s + 2;
                v.x = strtof(e, &e); v.y = strtof(e, &e);
                file_vt.push_back(v);
            } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
                std::vector<Key> poly;
                char *e = s + 1;
                for (;;) {
                    while (*e == ' ' || *e == '\t') ++e;
                    if (*e == '\0' || *e == '\n' || *e == '\r' || *e == '#') break;
                    Key k{-1, -1, -1};
                    k.v = ResolveIndex(strtol(e, &e, 10), file_v.size());
                    if (*e == '/') {
                        ++e;
                        if (*e != '/') k.vt = ResolveIndex(strtol(e, &e, 10), file_vt.size());
                        if (*e == '/') {
                            ++e;
                            k.vn = ResolveIndex(strtol(e, &e, 10), file_vn.size());
                        }
                    }
                    if (k.v < 0 || k.v >= static_cast<int>(file_v.size())) {
                        fclose(f);
                        error_ = "bad vertex index in '" + path + "'";
                        return false;
                    }
                    poly.push_back(k);
                }
                for (size_t i = 1; i + 1 < poly.size(); ++i) { // aiProcess_Triangulate (fan)
                    corners.push_back(poly[0]);
                    corners.push_back(poly[i]);
                    corners.push_back(poly[i + 1]);
                }
            }
        }
        fclose(f);
        return Assemble(path, flags, file_v, file_vn, file_vt, corners);
    }

    // Stanford PLY, ASCII flavour (resources/scene/box/models/bun_zipper_1.ply): `element vertex` with x y z and optionally
    // nx ny nz / s t (or u v) among its properties, `element face` as index lists.  Per-vertex attributes: a corner's three
    // indices coincide.
    bool LoadPly(const std::string &path, unsigned int flags) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) {
            error_ = "cannot open '" + path + "'";
            return false;
        }
        std::vector<char> line(1 << 16);
        size_t num_vertices = 0, num_faces = 0;
        std::vector<std::string> vertex_props;
        int element = 0; // 1 = vertex, 2 = face, 3 = something else
        bool ascii = false, header_done = false;
        std::vector<std::pair<int, size_t>> element_order; // (kind, count) in file order
        while (fgets(line.data(), static_cast<int>(line.size()), f)) {
            char word[64] = "", a[64] = "", b[64] = "", c[64] = "";
            const int n = sscanf(line.data(), "%63s %63s %63s %63s", word, a, b, c);
            if (n < 1) continue;
            if (!strcmp(word, "format")) ascii = !strcmp(a, "ascii");
            else if (!strcmp(word, "element")) {
                const size_t count = static_cast<size_t>(strtoull(b, nullptr, 10));
                element = !strcmp(a, "vertex") ? 1 : (!strcmp(a, "face") ? 2 : 3);
                if (element == 1) num_vertices = count;
                if (element == 2) num_faces = count;
                element_order.push_back({element, count});
            } else if (!strcmp(word, "property") && element == 1 && n >= 3) vertex_props.push_back(b);
            else if (!strcmp(word, "end_header")) {
                header_done = true;
                break;
            }
        }
        if (!header_done || !ascii) {
            fclose(f);
            error_ = "assimp shim: '" + path + "' is not an ASCII PLY file";
            return false;
        }
        auto column = [&](const char *name, const char *alt = nullptr) {
            for (size_t i = 0; i < vertex_props.size(); ++i)
                if (vertex_props[i] == name || (alt && vertex_props[i] == alt)) return static_cast<int>(i);
            return -1;
        };
        const int cx = column("x"), cy = column("y"), cz = column("z"), cnx = column("nx"), cny = column("ny"), cnz = column("nz"),
                  cu = column("s", "u"), cv = column("t", "v");
        if (cx < 0 || cy < 0 || cz < 0) {
            fclose(f);
            error_ = "assimp shim: no x / y / z vertex properties in '" + path + "'";
            return false;
        }
        std::vector<V3> file_v, file_vn, file_vt;
        std::vector<Key> corners;
        std::vector<float> row(vertex_props.size());
        for (const auto &el : element_order) {
            for (size_t i = 0; i < el.second; ++i) {
                if (!fgets(line.data(), static_cast<int>(line.size()), f)) {
                    fclose(f);
                    error_ = "assimp shim: '" + path + "' ends early";
                    return false;
                }
                char *e = line.data();
                if (el.first == 1) {
                    for (float &x : row) x = strtof(e, &e);
                    file_v.push_back({row[cx], row[cy], row[cz]});
                    if (cnx >= 0 && cny >= 0 && cnz >= 0) file_vn.push_back({row[cnx], row[cny], row[cnz]});
                    if (cu >= 0 && cv >= 0) file_vt.push_back({row[cu], row[cv], 0.0f});
                } else if (el.first == 2) {
                    const long count = strtol(e, &e, 10);
                    std::vector<int> poly;
                    for (long k = 0; k < count; ++k) poly.push_back(static_cast<int>(strtol(e, &e, 10)));
                    for (size_t k = 1; k + 1 < poly.size(); ++k) // aiProcess_Triangulate (fan)
                        for (int id : {poly[0], poly[k], poly[k + 1]}) {
                            if (id < 0 || id >= static_cast<int>(num_vertices)) {
                                fclose(f);
                                error_ = "bad vertex index in '" + path + "'";
                                return false;
                            }
                            corners.push_back({id, file_vt.empty() ? -1 : id, file_vn.empty() ? -1 : id});
                        }
                }
            }
        }
        fclose(f);
        (void)num_faces;
        return Assemble(path, flags, file_v, file_vn, file_vt, corners);
    }

    // The post-process steps the reference asks for (model_loader.cpp:512-520), shared by both readers.
    bool Assemble(const std::string &path, unsigned int flags, const std::vector<V3> &file_v, const std::vector<V3> &file_vn,
                  const std::vector<V3> &file_vt, const std::vector<Key> &corners) {
        if (corners.empty()) {
            error_ = "no faces in '" + path + "'";
            return false;
        }

        bool has_vt = !file_vt.empty(), has_vn = !file_vn.empty();
        for (const Key &k : corners) {
            has_vt = has_vt && k.vt >= 0 && k.vt < static_cast<int>(file_vt.size());
            has_vn = has_vn && k.vn >= 0 && k.vn < static_cast<int>(file_vn.size());
        }

        // aiProcess_GenSmoothNormals: average of unit face normals per position.
        std::vector<V3> smooth;
        const bool gen_normals = !has_vn && (flags & aiProcess_GenSmoothNormals);
        if (gen_normals) {
            smooth.assign(file_v.size(), V3{0, 0, 0});
            for (size_t t = 0; t + 2 < corners.size(); t += 3) {
                V3 n = Cross(Sub(file_v[corners[t + 1].v], file_v[corners[t].v]),
                             Sub(file_v[corners[t + 2].v], file_v[corners[t].v]));
                if (!NormalizeSafe(&n)) continue;
                for (int c = 0; c < 3; ++c) {
                    V3 &acc = smooth[corners[t + c].v];
                    acc = {acc.x + n.x, acc.y + n.y, acc.z + n.z};
                }
            }
            for (V3 &n : smooth)
                if (!NormalizeSafe(&n)) n = {0, 0, 1};
        }

        // Join identical corners.
        std::unordered_map<Key, unsigned int, KeyHash> lookup;
        lookup.reserve(corners.size());
        positions_.clear(); normals_.clear(); uvs_.clear(); indices_.clear();
        for (Key k : corners) {
            if (!has_vt) k.vt = -1;
            if (!has_vn) k.vn = -1;
            auto it = lookup.find(k);
            if (it == lookup.end()) {
                const unsigned int id = static_cast<unsigned int>(positions_.size());
                it = lookup.emplace(k, id).first;
                positions_.push_back({file_v[k.v].x, file_v[k.v].y, file_v[k.v].z});
                if (has_vn) normals_.push_back({file_vn[k.vn].x, file_vn[k.vn].y, file_vn[k.vn].z});
                else if (gen_normals) normals_.push_back({smooth[k.v].x, smooth[k.v].y, smooth[k.v].z});
                if (has_vt) {
                    const float v = (flags & aiProcess_FlipUVs) ? 1.0f - file_vt[k.vt].y : file_vt[k.vt].y;
                    uvs_.push_back({file_vt[k.vt].x, v, 0.0f});
                }
            }
            indices_.push_back(it->second);
        }

        // aiProcess_CalcTangentSpace
        tangents_.clear(); bitangents_.clear();
        if ((flags & aiProcess_CalcTangentSpace) && !uvs_.empty() && !normals_.empty()) {
            std::vector<V3> tan(positions_.size(), V3{0, 0, 0}), bit(positions_.size(), V3{0, 0, 0});
            for (size_t t = 0; t + 2 < indices_.size(); t += 3) {
                const unsigned int i0 = indices_[t], i1 = indices_[t + 1], i2 = indices_[t + 2];
                const V3 p0{positions_[i0].x, positions_[i0].y, positions_[i0].z},
                         p1{positions_[i1].x, positions_[i1].y, positions_[i1].z},
                         p2{positions_[i2].x, positions_[i2].y, positions_[i2].z};
                const V3 v = Sub(p1, p0), w = Sub(p2, p0);
                float sx = uvs_[i1].x - uvs_[i0].x, sy = uvs_[i1].y - uvs_[i0].y;
                float tx = uvs_[i2].x - uvs_[i0].x, ty = uvs_[i2].y - uvs_[i0].y;
                const float dir = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
                if (sx * ty == sy * tx) { sx = 0.0f; sy = 1.0f; tx = 1.0f; ty = 0.0f; } // degenerate UVs
                const V3 ft = Scale(Sub(Scale(w, sy), Scale(v, ty)), dir);
                const V3 fb = Scale(Sub(Scale(w, sx), Scale(v, tx)), dir);
                for (unsigned int id : {i0, i1, i2}) {
                    tan[id] = {tan[id].x + ft.x, tan[id].y + ft.y, tan[id].z + ft.z};
                    bit[id] = {bit[id].x + fb.x, bit[id].y + fb.y, bit[id].z + fb.z};
                }
            }
            tangents_.resize(positions_.size());
            bitangents_.resize(positions_.size());
            for (size_t i = 0; i < positions_.size(); ++i) {
                V3 n{normals_[i].x, normals_[i].y, normals_[i].z};
                if (!NormalizeSafe(&n)) n = {0, 0, 1};
                V3 t = Sub(tan[i], Scale(n, Dot(tan[i], n)));
                if (!NormalizeSafe(&t)) t = AnyPerpendicular(n);
                V3 b = Sub(Sub(bit[i], Scale(n, Dot(bit[i], n))), Scale(t, Dot(bit[i], t)));
                if (!NormalizeSafe(&b)) { b = Cross(n, t); NormalizeSafe(&b); }
                tangents_[i] = {t.x, t.y, t.z};
                bitangents_[i] = {b.x, b.y, b.z};
            }
        }
        return true;
    }

    void Publish() {
        faces_.resize(indices_.size() / 3);
        for (size_t i = 0; i < faces_.size(); ++i) {
            faces_[i].mNumIndices = 3;
            faces_[i].mIndices = indices_.data() + 3 * i;
        }
        memset(&mesh_, 0, sizeof(mesh_));
        mesh_.mNumVertices = static_cast<unsigned int>(positions_.size());
        mesh_.mNumFaces = static_cast<unsigned int>(faces_.size());
        mesh_.mVertices = positions_.data();
        mesh_.mNormals = normals_.empty() ? nullptr : normals_.data();
        mesh_.mTangents = tangents_.empty() ? nullptr : tangents_.data();
        mesh_.mBitangents = bitangents_.empty() ? nullptr : bitangents_.data();
        mesh_.mTextureCoords[0] = uvs_.empty() ? nullptr : uvs_.data();
        mesh_.mFaces = faces_.data();

        mesh_ptr_ = &mesh_;
        mesh_index_ = 0;
        child_.mNumMeshes = 1;
        child_.mMeshes = &mesh_index_;
        child_.mNumChildren = 0;
        child_.mChildren = nullptr;
        child_ptr_ = &child_;
        root_.mNumMeshes = 0;
        root_.mMeshes = nullptr;
        root_.mNumChildren = 1;
        root_.mChildren = &child_ptr_;
        scene_.mFlags = 0;
        scene_.mRootNode = &root_;
        scene_.mNumMeshes = 1;
        scene_.mMeshes = &mesh_ptr_;
    }

    std::string error_;
    std::vector<aiVector3D> positions_, normals_, uvs_, tangents_, bitangents_;
    std::vector<unsigned int> indices_;
    std::vector<aiFace> faces_;
    aiMesh mesh_;
    aiMesh *mesh_ptr_ = nullptr;
    unsigned int mesh_index_ = 0;
    aiNode root_{}, child_{};
    aiNode *child_ptr_ = nullptr;
    aiScene scene_;
};

} // namespace Assimp

#endif
