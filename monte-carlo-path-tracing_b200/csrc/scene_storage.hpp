// scene_storage.hpp — owning storage behind a b200pt_scene_desc.
// Used by the pack loader (scene_pack.cpp) and by the reference-side glue
// (host/csrt_glue.hpp) that flattens a csrt::RendererConfig.
#pragma once
#include <vector>

#include "b200pt.h"

struct b200pt_scene {
    b200pt_scene_desc desc{};
    std::vector<b200pt_texture> textures;
    std::vector<float> pixels;
    std::vector<b200pt_bsdf> bsdfs;
    std::vector<b200pt_medium> media;
    std::vector<b200pt_instance> instances;
    std::vector<b200pt_emitter> emitters;
    std::vector<float> positions, normals, texcoords, tangents, bitangents;
    std::vector<uint32_t> indices;

    // Points desc at the vectors; call after the vectors stop changing.
    void Finalize() {
        b200pt_scene_desc &d = desc;
        d.abi_version = B200PT_ABI_VERSION;
        d.num_textures = textures.size();   d.textures = textures.data();
        d.num_pixels = pixels.size();       d.pixels = pixels.data();
        d.num_bsdfs = bsdfs.size();         d.bsdfs = bsdfs.data();
        d.num_media = media.size();         d.media = media.data();
        d.num_instances = instances.size(); d.instances = instances.data();
        d.num_emitters = emitters.size();   d.emitters = emitters.data();
        d.num_positions = positions.size() / 3;   d.positions = positions.data();
        d.num_normals = normals.size() / 3;       d.normals = normals.data();
        d.num_texcoords = texcoords.size() / 2;   d.texcoords = texcoords.data();
        d.num_tangents = tangents.size() / 3;     d.tangents = tangents.data();
        d.num_bitangents = bitangents.size() / 3; d.bitangents = bitangents.data();
        d.num_triangles = indices.size() / 3;     d.indices = indices.data();
    }
};
