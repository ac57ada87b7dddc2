// csrt_glue.hpp — reference-side binding: csrt::RendererConfig <-> b200pt_scene_desc.
//
// This is the stub a maintainer of the reference adds so that
// csrt::Renderer::Draw (src/renderer/renderer.cpp:678) can forward to libb200pt
// (see INTEGRATION.md).  It is compiled AGAINST the reference's headers
// (-I/root/reference/include) and is therefore only built where the reference
// tree exists: by oracle/Makefile (which uses it to turn parsed XML scenes into
// scene packs and to feed packs back into the reference's CPU renderer).
//
// Every field copy cites the reference struct it mirrors:
//   Camera::Info       include/csrt/renderer/camera.hpp:12-21
//   IntegratorInfo     include/csrt/renderer/integrators/integrator.hpp:16-27
//   TextureInfo        include/csrt/renderer/textures/texture.hpp:21-27
//   BsdfInfo           include/csrt/renderer/bsdfs/bsdf.hpp:40-58
//   MediumInfo         include/csrt/renderer/medium/medium.hpp:40-45
//   InstanceInfo       include/csrt/rtcore/instance.hpp:24-51
//   EmitterInfo        include/csrt/renderer/emitters/emitter.hpp:30-47
#pragma once

#include <cstring>
#include <memory>

#include "csrt/renderer/renderer.hpp"

#include "b200pt.h"
#include "../csrc/scene_storage.hpp"

namespace b200pt_glue {

inline void CopyMat4(const csrt::Mat4 &m, float *out) {
    for (int r = 0; r < 4; ++r) {
        out[r * 4 + 0] = m.rows[r].x;
        out[r * 4 + 1] = m.rows[r].y;
        out[r * 4 + 2] = m.rows[r].z;
        out[r * 4 + 3] = m.rows[r].w;
    }
}

inline csrt::Mat4 ToMat4(const float *m) {
    return csrt::Mat4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13],
                      m[14], m[15]);
}

inline void CopyVec3(const csrt::Vec3 &v, float *out) { out[0] = v.x, out[1] = v.y, out[2] = v.z; }
inline csrt::Vec3 ToVec3(const float *v) { return csrt::Vec3(v[0], v[1], v[2]); }

// csrt::RendererConfig -> owning flat scene (what b200pt_create consumes).
inline std::unique_ptr<b200pt_scene> FlattenConfig(const csrt::RendererConfig &cfg) {
    std::unique_ptr<b200pt_scene> s(new b200pt_scene());

    b200pt_camera &cam = s->desc.camera;
    cam.spp = cfg.camera.spp;
    cam.width = cfg.camera.width;
    cam.height = cfg.camera.height;
    cam.fov_x = cfg.camera.fov_x;
    CopyVec3(cfg.camera.eye, cam.eye);
    CopyVec3(cfg.camera.look_at, cam.look_at);
    CopyVec3(cfg.camera.up, cam.up);

    b200pt_integrator &in = s->desc.integrator;
    in.type = static_cast<uint32_t>(cfg.integrator.type);
    in.hide_emitters = cfg.integrator.hide_emitters ? 1u : 0u;
    in.pdf_rr = cfg.integrator.pdf_rr;
    in.depth_rr = cfg.integrator.depth_rr;
    in.depth_max = cfg.integrator.depth_max;

    for (const csrt::TextureInfo &t : cfg.textures) {
        b200pt_texture o;
        memset(&o, 0, sizeof(o));
        o.type = static_cast<uint32_t>(t.type);
        CopyMat4(csrt::Mat4(), o.to_uv);
        switch (t.type) {
        case csrt::TextureType::kConstant:
            CopyVec3(t.constant.color, o.color0);
            break;
        case csrt::TextureType::kCheckerboard:
            CopyVec3(t.checkerboard.color0, o.color0);
            CopyVec3(t.checkerboard.color1, o.color1);
            CopyMat4(t.checkerboard.to_uv, o.to_uv);
            break;
        case csrt::TextureType::kBitmap:
            o.width = t.bitmap.width;
            o.height = t.bitmap.height;
            o.channels = t.bitmap.channel;
            CopyMat4(t.bitmap.to_uv, o.to_uv);
            o.pixel_offset = s->pixels.size();
            s->pixels.insert(s->pixels.end(), t.bitmap.data.begin(), t.bitmap.data.end());
            break;
        default:
            break;
        }
        s->textures.push_back(o);
    }

    for (const csrt::BsdfInfo &b : cfg.bsdfs) {
        b200pt_bsdf o;
        memset(&o, 0, sizeof(o));
        o.type = static_cast<uint32_t>(b.type);
        o.twosided = b.twosided ? 1u : 0u;
        o.id_opacity = b.id_opacity;
        o.id_bump_map = b.id_bump_map;
        o.id_radiance = o.id_diffuse_reflectance = o.id_roughness_u = o.id_roughness_v =
            o.id_specular_reflectance = o.id_specular_transmittance = B200PT_INVALID_ID;
        o.eta = 1.0f;
        o.area_light_weight = 1.0f;
        switch (b.type) {
        case csrt::BsdfType::kAreaLight:
            o.id_radiance = b.area_light.id_radiance;
            o.area_light_weight = b.area_light.weight;
            break;
        case csrt::BsdfType::kDiffuse:
            o.id_diffuse_reflectance = b.diffuse.id_diffuse_reflectance;
            break;
        case csrt::BsdfType::kRoughDiffuse:
            o.id_diffuse_reflectance = b.rough_diffuse.id_diffuse_reflectance;
            o.id_roughness_u = o.id_roughness_v = b.rough_diffuse.id_roughness;
            o.use_fast_approx = b.rough_diffuse.use_fast_approx ? 1u : 0u;
            break;
        case csrt::BsdfType::kConductor:
            o.id_roughness_u = b.conductor.id_roughness_u;
            o.id_roughness_v = b.conductor.id_roughness_v;
            o.id_specular_reflectance = b.conductor.id_specular_reflectance;
            CopyVec3(b.conductor.reflectivity, o.reflectivity);
            CopyVec3(b.conductor.edgetint, o.edgetint);
            break;
        case csrt::BsdfType::kDielectric:
        case csrt::BsdfType::kThinDielectric:
            o.id_roughness_u = b.dielectric.id_roughness_u;
            o.id_roughness_v = b.dielectric.id_roughness_v;
            o.id_specular_reflectance = b.dielectric.id_specular_reflectance;
            o.id_specular_transmittance = b.dielectric.id_specular_transmittance;
            o.eta = b.dielectric.eta;
            break;
        case csrt::BsdfType::kPlastic:
            o.id_roughness_u = o.id_roughness_v = b.plastic.id_roughness;
            o.id_diffuse_reflectance = b.plastic.id_diffuse_reflectance;
            o.id_specular_reflectance = b.plastic.id_specular_reflectance;
            o.eta = b.plastic.eta;
            break;
        default:
            break;
        }
        s->bsdfs.push_back(o);
    }

    for (const csrt::MediumInfo &m : cfg.media) {
        b200pt_medium o;
        memset(&o, 0, sizeof(o));
        o.type = static_cast<uint32_t>(m.type);
        CopyVec3(m.homogeneous.sigma_a, o.sigma_a);
        CopyVec3(m.homogeneous.sigma_s, o.sigma_s);
        o.phase_type = static_cast<uint32_t>(m.phase_func.type);
        CopyVec3(m.phase_func.g, o.g);
        s->media.push_back(o);
    }

    for (const csrt::InstanceInfo &i : cfg.instances) {
        b200pt_instance o;
        memset(&o, 0, sizeof(o));
        o.type = static_cast<uint32_t>(i.type);
        o.id_bsdf = i.id_bsdf;
        o.id_medium_int = i.id_medium_int;
        o.id_medium_ext = i.id_medium_ext;
        o.flip_normals = i.flip_normals ? 1u : 0u;
        CopyMat4(i.to_world, o.to_world);
        o.sphere_radius = i.sphere.radius;
        CopyVec3(i.sphere.center, o.sphere_center);
        o.cylinder_radius = i.cylinder.radius;
        CopyVec3(i.cylinder.p0, o.cylinder_p0);
        CopyVec3(i.cylinder.p1, o.cylinder_p1);
        const csrt::MeshesInfo &m = i.meshes;
        o.num_vertices = m.positions.size();
        o.num_triangles = m.indices.size();
        o.position_offset = o.normal_offset = o.texcoord_offset = o.tangent_offset = o.bitangent_offset =
            B200PT_NO_OFFSET;
        o.index_offset = s->indices.size() / 3;
        if (!m.positions.empty()) {
            o.position_offset = s->positions.size() / 3;
            for (const csrt::Vec3 &v : m.positions) s->positions.insert(s->positions.end(), {v.x, v.y, v.z});
        }
        if (!m.normals.empty()) {
            o.normal_offset = s->normals.size() / 3;
            for (const csrt::Vec3 &v : m.normals) s->normals.insert(s->normals.end(), {v.x, v.y, v.z});
        }
        if (!m.texcoords.empty()) {
            o.texcoord_offset = s->texcoords.size() / 2;
            for (const csrt::Vec2 &v : m.texcoords) s->texcoords.insert(s->texcoords.end(), {v.u, v.v});
        }
        if (!m.tangents.empty()) {
            o.tangent_offset = s->tangents.size() / 3;
            for (const csrt::Vec3 &v : m.tangents) s->tangents.insert(s->tangents.end(), {v.x, v.y, v.z});
        }
        if (!m.bitangents.empty()) {
            o.bitangent_offset = s->bitangents.size() / 3;
            for (const csrt::Vec3 &v : m.bitangents) s->bitangents.insert(s->bitangents.end(), {v.x, v.y, v.z});
        }
        for (const csrt::Uvec3 &t : m.indices) s->indices.insert(s->indices.end(), {t.x, t.y, t.z});
        s->instances.push_back(o);
    }

    for (const csrt::EmitterInfo &e : cfg.emitters) {
        b200pt_emitter o;
        memset(&o, 0, sizeof(o));
        o.type = static_cast<uint32_t>(e.type);
        o.id_texture = B200PT_INVALID_ID;
        CopyMat4(csrt::Mat4(), o.to_world);
        switch (e.type) {
        case csrt::EmitterType::kPoint:
            CopyVec3(e.point.position, o.position);
            CopyVec3(e.point.intensity, o.radiance);
            break;
        case csrt::EmitterType::kSpot:
            o.cutoff_angle = e.spot.cutoff_angle;
            o.beam_width = e.spot.beam_width;
            o.id_texture = e.spot.id_texture;
            CopyVec3(e.spot.intensity, o.radiance);
            CopyMat4(e.spot.to_world, o.to_world);
            break;
        case csrt::EmitterType::kDirectional:
            CopyVec3(e.directional.direction, o.direction);
            CopyVec3(e.directional.radiance, o.radiance);
            break;
        case csrt::EmitterType::kSun:
            o.cos_cutoff_angle = e.sun.cos_cutoff_angle;
            o.id_texture = e.sun.id_texture;
            CopyVec3(e.sun.direction, o.direction);
            CopyVec3(e.sun.radiance, o.radiance);
            break;
        case csrt::EmitterType::kEnvMap:
            o.id_texture = e.envmap.id_radiance;
            CopyMat4(e.envmap.to_world, o.to_world);
            break;
        case csrt::EmitterType::kConstant:
            CopyVec3(e.constant.radiance, o.radiance);
            break;
        default:
            break;
        }
        s->emitters.push_back(o);
    }

    s->Finalize();
    return s;
}

// b200pt_scene_desc -> csrt::RendererConfig (feeds a pack back into the reference renderer).
inline csrt::RendererConfig InflateScene(const b200pt_scene_desc &d) {
    csrt::RendererConfig cfg;
    cfg.backend_type = csrt::BackendType::kCpu;
    cfg.camera.spp = d.camera.spp;
    cfg.camera.width = d.camera.width;
    cfg.camera.height = d.camera.height;
    cfg.camera.fov_x = d.camera.fov_x;
    cfg.camera.eye = ToVec3(d.camera.eye);
    cfg.camera.look_at = ToVec3(d.camera.look_at);
    cfg.camera.up = ToVec3(d.camera.up);

    cfg.integrator.type = static_cast<csrt::IntegratorType>(d.integrator.type);
    cfg.integrator.hide_emitters = d.integrator.hide_emitters != 0;
    cfg.integrator.pdf_rr = d.integrator.pdf_rr;
    cfg.integrator.depth_rr = d.integrator.depth_rr;
    cfg.integrator.depth_max = d.integrator.depth_max;

    for (uint64_t i = 0; i < d.num_textures; ++i) {
        const b200pt_texture &t = d.textures[i];
        csrt::TextureInfo o;
        o.type = static_cast<csrt::TextureType>(t.type);
        switch (o.type) {
        case csrt::TextureType::kConstant:
            o.constant.color = ToVec3(t.color0);
            break;
        case csrt::TextureType::kCheckerboard:
            o.checkerboard.color0 = ToVec3(t.color0);
            o.checkerboard.color1 = ToVec3(t.color1);
            o.checkerboard.to_uv = ToMat4(t.to_uv);
            break;
        case csrt::TextureType::kBitmap: {
            o.bitmap.width = t.width;
            o.bitmap.height = t.height;
            o.bitmap.channel = t.channels;
            o.bitmap.to_uv = ToMat4(t.to_uv);
            const uint64_t n = static_cast<uint64_t>(t.width) * t.height * t.channels;
            o.bitmap.data.assign(d.pixels + t.pixel_offset, d.pixels + t.pixel_offset + n);
            break;
        }
        default:
            break;
        }
        cfg.textures.push_back(o);
    }

    for (uint64_t i = 0; i < d.num_bsdfs; ++i) {
        const b200pt_bsdf &b = d.bsdfs[i];
        csrt::BsdfInfo o;
        o.type = static_cast<csrt::BsdfType>(b.type);
        o.twosided = b.twosided != 0;
        o.id_opacity = b.id_opacity;
        o.id_bump_map = b.id_bump_map;
        switch (o.type) {
        case csrt::BsdfType::kAreaLight:
            o.area_light.id_radiance = b.id_radiance;
            o.area_light.weight = b.area_light_weight;
            break;
        case csrt::BsdfType::kDiffuse:
            o.diffuse.id_diffuse_reflectance = b.id_diffuse_reflectance;
            break;
        case csrt::BsdfType::kRoughDiffuse:
            o.rough_diffuse.use_fast_approx = b.use_fast_approx != 0;
            o.rough_diffuse.id_diffuse_reflectance = b.id_diffuse_reflectance;
            o.rough_diffuse.id_roughness = b.id_roughness_u;
            break;
        case csrt::BsdfType::kConductor:
            o.conductor.id_roughness_u = b.id_roughness_u;
            o.conductor.id_roughness_v = b.id_roughness_v;
            o.conductor.id_specular_reflectance = b.id_specular_reflectance;
            o.conductor.reflectivity = ToVec3(b.reflectivity);
            o.conductor.edgetint = ToVec3(b.edgetint);
            break;
        case csrt::BsdfType::kDielectric:
        case csrt::BsdfType::kThinDielectric:
            o.dielectric.id_roughness_u = b.id_roughness_u;
            o.dielectric.id_roughness_v = b.id_roughness_v;
            o.dielectric.id_specular_reflectance = b.id_specular_reflectance;
            o.dielectric.id_specular_transmittance = b.id_specular_transmittance;
            o.dielectric.eta = b.eta;
            break;
        case csrt::BsdfType::kPlastic:
            o.plastic.eta = b.eta;
            o.plastic.id_roughness = b.id_roughness_u;
            o.plastic.id_diffuse_reflectance = b.id_diffuse_reflectance;
            o.plastic.id_specular_reflectance = b.id_specular_reflectance;
            break;
        default:
            break;
        }
        cfg.bsdfs.push_back(o);
    }

    for (uint64_t i = 0; i < d.num_media; ++i) {
        const b200pt_medium &m = d.media[i];
        csrt::MediumInfo o;
        o.type = static_cast<csrt::MediumType>(m.type);
        o.homogeneous.sigma_a = ToVec3(m.sigma_a);
        o.homogeneous.sigma_s = ToVec3(m.sigma_s);
        o.phase_func.type = static_cast<csrt::PhaseFunctionType>(m.phase_type);
        o.phase_func.g = ToVec3(m.g);
        cfg.media.push_back(o);
    }

    for (uint64_t i = 0; i < d.num_instances; ++i) {
        const b200pt_instance &in = d.instances[i];
        csrt::InstanceInfo o;
        o.type = static_cast<csrt::InstanceType>(in.type);
        o.id_bsdf = in.id_bsdf;
        o.id_medium_int = in.id_medium_int;
        o.id_medium_ext = in.id_medium_ext;
        o.flip_normals = in.flip_normals != 0;
        o.to_world = ToMat4(in.to_world);
        o.sphere.radius = in.sphere_radius;
        o.sphere.center = ToVec3(in.sphere_center);
        o.cylinder.radius = in.cylinder_radius;
        o.cylinder.p0 = ToVec3(in.cylinder_p0);
        o.cylinder.p1 = ToVec3(in.cylinder_p1);
        const uint64_t nv = in.num_vertices;
        if (in.position_offset != B200PT_NO_OFFSET)
            for (uint64_t v = 0; v < nv; ++v) o.meshes.positions.push_back(ToVec3(d.positions + 3 * (in.position_offset + v)));
        if (in.normal_offset != B200PT_NO_OFFSET)
            for (uint64_t v = 0; v < nv; ++v) o.meshes.normals.push_back(ToVec3(d.normals + 3 * (in.normal_offset + v)));
        if (in.texcoord_offset != B200PT_NO_OFFSET)
            for (uint64_t v = 0; v < nv; ++v) {
                const float *uv = d.texcoords + 2 * (in.texcoord_offset + v);
                o.meshes.texcoords.push_back(csrt::Vec2(uv[0], uv[1]));
            }
        if (in.tangent_offset != B200PT_NO_OFFSET)
            for (uint64_t v = 0; v < nv; ++v) o.meshes.tangents.push_back(ToVec3(d.tangents + 3 * (in.tangent_offset + v)));
        if (in.bitangent_offset != B200PT_NO_OFFSET)
            for (uint64_t v = 0; v < nv; ++v)
                o.meshes.bitangents.push_back(ToVec3(d.bitangents + 3 * (in.bitangent_offset + v)));
        for (uint64_t t = 0; t < in.num_triangles; ++t) {
            const uint32_t *idx = d.indices + 3 * (in.index_offset + t);
            o.meshes.indices.push_back(csrt::Uvec3(idx[0], idx[1], idx[2]));
        }
        cfg.instances.push_back(o);
    }

    for (uint64_t i = 0; i < d.num_emitters; ++i) {
        const b200pt_emitter &e = d.emitters[i];
        csrt::EmitterInfo o;
        o.type = static_cast<csrt::EmitterType>(e.type);
        switch (o.type) {
        case csrt::EmitterType::kPoint:
            o.point.position = ToVec3(e.position);
            o.point.intensity = ToVec3(e.radiance);
            break;
        case csrt::EmitterType::kSpot:
            o.spot = csrt::SpotLightInfo();
            o.spot.cutoff_angle = e.cutoff_angle;
            o.spot.beam_width = e.beam_width;
            o.spot.id_texture = e.id_texture;
            o.spot.intensity = ToVec3(e.radiance);
            o.spot.to_world = ToMat4(e.to_world);
            break;
        case csrt::EmitterType::kDirectional:
            o.directional.direction = ToVec3(e.direction);
            o.directional.radiance = ToVec3(e.radiance);
            break;
        case csrt::EmitterType::kSun:
            o.sun.cos_cutoff_angle = e.cos_cutoff_angle;
            o.sun.id_texture = e.id_texture;
            o.sun.direction = ToVec3(e.direction);
            o.sun.radiance = ToVec3(e.radiance);
            break;
        case csrt::EmitterType::kEnvMap:
            o.envmap = csrt::EnvMapInfo();
            o.envmap.id_radiance = e.id_texture;
            o.envmap.to_world = ToMat4(e.to_world);
            break;
        case csrt::EmitterType::kConstant:
            o.constant.radiance = ToVec3(e.radiance);
            break;
        default:
            break;
        }
        cfg.emitters.push_back(o);
    }
    return cfg;
}

} // namespace b200pt_glue
