// textures.cuh — texture lookups (device): constant / checkerboard / bitmap colour, finite-difference gradient for
// bump maps, stochastic alpha test for opacity masks.  Restates src/renderer/textures/{texture,bitmap,checkboard,
// constant_texture}.cpp.
#pragma once
#include "b200pt.h"
#include "vecmath.cuh"

namespace b200pt {

// ---------------------------------------------------------------------------------------------
// Textures (textures/texture.cpp:63-113)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 BitmapColor(const DeviceScene &s, const DTexture &t, V2 uv) { // bitmap.cpp:6-56
    const V3 p = XformPoint(t.to_uv, mk3(uv.u, uv.v, 0.0f));
    float x = p.x * t.width, y = p.y * t.height;
    while (x < 0) x += t.width;
    while (x > t.width - 1) x -= t.width;
    while (y < 0) y += t.height;
    while (y > t.height - 1) y -= t.height;
    const uint32_t x0 = static_cast<uint32_t>(x), y0 = static_cast<uint32_t>(y);
    const float tx = x - x0, ty = y - y0;
    const uint32_t x1 = (tx > 0.0f) ? x0 + 1 : x0, y1 = (ty > 0.0f) ? y0 + 1 : y0;
    const float *data = s.pixels + t.pixel_offset;
    if (t.channels == 1) {
        const float c00 = __ldg(data + (x0 + t.width * y0)), c01 = __ldg(data + (x0 + t.width * y1)),
                    c10 = __ldg(data + (x1 + t.width * y0)), c11 = __ldg(data + (x1 + t.width * y1));
        return mk3(Lerp(Lerp(c00, c01, ty), Lerp(c10, c11, ty), tx));
    }
    auto px = [&](uint32_t xx, uint32_t yy) {
        const uint32_t o = (xx + t.width * yy) * t.channels;
        return mk3(__ldg(data + o), __ldg(data + o + 1), __ldg(data + o + 2));
    };
    const V3 c0 = Lerp(px(x0, y0), px(x0, y1), ty), c1 = Lerp(px(x1, y0), px(x1, y1), ty);
    return Lerp(c0, c1, tx);
}

__device__ __forceinline__ V3 CheckerboardColor(const DTexture &t, V2 uv) { // checkboard.cpp:6-21
    V3 p = XformPoint(t.to_uv, mk3(uv.u, uv.v, 0.0f));
    while (p.x > 1) p.x -= 1;
    while (p.x < 0) p.x += 1;
    while (p.y > 1) p.y -= 1;
    while (p.y < 0) p.y += 1;
    const int x = 2 * (static_cast<int>(p.x * 2) % 2) - 1, y = 2 * (static_cast<int>(p.y * 2) % 2) - 1;
    return (x * y == 1) ? mk3(t.color0) : mk3(t.color1);
}

__device__ __forceinline__ V3 TexColor(const DeviceScene &s, uint32_t id, V2 uv) {
    const DTexture &t = s.textures[id];
    switch (t.type) {
    case B200PT_TEX_CONSTANT: return mk3(t.color0);
    case B200PT_TEX_CHECKERBOARD: return CheckerboardColor(t, uv);
    case B200PT_TEX_BITMAP: return BitmapColor(s, t, uv);
    }
    return mk3(0.0f);
}

// texture.cpp:79-95, bitmap.cpp:58-68, checkboard.cpp:23-33 (constant textures have zero gradient)
__device__ __forceinline__ V2 TexGradient(const DeviceScene &s, uint32_t id, V2 uv) {
    if (s.textures[id].type == B200PT_TEX_CONSTANT) return {0.0f, 0.0f};
    constexpr float delta = 1e-4f, norm = 1.0f / delta;
    const float value = Length(TexColor(s, id, uv)), value_u = Length(TexColor(s, id, {uv.u + delta, uv.v})),
                value_v = Length(TexColor(s, id, {uv.u, uv.v + delta}));
    return {(value_u - value) * norm, (value_v - value) * norm};
}


// texture.cpp:97-113: IsTransparent — constant: color.x < xi (constant_texture.cpp:18-23); bitmap: bilinear alpha of a
// 4-channel image < xi (bitmap.cpp:70-100); checkerboard: never.
__device__ __forceinline__ bool TexIsTransparent(const DeviceScene &s, uint32_t id, V2 uv, float xi) {
    const DTexture &t = s.textures[id];
    if (t.type == B200PT_TEX_CONSTANT) return t.color0.x < xi;
    if (t.type != B200PT_TEX_BITMAP || t.channels != 4) return false;
    const V3 p = XformPoint(t.to_uv, mk3(uv.u, uv.v, 0.0f));
    float x = p.x * t.width, y = p.y * t.height;
    while (x < 0) x += t.width;
    while (x > t.width - 1) x -= t.width;
    while (y < 0) y += t.height;
    while (y > t.height - 1) y -= t.height;
    const uint32_t x0 = static_cast<uint32_t>(x), y0 = static_cast<uint32_t>(y);
    const float tx = x - x0, ty = y - y0;
    const uint32_t x1 = (tx > 0.0f) ? x0 + 1 : x0, y1 = (ty > 0.0f) ? y0 + 1 : y0;
    const float *data = s.pixels + t.pixel_offset;
    const float c00 = __ldg(data + (x0 + t.width * y0) * 4 + 3), c01 = __ldg(data + (x0 + t.width * y1) * 4 + 3),
                c10 = __ldg(data + (x1 + t.width * y0) * 4 + 3), c11 = __ldg(data + (x1 + t.width * y1) * 4 + 3);
    return Lerp(Lerp(c00, c01, ty), Lerp(c10, c11, ty), tx) < xi;
}

} // namespace b200pt
