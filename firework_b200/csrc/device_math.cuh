// Device-side f32 vector helpers and the counter-based RNG.
//
// This translation unit is compiled with --fmad=false: the reference is Rust, which never contracts
// a*b+c into a fused multiply-add, and first-hit object ids / t must match the reference semantics
// bit for bit (DESIGN.md "Arithmetic fidelity").  Every helper below spells out the reference's
// operation order (ultraviolet 0.5.1 Vec3: component-wise f32, dot = (x*x')+(y*y')+(z*z')).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace fw {

#define FW_DEV __device__ __forceinline__

FW_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
FW_DEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
FW_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
FW_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
FW_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
FW_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
FW_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
FW_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
FW_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
FW_DEV float dot3(float3 a, float3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
FW_DEV float mag_sq3(float3 a) { return (a.x * a.x) + (a.y * a.y) + (a.z * a.z); }
FW_DEV float mag3(float3 a) { return sqrtf(mag_sq3(a)); }
FW_DEV float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
FW_DEV float3 normalized3(float3 a) {
    float m = mag3(a);
    return f3(a.x / m, a.y / m, a.z / m);
}
FW_DEV float comp3(float3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// Mat3 (three column float4s) * Vec3
FW_DEV float3 mat_mul(float4 c0, float4 c1, float4 c2, float3 v) {
    return f3(c0.x * v.x + c1.x * v.y + c2.x * v.z, c0.y * v.x + c1.y * v.y + c2.y * v.z,
              c0.z * v.x + c1.z * v.y + c2.z * v.z);
}

constexpr float FW_PI = 3.14159265358979323846f;
constexpr float FW_FLT_MAX = 3.40282347e+38f;

// ---- Philox4x32-10 (Salmon et al., SC'11), counter = (block, bounce<<8 | kind, sample, pixel) -----------
enum StreamKind : uint32_t { STREAM_CAMERA = 0, STREAM_SCATTER = 1, STREAM_MEDIUM = 2 };

FW_DEV uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
FW_DEV float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

struct RngKey {  // identifies the path vertex whose draws are being made
    uint2 seed;
    uint32_t pixel, sample, bounce;
};

// Sequential uniforms of one (pixel, sample, bounce, kind) stream: draw i = lane i%4 of block i/4.
struct PhiloxStream {
    RngKey key;
    uint32_t kind, idx;
    uint4 buf;
    FW_DEV PhiloxStream(const RngKey& k, uint32_t kind_) : key(k), kind(kind_), idx(0) {}
    FW_DEV float next() {
        uint32_t lane = idx & 3u;
        if (lane == 0) buf = philox4x32_10(make_uint4(idx >> 2, (key.bounce << 8) | kind, key.sample, key.pixel), key.seed);
        ++idx;
        uint32_t v = lane == 0 ? buf.x : (lane == 1 ? buf.y : (lane == 2 ? buf.z : buf.w));
        return u32_to_unit(v);
    }
};
// The free-path draw of ConstantMedium object `id` (volume.rs:67): order-independent keyed draw.
FW_DEV float philox_medium_draw(const RngKey& k, uint32_t id) {
    uint4 o = philox4x32_10(make_uint4(id, (k.bounce << 8) | (uint32_t)STREAM_MEDIUM, k.sample, k.pixel), k.seed);
    return u32_to_unit(o.x);
}
// Explicit uniforms (the single-scatter-step parity gate).
struct ArrayStream {
    const float* u;
    int n, i;
    bool overrun;
    FW_DEV ArrayStream(const float* u_, int n_) : u(u_), n(n_), i(0), overrun(false) {}
    FW_DEV float next() {
        if (i >= n) { overrun = true; return 0.5f; }
        return u[i++];
    }
};

}  // namespace fw
