// vkrt-f32 arithmetic for the device side.
//
// Contract (DESIGN.md "Arithmetic"): every operation is ONE IEEE-754 binary32 round-to-nearest
// operation; a fused multiply-add happens only where fmaf()/fma_() is written; sin/cos/exp2/
// log2/pow are the polynomial routines below.  The translation unit must be compiled with
//   -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
// so that nvcc neither contracts a*b+c nor substitutes approximate div/sqrt.  The build
// defines VKRT_FMAD_OFF next to those flags; refuse to compile without it.
#pragma once
#ifndef VKRT_FMAD_OFF
#error "compile with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -DVKRT_FMAD_OFF"
#endif
#include <cuda_runtime.h>
#include <stdint.h>

namespace vkrt {

#define VKRT_DEV __device__ __forceinline__

VKRT_DEV float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VKRT_DEV uint32_t f2u(float f) { return __float_as_uint(f); }
VKRT_DEV float u2f(uint32_t u) { return __uint_as_float(u); }

// GLSL max/min with "a NaN x yields y" (what FMNMX does); written with compares so that the
// sign of a zero result is the same on every implementation of this contract.
VKRT_DEV float gl_max(float x, float y) { return (x != x) ? y : (x < y ? y : x); }
VKRT_DEV float gl_min(float x, float y) { return (x != x) ? y : (y < x ? y : x); }
VKRT_DEV float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }
VKRT_DEV float gl_sign(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
VKRT_DEV float gl_abs(float x) { return u2f(f2u(x) & 0x7fffffffu); }
VKRT_DEV float gl_mix(float x, float y, float a) { return fma_(y, a, x * (1.0f - a)); }

#define VKRT_PI 3.14159265359f               /* Tracer.comp:176 */
#define VKRT_TWO_PI (2.0f * 3.14159265359f)  /* constant-folded in binary32 */

VKRT_DEV float rne(float x) { return (x + 12582912.0f) - 12582912.0f; }

VKRT_DEV void sincos_(float x, float &s, float &c)
{
    const float q = rne(x * 6.366197467e-01f);
    const int n = (int)q;
    float r = fma_(q, -1.570796371e+00f, x);
    r = fma_(q, 4.371138829e-08f, r);
    const float r2 = r * r;
    float sp = 2.723468469e-06f;
    sp = fma_(sp, r2, -1.983996626e-04f);
    sp = fma_(sp, r2, 8.333331905e-03f);
    sp = fma_(sp, r2, -1.666666716e-01f);
    const float sr = fma_(r * r2, sp, r);
    float cp = -2.728823461e-07f;
    cp = fma_(cp, r2, 2.480049989e-05f);
    cp = fma_(cp, r2, -1.388888806e-03f);
    cp = fma_(cp, r2, 4.166666791e-02f);
    const float cr = fma_(r2 * r2, cp, fma_(-0.5f, r2, 1.0f));
    const float s0 = (n & 1) ? cr : sr;
    const float c0 = (n & 1) ? sr : cr;
    s = (n & 2) ? -s0 : s0;
    c = ((n + 1) & 2) ? -c0 : c0;
}

VKRT_DEV float log2_(float x)
{
    uint32_t b = f2u(x);
    if (b == 0u || b == 0x80000000u) return __int_as_float(0xff800000);             // -inf
    if (b >> 31) return __int_as_float(0x7fc00000);                                // nan
    if (b >= 0x7f800000u) return x;
    float bias = 0.0f;
    if (b < 0x00800000u) { x = x * 8388608.0f; b = f2u(x); bias = -23.0f; }
    const int e = (int)(b - 0x3f2aaaabu) >> 23;
    const float m = u2f(b - ((uint32_t)e << 23));
    const float f = m - 1.0f;
    float p = 2.032371908e-01f;
    p = fma_(p, f, -2.222205549e-01f);
    p = fma_(p, f, 2.016990036e-01f);
    p = fma_(p, f, -2.367866337e-01f);
    p = fma_(p, f, 2.887182832e-01f);
    p = fma_(p, f, -3.607895672e-01f);
    p = fma_(p, f, 4.808958173e-01f);
    p = fma_(p, f, -7.213465571e-01f);
    p = fma_(p, f, 1.442695022e+00f);
    return fma_(p, f, (float)e + bias);
}

VKRT_DEV float exp2_(float x)
{
    if (x != x) return x;
    if (x >= 128.0f) return __int_as_float(0x7f800000);
    if (x < -150.0f) return 0.0f;
    const float nf = rne(x);
    const float f = x - nf;
    const int n = (int)nf;
    float p = 1.546973217e-04f;
    p = fma_(p, f, 1.340043265e-03f);
    p = fma_(p, f, 9.618025273e-03f);
    p = fma_(p, f, 5.550327152e-02f);
    p = fma_(p, f, 2.402265072e-01f);
    p = fma_(p, f, 6.931471825e-01f);
    p = fma_(p, f, 1.0f);
    const int n1 = n >> 1, n2 = n - n1;
    return (p * u2f((uint32_t)(n1 + 127) << 23)) * u2f((uint32_t)(n2 + 127) << 23);
}

VKRT_DEV float pow_(float x, float y) { return exp2_(y * log2_(x)); }

struct V3 { float x, y, z; };
VKRT_DEV V3 v3(float a) { return {a, a, a}; }
VKRT_DEV V3 v3(float a, float b, float c) { return {a, b, c}; }
VKRT_DEV V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
VKRT_DEV V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
VKRT_DEV V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
VKRT_DEV V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
VKRT_DEV V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
VKRT_DEV V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
// vec3 / scalar: one correctly rounded reciprocal, three multiplies (part of the arithmetic contract)
VKRT_DEV V3 operator/(V3 a, float s) { const float inv = 1.0f / s; return {a.x * inv, a.y * inv, a.z * inv}; }
VKRT_DEV V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
VKRT_DEV float dot3(V3 a, V3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
VKRT_DEV V3 cross3(V3 a, V3 b)
{
    return {fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x))};
}
VKRT_DEV float length3(V3 a) { return sqrtf(dot3(a, a)); }
VKRT_DEV V3 normalize3(V3 a) { const float inv = 1.0f / sqrtf(dot3(a, a)); return a * inv; }
VKRT_DEV V3 madd3(float t, V3 d, V3 o) { return {fma_(t, d.x, o.x), fma_(t, d.y, o.y), fma_(t, d.z, o.z)}; }
VKRT_DEV V3 reflect3(V3 I, V3 N) { const float k = 2.0f * dot3(N, I); return madd3(-k, N, I); }
VKRT_DEV V3 refract3(V3 I, V3 N, float eta)
{
    const float d = dot3(N, I);
    const float k = 1.0f - (eta * eta) * (1.0f - d * d);
    if (k < 0.0f) return v3(0.0f);
    const float s = fma_(eta, d, sqrtf(k));
    return madd3(-s, N, eta * I);
}
VKRT_DEV V3 clamp3(V3 a, float lo, float hi) { return {gl_clamp(a.x, lo, hi), gl_clamp(a.y, lo, hi), gl_clamp(a.z, lo, hi)}; }
VKRT_DEV V3 mix3(V3 a, V3 b, float t) { return {gl_mix(a.x, b.x, t), gl_mix(a.y, b.y, t), gl_mix(a.z, b.z, t)}; }
VKRT_DEV V3 xyz(float4 v) { return {v.x, v.y, v.z}; }

// ---- counter-based RNG (replaces Tracer.comp:221-234; DESIGN.md "RNG") -----------------------
__host__ __device__ inline uint32_t pcg(uint32_t v)
{
    const uint32_t state = v * 747796405u + 2891336453u;
    const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
__host__ inline uint32_t frame_key(uint64_t seed, uint32_t frame_seed_bits, uint32_t frame_index)
{
    const uint32_t a = pcg(frame_index + 0x9E3779B9u);
    const uint32_t b = pcg(frame_seed_bits ^ a);
    const uint32_t c = pcg((uint32_t)(seed >> 32) ^ b);
    return pcg((uint32_t)seed ^ c);
}
VKRT_DEV uint32_t sample_key(uint32_t fkey, uint32_t pixel, uint32_t sample)
{
    const uint32_t kp = pcg(pixel + fkey);
    return pcg(kp ^ (sample * 0x9E3779B9u));
}
VKRT_DEV float u01(uint32_t skey, uint32_t dim)
{
    return (float)(pcg(skey + dim) >> 8) * 5.9604644775390625e-08f;
}
enum { DIMS_PER_BOUNCE = 32, SLOT_R2 = 0, SLOT_PHI = 1, SLOT_RR = 2, SLOT_LIGHT = 3, MAX_LIGHTS = 14 };
#define VKRT_DITHER_SAMPLE 0xFFFFFFFFu

} // namespace vkrt
