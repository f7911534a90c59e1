/*
 * glsl_shim.h -- ORACLE support (test infrastructure, never shipped).
 *
 * Just enough of GLSL 3.30 as C++ for the reference's fragment shaders of the
 * tile-production path to compile UNCHANGED as C++ (oracle/Makefile, target
 * `ref`): upsampleShader.glsl, normalShader.glsl, upsampleOrthoShader.glsl and
 * their example variants.  The shader text itself is read from the reference
 * checkout at build time; nothing of it lives in this repository.
 *
 *   vec2/vec3/vec4 with the swizzles the shaders use, ivec4, bvec2, mat3, mat4
 *   (column-major, M[c] is column c), component-wise operators, and the
 *   built-ins floor fract mod abs min max clamp mix smoothstep dot cross
 *   length normalize sqrt equal all, each spelled as the GLSL 3.30 spec
 *   (section 8) defines it, one IEEE fp32 operation per GLSL operation, no
 *   contraction (the library is compiled with -ffp-contract=off and
 *   -fsingle-precision-constant so that `0.5` is a float as in GLSL).
 *
 *   textureLod() over caller-supplied float arrays, OpenGL 3.3 spec 3.8.8:
 *   NEAREST  texel floor(u * size)
 *   LINEAR   i0 = floor(u*size - 0.5), weights from frac(u*size - 0.5),
 *            either exact fp32 or quantised to 8 fractional bits (what GPUs'
 *            texture units do; the spec allows either) -- ref_sampler::subtexel_bits
 *   wrap     CLAMP_TO_EDGE or REPEAT
 *
 * Two choices are the shim's, not the shader text's (GLSL leaves them open):
 *   normalize(v) = v * (1 / sqrt(dot(v,v)))  (the inversesqrt form)
 *   M * v        = M[0]*v.x + M[1]*v.y + ... summed left to right per component
 */
#ifndef GLSL_SHIM_H
#define GLSL_SHIM_H

#include <cmath>
#include <cstddef>

namespace glsl {

struct vec2; struct vec3; struct vec4;

/* a swizzle view inside a vector's storage: N floats of the parent, picks I... */
template <class V, int N, int... I> struct Swz {
    float d[N];
    operator V() const { return V(d[I]...); }
};

struct vec2 {
    union {
        struct { float x, y; };
        struct { float r, g; };
        Swz<vec2, 2, 0, 1> xy;
        Swz<vec2, 2, 1, 0> yx;
        Swz<vec4, 2, 0, 1, 0, 1> xyxy;
    };
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        Swz<vec2, 3, 0, 1> xy;
        Swz<vec3, 3, 0, 1, 2> xyz;
        Swz<vec3, 3, 0, 1, 2> rgb;
        Swz<vec3, 3, 2, 2, 2> zzz;
    };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(const vec2 &v, float c) : x(v.x), y(v.y), z(c) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        Swz<vec2, 4, 0, 1> xy;
        Swz<vec2, 4, 2, 3> zw;
        Swz<vec2, 4, 2, 1> zy;
        Swz<vec2, 4, 0, 3> xw;
        Swz<vec2, 4, 2, 2> zz;
        Swz<vec3, 4, 0, 1, 2> xyz;
        Swz<vec3, 4, 0, 1, 2> rgb;
        Swz<vec3, 4, 0, 1, 3> xyw;
        Swz<vec4, 4, 0, 1, 0, 1> xyxy;
        Swz<vec4, 4, 2, 0, 2, 0> zxzx;
        Swz<vec4, 4, 3, 3, 1, 1> wwyy;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec2 &u, const vec2 &v) : x(u.x), y(u.y), z(v.x), w(v.y) {}
    vec4(const vec2 &u, float c, float d) : x(u.x), y(u.y), z(c), w(d) {}
    vec4(const vec3 &u, float d) : x(u.x), y(u.y), z(u.z), w(d) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

struct ivec4 {
    int x, y, z, w;
    ivec4() : x(0), y(0), z(0), w(0) {}
    ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
};
struct bvec2 { bool x, y; };

/* ---- component-wise operators (GLSL 3.30 spec 5.9) ---- */
#define GLSL_OPS2(OP) \
    inline vec2 operator OP(const vec2 &a, const vec2 &b) { return vec2(a.x OP b.x, a.y OP b.y); } \
    inline vec2 operator OP(const vec2 &a, float b) { return vec2(a.x OP b, a.y OP b); } \
    inline vec2 operator OP(float a, const vec2 &b) { return vec2(a OP b.x, a OP b.y); }
#define GLSL_OPS3(OP) \
    inline vec3 operator OP(const vec3 &a, const vec3 &b) { return vec3(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
    inline vec3 operator OP(const vec3 &a, float b) { return vec3(a.x OP b, a.y OP b, a.z OP b); } \
    inline vec3 operator OP(float a, const vec3 &b) { return vec3(a OP b.x, a OP b.y, a OP b.z); }
#define GLSL_OPS4(OP) \
    inline vec4 operator OP(const vec4 &a, const vec4 &b) { return vec4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
    inline vec4 operator OP(const vec4 &a, float b) { return vec4(a.x OP b, a.y OP b, a.z OP b, a.w OP b); } \
    inline vec4 operator OP(float a, const vec4 &b) { return vec4(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }
GLSL_OPS2(+) GLSL_OPS2(-) GLSL_OPS2(*) GLSL_OPS2(/)
GLSL_OPS3(+) GLSL_OPS3(-) GLSL_OPS3(*) GLSL_OPS3(/)
GLSL_OPS4(+) GLSL_OPS4(-) GLSL_OPS4(*) GLSL_OPS4(/)
#undef GLSL_OPS2
#undef GLSL_OPS3
#undef GLSL_OPS4
inline vec3 &operator*=(vec3 &a, const vec3 &b) { a = a * b; return a; }
inline vec2 operator-(const vec2 &a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }

/* ---- matrices: column-major, M[c] = column c (spec 5.4.2, 5.6) ---- */
struct mat4 {
    vec4 c[4];
    mat4() {}
    mat4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3,
         float c0, float c1, float c2, float c3, float d0, float d1, float d2, float d3)
    { c[0] = vec4(a0, a1, a2, a3); c[1] = vec4(b0, b1, b2, b3); c[2] = vec4(c0, c1, c2, c3); c[3] = vec4(d0, d1, d2, d3); }
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    mat3() {}
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
/* linear-algebraic M * v: the sum of the columns scaled by the components of v */
inline vec4 operator*(const mat4 &m, const vec4 &v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
inline vec3 operator*(const mat3 &m, const vec3 &v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }

/* ---- built-in functions (spec 8.3, 8.4, 8.6) ---- */
inline float floor(float x) { return ::floorf(x); }
inline vec2 floor(const vec2 &v) { return vec2(::floorf(v.x), ::floorf(v.y)); }
inline vec4 floor(const vec4 &v) { return vec4(::floorf(v.x), ::floorf(v.y), ::floorf(v.z), ::floorf(v.w)); }
inline float fract(float x) { return x - ::floorf(x); }
inline vec2 fract(const vec2 &v) { return vec2(fract(v.x), fract(v.y)); }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }
inline vec2 mod(const vec2 &v, float y) { return vec2(mod(v.x, y), mod(v.y, y)); }
inline float abs(float x) { return ::fabsf(x); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float smoothstep(float e0, float e1, float x)
{
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float dot(const vec2 &a, const vec2 &b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4 &a, const vec4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const vec2 &v) { return ::sqrtf(dot(v, v)); }
inline float length(const vec3 &v) { return ::sqrtf(dot(v, v)); }
inline vec3 cross(const vec3 &a, const vec3 &b)
{
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline vec3 normalize(const vec3 &v) { return v * (1.0f / ::sqrtf(dot(v, v))); }
inline bvec2 equal(const vec2 &a, const vec2 &b) { bvec2 r = { a.x == b.x, a.y == b.y }; return r; }
inline bool all(const bvec2 &b) { return b.x && b.y; }

/* ---- samplers ---- */
enum { REF_NEAREST = 0, REF_LINEAR = 1 };
enum { REF_CLAMP = 0, REF_REPEAT = 1 };
struct ref_sampler {
    const float *data;  /* layers * h * w * channels floats, x fastest */
    const unsigned char *bytes; /* or: an unorm8 texture, texel value c / 255 (OpenGL 3.3 spec 2.1.2) */
    int w, h, layers, channels;
    int filter, wrap;
    int subtexel_bits;  /* LINEAR weights: 0 = exact fp32, n = rounded to n fractional bits */
    float missing[4];   /* value of the channels the storage does not have (0,0,0,1) */
    ref_sampler() : data(NULL), bytes(NULL), w(0), h(0), layers(1), channels(1), filter(REF_NEAREST), wrap(REF_CLAMP), subtexel_bits(8)
    { missing[0] = missing[1] = missing[2] = 0.0f; missing[3] = 1.0f; }
    int wrapi(int i, int n) const
    {
        if (wrap == REF_REPEAT) { i %= n; return i < 0 ? i + n : i; }
        return i < 0 ? 0 : (i >= n ? n - 1 : i);
    }
    vec4 texel(int i, int j, int l) const
    {
        vec4 r(missing[0], missing[1], missing[2], missing[3]);
        if (data == NULL && bytes == NULL) return vec4(0.0f);       /* unbound texture */
        l = l < 0 ? 0 : (l >= layers ? layers - 1 : l);
        const size_t o = (((size_t) l * h + wrapi(j, h)) * w + wrapi(i, w)) * channels;
        for (int c = 0; c < channels; ++c) r[c] = data != NULL ? data[o + c] : (float) bytes[o + c] / 255.0f;
        return r;
    }
    float weight(float f) const
    {
        if (subtexel_bits <= 0) return f;
        const float q = (float) (1 << subtexel_bits);
        return ::floorf(f * q + 0.5f) / q;
    }
    vec4 sample(float u, float v, int l) const
    {
        if (filter == REF_NEAREST) {
            return texel((int) ::floorf(u * (float) w), (int) ::floorf(v * (float) h), l);
        }
        float fu = u * (float) w - 0.5f, fv = v * (float) h - 0.5f;
        float i0f = ::floorf(fu), j0f = ::floorf(fv);
        int i0 = (int) i0f, j0 = (int) j0f;
        float a = weight(fu - i0f), b = weight(fv - j0f);
        /* spec eq. 3.26, left to right */
        return ((1.0f - a) * (1.0f - b)) * texel(i0, j0, l) + (a * (1.0f - b)) * texel(i0 + 1, j0, l)
             + ((1.0f - a) * b) * texel(i0, j0 + 1, l) + (a * b) * texel(i0 + 1, j0 + 1, l);
    }
};
typedef ref_sampler sampler2D;
typedef ref_sampler sampler2DArray;
inline vec4 textureLod(const sampler2D &s, const vec2 &uv, float) { return s.sample(uv.x, uv.y, 0); }
/* array layer = floor(layer + 0.5), clamped (spec 3.8.8) */
inline vec4 textureLod(const sampler2DArray &s, const vec3 &uvl, float) { return s.sample(uvl.x, uvl.y, (int) ::floorf(uvl.z + 0.5f)); }

} /* namespace glsl */

/* storage qualifiers: a shader's uniforms, inputs and outputs become the
 * (thread_local: tiles may be produced from several threads) variables of the
 * namespace the shader text is included into */
#define uniform thread_local
#define in thread_local
#define out thread_local
#define layout(x)
#define main shader_main
#define _FRAGMENT_ 1

#endif
