// hlsl_shim.hpp -- just enough of HLSL's vector/matrix types, intrinsics and resource objects, in C++, for the REFERENCE'S
// OWN shader source (package/Shaders/*.hlsl|.compute|.shader, read where it lies under /root/reference and lightly
// pre-processed by build_ref_hlsl.py) to compile with g++ and run on the CPU.  TEST INFRASTRUCTURE ONLY: the result,
// oracle/_ref/libref_hlsl.so, exists to pin oracle/gs_oracle.c against the reference's code itself.
//
// What is the reference's and what is this file's: every expression, constant, operation order and branch comes from the
// reference source.  This file supplies what the GPU / driver supplies there: IEEE float32 scalar arithmetic (compile with
// -ffp-contract=off), the intrinsics (sqrt, rcp = 1/x, normalize = v / sqrt(dot), exp/log from libm, f32tof16 round to
// nearest even, dot / mul summed left to right ...), buffer and texture loads, and Unity's engine globals (UnityCG.cginc
// is not vendored in the reference).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace hlsl {

typedef uint32_t uint;
struct float2; struct float3; struct float4; struct uint2; struct uint3; struct uint4;
template <class T, int N> struct VecOf;
template <> struct VecOf<float, 2> { typedef float2 type; };
template <> struct VecOf<float, 3> { typedef float3 type; };
template <> struct VecOf<float, 4> { typedef float4 type; };
template <> struct VecOf<uint, 2> { typedef uint2 type; };
template <> struct VecOf<uint, 3> { typedef uint3 type; };
template <> struct VecOf<uint, 4> { typedef uint4 type; };

// A swizzle is a view of some components of the vector it is a union member of (all union members start at offset 0).
template <class T, int NSRC, int N, int A, int B, int C, int D>
struct Swz {
  typedef typename VecOf<T, N>::type V;
  T &at(int i) { return reinterpret_cast<T *>(this)[i == 0 ? A : i == 1 ? B : i == 2 ? C : D]; }
  const T &at(int i) const { return reinterpret_cast<const T *>(this)[i == 0 ? A : i == 1 ? B : i == 2 ? C : D]; }
  operator V() const { V r; for (int i = 0; i < N; ++i) r[i] = at(i); return r; }
  Swz &operator=(const V &v) { V t = v; for (int i = 0; i < N; ++i) at(i) = t[i]; return *this; }
  Swz &operator=(const Swz &o) { return *this = (V)o; }
  template <int NS2, int A2, int B2, int C2, int D2>
  Swz &operator=(const Swz<T, NS2, N, A2, B2, C2, D2> &o) { return *this = (typename VecOf<T, N>::type)o; }
  Swz &operator+=(const V &v) { return *this = (V)(*this) + v; }
  Swz &operator-=(const V &v) { return *this = (V)(*this) - v; }
  Swz &operator*=(const V &v) { return *this = (V)(*this) * v; }
  Swz &operator/=(const V &v) { return *this = (V)(*this) / v; }
};

#include "hlsl_swizzles.inc"

#define HLSL_VEC_COMMON(NAME, T, N)                                         \
  T &operator[](int i) { return (&x)[i]; }                                  \
  const T &operator[](int i) const { return (&x)[i]; }                      \
  NAME(const NAME &o) { for (int i = 0; i < N; ++i) (&x)[i] = (&o.x)[i]; }  \
  NAME &operator=(const NAME &o) { for (int i = 0; i < N; ++i) (&x)[i] = (&o.x)[i]; return *this; }

struct float2 {
  union { struct { float x, y; }; struct { float r, g; }; HLSL_SWIZZLES_FLOAT2 };
  float2() : x(0), y(0) {}
  float2(float s) : x(s), y(s) {}
  float2(float a, float b) : x(a), y(b) {}
  HLSL_VEC_COMMON(float2, float, 2)
};
struct float3 {
  union { struct { float x, y, z; }; struct { float r, g, b; }; HLSL_SWIZZLES_FLOAT3 };
  float3() : x(0), y(0), z(0) {}
  float3(float s) : x(s), y(s), z(s) {}
  float3(float a, float b, float c) : x(a), y(b), z(c) {}
  float3(const float2 &a, float c) : x(a.x), y(a.y), z(c) {}
  HLSL_VEC_COMMON(float3, float, 3)
};
struct float4 {
  union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; HLSL_SWIZZLES_FLOAT4 };
  float4() : x(0), y(0), z(0), w(0) {}
  float4(float s) : x(s), y(s), z(s), w(s) {}
  float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  float4(const float3 &a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
  float4(const float2 &a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
  float4(const float2 &a, const float2 &b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
  HLSL_VEC_COMMON(float4, float, 4)
};
struct uint2 {
  union { struct { uint x, y; }; HLSL_SWIZZLES_UINT2 };
  uint2() : x(0), y(0) {}
  uint2(uint s) : x(s), y(s) {}
  uint2(uint a, uint b) : x(a), y(b) {}
  HLSL_VEC_COMMON(uint2, uint, 2)
};
struct uint3 {
  union { struct { uint x, y, z; }; HLSL_SWIZZLES_UINT3 };
  uint3() : x(0), y(0), z(0) {}
  uint3(uint s) : x(s), y(s), z(s) {}
  uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
  HLSL_VEC_COMMON(uint3, uint, 3)
};
struct uint4 {
  union { struct { uint x, y, z, w; }; HLSL_SWIZZLES_UINT4 };
  uint4() : x(0), y(0), z(0), w(0) {}
  uint4(uint s) : x(s), y(s), z(s), w(s) {}
  uint4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
  HLSL_VEC_COMMON(uint4, uint, 4)
};
typedef float half;
typedef float2 half2;
typedef float3 half3;
typedef float4 half4;
typedef uint3 int3;   // only used as a texel coordinate

struct bool2 { bool v[2]; };
struct bool3 { bool v[3]; };
struct bool4 { bool v[4]; };
template <int N> struct BoolOf;
template <> struct BoolOf<2> { typedef bool2 type; };
template <> struct BoolOf<3> { typedef bool3 type; };
template <> struct BoolOf<4> { typedef bool4 type; };

#define HLSL_FLOAT_OPS(V, N)                                                                                                       \
  inline V operator+(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }                    \
  inline V operator-(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }                    \
  inline V operator*(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }                    \
  inline V operator/(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }                    \
  inline V operator+(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }                          \
  inline V operator-(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }                          \
  inline V operator*(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }                          \
  inline V operator/(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }                          \
  inline V operator+(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = s + a[i]; return r; }                          \
  inline V operator-(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }                          \
  inline V operator*(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }                          \
  inline V operator/(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }                          \
  inline V operator-(const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                                      \
  inline V &operator+=(V &a, const V &b) { a = a + b; return a; }                                                                  \
  inline V &operator-=(V &a, const V &b) { a = a - b; return a; }                                                                  \
  inline V &operator*=(V &a, const V &b) { a = a * b; return a; }                                                                  \
  inline V &operator/=(V &a, const V &b) { a = a / b; return a; }                                                                  \
  inline BoolOf<N>::type operator<=(const V &a, const V &b) { BoolOf<N>::type r; for (int i = 0; i < N; ++i) r.v[i] = a[i] <= b[i]; return r; } \
  inline BoolOf<N>::type operator<(const V &a, const V &b) { BoolOf<N>::type r; for (int i = 0; i < N; ++i) r.v[i] = a[i] < b[i]; return r; }   \
  inline BoolOf<N>::type operator>=(const V &a, const V &b) { BoolOf<N>::type r; for (int i = 0; i < N; ++i) r.v[i] = a[i] >= b[i]; return r; } \
  inline BoolOf<N>::type operator>(const V &a, const V &b) { BoolOf<N>::type r; for (int i = 0; i < N; ++i) r.v[i] = a[i] > b[i]; return r; }   \
  inline bool all(const BoolOf<N>::type &b) { bool r = true; for (int i = 0; i < N; ++i) r = r && b.v[i]; return r; }              \
  inline bool any(const BoolOf<N>::type &b) { bool r = false; for (int i = 0; i < N; ++i) r = r || b.v[i]; return r; }             \
  inline V abs(const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = std::fabs(a[i]); return r; }                                  \
  inline V sqrt(const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = std::sqrt(a[i]); return r; }                                 \
  inline V log(const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = std::log(a[i]); return r; }                                   \
  inline V exp(const V &a) { V r; for (int i = 0; i < N; ++i) r[i] = std::exp(a[i]); return r; }                                   \
  inline V min(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] < b[i] ? a[i] : b[i]; return r; }            \
  inline V max(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] > b[i] ? a[i] : b[i]; return r; }            \
  inline V lerp(const V &a, const V &b, const V &t) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + t[i] * (b[i] - a[i]); return r; } \
  inline float dot(const V &a, const V &b) { float r = a[0] * b[0]; for (int i = 1; i < N; ++i) r = r + a[i] * b[i]; return r; }   \
  inline float length(const V &a) { return std::sqrt(dot(a, a)); }                                                                 \
  inline V normalize(const V &a) { return a / length(a); }

HLSL_FLOAT_OPS(float2, 2)
HLSL_FLOAT_OPS(float3, 3)
HLSL_FLOAT_OPS(float4, 4)

// ---- scalar intrinsics ----
inline float abs(float a) { return std::fabs(a); }
inline float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float rcp(float a) { return 1.0f / a; }
inline float rsqrt(float a) { return 1.0f / std::sqrt(a); }
inline float min(float a, float b) { return a < b ? a : b; }
inline float max(float a, float b) { return a > b ? a : b; }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float saturate(float v) { return (v > 0.0f) ? ((v < 1.0f) ? v : 1.0f) : 0.0f; }  // NaN -> 0
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float round(float a) { return std::nearbyint(a); }   // HLSL round: to nearest even
inline float exp(float a) { return std::exp(a); }
inline float log(float a) { return std::log(a); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float floor(float a) { return std::floor(a); }
inline float asfloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float asfloat(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint asuint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline float3 cross(const float3 &a, const float3 &b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

inline uint f32tof16(float f) {  // round to nearest even (the D3D conversion rule)
  uint x = asuint(f), sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) return sign | (ax > 0x7f800000u ? 0x7e00u : 0x7c00u);
  if (ax >= 0x477ff000u) return sign | 0x7c00u;
  if (ax < 0x33000001u) return sign;
  int e = (int)(ax >> 23) - 127;
  uint m = (ax & 0x7fffffu) | 0x800000u, shift, h;
  if (e < -14) { shift = (uint)(13 + (-14 - e)); h = 0; } else { shift = 13; h = (uint)(e + 15) << 10; m &= 0x7fffffu; }
  uint q = m >> shift, rem = m & ((1u << shift) - 1u), half_ = 1u << (shift - 1);
  h += q;
  if (rem > half_ || (rem == half_ && (h & 1u))) h += 1;
  return sign | h;
}
inline float f16tof32(uint h) {
  h &= 0xffffu;
  uint sign = (h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3ffu;
  if (e == 0) { if (m == 0) return asfloat(sign); float v = (float)m * 5.9604644775390625e-08f; return sign ? -v : v; }
  if (e == 31) return asfloat(sign | 0x7f800000u | (m << 13));
  return asfloat(sign | ((e + 112u) << 23) | (m << 13));
}

// ---- matrices (row-major element access like HLSL's _mRC; mul(M, v) treats v as a column) ----
struct RowRef { float *p; float &operator[](int c) { return p[c]; } };
struct float4x4 {
  union {
    float m[4][4];
    struct { float _m00, _m01, _m02, _m03, _m10, _m11, _m12, _m13, _m20, _m21, _m22, _m23, _m30, _m31, _m32, _m33; };
    struct { float _11, _12, _13, _14, _21, _22, _23, _24, _31, _32, _33, _34, _41, _42, _43, _44; };
  };
  float4x4() { std::memset(m, 0, sizeof(m)); }
  RowRef operator[](int r) { return RowRef{m[r]}; }
};
struct float3x3 {
  union {
    float m[3][3];
    struct { float _m00, _m01, _m02, _m10, _m11, _m12, _m20, _m21, _m22; };
    struct { float _11, _12, _13, _21, _22, _23, _31, _32, _33; };
  };
  float3x3() { std::memset(m, 0, sizeof(m)); }
  float3x3(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    m[0][0] = a; m[0][1] = b; m[0][2] = c; m[1][0] = d; m[1][1] = e; m[1][2] = f; m[2][0] = g; m[2][1] = h; m[2][2] = i;
  }
  explicit float3x3(const float4x4 &o) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = o.m[r][c]; }
  RowRef operator[](int r) { return RowRef{m[r]}; }
};
inline float3x3 transpose(const float3x3 &a) { float3x3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
inline float3x3 mul(const float3x3 &a, const float3x3 &b) {
  float3x3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
inline float3 mul(const float3x3 &a, const float3 &v) {
  return float3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
                a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
inline float4 mul(const float4x4 &a, const float4 &v) {
  float4 r;
  for (int i = 0; i < 4; ++i) r[i] = a.m[i][0] * v.x + a.m[i][1] * v.y + a.m[i][2] * v.z + a.m[i][3] * v.w;
  return r;
}

// ---- resources ----
struct ByteAddressBuffer {
  const uint8_t *p = nullptr;
  uint Load(uint a) const { uint v; std::memcpy(&v, p + a, 4); return v; }
  uint2 Load2(uint a) const { return uint2(Load(a), Load(a + 4)); }
  uint3 Load3(uint a) const { return uint3(Load(a), Load(a + 4), Load(a + 8)); }
  uint4 Load4(uint a) const { return uint4(Load(a), Load(a + 4), Load(a + 8), Load(a + 12)); }
};
typedef ByteAddressBuffer RWByteAddressBuffer;   // the kernels compiled here only read through it
template <class T> struct StructuredBuffer {
  T *p = nullptr;
  T &operator[](uint i) const { return p[i]; }
};
template <class T> struct RWStructuredBuffer : StructuredBuffer<T> {};
struct Texture2D {   // the colour texture: `fetch` does what the texture unit does for the asset's GraphicsFormat
  float4 (*fetch)(uint x, uint y) = nullptr;
  float4 Load(const uint3 &c) const { return fetch(c.x, c.y); }
};

}  // namespace hlsl
