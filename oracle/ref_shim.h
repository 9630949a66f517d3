// oracle/ref_shim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Host shim (our own code) that lets plain g++ compile the reference's hand-written CUDA math
// headers *where they lie* under /root/reference (no reference source is copied into this repo).
// It supplies (a) host stand-ins for the CUDA device intrinsics those headers call and (b) the
// component-wise float3 operators that the reference normally gets from the Slang-generated
// prelude `threedgutSlang.cuh` (slangc is not available in this image).
// CUDA's own headers are included first as an ordinary host compiler would see them; only then is
// __CUDACC__ defined so that the reference's `#ifdef __CUDACC__` guards open.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
using std::min;
using std::max;
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __saturatef(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float atomicAdd(float* a, float v) { float o = *a; *a += v; return o; }
static inline int atomicMin(int* a, int v) { int o = *a; *a = std::min(o, v); return o; }
static inline int atomicMax(int* a, int v) { int o = *a; *a = std::max(o, v); return o; }
static inline unsigned atomicMin(unsigned* a, unsigned v) { unsigned o = *a; *a = std::min(o, v); return o; }
static inline unsigned atomicMax(unsigned* a, unsigned v) { unsigned o = *a; *a = std::max(o, v); return o; }
// float3 (x) float3 operators: supplied by the Slang CUDA prelude in the real build.
static inline float3 operator+(const float3& a, const float3& b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(const float3& a, const float3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(const float3& a, const float3& b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator/(const float3& a, const float3& b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float3 operator-(const float3& a) { return make_float3(-a.x, -a.y, -a.z); }
static inline float3 make_float3(float a) { return make_float3(a, a, a); }
static inline float4 make_float4(float a) { return make_float4(a, a, a, a); }
#define __CUDACC__ 1
