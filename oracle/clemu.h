/*
 * clemu.h -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * A minimal OpenCL-C-on-C++ emulation layer: just enough of the OpenCL C language
 * (vector types with swizzles, image reads/writes, samplers, work-item functions,
 * atomic_inc, convert_*) that the reference's kernel sources under
 * /root/reference/kernel/ *.cl compile UNMODIFIED IN PLACE with g++ (oracle/Makefile pipes
 * each file through a two-rule sed -- vector literals "(int2)(a,b)" -> "clemu_make_int2(a,b)"
 * and "void kernel" -> "kernel void" -- straight into the compiler; nothing is copied).
 *
 * The result (oracle/_ref/libvp_clref.so) is "the reference kernels executed by an
 * emulated OpenCL runtime".  What this layer has to DEFINE, because the OpenCL spec leaves
 * it to the runtime, is stated once here and mirrored by vp_oracle.c:
 *   - read_imageui + CLK_FILTER_LINEAR on an integer image: bilinear per the OpenCL 1.2
 *     spec formula in fp32, result converted by mode (RTE default | TRUNC | NEAREST);
 *   - an int2 coordinate with a LINEAR sampler (BGR branch of quad2nv12/quad2rgba): texel;
 *   - native_sqrt: correctly rounded sqrtf;  no FMA contraction (-ffp-contract=off);
 *   - work-items execute in raster order (dim 0 fastest) on one thread when threads == 1.
 */
#ifndef CLEMU_H
#define CLEMU_H

#include <cmath>
#include <cstddef>
#include <cstdint>

#define CL_VERSION_1_0 100 /* keeps the .cl files from including their IDE shim clstd.h */

typedef unsigned char uchar;
typedef unsigned int uint;

/* ---- vector types ---- */
template <typename T>
struct clemu_vec2 {
	union { struct { T x, y; }; struct { T s0, s1; }; };
};
template <typename T>
struct clemu_vec3 {
	union { struct { T x, y, z; }; struct { T r, g, b; }; };
};
template <typename T>
struct clemu_vec4 {
	union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; };
};

typedef clemu_vec2<int> int2;
typedef clemu_vec2<float> float2;
typedef clemu_vec3<float> float3;
typedef clemu_vec4<float> float4;
typedef clemu_vec4<uint> uint4;

template <typename A, typename B> static inline int2 clemu_make_int2(A a, B b) { int2 v; v.x = (int)a; v.y = (int)b; return v; }
template <typename A, typename B> static inline float2 clemu_make_float2(A a, B b) { float2 v; v.x = (float)a; v.y = (float)b; return v; }
template <typename A, typename B, typename C> static inline float3 clemu_make_float3(A a, B b, C c) { float3 v; v.x = (float)a; v.y = (float)b; v.z = (float)c; return v; }
template <typename A, typename B, typename C, typename D> static inline uint4 clemu_make_uint4(A a, B b, C c, D d) { uint4 v; v.x = (uint)a; v.y = (uint)b; v.z = (uint)c; v.w = (uint)d; return v; }
template <typename A, typename B, typename C, typename D> static inline float4 clemu_make_float4(A a, B b, C c, D d) { float4 v; v.x = (float)a; v.y = (float)b; v.z = (float)c; v.w = (float)d; return v; }

/* element-wise operators actually used by the kernels (OpenCL C 6.3) */
static inline int2 operator+(int2 a, int2 b) { return clemu_make_int2(a.x + b.x, a.y + b.y); }
static inline int2& operator/=(int2& a, int s) { a.x /= s; a.y /= s; return a; }

static inline float2 operator*(float2 a, float2 b) { return clemu_make_float2(a.x * b.x, a.y * b.y); }
static inline float2 operator*(float s, float2 a) { return clemu_make_float2(s * a.x, s * a.y); }
static inline float2 operator/(float2 a, float s) { return clemu_make_float2(a.x / s, a.y / s); }
static inline float2 operator+(float2 a, float2 b) { return clemu_make_float2(a.x + b.x, a.y + b.y); }

static inline float3& operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }

static inline float4 operator-(float4 a, float4 b) { return clemu_make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline float4 operator*(float4 a, float4 b) { return clemu_make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline float4& operator*=(float4& a, float4 b) { a = a * b; return a; }
/* vector / scalar: the scalar is converted to the element type first (6.3.a) */
static inline float4 operator/(float4 a, int s) { const float f = (float)s; return clemu_make_float4(a.x / f, a.y / f, a.z / f, a.w / f); }

static inline uint4 operator*(uint4 a, uint4 b) { return clemu_make_uint4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline uint4& operator+=(uint4& a, uint4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; return a; }
static inline uint4 operator/(uint4 a, int s) { const uint u = (uint)s; return clemu_make_uint4(a.x / u, a.y / u, a.z / u, a.w / u); }

static inline float4 convert_float4(uint4 v) { return clemu_make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
static inline float4 native_sqrt(float4 v) { return clemu_make_float4(sqrtf(v.x), sqrtf(v.y), sqrtf(v.z), sqrtf(v.w)); }
static inline float min(float a, float b) { return b < a ? b : a; } /* 6.12.4 */

static inline uchar convert_uchar_sat(int v) { return (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
static inline uchar convert_uchar_sat(uint v) { return (uchar)(v > 255u ? 255u : v); }
static inline uchar convert_uchar_sat(float v) /* RTZ, saturating, NaN -> 0 (6.2.3.3) */
{
	if (!(v > 0.0f)) return 0;
	if (v >= 255.0f) return 255;
	return (uchar)(int)v;
}

/* ---- work-item functions ---- */
struct clemu_workitem { size_t gid[3]; size_t gsize[3]; };
extern thread_local clemu_workitem clemu_wi;
static inline size_t get_global_id(uint d) { return clemu_wi.gid[d]; }
static inline size_t get_global_size(uint d) { return clemu_wi.gsize[d]; }

/* ---- atomics ---- */
static inline int atomic_inc(volatile int* p) { return __atomic_fetch_add(p, 1, __ATOMIC_RELAXED); }

/* ---- images and samplers ---- */
enum { CLEMU_U8 = 0, CLEMU_RGBA8 = 1, CLEMU_F32 = 2 };
struct clemu_image { void* data; int width, height, type; };
typedef const clemu_image* image2d_t;

typedef unsigned int sampler_t;
enum {
	CLK_NORMALIZED_COORDS_FALSE = 0,
	CLK_ADDRESS_NONE = 0, CLK_ADDRESS_CLAMP_TO_EDGE = 2,
	CLK_FILTER_NEAREST = 0x10, CLK_FILTER_LINEAR = 0x20
};

enum { CLEMU_BILINEAR_RTE = 0, CLEMU_BILINEAR_TRUNC = 1, CLEMU_NEAREST = 2 };
extern int clemu_linear_mode;

static inline int get_image_width(image2d_t i) { return i->width; }
static inline int get_image_height(image2d_t i) { return i->height; }

static inline int clemu_clamp(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }
static inline int clemu_sat_index(float f, int n)
{
	f = fminf(fmaxf(f, -1.0f), (float)n);
	return clemu_clamp((int)f, n);
}

static inline uint4 clemu_texel_ui(image2d_t img, int i, int j)
{
	if (img->type == CLEMU_RGBA8) {
		const uchar* p = (const uchar*)img->data + 4 * ((size_t)i + (size_t)j * img->width);
		return clemu_make_uint4(p[0], p[1], p[2], p[3]);
	}
	/* CL_R: (r, 0, 0, 1) */
	return clemu_make_uint4(((const uchar*)img->data)[(size_t)i + (size_t)j * img->width], 0, 0, 1);
}

static inline uint4 read_imageui(image2d_t img, sampler_t s, int2 c)
{
	(void)s; /* every sampler in the reference is either CLAMP_TO_EDGE or never out of range */
	return clemu_texel_ui(img, clemu_clamp(c.x, img->width), clemu_clamp(c.y, img->height));
}

static inline uint clemu_round(float v)
{
	if (!(v >= 0.0f)) return 0;
	if (v >= 255.0f) return 255;
	return clemu_linear_mode == CLEMU_BILINEAR_TRUNC ? (uint)v : (uint)rintf(v);
}

static inline uint4 read_imageui(image2d_t img, sampler_t s, float2 c)
{
	const int w = img->width, h = img->height;
	if (!(s & CLK_FILTER_LINEAR) || clemu_linear_mode == CLEMU_NEAREST)
		return clemu_texel_ui(img, clemu_sat_index(floorf(c.x), w), clemu_sat_index(floorf(c.y), h));
	/* OpenCL 1.2 spec 8.2, CLK_FILTER_LINEAR on a 2D image */
	const float fu = c.x - 0.5f, fv = c.y - 0.5f;
	const float fi = floorf(fu), fj = floorf(fv);
	const float a = fu - fi, b = fv - fj;
	const int i0 = clemu_sat_index(fi, w), i1 = clemu_sat_index(fi + 1.0f, w);
	const int j0 = clemu_sat_index(fj, h), j1 = clemu_sat_index(fj + 1.0f, h);
	const float4 t00 = convert_float4(clemu_texel_ui(img, i0, j0)), t10 = convert_float4(clemu_texel_ui(img, i1, j0));
	const float4 t01 = convert_float4(clemu_texel_ui(img, i0, j1)), t11 = convert_float4(clemu_texel_ui(img, i1, j1));
	const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	uint4 r;
	r.x = clemu_round(((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x);
	r.y = clemu_round(((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y);
	r.z = clemu_round(((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z);
	r.w = clemu_round(((w00 * t00.w + w10 * t10.w) + w01 * t01.w) + w11 * t11.w);
	return r;
}

static inline float4 read_imagef(image2d_t img, sampler_t s, int2 c)
{
	(void)s;
	const int i = clemu_clamp(c.x, img->width), j = clemu_clamp(c.y, img->height);
	return clemu_make_float4(((const float*)img->data)[(size_t)i + (size_t)j * img->width], 0.f, 0.f, 1.f);
}

static inline void write_imageui(image2d_t img, int2 c, uint4 v)
{
	if (c.x < 0 || c.y < 0 || c.x >= img->width || c.y >= img->height) return;
	if (img->type == CLEMU_RGBA8) {
		uchar* p = (uchar*)img->data + 4 * ((size_t)c.x + (size_t)c.y * img->width);
		p[0] = convert_uchar_sat(v.x); p[1] = convert_uchar_sat(v.y);
		p[2] = convert_uchar_sat(v.z); p[3] = convert_uchar_sat(v.w);
	} else {
		((uchar*)img->data)[(size_t)c.x + (size_t)c.y * img->width] = convert_uchar_sat(v.x);
	}
}
/* raw2quad.cl passes a scalar colour (implicit scalar -> vector widening) */
static inline void write_imageui(image2d_t img, int2 c, uint v) { write_imageui(img, c, clemu_make_uint4(v, v, v, v)); }

static inline void write_imagef(image2d_t img, int2 c, float v)
{
	if (c.x < 0 || c.y < 0 || c.x >= img->width || c.y >= img->height) return;
	((float*)img->data)[(size_t)c.x + (size_t)c.y * img->width] = v;
}
static inline void write_imagef(image2d_t img, int2 c, float4 v) { write_imagef(img, c, v.x); }

/* ---- address-space / access qualifiers and the kernel keyword ---- */
#ifndef CLEMU_NO_KEYWORDS
#define kernel extern "C"
#define global
#define read_only
#define write_only
#endif

#endif /* CLEMU_H */
