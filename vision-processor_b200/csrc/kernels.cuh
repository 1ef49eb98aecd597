/*
 * kernels.cuh -- sm_100a device code of libvp_b200.so.
 *
 * Each kernel cites the reference kernel (path:line under TIGERs-Mannheim/vision-processor) whose
 * results it reproduces.  None of this is a translation of the OpenCL sources: the Bayer planes are
 * never materialised on the fused path (texels are gathered straight from the raw frame), the
 * field->image projection is evaluated once per camera geometry into an L2-resident coordinate
 * table, the summed-area table is an exact int32 warp-shuffle row scan plus a blocked column scan,
 * NV12 is produced per 2x2 block, and the blob list is compacted deterministically in raster order
 * (count -> rank -> emit) instead of through a racing atomic counter.
 *
 * Canonical arithmetic (the one the CPU checker under oracle/ restates): fp32, every operation
 * individually rounded to nearest-even, no FMA contraction (explicit __f*_rn intrinsics and
 * -fmad=false), IEEE division and square root.
 */
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <type_traits>

#include "vp_b200.h"
#include "device_util.cuh"

namespace vpk {


/* ------------------------------------------------------------------------------------------------
 * field -> image projection, kernel/resampling.cl:29-47
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ float2 field2image(const vp_camera_model& m, float fx, float fy, float fz)
{
	const float vx = __fsub_rn(fx, m.c[0]);
	const float vy = __fsub_rn(fy, m.c[1]);
	const float vz = __fsub_rn(fz, m.c[2]);
	const float rx = __fadd_rn(__fadd_rn(__fmul_rn(m.r[0], vx), __fmul_rn(m.r[1], vy)), __fmul_rn(m.r[2], vz));
	const float ry = __fadd_rn(__fadd_rn(__fmul_rn(m.r[3], vx), __fmul_rn(m.r[4], vy)), __fmul_rn(m.r[5], vz));
	const float rz = __fadd_rn(__fadd_rn(__fmul_rn(m.r[6], vx), __fmul_rn(m.r[7], vy)), __fmul_rn(m.r[8], vz));
	const float nx = __fdiv_rn(rx, rz);
	const float ny = __fdiv_rn(ry, rz);
	float ux = nx, uy = ny;
#pragma unroll
	for (int i = 0; i < 8; i++) { /* resampling.cl:40 */
		const float q = __fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy));
		const float dr = __fadd_rn(1.0f, __fmul_rn(m.d, q));
		ux = __fdiv_rn(nx, dr);
		uy = __fdiv_rn(ny, dr);
	}
	return make_float2(__fadd_rn(__fmul_rn(m.f, ux), m.p[0]), __fadd_rn(__fmul_rn(m.f, uy), m.p[1]));
}

/* Coordinate table of one camera geometry: image position of every flat pixel (resampling.cl:53).
 * Evaluated once per geometry change; the per-frame kernel reads it back from L2. */
__global__ void k_coord_table(float2* __restrict__ lut, vp_camera_model m, float height, float scale, float offx, float offy,
                              int wf, int hf)
{
	const int gx = blockIdx.x * blockDim.x + threadIdx.x;
	const int gy = blockIdx.y * blockDim.y + threadIdx.y;
	if (gx >= wf || gy >= hf)
		return;
	const float X = __fadd_rn(__fmul_rn((float)gx, scale), offx);
	const float Y = __fadd_rn(__fmul_rn((float)gy, scale), offy);
	lut[(size_t)gy * wf + gx] = field2image(m, X, Y, height);
}

/* ------------------------------------------------------------------------------------------------
 * texel sources: quad plane c, texel (i, j) -> byte
 * ---------------------------------------------------------------------------------------------- */
struct SrcPlanes { /* four U8 images as produced by raw2quad (stage API) */
	const uint8_t* ch[4];
	int w;
	__device__ __forceinline__ uint32_t tex(int c, int i, int j) const { return __ldg(ch[c] + (size_t)j * w + i); }
};
struct SrcBayer { /* raw Bayer frame: plane c texel (i,j) == raw[(2j + c/2) * 2wq + 2i + c%2]  (raw2quad.cl:31-37) */
	const uint8_t* raw;
	int row; /* 2*wq */
	__device__ __forceinline__ uint32_t tex(int c, int i, int j) const
	{
		return __ldg(raw + (size_t)(2 * j + (c >> 1)) * row + 2 * i + (c & 1));
	}
};
struct SrcBGR { /* interleaved BGR frame (raw2quad.cl:23-29) */
	const uint8_t* raw;
	int w;
	__device__ __forceinline__ uint32_t tex(int c, int i, int j) const { return __ldg(raw + 3 * ((size_t)j * w + i) + c); }
};

/* one axis of the OpenCL 1.2 LINEAR filter (spec 8.2): i0 = floor(u - 0.5), a = frac(u - 0.5), clamp after */
struct Axis {
	int i0, i1;
	float a, oma;
};
template <int MODE>
__device__ __forceinline__ Axis axis_setup(float u, int n)
{
	Axis ax;
	if (MODE == MODE_NEAREST) {
		ax.i0 = ax.i1 = sat_index(floorf(u), n);
		ax.a = 0.f;
		ax.oma = 1.f;
		return ax;
	}
	const float fu = __fsub_rn(u, 0.5f);
	const float fi = floorf(fu);
	ax.a = __fsub_rn(fu, fi);
	ax.oma = __fsub_rn(1.0f, ax.a);
	ax.i0 = sat_index(fi, n);
	ax.i1 = sat_index(__fadd_rn(fi, 1.0f), n);
	return ax;
}

/* The same axis for |u| < 2^20 without conversion-pipe instructions (FRND/F2I): rint via the 1.5*2^23 magic add,
 * corrected to floor; the integer index comes out of the mantissa of the same add.  Bit-identical to axis_setup. */
__device__ __forceinline__ Axis axis_setup_fast(float u, int n)
{
	constexpr float M = 12582912.0f; /* 1.5 * 2^23 */
	Axis ax;
	const float fu = __fsub_rn(u, 0.5f);
	const float t = __fadd_rn(fu, M);
	const float r = __fsub_rn(t, M); /* rint(fu) */
	const bool dec = r > fu;
	const float fi = dec ? __fsub_rn(r, 1.0f) : r; /* floor(fu) */
	const int ii = __float_as_int(t) - 0x4B400000 - (dec ? 1 : 0);
	ax.a = __fsub_rn(fu, fi);
	ax.oma = __fsub_rn(1.0f, ax.a);
	ax.i0 = min(max(ii, 0), n - 1);
	ax.i1 = min(max(ii + 1, 0), n - 1);
	return ax;
}

/* u8 -> fp32 without the conversion pipe: 0x4B000000 | b is the float 2^23 + b */
__device__ __forceinline__ float u8_to_float(uint32_t b) { return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f); }

/* bilinear blend of four byte texels, MODE_RTE only: round-to-nearest-even happens in the 2^23 add */
__device__ __forceinline__ uint32_t blend_rte(uint32_t b00, uint32_t b10, uint32_t b01, uint32_t b11, const Axis& x, const Axis& y)
{
	const float w00 = __fmul_rn(x.oma, y.oma), w10 = __fmul_rn(x.a, y.oma);
	const float w01 = __fmul_rn(x.oma, y.a), w11 = __fmul_rn(x.a, y.a);
	const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, u8_to_float(b00)), __fmul_rn(w10, u8_to_float(b10))), __fmul_rn(w01, u8_to_float(b01))),
	                            __fmul_rn(w11, u8_to_float(b11)));
	/* val is finite and in [0, 256): 2^23 + val rounds to nearest-even integer; saturate like the generic path */
	return min(__float_as_uint(__fadd_rn(val, 8388608.0f)) - 0x4B000000u, 255u);
}

/* read_imageui(plane c, LINEAR|UNNORMALIZED|CLAMP_TO_EDGE, (u,v)).x */
template <int MODE, class Src>
__device__ __forceinline__ uint32_t sample(const Src& s, int c, const Axis& x, const Axis& y)
{
	if (MODE == MODE_NEAREST)
		return s.tex(c, x.i0, y.i0);
	const float t00 = (float)s.tex(c, x.i0, y.i0);
	const float t10 = (float)s.tex(c, x.i1, y.i0);
	const float t01 = (float)s.tex(c, x.i0, y.i1);
	const float t11 = (float)s.tex(c, x.i1, y.i1);
	const float w00 = __fmul_rn(x.oma, y.oma), w10 = __fmul_rn(x.a, y.oma);
	const float w01 = __fmul_rn(x.oma, y.a), w11 = __fmul_rn(x.a, y.a);
	const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, t00), __fmul_rn(w10, t10)), __fmul_rn(w01, t01)), __fmul_rn(w11, t11));
	/* negative / NaN -> 0, >= 255 -> 255: the saturating conversions do exactly that */
	return min(MODE == MODE_TRUNC ? __float2uint_rz(val) : __float2uint_rn(val), 255u);
}

/* demosaic taps shared by resampling.cl:56-81, quad2nv12.cl:27-51, quad2rgba.cl:27-51 */
template <int FMT, int MODE, class Src>
__device__ __forceinline__ void demosaic(const Src& s, int wq, int hq, float px, float py, uint32_t& r, uint32_t& g, uint32_t& b)
{
	if (FMT == FMT_BGR) {
		const Axis x = axis_setup<MODE>(px, wq), y = axis_setup<MODE>(py, hq);
		r = sample<MODE>(s, 2, x, y);
		g = sample<MODE>(s, 1, x, y);
		b = sample<MODE>(s, 0, x, y);
		return;
	}
	const Axis xp = axis_setup<MODE>(__fadd_rn(px, 0.25f), wq), xn = axis_setup<MODE>(__fsub_rn(px, 0.25f), wq);
	const Axis yp = axis_setup<MODE>(__fadd_rn(py, 0.25f), hq), yn = axis_setup<MODE>(__fsub_rn(py, 0.25f), hq);
	/* in both Bayer orders plane 0 is tapped at (+,+), 1 at (-,+), 2 at (+,-), 3 at (-,-) */
	const uint32_t v0 = sample<MODE>(s, 0, xp, yp);
	const uint32_t v1 = sample<MODE>(s, 1, xn, yp);
	const uint32_t v2 = sample<MODE>(s, 2, xp, yn);
	const uint32_t v3 = sample<MODE>(s, 3, xn, yn);
	if (FMT == FMT_RGGB) { /* resampling.cl:65-70 */
		r = v0;
		g = v1 / 2 + v2 / 2;
		b = v3;
	} else { /* GRBG, resampling.cl:74-80 */
		r = v1;
		g = v0 / 2 + v3 / 2;
		b = v2;
	}
}

/* the BGR branch of quad2nv12/quad2rgba hands an int2 to the sampler: direct texel (quad2nv12.cl:27-33) */
template <int FMT, int MODE, class Src>
__device__ __forceinline__ void demosaic_quad(const Src& s, int wq, int hq, int x, int y, uint32_t& r, uint32_t& g, uint32_t& b)
{
	if (FMT == FMT_BGR) {
		r = s.tex(2, x, y);
		g = s.tex(1, x, y);
		b = s.tex(0, x, y);
	} else {
		demosaic<FMT, MODE>(s, wq, hq, (float)x, (float)y, r, g, b);
	}
}

__device__ __forceinline__ uint32_t drgb(uint32_t r, uint32_t g, uint32_t b)
{
	/* resampling.cl:86-91, uint32 arithmetic; results are in [0,255] */
	const uint32_t dr = (2u * r - g - b + 510u) / 4u;
	const uint32_t dg = (2u * g - b - r + 510u) / 4u;
	const uint32_t db = (2u * b - r - g + 510u) / 4u;
	return dr | (dg << 8) | (db << 16) | 0xFF000000u;
}

template <class Src> struct SrcTraits { static constexpr bool is_bayer = false; };
template <> struct SrcTraits<SrcBayer> { static constexpr bool is_bayer = true; };

template <class Src>
__device__ __forceinline__ Src src_frame(Src s, size_t byte_offset);
template <>
__device__ __forceinline__ SrcPlanes src_frame(SrcPlanes s, size_t o)
{
	for (int c = 0; c < 4; c++) s.ch[c] += o;
	return s;
}
template <>
__device__ __forceinline__ SrcBayer src_frame(SrcBayer s, size_t o) { s.raw += o; return s; }
template <>
__device__ __forceinline__ SrcBGR src_frame(SrcBGR s, size_t o) { s.raw += o; return s; }

/* ------------------------------------------------------------------------------------------------
 * K1 reproject: (raw2quad.cl:21-39 +) resampling.cl:52-99.  One flat pixel per thread, frame = blockIdx.y.
 * Bayer + bilinear-RTE (the default) takes a fast path: texels are gathered straight from the raw frame with
 * 32-bit offsets, all float<->int conversions are done with magic adds on the FMA/ALU pipes.  Coordinates that are
 * not comfortably finite (|p| >= 2^20: only degenerate geometries) take the generic path, bit-identical by construction.
 * ---------------------------------------------------------------------------------------------- */
template <int FMT>
__device__ __forceinline__ uint32_t reproject_bayer_rte_fast(const uint8_t* __restrict__ raw, int wq, int hq, float px, float py)
{
	const Axis xp = axis_setup_fast(__fadd_rn(px, 0.25f), wq), xn = axis_setup_fast(__fsub_rn(px, 0.25f), wq);
	const Axis yp = axis_setup_fast(__fadd_rn(py, 0.25f), hq), yn = axis_setup_fast(__fsub_rn(py, 0.25f), hq);
	const uint32_t row = 2u * (uint32_t)wq;
	/* byte offsets of quad (i, j): 2j*row + 2i; plane c adds (c/2)*row + c%2 */
	const uint32_t yp0 = 2u * (uint32_t)yp.i0 * row, yp1 = 2u * (uint32_t)yp.i1 * row;
	const uint32_t yn0 = 2u * (uint32_t)yn.i0 * row, yn1 = 2u * (uint32_t)yn.i1 * row;
	const uint32_t xp0 = 2u * (uint32_t)xp.i0, xp1 = 2u * (uint32_t)xp.i1, xn0 = 2u * (uint32_t)xn.i0, xn1 = 2u * (uint32_t)xn.i1;
	const uint8_t* p0 = raw;           /* plane 0 at (+,+) */
	const uint8_t* p1 = raw + 1;       /* plane 1 at (-,+) */
	const uint8_t* p2 = raw + row;     /* plane 2 at (+,-) */
	const uint8_t* p3 = raw + row + 1; /* plane 3 at (-,-) */
	const uint32_t v0 = blend_rte(__ldg(p0 + (yp0 + xp0)), __ldg(p0 + (yp0 + xp1)), __ldg(p0 + (yp1 + xp0)), __ldg(p0 + (yp1 + xp1)), xp, yp);
	const uint32_t v1 = blend_rte(__ldg(p1 + (yp0 + xn0)), __ldg(p1 + (yp0 + xn1)), __ldg(p1 + (yp1 + xn0)), __ldg(p1 + (yp1 + xn1)), xn, yp);
	const uint32_t v2 = blend_rte(__ldg(p2 + (yn0 + xp0)), __ldg(p2 + (yn0 + xp1)), __ldg(p2 + (yn1 + xp0)), __ldg(p2 + (yn1 + xp1)), xp, yn);
	const uint32_t v3 = blend_rte(__ldg(p3 + (yn0 + xn0)), __ldg(p3 + (yn0 + xn1)), __ldg(p3 + (yn1 + xn0)), __ldg(p3 + (yn1 + xn1)), xn, yn);
	if (FMT == FMT_RGGB)
		return drgb(v0, v1 / 2 + v2 / 2, v3);
	return drgb(v1, v0 / 2 + v3 / 2, v2);
}

template <int FMT, int MODE, class Src>
__global__ void __launch_bounds__(256) k_reproject(Src src, size_t src_frame_stride, const float2* __restrict__ lut,
                                                    uint32_t* __restrict__ flat, int wq, int hq, int nf)
{
	const int idx = blockIdx.x * 256 + threadIdx.x;
	if (idx >= nf)
		return;
	const Src s = src_frame(src, (size_t)blockIdx.y * src_frame_stride);
	const float2 pos = __ldg(lut + idx);
	uint32_t out;
	if constexpr (MODE == MODE_RTE && FMT != FMT_BGR && SrcTraits<Src>::is_bayer) {
		if (fabsf(pos.x) < 1048576.0f && fabsf(pos.y) < 1048576.0f) {
			out = reproject_bayer_rte_fast<FMT>(s.raw, wq, hq, pos.x, pos.y);
		} else {
			uint32_t r, g, b;
			demosaic<FMT, MODE>(s, wq, hq, pos.x, pos.y, r, g, b);
			out = drgb(r, g, b);
		}
	} else {
		uint32_t r, g, b;
		demosaic<FMT, MODE>(s, wq, hq, pos.x, pos.y, r, g, b);
		out = drgb(r, g, b);
	}
	flat[(size_t)blockIdx.y * nf + idx] = out;
}

/* ------------------------------------------------------------------------------------------------
 * K1, shared-memory staged forms (Bayer, bilinear-RTE): the hot kernels of the fused path.
 *
 * The direct kernel above spends most of its issue slots on address arithmetic, clamps and u8->fp32 conversions of
 * 16 gathered texels per pixel.  The staged kernels give a CTA a 64x16 tile of the flat image; the quads its pixels can touch
 * (known per camera geometry: k_tile_table) are staged into shared memory -- 16-byte coalesced copies of raw Bayer rows,
 * de-interleaved on the way in, edge texels replicated so that CLAMP_TO_EDGE needs no per-tap clamp -- and every tap
 * becomes one LDS with an immediate offset.  Tiles whose footprint does not fit (extreme perspective) or whose coordinates are
 * not finite fall back to the direct path pixel by pixel; results are identical.
 * ---------------------------------------------------------------------------------------------- */
constexpr int FT_W = 64, FT_H = 16;   /* flat tile */
constexpr int TQ_W = 80, TQ_H = 24;   /* staged quads per plane (capacity) */

struct TileEntry {
	int ib, jb;     /* quad coordinate of staged texel (0,0); ib is a multiple of 8 */
	int height;     /* staged quad rows needed (<= TQ_H) */
	int flags;      /* bit 0: footprint fits and coordinates are finite; bit 1: staging with 16-byte vectors is possible */
};

/* floor(u - 0.5) exactly as axis_setup_fast computes it */
__device__ __forceinline__ int axis_floor_fast(float u)
{
	constexpr float M = 12582912.0f;
	const float fu = __fsub_rn(u, 0.5f);
	const float t = __fadd_rn(fu, M);
	const float r = __fsub_rn(t, M);
	return __float_as_int(t) - 0x4B400000 - (r > fu ? 1 : 0);
}

/* once per geometry: footprint of every flat tile in quad coordinates */
__global__ void __launch_bounds__(256) k_tile_table(const float2* __restrict__ lut, TileEntry* __restrict__ table, int wf, int hf, int wq, int hq)
{
	__shared__ int red[4][8];
	const int tx = blockIdx.x, ty = blockIdx.y;
	int imin = INT_MAX, imax = INT_MIN, jmin = INT_MAX, jmax = INT_MIN;
	bool finite = true;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int gx = tx * FT_W + (threadIdx.x & 63), gy = ty * FT_H + (threadIdx.x >> 6) + 4 * k;
		if (gx < wf && gy < hf) {
			const float2 pos = __ldg(lut + (size_t)gy * wf + gx);
			if (fabsf(pos.x) < 1048576.0f && fabsf(pos.y) < 1048576.0f) {
				imin = min(imin, axis_floor_fast(__fsub_rn(pos.x, 0.25f)));
				imax = max(imax, axis_floor_fast(__fadd_rn(pos.x, 0.25f)) + 1);
				jmin = min(jmin, axis_floor_fast(__fsub_rn(pos.y, 0.25f)));
				jmax = max(jmax, axis_floor_fast(__fadd_rn(pos.y, 0.25f)) + 1);
			} else {
				finite = false;
			}
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
		imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));
		jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
		jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
	}
	const int all_finite = __syncthreads_and(finite);
	if ((threadIdx.x & 31) == 0) {
		const int wp = threadIdx.x >> 5;
		red[0][wp] = imin; red[1][wp] = imax; red[2][wp] = jmin; red[3][wp] = jmax;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int k = 1; k < 8; k++) {
			red[0][0] = min(red[0][0], red[0][k]); red[1][0] = max(red[1][0], red[1][k]);
			red[2][0] = min(red[2][0], red[2][k]); red[3][0] = max(red[3][0], red[3][k]);
		}
		TileEntry e;
		const int i0 = red[0][0], i1 = red[1][0], j0 = red[2][0], j1 = red[3][0];
		const bool any = i0 <= i1 && j0 <= j1; /* false for an empty tile */
		e.ib = any ? (i0 >= 0 ? i0 / 8 * 8 : -((-i0 + 7) / 8 * 8)) : 0;
		e.jb = any ? j0 : 0;
		e.height = any ? j1 - j0 + 1 : 0;
		const bool fits = any && all_finite && (long long)i1 - e.ib + 1 <= TQ_W && e.height <= TQ_H;
		const bool vec = fits && (wq % 8) == 0; /* raw rows are 16-byte aligned and no 8-quad vector straddles the image edge */
		e.flags = (fits ? 1 : 0) | (vec ? 2 : 0);
		table[ty * gridDim.x + tx] = e;
	}
}


/* two tile-relative axes at once (u.x, u.y share the plane dimension): indices into the staged planes, no clamp
 * (edge texels are replicated in the tile); bit-identical to axis_setup */
__device__ __forceinline__ void axis_staged2(float2 u, int origin_magic, int& i0, int& i1, float2& a, float2& oma)
{
	constexpr float M = 12582912.0f;
	const float2 fu = add2(u, make_float2(-0.5f, -0.5f));
	const float2 t = add2(fu, make_float2(M, M));
	const float2 r = add2(t, make_float2(-M, -M)); /* rint(fu) */
	const bool d0 = r.x > fu.x, d1 = r.y > fu.y;
	const float2 fi = add2(r, make_float2(d0 ? -1.0f : 0.0f, d1 ? -1.0f : 0.0f)); /* floor(fu) */
	i0 = __float_as_int(t.x) - origin_magic - (d0 ? 1 : 0);
	i1 = __float_as_int(t.y) - origin_magic - (d1 ? 1 : 0);
	a = sub2(fu, fi);
	oma = sub2(make_float2(1.0f, 1.0f), a);
}

/* ------------------------------------------------------------------------------------------------
 * K1, frame-invariant form: the default of the batched path.
 *
 * Everything resampling.cl:52-80 computes before it touches a texel -- the projected position, the four filter axes, the
 * sixteen weight products w = (1-a|a)*(1-b|b) and the tap addresses -- depends on the camera geometry only, not on the
 * frame.  A CTA therefore owns one 64x16 flat tile for a CHUNK of frames of the same camera: each thread derives the
 * weights (as packed fp32x2 pairs) and staged-plane offsets of its four pixels once, keeps them in registers, and so
 * does the staging plan (which 16-byte raw vectors go where).  Per frame what remains is the arithmetic on the data:
 * 16 LDS + 8 FMUL2 + 12 FADD + 2 FADD2 per pixel and the dRGB pack.  The raw vectors of frame f+1 are fetched before
 * frame f is blended.  Same operations in the same order as the direct kernel / the oracle: bit-identical.
 *
 * The integer tail works on the biased words 0x4B000000 + v that the 2^23 add leaves behind (val in [0, 255.5) for
 * weights in [0,1]): (x>>1) keeps the bias halved exactly (it is even), and 2r-g-b cancels it (resampling.cl:86-91).
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t drgb_biased(uint32_t r, uint32_t g, uint32_t b)
{
	const uint32_t s = r + g + b; /* bias 3B; 3x - s cancels it */
	const uint32_t dr = (3u * r - s + 510u) >> 2, dg = (3u * g - s + 510u) >> 2, db = (3u * b - s + 510u) >> 2;
	return dr | (dg << 8) | (db << 16) | 0xFF000000u;
}

constexpr int HP = TQ_W + 4;               /* row pitch of the staged planes: an odd number of 16-byte chunks (see the staging plan) */
constexpr int HPLANE = TQ_H * HP;          /* floats per plane */
constexpr int HT = 4 * HPLANE;             /* floats per plane buffer */
constexpr int HNV = TQ_W / 8;              /* 16-byte raw vectors per staged row */
constexpr int HRING = HNV * 2 * TQ_H * 16; /* bytes of one raw stage */
constexpr size_t HOIST_SMEM = 2 * (size_t)HT * 4 + 2 * (size_t)HRING;


/* the per-frame arithmetic of one thread's PX pixels (rows ly, ly+RS, ly+2RS, ... of the tile, RS = 16/PX) */
template <int FMT, bool FULL, int PX>
__device__ __forceinline__ void hoist_blend(const float* __restrict__ T, const float2 (&W)[PX][8], const int (&O)[PX][4], uint32_t* __restrict__ out, int wf,
                                            bool okx, int rows_ok, unsigned long long one2)
{
	constexpr int RS = FT_H / PX;
#pragma unroll
	for (int k = 0; k < PX; k++) {
		const float* t0 = T + O[k][0];
		const float* t1 = T + O[k][1];
		const float* t2 = T + O[k][2];
		const float* t3 = T + O[k][3];
		const float2 p00 = mul2(W[k][0], make_float2(t0[0], t1[0]));
		const float2 p10 = mul2(W[k][1], make_float2(t0[1], t1[1]));
		const float2 p01 = mul2(W[k][2], make_float2(t0[HP], t1[HP]));
		const float2 p11 = mul2(W[k][3], make_float2(t0[HP + 1], t1[HP + 1]));
		const float2 r00 = mul2(W[k][4], make_float2(t2[0], t3[0]));
		const float2 r10 = mul2(W[k][5], make_float2(t2[1], t3[1]));
		const float2 r01 = mul2(W[k][6], make_float2(t2[HP], t3[HP]));
		const float2 r11 = mul2(W[k][7], make_float2(t2[HP + 1], t3[HP + 1]));
		/* ((p00 + p10) + p01) + p11, every sum rounded on its own (resampling.cl's filter) */
		const float2 va = add2_opaque(p11, add2_opaque(p01, add2_opaque(p10, p00, one2), one2), one2);
		const float2 vb = add2_opaque(r11, add2_opaque(r01, add2_opaque(r10, r00, one2), one2), one2);
		const float2 ra = add2(va, make_float2(8388608.0f, 8388608.0f)); /* RNE to integer in the mantissa */
		const float2 rb = add2(vb, make_float2(8388608.0f, 8388608.0f));
		const uint32_t v0 = __float_as_uint(ra.x), v1 = __float_as_uint(ra.y), v2 = __float_as_uint(rb.x), v3 = __float_as_uint(rb.y);
		const uint32_t px = FMT == FMT_RGGB ? drgb_biased(v0, (v1 >> 1) + (v2 >> 1), v3) : drgb_biased(v1, (v0 >> 1) + (v3 >> 1), v2);
		if (FULL || (okx && RS * k < rows_ok))
			out[(uint32_t)(RS * k) * (uint32_t)wf] = px;
	}
}

/* 16 raw bytes (8 quads of one raw row = planes 2s | 2s+1 interleaved) -> 8 + 8 fp32 texels in the staged planes */
__device__ __forceinline__ void hoist_convert(uint4 qq, int edge, float* __restrict__ d0)
{
	if (edge) { /* replicate the edge quad's two bytes over the whole vector (CLAMP_TO_EDGE left / right of the image) */
		const uint32_t eq = edge == 1 ? (qq.x & 0xFFFFu) : (qq.w >> 16);
		qq.x = qq.y = qq.z = qq.w = eq * 0x00010001u;
	}
	float* d1 = d0 + HPLANE;
	const uint32_t wds[4] = { qq.x, qq.y, qq.z, qq.w };
	float2 e0[4], e1[4];
#pragma unroll
	for (int k = 0; k < 4; k++) { /* bytes 0,2 -> plane 2s; bytes 1,3 -> plane 2s+1; 0x4B0000bb is the float 2^23 + b */
		const float2 a = make_float2(__uint_as_float(__byte_perm(wds[k], 0x4B000000u, 0x7440)), __uint_as_float(__byte_perm(wds[k], 0x4B000000u, 0x7442)));
		const float2 b = make_float2(__uint_as_float(__byte_perm(wds[k], 0x4B000000u, 0x7441)), __uint_as_float(__byte_perm(wds[k], 0x4B000000u, 0x7443)));
		e0[k] = add2(a, make_float2(-8388608.0f, -8388608.0f));
		e1[k] = add2(b, make_float2(-8388608.0f, -8388608.0f));
	}
	reinterpret_cast<float4*>(d0)[0] = make_float4(e0[0].x, e0[0].y, e0[1].x, e0[1].y);
	reinterpret_cast<float4*>(d0)[1] = make_float4(e0[2].x, e0[2].y, e0[3].x, e0[3].y);
	reinterpret_cast<float4*>(d1)[0] = make_float4(e1[0].x, e1[0].y, e1[1].x, e1[1].y);
	reinterpret_cast<float4*>(d1)[1] = make_float4(e1[2].x, e1[2].y, e1[3].x, e1[3].y);
}


/* Pipeline of one CTA over its frames 0..n-1 (one barrier per frame, everything addressed statically by frame parity):
 *   raw ring R[0], R[1]  (cp.async; a thread copies and later converts ITS OWN vectors: no cross-thread hand-over)
 *   plane buffers T[0], T[1]  (fp32, ping-pong)
 *   iteration f:  issue copy(f+2) -> R[f&1];  blend(f) from T[f&1];  wait copy(f+1);  convert(f+1): R[(f+1)&1] -> T[(f+1)&1];  barrier
 * R[f&1] is free at the start of iteration f (this thread converted frame f out of it in iteration f-1); T[(f+1)&1] was
 * last read by blend(f-1), which every warp has left before the barrier that ended iteration f-1.  A warp that has
 * finished blending goes straight on to converting.
 * Staging plan: lanes 2k and 2k+1 take vector k of two staged rows ONE pitch apart; the pitch is 21 chunks of 16 bytes,
 * so the eight lanes of a store phase hit eight different bank groups (a pitch of 20 gave 2-way conflicts). */
/* PX = pixels per thread: 4 (256 threads, 128 registers, 2 CTAs/SM) or 2 (512 threads, 64 registers, 2 CTAs/SM = twice the warps) */
template <int FMT, int PX>
__global__ void __launch_bounds__(1024 / PX, 2) k_reproject_hoist(const uint8_t* __restrict__ raw0, size_t frame_stride, const float2* __restrict__ lut,
                                                              const TileEntry* __restrict__ table, uint32_t* __restrict__ flat, int wq, int hq,
                                                              int wf, int hf, int n_frames, int chunk, float one)
{
	extern __shared__ __align__(16) unsigned char hoist_smem[];
	float* const T = reinterpret_cast<float*>(hoist_smem);
	unsigned char* const ring = hoist_smem + 2 * (size_t)HT * 4;
	const int tx = blockIdx.x, ty = blockIdx.y;
	const int f0 = blockIdx.z * chunk;
	const int n = min(n_frames, f0 + chunk) - f0;
	const TileEntry e = table[ty * gridDim.x + tx];
	constexpr int NT = 1024 / PX, RS = FT_H / PX; /* threads per CTA; row distance between a thread's pixels */
	constexpr int NSLOT = (HNV * 2 * TQ_H + NT - 1) / NT; /* raw vectors per thread */
	const int tid = threadIdx.x;
	const int lx = tid & 63, ly = tid >> 6;
	const int gx = tx * FT_W + lx;
	const uint32_t nfl = (uint32_t)wf * (uint32_t)hf;
	const uint8_t* const raw = raw0 + (size_t)f0 * frame_stride;
	float2 pos[PX]; /* requested first: nothing below depends on them until the weights are derived */
#pragma unroll
	for (int k = 0; k < PX; k++) {
		const int gy = ty * FT_H + ly + RS * k;
		pos[k] = make_float2(0.f, 0.f);
		if (gx < wf && gy < hf)
			pos[k] = __ldg(lut + (gy * wf + gx));
	}

	if (!(e.flags & 1)) { /* footprint does not fit: direct gather, frame by frame */
#pragma unroll 1
		for (int f = 0; f < n; f++) {
			const uint8_t* rawf = raw + (size_t)f * frame_stride;
			uint32_t* out = flat + (size_t)(f0 + f) * nfl;
#pragma unroll 1
			for (int k = 0; k < PX; k++) {
				const int gy = ty * FT_H + ly + RS * k;
				if (gx < wf && gy < hf) {
					const float2 pos = __ldg(lut + (gy * wf + gx));
					uint32_t v;
					if (fabsf(pos.x) < 1048576.0f && fabsf(pos.y) < 1048576.0f) {
						v = reproject_bayer_rte_fast<FMT>(rawf, wq, hq, pos.x, pos.y);
					} else {
						uint32_t r, g, b;
						const SrcBayer s{ rawf, 2 * wq };
						demosaic<FMT, MODE_RTE>(s, wq, hq, pos.x, pos.y, r, g, b);
						v = drgb(r, g, b);
					}
					out[gy * wf + gx] = v;
				}
			}
		}
		return;
	}

	/* ---- frame-invariant staging plan: up to two 16-byte raw vectors per thread ---- */
	const int row_bytes = 2 * wq;
	const bool vec = (e.flags & 2) != 0;
	const int n_rr = 2 * e.height; /* raw rows to stage: rr = 2*j + s -> staged row j of planes (2s, 2s+1) */
	const uint8_t* s_src[NSLOT]; /* this thread's vector i in the next frame to copy */
	int s_dst[NSLOT], s_edge[NSLOT]; /* s_edge < 0: no vector */
#pragma unroll
	for (int i = 0; i < NSLOT; i++) {
		const int v = tid + NT * i;
		const int q4 = v / (4 * HNV), w = v - q4 * (4 * HNV);
		const int half = w >= 2 * HNV ? 1 : 0, w2 = w - half * 2 * HNV;
		const int cv = w2 >> 1, rr = 4 * q4 + half + 2 * (w2 & 1);
		s_src[i] = raw;
		s_dst[i] = 0;
		s_edge[i] = -1;
		if (vec && rr < n_rr) {
			const int qy = clampi(e.jb + (rr >> 1), 0, hq - 1);
			const int qx0 = e.ib + cv * 8;
			const int qxc = clampi(qx0, 0, wq - 8);
			s_src[i] = raw + ((2 * qy + (rr & 1)) * row_bytes + 2 * qxc);
			s_dst[i] = (rr & 1) * 2 * HPLANE + (rr >> 1) * HP + cv * 8;
			s_edge[i] = qx0 < 0 ? 1 : (qx0 != qxc ? 2 : 0);
		}
	}
	unsigned char* const my_ring = ring + tid * 16; /* vector i of stage st: my_ring + st*HRING + i*NT*16 */
	int to_copy = n;                                /* frames not yet requested */
	auto issue_copy = [&](auto STAGE) {             /* next frame into R[STAGE]; always commits, possibly an empty group */
		if (to_copy > 0) {
#pragma unroll
			for (int i = 0; i < NSLOT; i++) {
				if (s_edge[i] >= 0)
					cp_async16(my_ring + decltype(STAGE)::value * HRING + i * NT * 16, s_src[i]);
				s_src[i] += frame_stride;
			}
		}
		to_copy--;
		cp_async_commit();
	};
	const uint8_t* gather_src = raw; /* unaligned path only: the next frame to convert */
	auto convert = [&](auto PARC) { /* this thread's share of the next frame: R[PAR] -> T[PAR] */
		constexpr int PAR = decltype(PARC)::value;
		float* Tb = T + PAR * HT;
		if (vec) {
#pragma unroll
			for (int i = 0; i < NSLOT; i++)
				if (s_edge[i] >= 0)
					hoist_convert(*reinterpret_cast<const uint4*>(my_ring + PAR * HRING + i * NT * 16), s_edge[i], Tb + s_dst[i]);
		} else { /* raw rows not 16-byte aligned (wq % 8 != 0): per-texel gather with the edge replicated */
			const int tot = 4 * e.height * TQ_W;
			for (int v = tid; v < tot; v += NT) {
				const int rc = v / TQ_W, ii = v - rc * TQ_W; /* rc = 4*jj + c */
				const int jj = rc >> 2, c = rc & 3;
				const int qx = clampi(e.ib + ii, 0, wq - 1), qy = clampi(e.jb + jj, 0, hq - 1);
				Tb[c * HPLANE + jj * HP + ii] = u8_to_float(__ldg(gather_src + ((2 * qy + (c >> 1)) * row_bytes + 2 * qx + (c & 1))));
			}
			gather_src += frame_stride;
		}
	};
	issue_copy(IntC<0>{});
	issue_copy(IntC<1>{});

	/* ---- frame-invariant part: weights and tap offsets of this thread's four pixels (the copies are in flight) ---- */
	float2 W[PX][8];
	int O[PX][4];
	const int xmagic = 0x4B400000 + e.ib, ymagic = 0x4B400000 + e.jb;
#pragma unroll
	for (int k = 0; k < PX; k++) {
		const int gy = ty * FT_H + ly + RS * k;
		const bool ok = gx < wf && gy < hf;
		int ixp, ixn, iyp, iyn;
		float2 ax, ox, ay, oy; /* .x = the +0.25 axis, .y = the -0.25 axis */
		axis_staged2(add2(make_float2(pos[k].x, pos[k].x), make_float2(0.25f, -0.25f)), xmagic, ixp, ixn, ax, ox);
		axis_staged2(add2(make_float2(pos[k].y, pos[k].y), make_float2(0.25f, -0.25f)), ymagic, iyp, iyn, ay, oy);
		/* planes 0 (+,+) | 1 (-,+) share the +y axis, planes 2 (+,-) | 3 (-,-) the -y axis */
		const float2 ayp = make_float2(ay.x, ay.x), oyp = make_float2(oy.x, oy.x);
		const float2 ayn = make_float2(ay.y, ay.y), oyn = make_float2(oy.y, oy.y);
		W[k][0] = mul2(ox, oyp); W[k][1] = mul2(ax, oyp); W[k][2] = mul2(ox, ayp); W[k][3] = mul2(ax, ayp);
		W[k][4] = mul2(ox, oyn); W[k][5] = mul2(ax, oyn); W[k][6] = mul2(ox, ayn); W[k][7] = mul2(ax, ayn);
		O[k][0] = ok ? iyp * HP + ixp : 0; /* pixels outside the image read texel 0 and are not stored */
		O[k][1] = ok ? HPLANE + iyp * HP + ixn : 0;
		O[k][2] = ok ? 2 * HPLANE + iyn * HP + ixp : 0;
		O[k][3] = ok ? 3 * HPLANE + iyn * HP + ixn : 0;
	}
	const bool okx = gx < wf;
	const int rows_ok = hf - (ty * FT_H + ly);                        /* pixel k is inside the image iff RS*k < rows_ok */
	const bool full = (tx + 1) * FT_W <= wf && (ty + 1) * FT_H <= hf; /* CTA-uniform: no per-pixel predicates */
	uint32_t* out = flat + (size_t)f0 * nfl + ((uint32_t)(ty * FT_H + ly) * (uint32_t)wf + (uint32_t)gx); /* pixel k adds RS*k*wf */
	const unsigned long long one2 = f2_bits(make_float2(one, one));

	cp_async_wait<1>();
	convert(IntC<0>{});
	__syncthreads();
	int left = n; /* frames not yet blended */
	auto iteration = [&](auto PARC) {
		constexpr int PAR = decltype(PARC)::value;
		issue_copy(PARC);
		if (full)
			hoist_blend<FMT, true, PX>(T + PAR * HT, W, O, out, wf, true, 16, one2);
		else
			hoist_blend<FMT, false, PX>(T + PAR * HT, W, O, out, wf, okx, rows_ok, one2);
		out += nfl;
		left--;
		cp_async_wait<1>();
		if (left > 0)
			convert(IntC<PAR ^ 1>{});
		__syncthreads();
	};
#pragma unroll 1
	while (true) {
		iteration(IntC<0>{});
		if (left == 0) break;
		iteration(IntC<1>{});
		if (left == 0) break;
	}
}

/* ------------------------------------------------------------------------------------------------
 * K1, four frames at a time (the default of the batched path).
 *
 * k_reproject_hoist is bounded by the shared-memory pipe: 16 fp32 tap loads per pixel and frame.  The weights and the tap
 * addresses are the same for every frame of the camera, so the staged planes here hold, per texel, the BYTES OF FOUR
 * FRAMES in one 32-bit word: one LDS fetches a tap for four frames, one PRMT per frame turns its byte into a float and one
 * FMUL2 with a broadcast operand forms the products (frame A, frame B) x the scalar weight.  Per pixel and frame: 4 LDS instead
 * of 16; the staging does no arithmetic at all (a byte transpose of the four frames' raw vectors: 2 PRMT per staged word) and
 * stores a quarter of the bytes.  The arithmetic per frame is the canonical one -- same operations in the same order, every one
 * rounded on its own -- under a power-of-two scaling (below): bit-identical to k_reproject_hoist and to the oracle.
 *
 * Pipeline per quad of frames: convert(q) -> T | copy(q+1) issued (cp.async, own vectors, lands under the blend) | barrier |
 * blend(q) | barrier.  ONE plane buffer, ONE raw stage, the eight axis values of a pixel in registers (the sixteen weights are
 * formed in the blend), two CTAs per SM: each of the alternatives -- a second plane buffer (one barrier per quad), a second raw
 * stage, the sixteen weights in registers, a third CTA per SM at 80 registers -- measured 2..15 % slower
 * (profiles/r01_hoist_sweeps.txt, profiles/r02_sweeps.txt).
 *
 * How a staged byte becomes a product: the byte is used as the fp32 DENORMAL b * 2^-149 (one PRMT against zero) and the weight
 * carries a factor 2^100 (folded into the x-axis values once per tile): mul.rn(w * 2^100, b * 2^-149) = RN(w * b) * 2^-49 exactly --
 * scaling by a power of two commutes with the rounding as long as nothing leaves the normal range, and the smallest non-zero
 * product is 2^-50 * 2^-49 (an axis weight is a - floor(a) or its complement for a coordinate a = u - 0.5 with u >= 0.25 or the
 * difference rounded away: a multiple of 2^-25 or zero).  Every sum and the final 2^23 * 2^-49 rounding add carry the same
 * factor, so the mantissa bits -- the rounded integer -- are those of the canonical arithmetic under another exponent: the
 * "bias" of the integer tail is 0x32800000 instead of 0x4B000000.  Blackwell multiplies denormal operands at full rate
 * (tools/ubench/pipes.cu).  Against fma(2^23 + b, w, -w * 2^23) (round 1) this removes one FMUL per tap and quad of frames and
 * turns the products from three-source FFMA2 into FMUL2; the sums are FADD2.FTZ (add2_ftz), which ptxas cannot contract.
 * ---------------------------------------------------------------------------------------------- */
constexpr size_t HOIST4_SMEM = (size_t)HT * 4 + 4 * (size_t)HRING;
constexpr float HOIST4_WSCALE = 1.2676506002282294e30f; /* 2^100 */
constexpr float HOIST4_ROUND = 1.4901161193847656e-08f; /* 2^23 * 2^-49 = 2^-26 */

/* The integer tail of resampling.cl:82-91 (green = mean of the two green planes, dRGB = (2x - y - z + 510) >> 2, packed RGBA)
 * for TWO frames at once in the 16-bit halves of a register.  The rounded sums leave bias + n with n <= 255 in the low mantissa
 * bits, so PRMT packs (n of frame A, n of frame B); all of the tail is integer arithmetic mod 2^32 on A + 2^16 B whose final
 * per-lane values (2x - y - z + 510) * 64 lie in [0, 65280]: whatever an intermediate borrows from or carries into the other lane
 * is returned by the end.  The factor 64 puts the quotient's eight bits (bits 2..9) into byte 1 of each lane, where PRMT picks
 * them up -- no shifts, no masks.  17 instructions per frame pair against 14 per frame on the biased words (drgb_biased). */
__device__ __forceinline__ void drgb_biased_pair(float2 r, float2 g1, float2 g2, float2 b, uint32_t& px_a, uint32_t& px_b)
{
	const uint32_t R = __byte_perm(__float_as_uint(r.x), __float_as_uint(r.y), 0x5410), B = __byte_perm(__float_as_uint(b.x), __float_as_uint(b.y), 0x5410);
	const uint32_t G1 = __byte_perm(__float_as_uint(g1.x), __float_as_uint(g1.y), 0x5410), G2 = __byte_perm(__float_as_uint(g2.x), __float_as_uint(g2.y), 0x5410);
	const uint32_t G = ((G1 & 0xFFFEFFFEu) + (G2 & 0xFFFEFFFEu)) >> 1; /* (g1 >> 1) + (g2 >> 1) in each lane: the sum is even, no bit crosses */
	const uint32_t t = 510u * 64u * 0x00010001u - (R + G + B) * 64u;
	const uint32_t dr = R * 192u + t, dg = G * 192u + t, db = B * 192u + t; /* (3x - s + 510) << 6 */
	const uint32_t rg = __byte_perm(dr, dg, 0x7351);           /* dr A, dg A, dr B, dg B */
	const uint32_t ba = __byte_perm(db, 0xFFFFFFFFu, 0x4341);  /* db A, 255, db B, 255 */
	px_a = __byte_perm(rg, ba, 0x5410);
	px_b = __byte_perm(rg, ba, 0x7632);
}

/* w * 2^100 times (byte LO, byte LO+1) of a staged word as denormals, two frames at once: bit for bit mul.rn(w, float(b)) * 2^-49 */
template <int LO>
__device__ __forceinline__ float2 weighted_pair(uint32_t wd, float w)
{
	return mul2(make_float2(__uint_as_float(__byte_perm(wd, 0u, 0x4440 + LO)), __uint_as_float(__byte_perm(wd, 0u, 0x4441 + LO))), make_float2(w, w));
}

template <int FMT, bool FULL>
__device__ __forceinline__ void hoist4_blend(const uint32_t* __restrict__ T, const float2 (&W)[4][4], const int (&O)[4][4], uint32_t* __restrict__ out,
                                             uint32_t nfl, int wf, bool okx, int rows_ok, int n_valid)
{
	const size_t nfl4 = (size_t)nfl * 4, wf4 = (size_t)wf * 4;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		float2 ab[4], cd[4]; /* channel c: (frame A, frame B) and (frame C, frame D) */
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const uint32_t* t = T + O[k][c];
#pragma unroll
			for (int tap = 0; tap < 4; tap++) {
				const uint32_t wd = t[(tap & 1) + (tap >> 1) * HP];
				/* W[k] holds the eight axis values of the pixel (ox, ax | oy, ay, each for the +0.25 and the -0.25 plane): the weight
				 * is formed here, with the same single rounding as a weight kept in a register */
				const float2 xw = W[k][tap & 1];        /* (+0.25 axis, -0.25 axis): ox for tap 0/2, ax for tap 1/3 */
				const float2 yw = W[k][2 + (tap >> 1)]; /* (y axis of channels 0/1, y axis of channels 2/3): oy, ay */
				float w; /* volatile: the product is frame-invariant and would be hoisted out of the frame loop again (64 registers) */
				asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(w) : "f"((c & 1) ? xw.y : xw.x), "f"((c >> 1) ? yw.y : yw.x));
				const float2 pab = weighted_pair<0>(wd, w);
				const float2 pcd = weighted_pair<2>(wd, w);
				/* ((p00 + p10) + p01) + p11, every sum rounded on its own */
				ab[c] = tap == 0 ? pab : add2_ftz(pab, ab[c]);
				cd[c] = tap == 0 ? pcd : add2_ftz(pcd, cd[c]);
			}
			ab[c] = add2(ab[c], make_float2(HOIST4_ROUND, HOIST4_ROUND)); /* RNE to integer in the mantissa */
			cd[c] = add2(cd[c], make_float2(HOIST4_ROUND, HOIST4_ROUND));
		}
		const bool in = FULL || (okx && 4 * k < rows_ok);
		/* the store address walks frame by frame and then to the next row with 64-bit ADDS (IADD3 + IADD3.X: the cheap integer
		 * path) instead of a scaled 64-bit address per store (LEA + LEA.HI.X on the ALU pipe that the PRMTs saturate) */
		unsigned char* o = reinterpret_cast<unsigned char*>(out) + (size_t)(4 * k) * wf4;
		constexpr int CR = FMT == FMT_RGGB ? 0 : 1, CG1 = FMT == FMT_RGGB ? 1 : 0, CG2 = FMT == FMT_RGGB ? 2 : 3, CB = FMT == FMT_RGGB ? 3 : 2;
#pragma unroll
		for (int j = 0; j < 4; j += 2) {
			uint32_t pa, pb;
			if (j == 0)
				drgb_biased_pair(ab[CR], ab[CG1], ab[CG2], ab[CB], pa, pb);
			else
				drgb_biased_pair(cd[CR], cd[CG1], cd[CG2], cd[CB], pa, pb);
			if (in && j < n_valid)
				*reinterpret_cast<uint32_t*>(o) = pa;
			o += nfl4;
			if (in && j + 1 < n_valid)
				*reinterpret_cast<uint32_t*>(o) = pb;
			o += nfl4;
		}
	}
}

template <int FMT>
__global__ void __launch_bounds__(256, 2) k_reproject_hoist4(const uint8_t* __restrict__ raw0, size_t frame_stride, const float2* __restrict__ lut,
                                                               const TileEntry* __restrict__ table, uint32_t* __restrict__ flat, int wq, int hq,
                                                               int wf, int hf, int n_frames, int chunk)
{
	extern __shared__ __align__(16) unsigned char hoist_smem[];
	uint32_t* const T0 = reinterpret_cast<uint32_t*>(hoist_smem);
	unsigned char* const ring = hoist_smem + (size_t)HT * 4;
	const int tx = blockIdx.x, ty = blockIdx.y;
	const int f0 = blockIdx.z * chunk;
	const int n = min(n_frames, f0 + chunk) - f0;
	const TileEntry e = table[ty * gridDim.x + tx];
	const int tid = threadIdx.x;
	const int lx = tid & 63, ly = tid >> 6;
	const int gx = tx * FT_W + lx;
	const uint32_t nfl = (uint32_t)wf * (uint32_t)hf;
	const uint8_t* const raw = raw0 + (size_t)f0 * frame_stride;
	float2 pos[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int gy = ty * FT_H + ly + 4 * k;
		pos[k] = make_float2(0.f, 0.f);
		if (gx < wf && gy < hf)
			pos[k] = __ldg(lut + (gy * wf + gx));
	}

	if (!(e.flags & 1)) { /* footprint does not fit: direct gather, frame by frame */
#pragma unroll 1
		for (int f = 0; f < n; f++) {
			const uint8_t* rawf = raw + (size_t)f * frame_stride;
			uint32_t* out = flat + (size_t)(f0 + f) * nfl;
#pragma unroll 1
			for (int k = 0; k < 4; k++) {
				const int gy = ty * FT_H + ly + 4 * k;
				if (gx < wf && gy < hf) {
					const float2 q = __ldg(lut + (gy * wf + gx));
					uint32_t v;
					if (fabsf(q.x) < 1048576.0f && fabsf(q.y) < 1048576.0f) {
						v = reproject_bayer_rte_fast<FMT>(rawf, wq, hq, q.x, q.y);
					} else {
						uint32_t r, g, b;
						const SrcBayer s{ rawf, 2 * wq };
						demosaic<FMT, MODE_RTE>(s, wq, hq, q.x, q.y, r, g, b);
						v = drgb(r, g, b);
					}
					out[gy * wf + gx] = v;
				}
			}
		}
		return;
	}

	/* ---- frame-invariant staging plan: up to two 16-byte raw vectors per thread and frame ---- */
	const int row_bytes = 2 * wq;
	const bool vec = (e.flags & 2) != 0;
	const int n_rr = 2 * e.height;
	int s_src[2], s_dst[2], s_edge[2]; /* s_edge < 0: no vector */
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const int v = tid + 256 * i;
		const int q4 = v / (4 * HNV), w = v - q4 * (4 * HNV);
		const int half = w >= 2 * HNV ? 1 : 0, w2 = w - half * 2 * HNV;
		const int cv = w2 >> 1, rr = 4 * q4 + half + 2 * (w2 & 1);
		s_src[i] = 0;
		s_dst[i] = 0;
		s_edge[i] = -1;
		if (vec && rr < n_rr) {
			const int qy = clampi(e.jb + (rr >> 1), 0, hq - 1);
			const int qx0 = e.ib + cv * 8;
			const int qxc = clampi(qx0, 0, wq - 8);
			s_src[i] = (2 * qy + (rr & 1)) * row_bytes + 2 * qxc;
			s_dst[i] = (rr & 1) * 2 * HPLANE + (rr >> 1) * HP + cv * 8;
			s_edge[i] = qx0 < 0 ? 1 : (qx0 != qxc ? 2 : 0);
		}
	}
	unsigned char* const my_ring = ring + tid * 16; /* frame j, vector i: my_ring + j*HRING + i*4096 */
	const int n_quads = (n + 3) >> 2;
	const uint8_t* next_src = raw; /* first frame of the next quad to copy */
	auto issue_copy = [&](int q) { /* frames 4q..4q+3 (the last frame repeated past the end) into the raw stage; always commits */
		if (vec && q < n_quads) {
			unsigned char* dst = my_ring;
			const int left = n - 4 * q; /* >= 1 */
			const uint8_t* src = next_src;
#pragma unroll
			for (int j = 0; j < 4; j++) {
#pragma unroll
				for (int i = 0; i < 2; i++)
					if (s_edge[i] >= 0)
						cp_async16(dst + j * HRING + i * 4096, src + s_src[i]);
				if (j + 1 < left)
					src += frame_stride;
			}
			next_src += 4 * frame_stride;
		}
		cp_async_commit();
	};
	auto convert = [&](int q) { /* this thread's vectors of quad q: byte transpose of the four frames -> T */
		uint32_t* const T = T0;
		if (vec) {
#pragma unroll
			for (int i = 0; i < 2; i++) {
				if (s_edge[i] < 0)
					continue;
				uint32_t fr[4][4];
#pragma unroll
				for (int j = 0; j < 4; j++) {
					uint4 qq = *reinterpret_cast<const uint4*>(my_ring + j * HRING + i * 4096);
					if (s_edge[i]) { /* replicate the edge quad's two bytes over the whole vector */
						const uint32_t eq = s_edge[i] == 1 ? (qq.x & 0xFFFFu) : (qq.w >> 16);
						qq.x = qq.y = qq.z = qq.w = eq * 0x00010001u;
					}
					fr[j][0] = qq.x; fr[j][1] = qq.y; fr[j][2] = qq.z; fr[j][3] = qq.w;
				}
				uint32_t pe[8], po[8]; /* plane 2s (even bytes) and 2s+1 (odd bytes), 8 texels each, byte j of a word = frame j */
#pragma unroll
				for (int w = 0; w < 4; w++) {
					const uint32_t ab_lo = __byte_perm(fr[0][w], fr[1][w], 0x5140), ab_hi = __byte_perm(fr[0][w], fr[1][w], 0x7362);
					const uint32_t cd_lo = __byte_perm(fr[2][w], fr[3][w], 0x5140), cd_hi = __byte_perm(fr[2][w], fr[3][w], 0x7362);
					pe[2 * w] = __byte_perm(ab_lo, cd_lo, 0x5410);     /* byte 0 of the four frames */
					po[2 * w] = __byte_perm(ab_lo, cd_lo, 0x7632);     /* byte 1 */
					pe[2 * w + 1] = __byte_perm(ab_hi, cd_hi, 0x5410); /* byte 2 */
					po[2 * w + 1] = __byte_perm(ab_hi, cd_hi, 0x7632); /* byte 3 */
				}
				uint32_t* d0 = T + s_dst[i];
				uint32_t* d1 = d0 + HPLANE;
				reinterpret_cast<uint4*>(d0)[0] = make_uint4(pe[0], pe[1], pe[2], pe[3]);
				reinterpret_cast<uint4*>(d0)[1] = make_uint4(pe[4], pe[5], pe[6], pe[7]);
				reinterpret_cast<uint4*>(d1)[0] = make_uint4(po[0], po[1], po[2], po[3]);
				reinterpret_cast<uint4*>(d1)[1] = make_uint4(po[4], po[5], po[6], po[7]);
			}
		} else { /* raw rows not 16-byte aligned (wq % 8 != 0): per-texel gather with the edge replicated */
			const int tot = 4 * e.height * TQ_W;
			for (int v = tid; v < tot; v += 256) {
				const int rc = v / TQ_W, ii = v - rc * TQ_W; /* rc = 4*jj + c */
				const int jj = rc >> 2, c = rc & 3;
				const int qx = clampi(e.ib + ii, 0, wq - 1), qy = clampi(e.jb + jj, 0, hq - 1);
				const int o = (2 * qy + (c >> 1)) * row_bytes + 2 * qx + (c & 1);
				uint32_t wd = 0;
#pragma unroll
				for (int j = 0; j < 4; j++)
					wd |= (uint32_t)__ldg(raw + (size_t)min(4 * q + j, n - 1) * frame_stride + o) << (8 * j);
				T[c * HPLANE + jj * HP + ii] = wd;
			}
		}
	};
	issue_copy(0);

	/* ---- frame-invariant part: weights and tap offsets of this thread's four pixels (the first copy is in flight) ---- */
	float2 W[4][4];
	int O[4][4];
	const int xmagic = 0x4B400000 + e.ib, ymagic = 0x4B400000 + e.jb;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int gy = ty * FT_H + ly + 4 * k;
		const bool ok = gx < wf && gy < hf;
		int ixp, ixn, iyp, iyn;
		float2 ax, ox, ay, oy; /* .x = the +0.25 axis, .y = the -0.25 axis */
		axis_staged2(add2(make_float2(pos[k].x, pos[k].x), make_float2(0.25f, -0.25f)), xmagic, ixp, ixn, ax, ox);
		axis_staged2(add2(make_float2(pos[k].y, pos[k].y), make_float2(0.25f, -0.25f)), ymagic, iyp, iyn, ay, oy);
		/* the factor 2^100 of the weights rides on the x-axis values (exact: a power of two) */
		ox = mul2(ox, make_float2(HOIST4_WSCALE, HOIST4_WSCALE));
		ax = mul2(ax, make_float2(HOIST4_WSCALE, HOIST4_WSCALE));
		W[k][0] = ox; W[k][1] = ax; W[k][2] = oy; W[k][3] = ay; /* oy = (+y plane, -y plane), ay likewise */
		O[k][0] = ok ? iyp * HP + ixp : 0; /* pixels outside the image read texel 0 and are not stored */
		O[k][1] = ok ? HPLANE + iyp * HP + ixn : 0;
		O[k][2] = ok ? 2 * HPLANE + iyn * HP + ixp : 0;
		O[k][3] = ok ? 3 * HPLANE + iyn * HP + ixn : 0;
	}
	const bool okx = gx < wf;
	const int rows_ok = hf - (ty * FT_H + ly);
	const bool full = (tx + 1) * FT_W <= wf && (ty + 1) * FT_H <= hf;
	uint32_t* out = flat + (size_t)f0 * nfl + ((uint32_t)(ty * FT_H + ly) * (uint32_t)wf + (uint32_t)gx);
#pragma unroll 1
	for (int q = 0; q < n_quads; q++) {
		cp_async_wait<0>(); /* quad q has landed (this thread's vectors) */
		convert(q);
		issue_copy(q + 1); /* a thread refills its own ring slots as soon as it has converted them: the copy flies under the blend */
		__syncthreads();
		const int n_valid = min(4, n - 4 * q);
		if (full)
			hoist4_blend<FMT, true>(T0, W, O, out, nfl, wf, true, 16, n_valid);
		else
			hoist4_blend<FMT, false>(T0, W, O, out, nfl, wf, okx, rows_ok, n_valid);
		out += (size_t)4 * nfl;
		__syncthreads(); /* everybody has read the planes before the next quad is converted into them */
	}
}

/* ------------------------------------------------------------------------------------------------
 * raw2quad.cl:21-39 (stage API only; the fused path never materialises the planes)
 * ---------------------------------------------------------------------------------------------- */
__global__ void k_raw2quad_bayer(const uint8_t* __restrict__ raw, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1,
                                 uint8_t* __restrict__ c2, uint8_t* __restrict__ c3, int wq, int hq)
{
	/* four quads per thread when the row allows 8-byte loads / 4-byte stores */
	const int groups = (wq + 3) / 4;
	const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if (gidx >= groups || y >= hq)
		return;
	const int x0 = gidx * 4;
	const size_t row = 2 * (size_t)wq;
	const uint8_t* r0 = raw + 2 * (size_t)y * row + 2 * x0;
	const uint8_t* r1 = r0 + row;
	const size_t o = (size_t)y * wq + x0;
	if ((wq & 3) == 0) {
		const uint2 a = __ldg(reinterpret_cast<const uint2*>(r0));
		const uint2 b = __ldg(reinterpret_cast<const uint2*>(r1));
		*reinterpret_cast<uint32_t*>(c0 + o) = __byte_perm(a.x, a.y, 0x6420);
		*reinterpret_cast<uint32_t*>(c1 + o) = __byte_perm(a.x, a.y, 0x7531);
		*reinterpret_cast<uint32_t*>(c2 + o) = __byte_perm(b.x, b.y, 0x6420);
		*reinterpret_cast<uint32_t*>(c3 + o) = __byte_perm(b.x, b.y, 0x7531);
	} else {
		for (int k = 0; k < 4 && x0 + k < wq; k++) {
			c0[o + k] = r0[2 * k];
			c1[o + k] = r0[2 * k + 1];
			c2[o + k] = r1[2 * k];
			c3[o + k] = r1[2 * k + 1];
		}
	}
}

__global__ void k_raw2quad_bgr(const uint8_t* __restrict__ raw, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1,
                               uint8_t* __restrict__ c2, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	c0[i] = __ldg(raw + 3 * (size_t)i);
	c1[i] = __ldg(raw + 3 * (size_t)i + 1);
	c2[i] = __ldg(raw + 3 * (size_t)i + 2);
}

/* ------------------------------------------------------------------------------------------------
 * gradient dot product, gradientDot.cl:22-30.
 * sum_c (R_c - L_c)(U_c - D_c) = R.U - R.D - L.U + L.D as four byte dot products (alpha masked off R and L);
 * every value is an integer of magnitude <= 195075, so the int32 result converted to fp32 equals the
 * reference's float arithmetic bit for bit.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int grad_dot_px(uint32_t R, uint32_t L, uint32_t U, uint32_t D)
{
	R &= 0x00FFFFFFu;
	L &= 0x00FFFFFFu;
	const uint32_t pos = __dp4a(L, D, __dp4a(R, U, 0u));
	const uint32_t neg = __dp4a(L, U, __dp4a(R, D, 0u));
	return (int)pos - (int)neg;
}

/* stage API: gradient only (float out) */
__global__ void k_gradient_dot(const uint32_t* __restrict__ in, float* __restrict__ out, int w, int h, int o)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= w || y >= h)
		return;
	const uint32_t* row = in + (size_t)y * w;
	const uint32_t R = __ldg(row + min(x + o, w - 1)), L = __ldg(row + max(x - o, 0));
	const uint32_t U = __ldg(in + (size_t)min(y + o, h - 1) * w + x), D = __ldg(in + (size_t)max(y - o, 0) * w + x);
	out[(size_t)y * w + x] = (float)grad_dot_px(R, L, U, D);
}

/* K2a: gradient + exact row prefix sums.  One warp per row; per 128-pixel segment each lane owns 4
 * consecutive pixels (16-byte loads/stores when wf % 4 == 0), local prefix + warp-shuffle scan + running
 * carry.  Writes gradDot (fp32, API output) and the int32 row sums (internal).  flag[frame] is raised when a
 * row sum leaves the exact range of fp32 (satHorizontal.cl:26-31 would start rounding). */
/* gradient + exact row prefix sums of one image row by one warp; `srow` may point to global or shared memory.
 * Returns true if a row sum left the exactness bound. */

/* (the two helpers below restate the body of row_gradscan for k_grad_rowscan_wide; row_gradscan keeps its own copy because
 * ptxas allocates 8 registers more for it when it is written in terms of them, which costs the batch path occupancy) */
/* gradDot of the four pixels x0..x0+3 of row y (gradientDot.cl:22-30; pixels beyond the row end give 0); `vec`: rows are
 * 16-byte aligned, `pairs`: the R/L taps of the four pixels are two aligned 8-byte pairs each */
__device__ __forceinline__ void grad_dot4(const uint32_t* __restrict__ row, const uint32_t* __restrict__ up, const uint32_t* __restrict__ dn, int x0, int wf,
                                          int o, bool vec, bool pairs, int (&g)[4])
{
	uint32_t U[4], D[4], Rr[4], Ll[4];
	if (vec) {
		const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(up + x0));
		const uint4 d4 = __ldg(reinterpret_cast<const uint4*>(dn + x0));
		U[0] = u4.x; U[1] = u4.y; U[2] = u4.z; U[3] = u4.w;
		D[0] = d4.x; D[1] = d4.y; D[2] = d4.z; D[3] = d4.w;
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int x = min(x0 + k, wf - 1);
			U[k] = __ldg(up + x);
			D[k] = __ldg(dn + x);
		}
	}
	if (pairs && x0 - o >= 0 && x0 + 3 + o <= wf - 1) { /* no tap of these 4 pixels clamps */
		const uint2 r0 = __ldg(reinterpret_cast<const uint2*>(row + x0 + o)), r1 = __ldg(reinterpret_cast<const uint2*>(row + x0 + o + 2));
		const uint2 l0 = __ldg(reinterpret_cast<const uint2*>(row + x0 - o)), l1 = __ldg(reinterpret_cast<const uint2*>(row + x0 - o + 2));
		Rr[0] = r0.x; Rr[1] = r0.y; Rr[2] = r1.x; Rr[3] = r1.y;
		Ll[0] = l0.x; Ll[1] = l0.y; Ll[2] = l1.x; Ll[3] = l1.y;
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int x = x0 + k;
			Rr[k] = __ldg(row + min(x + o, wf - 1));
			Ll[k] = __ldg(row + clampi(x - o, 0, wf - 1));
		}
	}
#pragma unroll
	for (int k = 0; k < 4; k++)
		g[k] = x0 + k < wf ? grad_dot_opaque(Rr[k], Ll[k], U[k], D[k]) : 0;
}

/* gradDot and row prefix sums of four pixels to memory; returns true if a prefix left the exactness bound */
template <class SumT>
__device__ __forceinline__ bool store_grad4(float* __restrict__ grow, SumT* __restrict__ srow, int x0, int wf, bool vec, const int (&g)[4], int base)
{
	const int p1 = g[0] + g[1], p2 = p1 + g[2], p3 = p2 + g[3];
	const int s0 = base + g[0], s1 = base + p1, s2 = base + p2, s3 = base + p3;
	const bool bad = max(max(abs(s0), abs(s1)), max(abs(s2), abs(s3))) >= SAT_EXACT_LIMIT;
	if (vec) {
		*reinterpret_cast<float4*>(grow + x0) = make_float4(small_int_to_float(g[0]), small_int_to_float(g[1]), small_int_to_float(g[2]), small_int_to_float(g[3]));
		if constexpr (std::is_integral<SumT>::value)
			*reinterpret_cast<int4*>(srow + x0) = make_int4(s0, s1, s2, s3);
		else /* fp32: exact below 2^22; anything at or above SAT_EXACT_LIMIT raises the flag and is never used */
			*reinterpret_cast<float4*>(srow + x0) = make_float4(small_int_to_float(s0), small_int_to_float(s1), small_int_to_float(s2), small_int_to_float(s3));
	} else {
		const int s[4] = { s0, s1, s2, s3 };
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (x0 + k < wf) {
				grow[x0 + k] = (float)g[k];
				srow[x0 + k] = (SumT)s[k];
			}
	}
	return bad;
}

template <class SumT>
__device__ __forceinline__ bool row_gradscan(const uint32_t* __restrict__ img, int y, int wf, int hf, int o, int lane, float* __restrict__ grow,
                                             SumT* __restrict__ srow)
{
	const uint32_t* row = img + y * wf;
	const uint32_t* up = img + min(y + o, hf - 1) * wf;
	const uint32_t* dn = img + max(y - o, 0) * wf;
	const bool vec = (wf & 3) == 0;
	const bool pairs = vec && (o & 1) == 0; /* R/L taps of a lane's 4 pixels are two aligned 8-byte pairs each */
	int carry = 0;
	bool bad = false;
	for (int x0 = lane * 4; x0 - lane * 4 < wf; x0 += 128) {
		int g[4] = { 0, 0, 0, 0 };
		if (x0 < wf) {
			uint32_t U[4], D[4], Rr[4], Ll[4];
			if (vec) {
				const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(up + x0));
				const uint4 d4 = __ldg(reinterpret_cast<const uint4*>(dn + x0));
				U[0] = u4.x; U[1] = u4.y; U[2] = u4.z; U[3] = u4.w;
				D[0] = d4.x; D[1] = d4.y; D[2] = d4.z; D[3] = d4.w;
			} else {
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const int x = min(x0 + k, wf - 1);
					U[k] = __ldg(up + x);
					D[k] = __ldg(dn + x);
				}
			}
			if (pairs && x0 - o >= 0 && x0 + 3 + o <= wf - 1) { /* no tap of these 4 pixels clamps */
				const uint2 r0 = __ldg(reinterpret_cast<const uint2*>(row + x0 + o)), r1 = __ldg(reinterpret_cast<const uint2*>(row + x0 + o + 2));
				const uint2 l0 = __ldg(reinterpret_cast<const uint2*>(row + x0 - o)), l1 = __ldg(reinterpret_cast<const uint2*>(row + x0 - o + 2));
				Rr[0] = r0.x; Rr[1] = r0.y; Rr[2] = r1.x; Rr[3] = r1.y;
				Ll[0] = l0.x; Ll[1] = l0.y; Ll[2] = l1.x; Ll[3] = l1.y;
			} else {
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const int x = x0 + k;
					Rr[k] = __ldg(row + min(x + o, wf - 1));
					Ll[k] = __ldg(row + clampi(x - o, 0, wf - 1));
				}
			}
#pragma unroll
			for (int k = 0; k < 4; k++)
				g[k] = x0 + k < wf ? grad_dot_opaque(Rr[k], Ll[k], U[k], D[k]) : 0;
		}
		int p1 = g[0] + g[1], p2 = p1 + g[2], p3 = p2 + g[3];
		int incl = p3;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d)
				incl += t;
		}
		const int base = carry + incl - p3; /* exclusive prefix of this lane's 4 pixels */
		carry += __shfl_sync(0xffffffffu, incl, 31);
		if (x0 < wf) {
			const int s0 = base + g[0], s1 = base + p1, s2 = base + p2, s3 = base + p3;
			bad |= max(max(abs(s0), abs(s1)), max(abs(s2), abs(s3))) >= SAT_EXACT_LIMIT;
			if (vec) {
				*reinterpret_cast<float4*>(grow + x0) = make_float4(small_int_to_float(g[0]), small_int_to_float(g[1]), small_int_to_float(g[2]), small_int_to_float(g[3]));
				if constexpr (std::is_integral<SumT>::value)
					*reinterpret_cast<int4*>(srow + x0) = make_int4(s0, s1, s2, s3);
				else /* fp32: exact below 2^22; anything at or above SAT_EXACT_LIMIT raises the flag and is never used */
					*reinterpret_cast<float4*>(srow + x0) = make_float4(small_int_to_float(s0), small_int_to_float(s1), small_int_to_float(s2), small_int_to_float(s3));
			} else {
				const int s[4] = { s0, s1, s2, s3 };
#pragma unroll
				for (int k = 0; k < 4; k++)
					if (x0 + k < wf) {
						grow[x0 + k] = (float)g[k];
						srow[x0 + k] = (SumT)s[k];
					}
			}
		}
	}
	return bad;
}

constexpr int ROWSCAN_WARPS = 8;
template <class SumT>
__global__ void __launch_bounds__(ROWSCAN_WARPS * 32) k_grad_rowscan(const uint32_t* __restrict__ flat, float* __restrict__ grad,
                                                                     SumT* __restrict__ rowsum, int wf, int hf, int o,
                                                                     int* __restrict__ flag)
{
	const int lane = threadIdx.x & 31;
	const int y = blockIdx.x * ROWSCAN_WARPS + (threadIdx.x >> 5);
	if (y >= hf)
		return;
	const size_t fbase = (size_t)blockIdx.y * wf * hf;
	if (row_gradscan(flat + fbase, y, wf, hf, o, lane, grad + fbase + y * wf, rowsum + fbase + y * wf))
		flag[blockIdx.y] = 1;
}

/* K2a for a LONE frame (latency path): one CTA per row, warp w takes pixels [128w, 128w+128) -- every load of the row is
 * in flight at once and the row sum is a two-level scan (lanes, then warps through shared memory) instead of a carry
 * handed from segment to segment.  A single frame offers k_grad_rowscan one warp per row = ~7 warps per SM, each
 * walking its row in ~10 dependent steps; here the same work is one memory round trip deep.  Integer sums: the result
 * does not depend on the order, bit-identical to k_grad_rowscan. */
constexpr int ROWWIDE_MAX_WARPS = 32;
template <class SumT>
__global__ void __launch_bounds__(ROWWIDE_MAX_WARPS * 32) k_grad_rowscan_wide(const uint32_t* __restrict__ flat, float* __restrict__ grad,
                                                                             SumT* __restrict__ rowsum, int wf, int hf, int o, int* __restrict__ flag)
{
	__shared__ int totals[ROWWIDE_MAX_WARPS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
	const int y = blockIdx.x;
	const size_t fbase = (size_t)blockIdx.y * wf * hf;
	const uint32_t* img = flat + fbase;
	const bool vec = (wf & 3) == 0;
	const bool pairs = vec && (o & 1) == 0;
	const int x0 = warp * 128 + lane * 4;
	int g[4] = { 0, 0, 0, 0 };
	if (x0 < wf)
		grad_dot4(img + y * wf, img + min(y + o, hf - 1) * wf, img + max(y - o, 0) * wf, x0, wf, o, vec, pairs, g);
	const int p3 = ((g[0] + g[1]) + g[2]) + g[3];
	int incl = p3;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d)
			incl += t;
	}
	if (lane == 31)
		totals[warp] = incl;
	__syncthreads();
	int before = lane < warp && lane < n_warps ? totals[lane] : 0; /* warps to the left of this one */
#pragma unroll
	for (int d = 16; d; d >>= 1)
		before += __shfl_xor_sync(0xffffffffu, before, d);
	bool bad = false;
	if (x0 < wf)
		bad = store_grad4(grad + fbase + y * wf, rowsum + fbase + y * wf, x0, wf, vec, g, before + incl - p3);
	if (bad)
		flag[blockIdx.y] = 1;
}

/* K2b: column prefix sums of the row sums -> SAT (satVertical.cl:22-31), exact in int32 (a
 * wrapped value can only appear after an exactly detected excursion beyond 2^24).  One CTA = 32 columns x all
 * rows; warp w owns rows [w*rpw, (w+1)*rpw) held in registers, warp totals are exchanged through shared memory. */
template <int RPW>
__global__ void __launch_bounds__(1024) k_colscan(const int32_t* __restrict__ rowsum, float* __restrict__ sat, int wf, int hf,
                                                  int rpw, int* __restrict__ flag)
{
	__shared__ int tot[32][33];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x = blockIdx.x * 32 + lane;
	const size_t fbase = (size_t)blockIdx.y * wf * hf;
	const int32_t* src = rowsum + fbase;
	float* dst = sat + fbase;
	const int y0 = warp * rpw;
	int v[RPW];
	int sum = 0;
#pragma unroll
	for (int k = 0; k < RPW; k++) {
		const int y = y0 + k;
		v[k] = (k < rpw && y < hf && x < wf) ? __ldg(src + (y * wf + x)) : 0;
	}
#pragma unroll
	for (int k = 0; k < RPW; k++) {
		sum += v[k];
		v[k] = sum;
	}
	tot[warp][lane] = sum;
	__syncthreads();
	int off = 0;
	for (int w = 0; w < warp; w++)
		off += tot[w][lane];
	bool bad = false;
#pragma unroll
	for (int k = 0; k < RPW; k++) {
		const int y = y0 + k;
		if (k < rpw && y < hf && x < wf) {
			const int s = v[k] + off;
			bad |= abs(s) >= SAT_EXACT_LIMIT;
			dst[y * wf + x] = (float)s;
		}
	}
	if (bad)
		flag[blockIdx.y] = 1;
}

/* Sequential-order SAT (bit-exact for ANY fp32 input): the stage API's sat_horizontal / sat_vertical, and the
 * fallback of the fused path for frames whose sums leave the exact range (flag != NULL: skip frames whose
 * flag is 0).  Horizontal: one warp per row, 32 coalesced values at a time, the running sum is carried through
 * the lanes in order, so the additions happen in exactly the order of satHorizontal.cl:26-31. */
__global__ void __launch_bounds__(256) k_sat_h_seq(const float* __restrict__ in, float* __restrict__ out, int w, int h,
                                                   const int* __restrict__ flag)
{
	if (flag && flag[blockIdx.y] == 0)
		return;
	const int lane = threadIdx.x & 31;
	const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
	if (y >= h)
		return;
	const size_t base = (size_t)blockIdx.y * w * h + (size_t)y * w;
	float sum = 0.f;
	for (int x0 = 0; x0 < w; x0 += 32) {
		const int x = x0 + lane;
		const float val = x < w ? in[base + x] : 0.f;
		float mine = 0.f;
		const int n = min(32, w - x0);
		for (int k = 0; k < n; k++) {
			sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, val, k));
			if (lane == k)
				mine = sum;
		}
		if (x < w)
			out[base + x] = mine;
	}
}

/* Vertical: one thread per column, rows in order (satVertical.cl:26-31); coalesced across the warp. */
__global__ void __launch_bounds__(128) k_sat_v_seq(const float* __restrict__ in, float* __restrict__ out, int w, int h,
                                                   const int* __restrict__ flag)
{
	if (flag && flag[blockIdx.y] == 0)
		return;
	const int x = blockIdx.x * 128 + threadIdx.x;
	if (x >= w)
		return;
	const size_t base = (size_t)blockIdx.y * w * h + x;
	float sum = 0.f;
	int y = 0;
	for (; y + 8 <= h; y += 8) {
		float t[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
			t[k] = in[base + (size_t)(y + k) * w];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			sum = __fadd_rn(sum, t[k]);
			out[base + (size_t)(y + k) * w] = sum;
		}
	}
	for (; y < h; y++) {
		sum = __fadd_rn(sum, in[base + (size_t)y * w]);
		out[base + (size_t)y * w] = sum;
	}
}

/* ------------------------------------------------------------------------------------------------
 * circularity, satBlobCenter.cl:22-42
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ float circle_px(const float* __restrict__ sat, int w, int h, int x, int y, int r, float div)
{
	const float* rp = sat + (size_t)clampi(y + r, 0, h - 1) * w;
	const float* r1 = sat + (size_t)clampi(y + 1, 0, h - 1) * w;
	const float* m1 = sat + (size_t)clampi(y - 1, 0, h - 1) * w;
	const float* mr = sat + (size_t)clampi(y - r, 0, h - 1) * w;
	const int xp = clampi(x + r, 0, w - 1), x1 = clampi(x + 1, 0, w - 1);
	const int xm = clampi(x - 1, 0, w - 1), xr = clampi(x - r, 0, w - 1);
	const float pp = __fadd_rn(__fsub_rn(__fsub_rn(__ldg(rp + xp), __ldg(r1 + xp)), __ldg(rp + x1)), __ldg(r1 + x1));
	const float pn = __fadd_rn(__fsub_rn(__fsub_rn(__ldg(mr + xp), __ldg(m1 + xp)), __ldg(mr + x1)), __ldg(m1 + x1));
	const float np = __fadd_rn(__fsub_rn(__fsub_rn(__ldg(rp + xr), __ldg(r1 + xr)), __ldg(rp + xm)), __ldg(r1 + xm));
	const float nn = __fadd_rn(__fsub_rn(__fsub_rn(__ldg(mr + xr), __ldg(m1 + xr)), __ldg(mr + xm)), __ldg(m1 + xm));
	return __fdiv_rn(min_cl(min_cl(pp, nn), min_cl(pn, np)), div);
}

__global__ void __launch_bounds__(256) k_circle(const float* __restrict__ sat, float* __restrict__ out, int w, int h, int r)
{
	const int x = blockIdx.x * 64 + (threadIdx.x & 63);
	const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
	if (x >= w || y >= h)
		return;
	const size_t fbase = (size_t)blockIdx.z * w * h;
	out[fbase + (size_t)y * w + x] = circle_px(sat + fbase, w, h, x, y, r, (float)(r * r));
}

/* ------------------------------------------------------------------------------------------------
 * blob list, blobList.cl:36-102 -- deterministic raster-order compaction
 * ---------------------------------------------------------------------------------------------- */

/* classification of one pixel: 0 below threshold, 1 not a local peak, 2 rejected by score, 3 blob */
struct PeakCtx {
	float c, cnx, cpx, cny, cpy;
};
__device__ __forceinline__ int peak_class(const uint32_t* __restrict__ img, const float* __restrict__ circ, int w, int h, int x, int y,
                                          float thr, float min_score, int radius, bool need_score, PeakCtx& p)
{
	p.c = __ldg(circ + (size_t)y * w + x);
	if (p.c < thr) /* blobList.cl:39 */
		return 0;
	p.cnx = __ldg(circ + (size_t)y * w + max(x - 1, 0));
	p.cpx = __ldg(circ + (size_t)y * w + min(x + 1, w - 1));
	p.cny = __ldg(circ + (size_t)max(y - 1, 0) * w + x);
	p.cpy = __ldg(circ + (size_t)min(y + 1, h - 1) * w + x);
	if (p.cnx > p.c || p.cpx > p.c || p.cny > p.c || p.cpy > p.c) /* :47-55 */
		return 1;
	if (need_score) {
		const DiscStats d = disc_stats(img, w, h, x, y, radius);
		if (blob_score(d, p.c) < min_score) /* :79 */
			return 2;
	}
	return 3;
}

/* Compaction scratch (per frame): rowcount[hf] = blobs per row, masks[hf][ceil(wf/32)] = one bit per blob pixel.
 * Pass A (k_grad_circ / k_circ_stream_rs on the fused path, k_peaks_count for the stage API and the generic radius) classifies every pixel once, writes the
 * bit masks, counts blobs per row and accumulates counter[0..2]; pass B (k_peaks_emit) ranks the set bits in raster
 * order (first slot + blobs of earlier rows + blobs to the left) and writes the records with rank < max_matches. */

/* publish one warp's classification of 32 consecutive pixels of row y (segment index seg).  The mask words were
 * zeroed by k_peaks_prepare, so only segments that hold a blob are written. */
__device__ __forceinline__ void publish_segment(int cls, int lane, int32_t* __restrict__ rowcount_f, uint32_t* __restrict__ masks_f,
                                                int wpr, int y, int seg, int& n_blob, int& n_score, int& n_peak)
{
	const unsigned m3 = __ballot_sync(0xffffffffu, cls == 3);
	const unsigned m2 = __ballot_sync(0xffffffffu, cls == 2);
	const unsigned m1 = __ballot_sync(0xffffffffu, cls == 1);
	if (lane == 0 && m3) {
		masks_f[y * wpr + seg] = m3;
		atomicAdd(rowcount_f + y, __popc(m3));
	}
	n_blob += __popc(m3);
	n_score += __popc(m2);
	n_peak += __popc(m1);
}


/* stage API pass A: a warp = 32 consecutive pixels of a row (x0 multiple of 32) */
__global__ void __launch_bounds__(256) k_peaks_count(const uint32_t* __restrict__ img, const float* __restrict__ circ, int w, int h,
                                                     float thr, float min_score, int radius, int need_score,
                                                     int32_t* __restrict__ counter, int32_t* __restrict__ rowcount, uint32_t* __restrict__ masks, int wpr)
{
	const int x = blockIdx.x * 256 + threadIdx.x;
	const int y = blockIdx.y;
	const int f = blockIdx.z;
	const size_t fbase = (size_t)f * w * h;
	int cls = 0;
	if (x < w) {
		PeakCtx p;
		cls = peak_class(img + fbase, circ + fbase, w, h, x, y, thr, min_score, radius, need_score != 0, p);
	}
	int nb = 0, ns = 0, np = 0;
	publish_segment(cls, threadIdx.x & 31, rowcount + (size_t)f * h, masks + (size_t)f * h * wpr, wpr, y, x >> 5, nb, ns, np);
	publish_counters(threadIdx.x & 31, counter + 3 * f, nb, ns, np);
}

/* ------------------------------------------------------------------------------------------------
 * Streaming circularity from row sums (the row-sum flow: A/B alternative of gradcirc.cuh's fused kernel, and the path of gradient
 * offsets it does not stage); no summed-area table, no separate border pass.  A warp walks DOWN a strip of the image, lane =
 * column; rings of depth R+2 and a row loop unrolled by R+2 make every ring index a compile-time constant.
 *
 * satBlobCenter.cl:37-40 only ever uses the summed-area table through four box sums, and a box sum over columns
 * (u, u+K] x rows (v, v+K] is  sum_{y in (v, v+K]} [RS(u+K, y) - RS(u, y)]  with RS the row prefix sums that
 * k_grad_rowscan already produces.  So the column scan (satVertical.cl) and the materialised SAT are not needed: the
 * streaming kernel keeps a K-row sliding window of horizontal box sums per lane,
 *     hrow(t) = RS(u+K, t) - RS(u, t),     Q(u, v = t-K) = Q(u, v-1) + hrow(t) - hrow(t-K),
 * all exact integers in fp32 inside the exactness bound -- bit-identical to the SAT form by the argument at
 * SAT_EXACT_LIMIT.  What the SAT was also needed for is the bound itself (|SAT| < 2^22 everywhere): every lane
 * accumulates the column sum of RS over its segment's rows and the largest magnitude the running sum reached;
 * k_sat_check_rs combines the segments (|SAT(x,y)| <= |carry(x, seg)| + max|local|, conservative) and raises the frame's
 * flag.  Flagged frames (never seen on camera images) are redone afterwards in the reference's sequential order
 * (k_sat_check_rs, then k_fallback_frame).
 * ---------------------------------------------------------------------------------------------- */

constexpr int CIRC_STREAM_MAX_R = 12;

template <int R>
__global__ void __launch_bounds__(128, (R <= 5 ? 6 : (R <= 7 ? 5 : 4))) k_circ_stream_rs(const float* __restrict__ rs, float* __restrict__ circ_out, const uint32_t* __restrict__ flat,
                                                     int w, int h, int seg_rows, float thr, float min_score, int radius, int need_score,
                                                     const int* __restrict__ flag, int32_t* __restrict__ counter, int32_t* __restrict__ rowcount,
                                                     uint32_t* __restrict__ masks, int wpr, float* __restrict__ segsum, float* __restrict__ segmax)
{
	constexpr int K = R - 1, D = R + 2;
	constexpr int LO = R + 2;          /* first output lane: needs Q from lane-(R+1) and a circularity from lane-1 */
	constexpr int SWU = 32 - LO - 1;   /* output lanes LO..30 */
	constexpr float DIV = (float)(R * R);
	constexpr float RCP = 1.0f / DIV;
	const int lane = threadIdx.x & 31;
	const int strip = blockIdx.x * 4 + (threadIdx.x >> 5);
	const int f = blockIdx.z;
	const int xs = strip * SWU;
	if (xs >= w || flag[f] != 0)
		return; /* frames whose row sums left the bound are done by the fallback pass */
	const int x = xs + lane - LO;                       /* this lane's pixel column (may be outside the image) */
	const int ys = blockIdx.y * seg_rows, ye = min(ys + seg_rows, h);
	const size_t fbase = (size_t)f * w * h;
	const float* rsf = rs + fbase;
	float* circf = circ_out + fbase;
	const uint32_t* flatf = flat + fbase;
	const bool x_in = x >= 0 && x < w;
	const bool out_lane = lane >= LO && lane <= 30 && x < w; /* x >= 0 follows from lane >= LO */
	int nb = 0, ns = 0, npk = 0;
	int32_t* rcf = rowcount + f * h;
	uint32_t* mkf = masks + (size_t)f * h * wpr;

	auto publish = [&](int cls, int yy) { /* warp-uniform call */
		if (cls == 3) {
			atomicOr(mkf + (yy * wpr + (x >> 5)), 1u << (x & 31));
			atomicAdd(rcf + yy, 1);
		}
		nb += __popc(__ballot_sync(0xffffffffu, cls == 3));
		ns += __popc(__ballot_sync(0xffffffffu, cls == 2));
		npk += __popc(__ballot_sync(0xffffffffu, cls == 1));
	};

	/* Columns of this lane's box: (u, u+K] with u = x+1, both CLAMPED into the row.  With clamped taps the reference's
	 * S(a,b) - S(a,c) - S(d,b) + S(d,c) is still the sum over rows (c, b] of RS(a) - RS(d), so image borders need no special
	 * case: a clamped column pair just gives a narrower (or empty) box, and rows outside [1, h-1] contribute nothing
	 * (clamp(y-r) = 0 starts the box at row 1; clamp(y+r) = h-1 ends it at row h-1). */
	const int ua = clampi(x + 1, 0, w - 1), ub = clampi(x + 1 + K, 0, w - 1);
	const float* pa = rsf + ua;
	const float* pb = rsf + ub;
	const bool box_full = __all_sync(0xffffffffu, ub - ua == K); /* false only for the strips at the left/right image edge */
	/* L2 prefetch plan: the strip's RS columns span two 128-byte lines per row; lanes 0..D-1 take the first line of the D
	 * rows of a group, lanes 16..16+D-1 the second */
	const int pf_row = lane & 15;
	const float* pf = rsf + (clampi(xs - LO + 1 + 32 * (lane >> 4), 0, w - 1) + pf_row * w);
	float* pc = circf + (x_in ? x : 0);

	float la[D], lb[D], hn[D], hold[K > 0 ? K : 1], qa[D], qb[D], cr[D];
#pragma unroll
	for (int i = 0; i < D; i++)
		hn[i] = qa[i] = qb[i] = cr[i] = 0.f;
	float qrun = 0.f;                 /* sum of the last K hrow values */
	float colrun = 0.f, colmax = 0.f; /* column sum of RS(ua, .) over this segment's rows, and its largest magnitude on the way */
	const int t0 = ys - 1 - R;
	const int n_groups = (ye + R - t0 + D) / D; /* whole groups: the extra rows of the last one are computed and never used */

	/* D rows of RS in flight per lane.  Rows 1..h-1 give hrow = RS(ub) - RS(ua); for every other row both loads read the
	 * SAME address, so hrow comes out as exactly zero without a select in the consumer. */
	auto load_group = [&](int t) {
		if (t >= 1 && t + D - 1 <= h - 1 && box_full) { /* warp-uniform: every row counts and no lane's columns are clamped */
#pragma unroll
			for (int s = 0; s < D; s++) {
				const float* q = elem_ptr(pa, (unsigned)((t + s) * w));
				la[s] = __ldg(q);
				lb[s] = __ldg(q + K); /* one address per row, the second load at an immediate offset */
			}
		} else if (t >= 1 && t + D - 1 <= h - 1) {
#pragma unroll
			for (int s = 0; s < D; s++) {
				const unsigned o = (unsigned)((t + s) * w);
				la[s] = __ldg(elem_ptr(pa, o));
				lb[s] = __ldg(elem_ptr(pb, o));
			}
		} else {
#pragma unroll
			for (int s = 0; s < D; s++) {
				const int row = t + s;
				const unsigned o = (unsigned)(clampi(row, 0, h - 1) * w);
				la[s] = __ldg(elem_ptr(pa, o));
				lb[s] = __ldg(elem_ptr(row >= 1 && row <= h - 1 ? pb : pa, o));
			}
		}
	};

	/* One group = D rows, straight-line code: the shuffles and the short dependent chains of different rows overlap; ONE
	 * vote per group decides whether any pixel reaches the threshold at all (almost never), and only then are the rows
	 * classified one by one.  INNER groups (every row of the group is owned by the segment, both as a summed row and as an
	 * output row) carry no per-row range predicates. */
	auto group = [&](auto INNER_C, int g) {
		constexpr bool INNER = decltype(INNER_C)::value;
		const int t = t0 + g * D;
#pragma unroll
		for (int i = 0; i < K; i++)
			hold[i] = hn[D - K + i];
#pragma unroll
		for (int s = 0; s < D; s++) {
			hn[s] = __fsub_rn(lb[s], la[s]); /* hrow(t+s) */
			if (INNER || (t + s >= ys && t + s < ye)) { /* rows this segment owns */
				colrun = __fadd_rn(colrun, la[s]);
				colmax = fmaxf(colmax, fabsf(colrun));
			}
		}
		if (g + 1 < n_groups)
			load_group(t + D);
		if (INNER && t + 3 * D <= h && pf_row < D) /* pull the group after the next one into L2 with ONE instruction: lane = (row, line) */
			asm volatile("prefetch.global.L2 [%0];" ::"l"(elem_ptr(pf, (unsigned)((t + 2 * D) * w))));

		constexpr int DD = 4 * D;
		const int y0 = t - R; /* circularity row of step 0 */
		float crow[D + 2]; /* circularity rows y0-2 .. y0+D-1 */
		crow[0] = cr[(0 - R - 2 + DD) % D];
		crow[1] = cr[(0 - R - 1 + DD) % D];
#pragma unroll
		for (int s = 0; s < D; s++) {
			const int so = (s - K + DD) % D, sq = (s - 2 * R + DD) % D, sc = (s - R + DD) % D;
			const float h_old = s - K >= 0 ? hn[s - K >= 0 ? s - K : 0] : hold[s - K >= 0 ? 0 : s];
			qrun = __fadd_rn(qrun, __fsub_rn(hn[s], h_old)); /* Q row v = t+s-K: rows v+1 .. v+K */
			const float q = qrun;
			qa[so] = q;
			qb[so] = __shfl_up_sync(0xffffffffu, q, R + 1);
			const float m = fminf(fminf(qa[so], qb[sq]), fminf(__fsub_rn(0.0f, qa[sq]), __fsub_rn(0.0f, qb[so])));
			const float q0 = __fmul_rn(m, RCP);
			const float c = __fmaf_rn(__fmaf_rn(-q0, DIV, m), RCP, q0); /* == m / (R*R), satBlobCenter.cl:41 */
			const int y = y0 + s;
			cr[sc] = c;
			crow[s + 2] = c;
			if (out_lane && (INNER || (y >= ys && y < ye)))
				*elem_ptr(pc, (unsigned)(y * w)) = c;
		}
		/* rows classified by this group: yy = y0-1 .. y0+D-2, i.e. crow[1 .. D] */
		float mx = crow[1];
#pragma unroll
		for (int s = 2; s <= D; s++)
			mx = fmaxf(mx, crow[s]);
		if (__any_sync(0xffffffffu, out_lane && !(mx < thr))) {
#pragma unroll 1
			for (int i = 1; i <= D; i++) {
				const int yy = y0 + i - 2;
				float cm = crow[1], up = crow[0], dn = crow[2];
#pragma unroll
				for (int k = 2; k <= D; k++)
					if (i == k) {
						cm = crow[k];
						up = crow[k - 1];
						dn = crow[k + 1];
					}
				const bool cand = out_lane && yy >= ys && yy < ye && !(cm < thr);
				if (!__any_sync(0xffffffffu, cand))
					continue;
				const float cl = __shfl_up_sync(0xffffffffu, cm, 1), crr = __shfl_down_sync(0xffffffffu, cm, 1);
				const int cls = cand ? classify_px(flatf, w, h, x, yy, radius, thr, min_score, need_score, cm, x > 0 ? cl : cm, x < w - 1 ? crr : cm,
				                                   yy > 0 ? up : cm, yy < h - 1 ? dn : cm)
				                     : 0;
				publish(cls, yy);
			}
		}
	};

	load_group(t0);
#pragma unroll 1
	for (int g = 0; g < n_groups; g++) {
		const int t = t0 + g * D;
		/* summed rows t .. t+D-1 and output rows t-R .. t-R+D-1 all inside [ys, ye) */
		if (t - R >= ys && t + D <= ye)
			group(IntC<1>{}, g);
		else
			group(IntC<0>{}, g);
	}
	publish_counters(lane, counter + 3 * f, nb, ns, npk);
	if (x + 1 >= 0 && x + 1 <= w - 1) { /* column ua is really this lane's (not a clamped duplicate) */
		const size_t i = ((size_t)f * gridDim.y + blockIdx.y) * w + ua;
		segsum[i] = colrun;
		segmax[i] = colmax;
	}
}

__device__ __forceinline__ void store_match(uint8_t* __restrict__ dst, float mx, float my, const uint32_t color[3], uint32_t center, float circ,
                                            float score)
{
	/* 22-byte packed record, 2-byte aligned: eleven 16-bit stores (sparse: a few hundred records per frame) */
	uint16_t* d = reinterpret_cast<uint16_t*>(dst);
	const uint32_t ux = __float_as_uint(mx), uy = __float_as_uint(my), uc = __float_as_uint(circ), us = __float_as_uint(score);
	d[0] = (uint16_t)ux; d[1] = (uint16_t)(ux >> 16);
	d[2] = (uint16_t)uy; d[3] = (uint16_t)(uy >> 16);
	d[4] = (uint16_t)(color[0] | (color[1] << 8));
	d[5] = (uint16_t)(color[2] | ((center & 255u) << 8));
	d[6] = (uint16_t)(((center >> 8) & 255u) | (((center >> 16) & 255u) << 8));
	d[7] = (uint16_t)uc; d[8] = (uint16_t)(uc >> 16);
	d[9] = (uint16_t)us; d[10] = (uint16_t)(us >> 16);
}

/* disc statistics of one blob computed by the whole warp (blobList.cl:63-72): the (2r+1)^2 window is dealt out to the
 * lanes position by position, four independent loads in flight per lane; integer sums, so the order is irrelevant */
__device__ __forceinline__ DiscStats disc_stats_warp(const uint32_t* __restrict__ img, int w, int h, int x, int y, int radius, int lane)
{
	DiscStats d;
	d.n = 0;
	d.s1[0] = d.s1[1] = d.s1[2] = d.s2[0] = d.s2[1] = d.s2[2] = 0;
	const int sq = radius * radius, side = 2 * radius + 1, area = side * side;
	for (int i0 = 0; i0 < area; i0 += 128) {
		uint32_t v[4];
		bool in[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int i = i0 + 32 * k + lane;
			const int row = i / side;
			const int dy = row - radius, dx = i - row * side - radius;
			in[k] = i < area && dx * dx + dy * dy <= sq;
			v[k] = in[k] ? __ldg(img + (size_t)clampi(y + dy, 0, h - 1) * w + clampi(x + dx, 0, w - 1)) : 0u;
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (in[k]) {
#pragma unroll
				for (int c = 0; c < 3; c++) {
					const uint32_t cc = (v[k] >> (8 * c)) & 255u;
					d.s1[c] += cc;
					d.s2[c] += cc * cc;
				}
				d.n++;
			}
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		d.n += __shfl_xor_sync(0xffffffffu, d.n, o);
#pragma unroll
		for (int k = 0; k < 3; k++) {
			d.s1[k] += __shfl_xor_sync(0xffffffffu, d.s1[k], o);
			d.s2[k] += __shfl_xor_sync(0xffffffffu, d.s2[k], o);
		}
	}
	return d;
}

/* the record of one blob pixel, blobList.cl:83-101; called by a full warp, lane 0 stores */
__device__ __forceinline__ void emit_match(const uint32_t* __restrict__ im, const float* __restrict__ ci, int w, int h, int x, int y, int radius,
                                           uint8_t* __restrict__ dst, int lane)
{
	/* requested before the disc statistics so that their latency overlaps (every lane loads; lane 0 uses them) */
	const float c = __ldg(ci + (size_t)y * w + x);
	const float cnx = __ldg(ci + (size_t)y * w + max(x - 1, 0)), cpx = __ldg(ci + (size_t)y * w + min(x + 1, w - 1));
	const float cny = __ldg(ci + (size_t)max(y - 1, 0) * w + x), cpy = __ldg(ci + (size_t)min(y + 1, h - 1) * w + x);
	const uint32_t centre = __ldg(im + (size_t)y * w + x);
	const DiscStats d = disc_stats_warp(im, w, h, x, y, radius, lane);
	if (lane != 0)
		return;
	const float score = blob_score(d, c);
	const float mx = __fadd_rn((float)x, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(cnx, cpx)), __fadd_rn(__fsub_rn(cnx, __fmul_rn(2.0f, c)), cpx))); /* :93 */
	const float my = __fadd_rn((float)y, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(cny, cpy)), __fadd_rn(__fsub_rn(cny, __fmul_rn(2.0f, c)), cpy))); /* :94 */
	const uint32_t color[3] = { d.s1[0] / (uint32_t)d.n, d.s1[1] / (uint32_t)d.n, d.s1[2] / (uint32_t)d.n }; /* :85 */
	store_match(dst, mx, my, color, centre, c, score);
}

/* The exactness bound of the summed-area table of one frame from the per-segment column sums of k_circ_stream_rs:
 * |SAT(x, y)| <= |sum of the segments above| + max |running sum inside the segment| (conservative).  Columns x_begin,
 * x_begin + x_step, ... < w per thread; called by a whole CTA, true when one of its columns may have left the bound. */
__device__ __forceinline__ bool sat_bound_exceeded(const float* __restrict__ segsum, const float* __restrict__ segmax, int n_seg, int w, int f, int x_begin,
                                                   int x_step)
{
	bool bad = false;
	for (int x = x_begin; x < w; x += x_step) {
		const float* ss = segsum + (size_t)f * n_seg * w + x;
		const float* sm = segmax + (size_t)f * n_seg * w + x;
		float carry = 0.f;
#pragma unroll 8
		for (int k = 0; k < n_seg; k++) { /* the loads do not depend on the carry: eight segments in flight */
			bad |= !(__fadd_rn(fabsf(carry), sm[k * w]) < (float)SAT_EXACT_LIMIT);
			carry = __fadd_rn(carry, ss[k * w]);
		}
	}
	return __syncthreads_or(bad);
}

constexpr int EMIT_MAX_ROWS = 128;
/* pass B: one warp per row that holds at least one blob; blobs are visited in x order, each one by the whole warp.
 * Latency path (segsum != nullptr): the grid has ceil(w/256) more CTAs per frame, which evaluate the exactness bound of
 * the SAT next to the record warps and raise the frame's flag for the HOST to see -- the caller then redoes the frame in
 * sequential order (k_fallback_frame + this kernel again) instead of launching the check and the fallback for
 * every frame just to have them exit. */
__global__ void __launch_bounds__(256) k_peaks_emit(const uint32_t* __restrict__ img, const float* __restrict__ circ, int w, int h, int radius,
                                                    int max_matches, const int32_t* __restrict__ first_slot, const int32_t* __restrict__ rowcount,
                                                    const uint32_t* __restrict__ masks, int wpr, uint8_t* __restrict__ matches, size_t match_frame_stride,
                                                    const float* __restrict__ segsum = nullptr, const float* __restrict__ segmax = nullptr, int n_seg = 0,
                                                    int* __restrict__ flag = nullptr, GcCheck gc = GcCheck(), int rows_per_cta = 8)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int f = blockIdx.y;
	const int n_row_ctas = (h + rows_per_cta - 1) / rows_per_cta;
	if (blockIdx.x >= n_row_ctas) { /* only launched when segsum is given: one thread per column */
		/* the sibling column CTAs of this frame may raise the flag while this one runs: the CTA takes ONE snapshot, so that
		 * every warp reaches the barrier inside sat_bound_exceeded or none does */
		__shared__ int flag_seen;
		if (threadIdx.x == 0)
			flag_seen = flag[f];
		__syncthreads();
		if (flag_seen != 0)
			return;
		/* fused gradient + circularity flow: one more CTA per row segment, which evaluates the segment's bound from k_grad_circ's side outputs;
		 * row-sum flow: one CTA per 256 columns over k_circ_stream_rs's column sums of the row sums */
		const bool bad = gc.striptot ? sat_bound_exceeded_g(gc, w, h, f, blockIdx.x - n_row_ctas)
		                             : sat_bound_exceeded(segsum, segmax, n_seg, w, f, (blockIdx.x - n_row_ctas) * 256 + threadIdx.x, w);
		if (bad && threadIdx.x == 0)
			flag[f] = 2;
		return;
	}
	/* rank of the CTA's first blob: blobs of all earlier rows, summed by the whole CTA (coalesced, one round trip deep); then
	 * the CTA's own rows.  A batch gives a CTA many rows (few, fat CTAs: the grid of one-warp-per-row CTAs cost more to
	 * launch than to run), a lone frame eight (every row at once). */
	__shared__ int s_part[8], s_cnt[EMIT_MAX_ROWS], s_pre[EMIT_MAX_ROWS];
	const int row0 = blockIdx.x * rows_per_cta;
	const int32_t* rc = rowcount + (size_t)f * h;
	int part = 0;
	for (int k = threadIdx.x; k < row0; k += 256)
		part += rc[k];
#pragma unroll
	for (int d = 16; d; d >>= 1)
		part += __shfl_xor_sync(0xffffffffu, part, d);
	if (lane == 0)
		s_part[warp] = part;
	if (threadIdx.x < rows_per_cta)
		s_cnt[threadIdx.x] = row0 + threadIdx.x < h ? rc[row0 + threadIdx.x] : 0;
	__syncthreads();
	if (threadIdx.x == 0) {
		int run = first_slot[f];
		for (int k = 0; k < 8; k++)
			run += s_part[k];
		for (int t = 0; t < rows_per_cta; t++) {
			s_pre[t] = run;
			run += s_cnt[t];
		}
	}
	__syncthreads();
	const size_t fbase = (size_t)f * w * h;
	uint8_t* out = matches + (size_t)f * match_frame_stride;
	/* The CTA's blobs are dealt out to its warps one by one, in rank order (blob i of the CTA -> warp i % 8): the blobs of a
	 * robot share a few rows, and a warp per ROW would leave most warps idle behind the one that owns those rows.  A warp
	 * finds its blob from the counts: the row by a vote over s_pre, the pixel as the n-th set bit of the row's mask words. */
	const int first = s_pre[0], last_t = rows_per_cta - 1;
	const int n_here = s_pre[last_t] + s_cnt[last_t] - first;
	for (int i = warp; i < n_here && first + i < max_matches; i += 8) {
		const int rank = first + i;
		int t = 0;
		for (int t0 = 0; t0 < rows_per_cta; t0 += 32) { /* the row whose rank range holds `rank` */
			const int tt = t0 + lane;
			const bool hit = tt < rows_per_cta && s_cnt[tt] > 0 && rank >= s_pre[tt] && rank < s_pre[tt] + s_cnt[tt];
			const unsigned v = __ballot_sync(0xffffffffu, hit);
			if (v)
				t = t0 + __ffs(v) - 1;
		}
		const int y = row0 + t;
		int n = rank - s_pre[t]; /* n-th blob of the row, counted from the left */
		const uint32_t* mk = masks + ((size_t)f * h + y) * wpr;
		int x = -1;
		for (int base = 0; base < wpr && x < 0; base += 32) {
			const uint32_t mine = base + lane < wpr ? mk[base + lane] : 0u;
			const int c = __popc(mine);
			int incl = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int u = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d)
					incl += u;
			}
			const int total = __shfl_sync(0xffffffffu, incl, 31);
			if (n < total) {
				const unsigned v = __ballot_sync(0xffffffffu, n < incl); /* first lane whose inclusive count exceeds n */
				const int src = __ffs(v) - 1;
				const uint32_t word = __shfl_sync(0xffffffffu, mine, src);
				const int skip = n - (__shfl_sync(0xffffffffu, incl, src) - __shfl_sync(0xffffffffu, c, src));
				x = (base + src) * 32 + (int)__fns(word, 0, skip + 1);
			} else {
				n -= total;
			}
		}
		emit_match(img + fbase, circ + fbase, w, h, x, y, radius, out + 22 * (size_t)rank, lane);
	}
}

/* per-batch preparation of the compaction scratch: zero the row counts and the exactness flags; either zero the
 * counters (fused path, main.cpp:283-288) or remember counter[0] as the first output slot (stage API). */
__global__ void k_peaks_prepare(int32_t* __restrict__ counter, int32_t* __restrict__ first_slot, int32_t* __restrict__ rowcount,
                                int n_rows_total, int n_frames, int zero_counters, int* __restrict__ flag, uint32_t* __restrict__ masks, int n_mask_words)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	for (int k = i; k < n_mask_words; k += gridDim.x * blockDim.x)
		masks[k] = 0u;
	if (i < n_rows_total)
		rowcount[i] = 0;
	if (i < n_frames) {
		if (zero_counters) {
			counter[3 * i] = counter[3 * i + 1] = counter[3 * i + 2] = 0;
			first_slot[i] = 0;
		} else {
			first_slot[i] = counter[3 * i];
		}
		if (flag)
			flag[i] = 0;
	}
}

/* blobScore.cl:23-66 (dead kernel): per-pixel score map */
__global__ void __launch_bounds__(256) k_blob_score(const uint32_t* __restrict__ img, const float* __restrict__ circ, float* __restrict__ out,
                                                    int w, int h, float thr, int radius)
{
	const int x = blockIdx.x * 256 + threadIdx.x;
	const int y = blockIdx.y;
	if (x >= w)
		return;
	PeakCtx p;
	const int cls = peak_class(img, circ, w, h, x, y, thr, 0.f, radius, false, p);
	float v = -INFINITY;
	if (cls == 3)
		v = blob_score(disc_stats(img, w, h, x, y, radius), p.c);
	out[(size_t)y * w + x] = v;
}

/* blobCenter.cl:29-63 (dead kernel): brute-force quadrant means */
__global__ void __launch_bounds__(256) k_circularize(const float* __restrict__ in, float* __restrict__ out, int w, int h, int maxr)
{
	const int px = blockIdx.x * 64 + (threadIdx.x & 63);
	const int py = blockIdx.y * 4 + (threadIdx.x >> 6);
	if (px >= w || py >= h)
		return;
	const float sq = __fmul_rn(__fadd_rn((float)maxr, 0.5f), __fadd_rn((float)maxr, 0.5f));
	int n = 0;
	float pp = 0.f, pn = 0.f, np = 0.f, nn = 0.f;
	for (int y = 1; y <= maxr; y++)
		for (int x = 1; x <= maxr; x++)
			if ((float)(x * x + y * y) <= sq) {
				const int xl = clampi(px - x, 0, w - 1), xr = clampi(px + x, 0, w - 1);
				const int yu = clampi(py + y, 0, h - 1), yd = clampi(py - y, 0, h - 1);
				np = __fadd_rn(np, __ldg(in + (size_t)yu * w + xl));
				pp = __fadd_rn(pp, __ldg(in + (size_t)yu * w + xr));
				nn = __fadd_rn(nn, __ldg(in + (size_t)yd * w + xl));
				pn = __fadd_rn(pn, __ldg(in + (size_t)yd * w + xr));
				n++;
			}
	const float fn = (float)n;
	pp = __fdiv_rn(pp, fn);
	nn = __fdiv_rn(nn, fn);
	pn = __fdiv_rn(pn, fn);
	np = __fdiv_rn(np, fn);
	out[(size_t)py * w + px] = min_cl(min_cl(pp, nn), min_cl(-pn, -np));
}

/* ------------------------------------------------------------------------------------------------
 * NV12 / RGBA debug-stream conversions.  One thread per 2x2 block: four Y bytes and the block's UV pair,
 * taken from the bottom-right pixel (the last writer of the reference's racing stores in raster order).
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t nv12_y(uint32_t r, uint32_t g, uint32_t b)
{
	return min((66u * r + 129u * g + 25u * b) / 256u + 16u, 255u); /* rgba2nv12.cl:27 */
}
__device__ __forceinline__ uint32_t nv12_uv(uint32_t r, uint32_t g, uint32_t b)
{
	/* rgba2nv12.cl:29-30: int arithmetic, C division truncates toward zero */
	const int ri = (int)r, gi = (int)g, bi = (int)b;
	const int u = clampi((-38 * ri + -74 * gi + 112 * bi) / 256 + 128, 0, 255);
	const int v = clampi((112 * ri + -94 * gi + -18 * bi) / 256 + 128, 0, 255);
	return (uint32_t)u | ((uint32_t)v << 8);
}

/* frame = blockIdx.z: frames `in_stride` pixels / `out_stride` bytes apart (a batch of debug-stream views in one launch) */
__global__ void __launch_bounds__(256) k_rgba2nv12(const uint32_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h, size_t in_stride = 0,
                                                   size_t out_stride = 0)
{
	const int bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
	if (2 * bx >= w)
		return;
	in += blockIdx.z * in_stride;
	out += blockIdx.z * out_stride;
	const uint2 a = __ldg(reinterpret_cast<const uint2*>(in + (size_t)(2 * by) * w + 2 * bx));
	const uint2 b = __ldg(reinterpret_cast<const uint2*>(in + (size_t)(2 * by + 1) * w + 2 * bx));
#define VP_Y(p) nv12_y((p) & 255u, ((p) >> 8) & 255u, ((p) >> 16) & 255u)
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by) * w + 2 * bx) = (uint16_t)(VP_Y(a.x) | (VP_Y(a.y) << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by + 1) * w + 2 * bx) = (uint16_t)(VP_Y(b.x) | (VP_Y(b.y) << 8));
#undef VP_Y
	*reinterpret_cast<uint16_t*>(out + (size_t)w * h + (size_t)by * w + 2 * bx) =
		(uint16_t)nv12_uv(b.y & 255u, (b.y >> 8) & 255u, (b.y >> 16) & 255u);
}

__device__ __forceinline__ uint32_t f2y(float v)
{
	/* f2nv12.cl:24: convert_uchar_sat(v + 127.0f): round toward zero, saturate, NaN -> 0 */
	return min(__float2uint_rz(__fadd_rn(v, 127.0f)), 255u);
}
__global__ void __launch_bounds__(256) k_f2nv12(const float* __restrict__ in, uint8_t* __restrict__ out, int w, int h, size_t in_stride = 0,
                                                size_t out_stride = 0)
{
	const int bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
	if (2 * bx >= w)
		return;
	in += blockIdx.z * in_stride;
	out += blockIdx.z * out_stride;
	const float2 a = __ldg(reinterpret_cast<const float2*>(in + (size_t)(2 * by) * w + 2 * bx));
	const float2 b = __ldg(reinterpret_cast<const float2*>(in + (size_t)(2 * by + 1) * w + 2 * bx));
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by) * w + 2 * bx) = (uint16_t)(f2y(a.x) | (f2y(a.y) << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by + 1) * w + 2 * bx) = (uint16_t)(f2y(b.x) | (f2y(b.y) << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)w * h + (size_t)by * w + 2 * bx) = (uint16_t)(127u | (127u << 8)); /* :25 */
}

/* ---- wide variants (w % 8 == 0, 8-byte aligned views): one thread per 8 x 2 pixels, 16-byte loads, 8-byte stores, a
 * 1-D grid over (frame, block row, block column) so that no CTA is partly empty.  Same integer arithmetic as above. ---- */
__device__ __forceinline__ uint32_t nv12_y_px(uint32_t rgba)
{
	/* 66 r + 129 g + 25 b as one byte dot product (alpha x 0); <= 56100, so the min of nv12_y never binds */
	return (__dp4a(rgba, 0x00198142u, 0u) >> 8) + 16u;
}
__device__ __forceinline__ uint32_t nv12_uv_px(uint32_t rgba)
{
	return nv12_uv(rgba & 255u, (rgba >> 8) & 255u, (rgba >> 16) & 255u);
}

__global__ void __launch_bounds__(256) k_rgba2nv12_wide(const uint32_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h, size_t in_stride,
                                                        size_t out_stride, int n_frames)
{
	const int bw = w >> 3, bh = h >> 1;
	const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (long long)bw * bh * n_frames)
		return;
	const int bx = (int)(idx % bw);
	const long long t = idx / bw;
	const int by = (int)(t % bh), f = (int)(t / bh);
	const uint32_t* src = in + (size_t)f * in_stride + (size_t)(2 * by) * w + 8 * bx;
	uint8_t* dst = out + (size_t)f * out_stride;
	const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(src)), a1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
	const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(src + w)), b1 = __ldg(reinterpret_cast<const uint4*>(src + w) + 1);
	const uint32_t ya0 = nv12_y_px(a0.x) | (nv12_y_px(a0.y) << 8) | (nv12_y_px(a0.z) << 16) | (nv12_y_px(a0.w) << 24);
	const uint32_t ya1 = nv12_y_px(a1.x) | (nv12_y_px(a1.y) << 8) | (nv12_y_px(a1.z) << 16) | (nv12_y_px(a1.w) << 24);
	const uint32_t yb0 = nv12_y_px(b0.x) | (nv12_y_px(b0.y) << 8) | (nv12_y_px(b0.z) << 16) | (nv12_y_px(b0.w) << 24);
	const uint32_t yb1 = nv12_y_px(b1.x) | (nv12_y_px(b1.y) << 8) | (nv12_y_px(b1.z) << 16) | (nv12_y_px(b1.w) << 24);
	*reinterpret_cast<uint2*>(dst + (size_t)(2 * by) * w + 8 * bx) = make_uint2(ya0, ya1);
	*reinterpret_cast<uint2*>(dst + (size_t)(2 * by + 1) * w + 8 * bx) = make_uint2(yb0, yb1);
	/* UV of a 2x2 block from its bottom-right pixel (the last writer of rgba2nv12.cl:28-31 in raster order) */
	*reinterpret_cast<uint2*>(dst + (size_t)w * h + (size_t)by * w + 8 * bx) =
		make_uint2(nv12_uv_px(b0.y) | (nv12_uv_px(b0.w) << 16), nv12_uv_px(b1.y) | (nv12_uv_px(b1.w) << 16));
}

__global__ void __launch_bounds__(256) k_f2nv12_wide(const float* __restrict__ in, uint8_t* __restrict__ out, int w, int h, size_t in_stride,
                                                     size_t out_stride, int n_frames)
{
	const int bw = w >> 3, bh = h >> 1;
	const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (long long)bw * bh * n_frames)
		return;
	const int bx = (int)(idx % bw);
	const long long t = idx / bw;
	const int by = (int)(t % bh), f = (int)(t / bh);
	const float* src = in + (size_t)f * in_stride + (size_t)(2 * by) * w + 8 * bx;
	uint8_t* dst = out + (size_t)f * out_stride;
	const float4 a0 = __ldg(reinterpret_cast<const float4*>(src)), a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
	const float4 b0 = __ldg(reinterpret_cast<const float4*>(src + w)), b1 = __ldg(reinterpret_cast<const float4*>(src + w) + 1);
#define VP_Y4(v) (f2y((v).x) | (f2y((v).y) << 8) | (f2y((v).z) << 16) | (f2y((v).w) << 24))
	*reinterpret_cast<uint2*>(dst + (size_t)(2 * by) * w + 8 * bx) = make_uint2(VP_Y4(a0), VP_Y4(a1));
	*reinterpret_cast<uint2*>(dst + (size_t)(2 * by + 1) * w + 8 * bx) = make_uint2(VP_Y4(b0), VP_Y4(b1));
#undef VP_Y4
	*reinterpret_cast<uint2*>(dst + (size_t)w * h + (size_t)by * w + 8 * bx) = make_uint2(0x7F7F7F7Fu, 0x7F7F7F7Fu); /* f2nv12.cl:25 */
}

/* quad2nv12.cl:23-58 straight from a Bayer frame, default sampling (bilinear, round-to-nearest-even), wq % 8 == 0.
 *
 * The sampler is asked for integer positions +-0.25 (quad2nv12.cl:36-40), so every filter weight is 0.25 or 0.75 against
 * the texel to the left / above: plane c of pixel (x, y) is
 *     [ wl*t(x-1,y-1) + wr*t(x,y-1) ] * wt + [ wl*t(x-1,y) + wr*t(x,y) ] * wb,   (wl,wr), (wt,wb) in {(1,3),(3,1)} / 4
 * -- all products and sums are exact in fp32 (multiples of 1/16 below 256), so the float pipeline of the generic kernel
 * computes S/16 exactly with S an integer <= 4080, and its round-to-nearest-even is (S + 7 + ((S >> 4) & 1)) >> 4.
 * That integer form runs here on 16-bit lanes, two pixels per register: one thread = 8 x 2 output pixels from six 16-byte
 * raw vectors (+ a 2-byte halo each), about 33 instructions per pixel instead of ~150.  Even bytes of a raw row are the
 * planes tapped at x+0.25 (left 1, own 3), odd bytes those tapped at x-0.25 (left 3, own 1); even raw rows the planes
 * tapped at y+0.25, odd rows those at y-0.25 (resampling.cl:65-70 / :74-80). */
struct RawRowLanes {
	uint32_t e[4], o[4]; /* horizontally blended even-byte / odd-byte plane: word j = pixels (x0+2j, x0+2j+1), 16-bit lanes, <= 1020 */
};
__device__ __forceinline__ RawRowLanes raw_row_hblend(const uint8_t* __restrict__ row, int x0)
{
	const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + 2 * x0));
	const uint32_t halo = x0 > 0 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(row + 2 * x0 - 2)) : (v.x & 0xFFFFu); /* quad x0-1, clamped to quad 0 */
	const uint32_t wd[4] = { v.x, v.y, v.z, v.w };
	RawRowLanes r;
	uint32_t pe = (halo & 0xFFu) << 16, po = (halo >> 8) << 16; /* "previous word": the left neighbour sits in its high lane */
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const uint32_t ce = wd[j] & 0x00FF00FFu, co = __byte_perm(wd[j], 0u, 0x4341);
		const uint32_t le = __byte_perm(pe, ce, 0x5432), lo = __byte_perm(po, co, 0x5432); /* (t(x-1), t(x)) of the word's two pixels */
		r.e[j] = 3u * ce + le; /* x + 0.25: left 1/4, own 3/4 */
		r.o[j] = 3u * lo + co; /* x - 0.25: left 3/4, own 1/4 */
		pe = ce;
		po = co;
	}
	return r;
}
/* (S + 7 + ((S >> 4) & 1)) >> 4 on both lanes; the bits the word-wide shift drags across are left for the caller to mask */
__device__ __forceinline__ uint32_t rte16_lanes(uint32_t s) { return (s + 0x00070007u + ((s >> 4) & 0x00010001u)) >> 4; }

/* r, g, b of the two pixels of word j (16-bit lanes, 0..255) from the horizontally blended raw rows of the quad row above
 * (top_e: planes 0|1, top_o: planes 2|3) and of the own quad row */
template <int FMT>
__device__ __forceinline__ void demosaic_lanes(const RawRowLanes& top_e, const RawRowLanes& top_o, const RawRowLanes& own_e, const RawRowLanes& own_o, int j,
                                               uint32_t& r, uint32_t& g, uint32_t& b)
{
	/* y + 0.25 (even raw rows): above 1/4, own 3/4; y - 0.25 (odd raw rows): above 3/4, own 1/4 */
	const uint32_t v0 = rte16_lanes(3u * own_e.e[j] + top_e.e[j]), v1 = rte16_lanes(3u * own_e.o[j] + top_e.o[j]);
	const uint32_t v2 = rte16_lanes(3u * top_o.e[j] + own_o.e[j]), v3 = rte16_lanes(3u * top_o.o[j] + own_o.o[j]);
	if (FMT == FMT_RGGB) { /* g = v1 / 2 + v2 / 2: even lanes add without a carry and the shift brings in a zero */
		r = v0 & 0x00FF00FFu;
		g = ((v1 & 0x00FE00FEu) + (v2 & 0x00FE00FEu)) >> 1;
		b = v3 & 0x00FF00FFu;
	} else {
		r = v1 & 0x00FF00FFu;
		g = ((v0 & 0x00FE00FEu) + (v3 & 0x00FE00FEu)) >> 1;
		b = v2 & 0x00FF00FFu;
	}
}

/* quad2rgba.cl:23-53 straight from a Bayer frame, default sampling, wq % 8 == 0: the same integer demosaic, one thread per
 * 8 x 1 pixels, written as RGBA8 (two 16-byte stores) */
template <int FMT>
__global__ void __launch_bounds__(256) k_raw2rgba_wide(const uint8_t* __restrict__ raw, uint32_t* __restrict__ out, int wq, int hq)
{
	const int bw = wq >> 3;
	const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (long long)bw * hq)
		return;
	const int bx = (int)(idx % bw), y = (int)(idx / bw);
	const int x0 = 8 * bx, rb = 2 * wq, ya = max(y - 1, 0);
	const RawRowLanes top_e = raw_row_hblend(raw + (size_t)(2 * ya) * rb, x0), top_o = raw_row_hblend(raw + (size_t)(2 * ya + 1) * rb, x0);
	const RawRowLanes own_e = raw_row_hblend(raw + (size_t)(2 * y) * rb, x0), own_o = raw_row_hblend(raw + (size_t)(2 * y + 1) * rb, x0);
	uint32_t px[8];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		uint32_t r, g, b;
		demosaic_lanes<FMT>(top_e, top_o, own_e, own_o, j, r, g, b);
		const uint32_t rg = __byte_perm(r, g, 0x6240), ba = b | 0xFF00FF00u; /* (r0, g0, r1, g1), (b0, 255, b1, 255) */
		px[2 * j] = __byte_perm(rg, ba, 0x5410);     /* quad2rgba.cl:52 */
		px[2 * j + 1] = __byte_perm(rg, ba, 0x7632);
	}
	uint4* o = reinterpret_cast<uint4*>(out + (size_t)y * wq + x0);
	o[0] = make_uint4(px[0], px[1], px[2], px[3]);
	o[1] = make_uint4(px[4], px[5], px[6], px[7]);
}

template <int FMT>
__global__ void __launch_bounds__(256) k_raw2nv12_wide(const uint8_t* __restrict__ raw, uint8_t* __restrict__ out, int wq, int hq, size_t src_stride,
                                                       size_t out_stride, int n_frames)
{
	const int bw = wq >> 3, bh = hq >> 1;
	const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
	if (idx >= (long long)bw * bh * n_frames)
		return;
	const int bx = (int)(idx % bw);
	const long long t = idx / bw;
	const int by = (int)(t % bh), f = (int)(t / bh);
	const uint8_t* src = raw + (size_t)f * src_stride;
	uint8_t* dst = out + (size_t)f * out_stride;
	const int x0 = 8 * bx, rb = 2 * wq;
	const int qrow[3] = { max(2 * by - 1, 0), 2 * by, 2 * by + 1 };
	RawRowLanes top_e = raw_row_hblend(src + (size_t)(2 * qrow[0]) * rb, x0); /* planes 0 | 1 of the quad row above */
	RawRowLanes top_o = raw_row_hblend(src + (size_t)(2 * qrow[0] + 1) * rb, x0); /* planes 2 | 3 */
	uint32_t uvw[4];
#pragma unroll
	for (int k = 0; k < 2; k++) {
		const RawRowLanes own_e = raw_row_hblend(src + (size_t)(2 * qrow[k + 1]) * rb, x0);
		const RawRowLanes own_o = raw_row_hblend(src + (size_t)(2 * qrow[k + 1] + 1) * rb, x0);
		uint32_t yw[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			uint32_t r, g, b;
			demosaic_lanes<FMT>(top_e, top_o, own_e, own_o, j, r, g, b);
			yw[j] = (((66u * r + (129u * g + 25u * b)) >> 8) & 0x00FF00FFu) + 0x00100010u; /* <= 56100 per lane: no carry; rgba2nv12.cl:27 */
			if (k == 1) /* UV of the 2x2 block from its bottom-right pixel: the high lane of the lower row */
				uvw[j] = nv12_uv(r >> 16, g >> 16, b >> 16);
		}
		*reinterpret_cast<uint2*>(dst + (size_t)qrow[k + 1] * wq + x0) = make_uint2(__byte_perm(yw[0], yw[1], 0x6420), __byte_perm(yw[2], yw[3], 0x6420));
		top_e = own_e;
		top_o = own_o;
	}
	*reinterpret_cast<uint2*>(dst + (size_t)wq * hq + (size_t)by * wq + x0) = make_uint2(uvw[0] | (uvw[1] << 16), uvw[2] | (uvw[3] << 16));
}

template <int FMT, int MODE, class Src>
__global__ void __launch_bounds__(256) k_quad2nv12(Src s0, uint8_t* __restrict__ out, int wq, int hq, size_t src_stride = 0, size_t out_stride = 0)
{
	const int bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
	if (2 * bx >= wq)
		return;
	const Src s = src_frame(s0, blockIdx.z * src_stride);
	out += blockIdx.z * out_stride;
	uint32_t yv[4], r, g, b;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		demosaic_quad<FMT, MODE>(s, wq, hq, 2 * bx + (k & 1), 2 * by + (k >> 1), r, g, b);
		yv[k] = nv12_y(r, g, b);
	}
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by) * wq + 2 * bx) = (uint16_t)(yv[0] | (yv[1] << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by + 1) * wq + 2 * bx) = (uint16_t)(yv[2] | (yv[3] << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)wq * hq + (size_t)by * wq + 2 * bx) = (uint16_t)nv12_uv(r, g, b); /* k == 3 */
}

template <int FMT, int MODE, class Src>
__global__ void __launch_bounds__(256) k_quad2rgba(Src s, uint32_t* __restrict__ out, int wq, int hq)
{
	const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
	if (x >= wq)
		return;
	uint32_t r, g, b;
	demosaic_quad<FMT, MODE>(s, wq, hq, x, y, r, g, b);
	out[(size_t)y * wq + x] = r | (g << 8) | (b << 16) | 0xFF000000u; /* quad2rgba.cl:52 */
}

/* ------------------------------------------------------------------------------------------------
 * Blob list -> hypothesis hand-off: what src/main.cpp:297-325 does on the CPU per frame -- copy every
 * CLMatch into a `Match` with its position in field millimetres (Perspective::flat2field, Perspective.cpp:127-129) and
 * insert it into a KD-tree for the radius searches of src/blobs -- done per batch on the GPU: the 40-byte records come
 * out ready to use, plus a uniform grid over the visible field (cell list by counting order) in place of the tree:
 *   order[k]      blob indices sorted by (cell, index): cells in row-major order, raster order inside a cell
 *   cell_start[c] first k of cell c (n_cells + 1 entries)
 * One CTA per frame; up to FB_MAX blobs are ordered by a bitonic sort of the unique keys cell * FB_MAX + index.
 * ---------------------------------------------------------------------------------------------- */
constexpr int FB_MAX = 4096;
__global__ void __launch_bounds__(256) k_blobs_to_field(const uint8_t* __restrict__ matches, size_t match_frame_stride, const int32_t* __restrict__ counter,
                                                        int max_blobs, float scale, float off_x, float off_y, float cell_mm, int cells_x, int cells_y,
                                                        vp_field_match* __restrict__ out, int32_t* __restrict__ order, int32_t* __restrict__ cell_start)
{
	__shared__ uint32_t key[FB_MAX];
	const int f = blockIdx.x, tid = threadIdx.x;
	const int n = min(max(counter[3 * f], 0), max_blobs); /* main.cpp:301 */
	const int n_cells = cells_x * cells_y;
	const uint8_t* src = matches + (size_t)f * match_frame_stride;
	vp_field_match* dst = out + (size_t)f * max_blobs;
	int n_sort = 1;
	while (n_sort < n) n_sort <<= 1;
	for (int i = tid; i < n_sort; i += 256) {
		uint32_t k = 0xFFFFFFFFu;
		if (i < n) {
			const uint8_t* m = src + 22 * (size_t)i; /* CLMatch, main.cpp:33-41: floats at unaligned offsets */
			uint32_t w[6];
#pragma unroll
			for (int j = 0; j < 11; j += 2)
				w[j >> 1] = (uint32_t)m[2 * j] | ((uint32_t)m[2 * j + 1] << 8) | ((uint32_t)m[2 * j + 2] << 16) | ((uint32_t)m[2 * j + 3] << 24);
			/* bytes 0-3 x, 4-7 y, 8-10 color, 11-13 center, 14-17 circ, 18-21 score */
			const float x = __uint_as_float(w[0]), y = __uint_as_float(w[1]);
			vp_field_match r;
			r.pos[0] = __fadd_rn(__fmul_rn(x, scale), off_x); /* flat2field: pos * fieldScale + (extent[0], extent[2]) */
			r.pos[1] = __fadd_rn(__fmul_rn(y, scale), off_y);
			r.color[0] = m[8]; r.color[1] = m[9]; r.color[2] = m[10];
			r.center[0] = m[11]; r.center[1] = m[12]; r.center[2] = m[13];
			r.circ = __uint_as_float((uint32_t)m[14] | ((uint32_t)m[15] << 8) | ((uint32_t)m[16] << 16) | ((uint32_t)m[17] << 24));
			r.score = __uint_as_float((uint32_t)m[18] | ((uint32_t)m[19] << 8) | ((uint32_t)m[20] << 16) | ((uint32_t)m[21] << 24));
			dst[i] = r;
			/* cell of the blob inside the visible extent; NaN positions (plateau peaks, blobList.cl:93-94) go to cell 0 */
			const float cxf = floorf(__fdiv_rn(__fsub_rn(r.pos[0], off_x), cell_mm)), cyf = floorf(__fdiv_rn(__fsub_rn(r.pos[1], off_y), cell_mm));
			const int cx = cxf >= 0.f ? (cxf < (float)cells_x ? (int)cxf : cells_x - 1) : 0;
			const int cy = cyf >= 0.f ? (cyf < (float)cells_y ? (int)cyf : cells_y - 1) : 0;
			k = (uint32_t)(cy * cells_x + cx) * (uint32_t)FB_MAX + (uint32_t)i;
		}
		key[i] = k;
	}
	__syncthreads();
	for (int size = 2; size <= n_sort; size <<= 1) {
		for (int stride = size >> 1; stride > 0; stride >>= 1) {
			for (int t = tid; t < (n_sort >> 1); t += 256) {
				const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
				const bool up = (lo & size) == 0;
				const uint32_t a = key[lo], b = key[hi];
				if ((a > b) == up) {
					key[lo] = b;
					key[hi] = a;
				}
			}
			__syncthreads();
		}
	}
	int32_t* ord = order + (size_t)f * max_blobs;
	int32_t* cs = cell_start + (size_t)f * (n_cells + 1);
	for (int k = tid; k <= n; k += 256) {
		const int c_here = k < n ? (int)(key[k] / FB_MAX) : n_cells;   /* cell of sorted position k; the end sentinel closes the table */
		const int c_prev = k > 0 ? (int)(key[k - 1] / FB_MAX) : -1;
		if (k < n)
			ord[k] = (int)(key[k] % FB_MAX);
		for (int c = c_prev + 1; c <= c_here; c++)
			cs[c] = k;
	}
}

} /* namespace vpk */
