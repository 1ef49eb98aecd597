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
 * Canonical arithmetic (shared with oracle/vp_oracle.c): fp32, every operation
 * individually rounded to nearest-even, no FMA contraction (explicit __f*_rn intrinsics and
 * -fmad=false), IEEE division and square root.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vp_b200.h"

namespace vpk {

constexpr int FMT_RGGB = VP_FMT_RGGB8, FMT_GRBG = VP_FMT_GRBG8, FMT_BGR = VP_FMT_BGR8;
constexpr int MODE_RTE = VP_SAMPLE_BILINEAR_RTE, MODE_TRUNC = VP_SAMPLE_BILINEAR_TRUNC, MODE_NEAREST = VP_SAMPLE_NEAREST;
constexpr int SAT_EXACT_LIMIT = 1 << 24; /* |integer| < 2^24 is exact in fp32 */

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

/* float -> texel index: saturating, NaN -> 0 (fmaxf/fminf return the non-NaN operand) */
__device__ __forceinline__ int sat_index(float f, int n)
{
	f = fminf(fmaxf(f, -1.0f), (float)n);
	return clampi(__float2int_rz(f), 0, n - 1);
}

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
 * ---------------------------------------------------------------------------------------------- */
template <int FMT, int MODE, class Src>
__global__ void __launch_bounds__(256) k_reproject(Src src, size_t src_frame_stride, const float2* __restrict__ lut,
                                                    uint32_t* __restrict__ flat, int wq, int hq, int nf)
{
	const int idx = blockIdx.x * 256 + threadIdx.x;
	if (idx >= nf)
		return;
	const Src s = src_frame(src, (size_t)blockIdx.y * src_frame_stride);
	const float2 pos = __ldg(lut + idx);
	uint32_t r, g, b;
	demosaic<FMT, MODE>(s, wq, hq, pos.x, pos.y, r, g, b);
	flat[(size_t)blockIdx.y * nf + idx] = drgb(r, g, b);
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
constexpr int ROWSCAN_WARPS = 8;
__global__ void __launch_bounds__(ROWSCAN_WARPS * 32) k_grad_rowscan(const uint32_t* __restrict__ flat, float* __restrict__ grad,
                                                                     int32_t* __restrict__ rowsum, int wf, int hf, int o,
                                                                     int* __restrict__ flag)
{
	const int lane = threadIdx.x & 31;
	const int y = blockIdx.x * ROWSCAN_WARPS + (threadIdx.x >> 5);
	if (y >= hf)
		return;
	const size_t fbase = (size_t)blockIdx.y * wf * hf;
	const uint32_t* img = flat + fbase;
	const uint32_t* row = img + (size_t)y * wf;
	const uint32_t* up = img + (size_t)min(y + o, hf - 1) * wf;
	const uint32_t* dn = img + (size_t)max(y - o, 0) * wf;
	float* grow = grad + fbase + (size_t)y * wf;
	int32_t* srow = rowsum + fbase + (size_t)y * wf;
	const bool vec = (wf & 3) == 0;
	int carry = 0;
	bool bad = false;
	for (int x0 = lane * 4; x0 - lane * 4 < wf; x0 += 128) {
		int g[4] = { 0, 0, 0, 0 };
		if (x0 < wf) {
			uint32_t U[4], D[4];
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
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int x = x0 + k;
				const uint32_t R = __ldg(row + min(x + o, wf - 1)), L = __ldg(row + clampi(x - o, 0, wf - 1));
				g[k] = x < wf ? grad_dot_px(R, L, U[k], D[k]) : 0;
			}
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
			bad |= (abs(s0) >= SAT_EXACT_LIMIT) | (abs(s1) >= SAT_EXACT_LIMIT) | (abs(s2) >= SAT_EXACT_LIMIT) | (abs(s3) >= SAT_EXACT_LIMIT);
			if (vec) {
				*reinterpret_cast<float4*>(grow + x0) = make_float4((float)g[0], (float)g[1], (float)g[2], (float)g[3]);
				*reinterpret_cast<int4*>(srow + x0) = make_int4(s0, s1, s2, s3);
			} else {
				const int s[4] = { s0, s1, s2, s3 };
#pragma unroll
				for (int k = 0; k < 4; k++)
					if (x0 + k < wf) {
						grow[x0 + k] = (float)g[k];
						srow[x0 + k] = s[k];
					}
			}
		}
	}
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
	const int y0 = warp * rpw;
	int v[RPW];
	int sum = 0;
#pragma unroll
	for (int k = 0; k < RPW; k++) {
		const int y = y0 + k;
		v[k] = (k < rpw && y < hf && x < wf) ? __ldg(rowsum + fbase + (size_t)y * wf + x) : 0;
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
			sat[fbase + (size_t)y * wf + x] = (float)s;
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
	return __fdiv_rn(fminf(fminf(pp, nn), fminf(pn, np)), div);
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
struct DiscStats {
	uint32_t s1[3], s2[3];
	int n;
};

__device__ __forceinline__ DiscStats disc_stats(const uint32_t* __restrict__ img, int w, int h, int x, int y, int radius)
{
	DiscStats d;
	d.n = 0;
	d.s1[0] = d.s1[1] = d.s1[2] = d.s2[0] = d.s2[1] = d.s2[2] = 0;
	const int sq = radius * radius;
	for (int dy = -radius; dy <= radius; dy++) { /* blobList.cl:63-72 */
		const uint32_t* row = img + (size_t)clampi(y + dy, 0, h - 1) * w;
		for (int dx = -radius; dx <= radius; dx++)
			if (dx * dx + dy * dy <= sq) {
				const uint32_t v = __ldg(row + clampi(x + dx, 0, w - 1));
#pragma unroll
				for (int k = 0; k < 3; k++) {
					const uint32_t c = (v >> (8 * k)) & 255u;
					d.s1[k] += c;
					d.s2[k] += c * c;
				}
				d.n++;
			}
	}
	return d;
}

__device__ __forceinline__ float blob_score(const DiscStats& d, float c)
{
	const float fn = (float)d.n;
	float sd[3];
#pragma unroll
	for (int k = 0; k < 3; k++) { /* blobList.cl:76 (native_sqrt -> correctly rounded) */
		const float f1 = (float)d.s1[k];
		sd[k] = __fsqrt_rn(__fdiv_rn(__fsub_rn((float)d.s2[k], __fdiv_rn(__fmul_rn(f1, f1), fn)), fn));
	}
	return __fdiv_rn(c, __fadd_rn(__fadd_rn(sd[0], sd[1]), sd[2])); /* :78 */
}

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

/* scratch layout per frame: [0] initial counter[0] (first output slot), then hf row counts */
/* pass A: classify every pixel, count blobs per row, accumulate counter[0..2].
 * CTA = 8 rows x 32 columns?  No: one warp = 32 consecutive pixels of a row, 8 warps = 256 pixels of one row. */
__global__ void __launch_bounds__(256) k_peaks_count(const uint32_t* __restrict__ img, const float* __restrict__ circ, int w, int h,
                                                     float thr, float min_score, int radius, int need_score,
                                                     int32_t* __restrict__ counter, int32_t* __restrict__ rowcount)
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
	const unsigned m3 = __ballot_sync(0xffffffffu, cls == 3);
	const unsigned m2 = __ballot_sync(0xffffffffu, cls == 2);
	const unsigned m1 = __ballot_sync(0xffffffffu, cls == 1);
	if ((threadIdx.x & 31) == 0) {
		if (m3) {
			atomicAdd(rowcount + (size_t)f * h + y, __popc(m3));
			atomicAdd(counter + 3 * f + 0, __popc(m3)); /* blobList.cl:87 counts past maxMatches too */
		}
		if (m2)
			atomicAdd(counter + 3 * f + 1, __popc(m2)); /* :80 */
		if (m1)
			atomicAdd(counter + 3 * f + 2, __popc(m1)); /* :53 */
	}
}

__device__ __forceinline__ void store_match(uint8_t* __restrict__ dst, float mx, float my, const uint32_t color[3], uint32_t center, float circ,
                                            float score)
{
	/* 22-byte packed record, 2-byte aligned: ten 16-bit stores + two bytes would do; keep it simple and sparse */
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

/* pass B: one warp per row that holds at least one blob: rank = first slot + blobs of earlier rows + blobs to the
 * left in this row; the record is written only if rank < max_matches (blobList.cl:88). */
__global__ void __launch_bounds__(256) k_peaks_emit(const uint32_t* __restrict__ img, const float* __restrict__ circ, int w, int h,
                                                    float thr, float min_score, int radius, int need_score, int max_matches,
                                                    const int32_t* __restrict__ first_slot, const int32_t* __restrict__ rowcount,
                                                    uint8_t* __restrict__ matches, size_t match_frame_stride)
{
	const int lane = threadIdx.x & 31;
	const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
	const int f = blockIdx.y;
	if (y >= h)
		return;
	const int32_t* rc = rowcount + (size_t)f * h;
	if (rc[y] == 0)
		return;
	int before = 0;
	for (int k = lane; k < y; k += 32)
		before += rc[k];
#pragma unroll
	for (int d = 16; d; d >>= 1)
		before += __shfl_xor_sync(0xffffffffu, before, d);
	int rank0 = first_slot[f] + before;
	if (rank0 >= max_matches)
		return;
	const size_t fbase = (size_t)f * w * h;
	const uint32_t* im = img + fbase;
	const float* ci = circ + fbase;
	uint8_t* out = matches + (size_t)f * match_frame_stride;
	for (int x0 = 0; x0 < w && rank0 < max_matches; x0 += 32) {
		const int x = x0 + lane;
		int cls = 0;
		PeakCtx p;
		if (x < w)
			cls = peak_class(im, ci, w, h, x, y, thr, min_score, radius, need_score != 0, p);
		const unsigned m = __ballot_sync(0xffffffffu, cls == 3);
		if (cls == 3) {
			const int rank = rank0 + __popc(m & ((1u << lane) - 1u));
			if (rank < max_matches) {
				const DiscStats d = disc_stats(im, w, h, x, y, radius);
				const float score = blob_score(d, p.c);
				/* blobList.cl:93-94 */
				const float mx = __fadd_rn((float)x, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(p.cnx, p.cpx)),
				                                              __fadd_rn(__fsub_rn(p.cnx, __fmul_rn(2.0f, p.c)), p.cpx)));
				const float my = __fadd_rn((float)y, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(p.cny, p.cpy)),
				                                              __fadd_rn(__fsub_rn(p.cny, __fmul_rn(2.0f, p.c)), p.cpy)));
				const uint32_t color[3] = { d.s1[0] / (uint32_t)d.n, d.s1[1] / (uint32_t)d.n, d.s1[2] / (uint32_t)d.n }; /* :85 */
				store_match(out + 22 * (size_t)rank, mx, my, color, __ldg(im + (size_t)y * w + x), p.c, score);
			}
		}
		rank0 += __popc(m);
	}
}

/* per-batch preparation of the compaction scratch: zero the row counts and the exactness flags; either zero the
 * counters (fused path, main.cpp:283-288) or remember counter[0] as the first output slot (stage API). */
__global__ void k_peaks_prepare(int32_t* __restrict__ counter, int32_t* __restrict__ first_slot, int32_t* __restrict__ rowcount,
                                int n_rows_total, int n_frames, int zero_counters, int* __restrict__ flag)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
	out[(size_t)py * w + px] = fminf(fminf(pp, nn), fminf(-pn, -np));
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

__global__ void __launch_bounds__(256) k_rgba2nv12(const uint32_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h)
{
	const int bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
	if (2 * bx >= w)
		return;
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
__global__ void __launch_bounds__(256) k_f2nv12(const float* __restrict__ in, uint8_t* __restrict__ out, int w, int h)
{
	const int bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
	if (2 * bx >= w)
		return;
	const float2 a = __ldg(reinterpret_cast<const float2*>(in + (size_t)(2 * by) * w + 2 * bx));
	const float2 b = __ldg(reinterpret_cast<const float2*>(in + (size_t)(2 * by + 1) * w + 2 * bx));
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by) * w + 2 * bx) = (uint16_t)(f2y(a.x) | (f2y(a.y) << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)(2 * by + 1) * w + 2 * bx) = (uint16_t)(f2y(b.x) | (f2y(b.y) << 8));
	*reinterpret_cast<uint16_t*>(out + (size_t)w * h + (size_t)by * w + 2 * bx) = (uint16_t)(127u | (127u << 8)); /* :25 */
}

template <int FMT, int MODE, class Src>
__global__ void __launch_bounds__(256) k_quad2nv12(Src s, uint8_t* __restrict__ out, int wq, int hq)
{
	const int bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
	if (2 * bx >= wq)
		return;
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

} /* namespace vpk */
