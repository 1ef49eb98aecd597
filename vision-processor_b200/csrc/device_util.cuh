/*
 * device_util.cuh -- device helpers shared by the translation units of libvp_b200.so (kernels.cuh, gradcirc.cu): clamps,
 * packed fp32x2 arithmetic, integer -> float without the conversion pipe, the gradient dot product, disc statistics and
 * the peak classification of kernel/blobList.cl.  Canonical arithmetic as stated at the top of kernels.cuh.
 */
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <type_traits>

#include "vp_b200.h"

namespace vpk {

constexpr int FMT_RGGB = VP_FMT_RGGB8, FMT_GRBG = VP_FMT_GRBG8, FMT_BGR = VP_FMT_BGR8;
constexpr int MODE_RTE = VP_SAMPLE_BILINEAR_RTE, MODE_TRUNC = VP_SAMPLE_BILINEAR_TRUNC, MODE_NEAREST = VP_SAMPLE_NEAREST;
/* Exactness bound of the fast path.  While every row sum and every SAT value stays below 2^22 in magnitude, the
 * fp32 running sums of satHorizontal.cl / satVertical.cl are exact integers AND so is every intermediate of the
 * four 4-tap box sums of satBlobCenter.cl:37-40 (|a-b| < 2^23, |a-b-c| < 2^23+2^22, |a-b-c+d| < 2^24), hence any
 * summation order gives the reference's bits.  Frames that leave the bound are redone in the reference's order. */
constexpr int SAT_EXACT_LIMIT = 1 << 22;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

/* OpenCL C 6.12.4 min(x, y): "returns y if y < x, otherwise x" -- differs from fminf for (+0, -0) and NaN */
__device__ __forceinline__ float min_cl(float x, float y) { return y < x ? y : x; }

/* &base[idx] as ONE instruction (IMAD.WIDE.U32): nvcc otherwise expands pointer + 32-bit index into a 4-instruction
 * 64-bit add/shift sequence and rematerialises the base, which matters in kernels that are issue-bound */
template <class T>
__device__ __forceinline__ T* elem_ptr(T* base, unsigned idx)
{
	static_assert(sizeof(T) == 4, "4-byte elements");
	unsigned long long r;
	asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(idx), "l"(base));
	return reinterpret_cast<T*>(r);
}

/* float -> texel index: saturating, NaN -> 0 (fmaxf/fminf return the non-NaN operand) */
__device__ __forceinline__ int sat_index(float f, int n)
{
	f = fminf(fmaxf(f, -1.0f), (float)n);
	return clampi(__float2int_rz(f), 0, n - 1);
}

/* Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2, new on sm_100): two IEEE round-to-nearest operations per issue
 * slot.  The staged kernel is issue-bound, so the two x axes, the two y axes and two samples at a time are evaluated
 * as pairs.  a - b is written fma2(b, -1, a): the product is exact, so the single rounding equals the subtraction's. */
/* Explicit PTX.  CAUTION: ptxas 12.9 contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 regardless of -fmad
 * (also through an fma against -0) -- seen in SASS, and in 0.1 % of the pixels as a 1-LSB difference.  A packed product
 * must therefore never feed a packed add: products are packed, their accumulation is scalar (FMUL2 -> FADD is left alone). */
__device__ __forceinline__ unsigned long long f2_bits(float2 v)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
	return r;
}
__device__ __forceinline__ float2 bits_f2(unsigned long long r)
{
	float2 v;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
	return v;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
	unsigned long long r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
	return bits_f2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
	unsigned long long r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
	return bits_f2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
	unsigned long long r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(b)), "l"(f2_bits(make_float2(-1.0f, -1.0f))), "l"(f2_bits(a)));
	return bits_f2(r);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/* a + b as ONE packed instruction that ptxas cannot contract with the packed product feeding it: fma(a, 1, b) with a
 * 1.0 the compiler cannot see (a kernel argument).  rn(a*1 + b) == rn(a + b) bit for bit.  With a literal 1.0 -- or a
 * plain add.rn.f32x2 -- ptxas 12.9 folds the preceding mul.rn.f32x2 into an FFMA2 and the product loses its rounding. */
__device__ __forceinline__ float2 add2_opaque(float2 a, float2 b, unsigned long long one2)
{
	unsigned long long r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(one2), "l"(f2_bits(b)));
	return bits_f2(r);
}

/* a + b as a 2-operand packed add that ptxas cannot contract either: the .ftz flavour.  A multiply without .ftz and an add with
 * it have no common FFMA2, so the product keeps its own rounding -- and the add stays a FADD2 (two 64-bit sources), which the
 * scheduler pairs with the ALU pipe's PRMTs at ~1 instruction per clock where the three-source FFMA2 of add2_opaque reaches
 * 0.67 (tools/ubench/pipes.cu).  ONLY for operands and sums that are zero or normal (flushing then never happens): the
 * reprojection's products are >= 2^-99 or exactly 0, see the notes at k_reproject_hoist4. */
__device__ __forceinline__ float2 add2_ftz(float2 a, float2 b)
{
	unsigned long long r;
	asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
	return bits_f2(r);
}

template <int V> struct IntC { static constexpr int value = V; };

/* int -> fp32 for |v| < 2^22 on the integer and FMA pipes (I2F runs on the quarter-rate conversion pipe): adding v to the
 * bit pattern of 1.5 * 2^23 moves the float by v units in the last place, i.e. by exactly v */
__device__ __forceinline__ float small_int_to_float(int v) { return __fsub_rn(__int_as_float(v + 0x4B400000), 12582912.0f); }

/* `img` is a flat image produced by the reprojection kernels: alpha is 255 in every pixel, so the alpha terms of the four
 * byte dot products cancel (R.U + L.D - R.D - L.U) and need not be masked off as in grad_dot_px */
__device__ __forceinline__ int grad_dot_opaque(uint32_t R, uint32_t L, uint32_t U, uint32_t D)
{
	const uint32_t pos = __dp4a(L, D, __dp4a(R, U, 0u));
	const uint32_t neg = __dp4a(L, U, __dp4a(R, D, 0u));
	return (int)pos - (int)neg;
}

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

__device__ __forceinline__ void publish_counters(int lane, int32_t* __restrict__ counter_f, int n_blob, int n_score, int n_peak)
{
	if (lane == 0) {
		if (n_blob)
			atomicAdd(counter_f + 0, n_blob); /* blobList.cl:87 counts past maxMatches too */
		if (n_score)
			atomicAdd(counter_f + 1, n_score); /* :80 */
		if (n_peak)
			atomicAdd(counter_f + 2, n_peak); /* :53 */
	}
}

/* blobList.cl:79 for the (non-default) case that the score can reject: kept out of line, it is never hot */
static __device__ __noinline__ int classify_by_score(const uint32_t* __restrict__ img, int w, int h, int x, int y, int radius, float c, float min_score)
{
	return blob_score(disc_stats(img, w, h, x, y, radius), c) < min_score ? 2 : 3;
}

/* peak test of one pixel given its circularity and its four (already clamped) neighbours: blobList.cl:38-81 */
__device__ __forceinline__ int classify_px(const uint32_t* __restrict__ img, int w, int h, int x, int y, int radius, float thr, float min_score,
                                           int need_score, float cm, float lf, float rt, float up, float dn)
{
	if (cm < thr)
		return 0;
	if (lf > cm || rt > cm || up > cm || dn > cm)
		return 1;
	return need_score ? classify_by_score(img, w, h, x, y, radius, cm, min_score) : 3;
}

/* side outputs of k_grad_circ and what sat_bound_exceeded_g needs to read them (by value to the kernels that check the bound) */
struct GcCheck {
	const float* segsum = nullptr;
	const float* segmax = nullptr;
	const int32_t* striptot = nullptr; /* nullptr: not the fused gradient + circularity flow */
	int n_seg = 0, seg_rows = 0, sw = 0, n_strips = 0;
};
constexpr int GC_CHECK_MAX_STRIPS = 512;

/* The exactness bound of the reference's summed-area table for ONE ROW SEGMENT k of one frame from what k_grad_circ leaves
 * behind:
 *   S(c, k), A(c, k)   per column c and row segment k: the column sum of gradDot over the segment and the largest magnitude
 *                      the running sum reached on the way (segsum, segmax: n_seg x w floats per frame),
 *   P(s, y)            per strip s (the `sw` output columns of one warp) and row y: the sum of gradDot over the strip's columns
 *                      and the rows of y's segment up to y (striptot: n_strips x h int32 per frame; exact).
 * For a pixel (x, y) in strip s and segment k:
 *   SAT(x, y) = sum_{s' < s} [carry(s', k) + P(s', y)]                                      exact: the strips to the left
 *             + sum_{c in s, c <= x} [carry(c, k) + running column sum inside segment k]      carry(., k) = the sums of the segments above
 *   |SAT(x, y)| <= max_{y in k} |left(s, y)| + max_{x in s} |prefix of the column carries| + sum_{c in s} A(c, k).
 * Called by a whole CTA (a multiple of 32 threads, one CTA per frame and segment, nothing shared between them); every
 * thread returns the same answer: true when some |SAT| of the segment -- and with it possibly a row prefix sum,
 * |RS| <= 2 max |SAT| -- may have reached SAT_EXACT_LIMIT.  About twice the true maximum on camera-like frames.  fp32
 * throughout: integer values are exact below 2^24, and any term that is not is four times the limit already. */
__device__ __forceinline__ bool sat_bound_exceeded_g(const GcCheck& gc, int w, int h, int f, int k)
{
	__shared__ float strip_carry[GC_CHECK_MAX_STRIPS]; /* what the segments above add to strip s */
	__shared__ unsigned left_max[GC_CHECK_MAX_STRIPS]; /* max over the segment's rows of |SAT| along the LEFT edge of strip s (float bits) */
	__shared__ float in_strip[GC_CHECK_MAX_STRIPS];    /* the two in-strip terms */
	const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, n_warps = nt >> 5;
	const int n_seg = gc.n_seg, seg_rows = gc.seg_rows, n_strips = gc.n_strips;
	const int32_t* __restrict__ P = gc.striptot + (size_t)f * n_strips * h;
	const float* __restrict__ S = gc.segsum + (size_t)f * n_seg * w;
	const float* __restrict__ A = gc.segmax + (size_t)f * n_seg * w;
	const int y_begin = k * seg_rows, y_end = min(y_begin + seg_rows, h);
	for (int s = tid; s < n_strips; s += nt) {
		float c = 0.0f;
#pragma unroll 4
		for (int q = 0; q < k; q++) /* the last row of every segment above holds that segment's strip total */
			c += (float)P[(size_t)s * h + (min((q + 1) * seg_rows, h) - 1)];
		strip_carry[s] = c;
		left_max[s] = 0u;
	}
	/* in-strip terms: one warp per strip, two columns per lane (sw <= 64) */
	for (int s = wid; s < n_strips; s += n_warps) {
		const int x0 = s * gc.sw + 2 * lane;
		float c0 = 0.0f, c1 = 0.0f, a = 0.0f;
		const bool in0 = 2 * lane < gc.sw && x0 < w, in1 = 2 * lane + 1 < gc.sw && x0 + 1 < w;
#pragma unroll 4
		for (int q = 0; q < k; q++) {
			c0 += in0 ? S[(size_t)q * w + x0] : 0.0f;
			c1 += in1 ? S[(size_t)q * w + x0 + 1] : 0.0f;
		}
		a = (in0 ? A[(size_t)k * w + x0] : 0.0f) + (in1 ? A[(size_t)k * w + x0 + 1] : 0.0f);
		float incl = c0 + c1; /* inclusive prefix over the lanes' column pairs */
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const float t = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d)
				incl += t;
		}
		float pre_max = fmaxf(fabsf(incl - c1), fabsf(incl)); /* the prefixes ending at this lane's first and second column */
#pragma unroll
		for (int d = 16; d; d >>= 1) {
			pre_max = fmaxf(pre_max, __shfl_xor_sync(0xffffffffu, pre_max, d));
			a += __shfl_xor_sync(0xffffffffu, a, d);
		}
		if (lane == 0)
			in_strip[s] = pre_max + a;
	}
	__syncthreads();
	/* the summed-area table along the left edge of every strip, row by row: one thread per row walks the strips, eight loads
	 * in flight at a time */
	for (int y0 = y_begin; y0 < y_end; y0 += nt) {
		const int y = y0 + tid;
		const bool in = y < y_end;
		float acc = 0.0f;
		for (int s0 = 0; s0 + 1 < n_strips; s0 += 8) {
			int v[8];
#pragma unroll
			for (int j = 0; j < 8; j++)
				v[j] = in && s0 + j + 1 < n_strips ? __ldg(P + (size_t)(s0 + j) * h + y) : 0;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				if (s0 + j + 1 < n_strips) { /* CTA-uniform */
					acc += strip_carry[s0 + j] + (float)v[j];
					const unsigned m = __reduce_max_sync(0xffffffffu, in ? __float_as_uint(fabsf(acc)) : 0u);
					if (lane == 0 && m)
						atomicMax(left_max + s0 + j + 1, m);
				}
			}
		}
	}
	__syncthreads();
	bool bad = false;
	for (int s = tid; s < n_strips; s += nt)
		bad |= !(__uint_as_float(left_max[s]) + in_strip[s] < (float)SAT_EXACT_LIMIT);
	return __syncthreads_or(bad);
}

} // namespace vpk
