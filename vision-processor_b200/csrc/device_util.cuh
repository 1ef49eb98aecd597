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
	float* scratch = nullptr;          /* gc_check_scratch_words() 4-byte words per frame */
	int n_seg = 0, seg_rows = 0, sw = 0, n_strips = 0;
};
__host__ __device__ inline size_t gc_check_scratch_words(int n_strips, int n_seg, int w) { return (size_t)2 * n_strips * n_seg + (size_t)n_seg * w; }

/* The exactness bound of the reference's summed-area table for one frame from what k_grad_circ leaves behind:
 *   S(c, k), A(c, k)   per column c and row segment k: the column sum of gradDot over the segment and the largest magnitude
 *                      the running sum reached on the way (segsum, segmax: n_seg x w floats per frame),
 *   P(s, y)            per strip s (the `sw` output columns of one warp) and row y: the sum of gradDot over the strip's columns
 *                      and the rows of y's segment up to y (striptot: n_strips x h int32 per frame; exact).
 * For a pixel (x, y) in strip s and segment k:
 *   SAT(x, y) = sum_{s' < s} [carry(s', k) + P(s', y)]                                      exact: the strips to the left
 *             + sum_{c in s, c <= x} [carry(c, k) + running column sum inside segment k]      carry(., k) = the sums of the segments above
 *   |SAT(x, y)| <= max_{y in k} |left(s, y)| + max_{x in s} |prefix of the column carries| + sum_{c in s} A(c, k).
 * Called by a whole CTA (a multiple of 32 threads); every thread returns the same answer: true when some |SAT| -- and with
 * it possibly a row prefix sum, |RS| <= 2 max |SAT| -- may have reached SAT_EXACT_LIMIT.  About twice the true maximum on
 * camera-like frames.  fp32 throughout: integer values are exact below 2^24, and any term that is not is four times
 * the limit already. */
__device__ __forceinline__ bool sat_bound_exceeded_g(const GcCheck& gc, int w, int h, int f)
{
	const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
	const int n_seg = gc.n_seg, seg_rows = gc.seg_rows, n_strips = gc.n_strips;
	const int32_t* __restrict__ P = gc.striptot + (size_t)f * n_strips * h;
	const float* __restrict__ S = gc.segsum + (size_t)f * n_seg * w;
	const float* __restrict__ A = gc.segmax + (size_t)f * n_seg * w;
	float* __restrict__ strip_carry = gc.scratch + (size_t)f * gc_check_scratch_words(n_strips, n_seg, w); /* [s][k] */
	unsigned* __restrict__ left_max = reinterpret_cast<unsigned*>(strip_carry + (size_t)n_strips * n_seg); /* [s][k], bits of a non-negative float */
	float* __restrict__ col_carry = strip_carry + (size_t)2 * n_strips * n_seg;                             /* [k][x] */
	const int n_pairs = n_strips * n_seg;
	for (int s = tid; s < n_strips; s += nt) { /* what the segments above segment k add to strip s */
		float c = 0.0f;
		for (int k = 0; k < n_seg; k++) {
			strip_carry[s * n_seg + k] = c;
			left_max[s * n_seg + k] = 0u;
			c += (float)P[(size_t)s * h + (min((k + 1) * seg_rows, h) - 1)];
		}
	}
	for (int x = tid; x < w; x += nt) { /* ... and to column x */
		float c = 0.0f;
#pragma unroll 4
		for (int k = 0; k < n_seg; k++) {
			col_carry[(size_t)k * w + x] = c;
			c += S[(size_t)k * w + x];
		}
	}
	__syncthreads();
	/* the summed-area table along the right edge of every strip, row by row: one thread per row walks the strips */
	const bool warp_rows_share_segment = (seg_rows & 31) == 0;
	for (int y0 = 0; y0 < h; y0 += nt) {
		const int y = y0 + tid;
		const int k = min(y, h - 1) / seg_rows;
		float acc = 0.0f;
		for (int s = 0; s + 1 < n_strips; s++) {
			if (y < h)
				acc += strip_carry[s * n_seg + k] + (float)P[(size_t)s * h + y];
			const unsigned mine = y < h ? __float_as_uint(fabsf(acc)) : 0u;
			if (warp_rows_share_segment) { /* y0 and nt are multiples of 32: the 32 rows of a warp lie in one segment */
				const unsigned m = __reduce_max_sync(0xffffffffu, mine);
				if (lane == 0 && m)
					atomicMax(left_max + (s + 1) * n_seg + k, m);
			} else if (mine) {
				atomicMax(left_max + (s + 1) * n_seg + k, mine);
			}
		}
	}
	__syncthreads();
	bool bad = false;
	for (int p = tid; p < n_pairs; p += nt) {
		const int s = p / n_seg, k = p - s * n_seg;
		const int x0 = s * gc.sw, x1 = min(x0 + gc.sw, w);
		float pre = 0.0f, in_pre = 0.0f, in_a = 0.0f;
		for (int x = x0; x < x1; x++) {
			pre += col_carry[(size_t)k * w + x];
			in_pre = fmaxf(in_pre, fabsf(pre));
			in_a += A[(size_t)k * w + x];
		}
		bad |= !(__uint_as_float(left_max[p]) + in_pre + in_a < (float)SAT_EXACT_LIMIT);
	}
	return __syncthreads_or(bad);
}

} // namespace vpk
