/*
 * gradcirc.cuh -- K2 of the fused path: gradientDot.cl:22-30 + satHorizontal.cl + satVertical.cl + satBlobCenter.cl:22-42 + the
 * peak classification of blobList.cl:38-81 in ONE pass over the flat image.
 *
 * Round 1 ran two kernels here (gradient + row prefix sums, then circularity from the row sums) and paid for the hand-over:
 * the row sums were written and read back (10 MB per 1224x1024 frame) and every pixel cost ~92 instructions.  satBlobCenter
 * only uses the summed-area table through four box sums of gradDot, and a box is a K x K window (K = R-1): this kernel
 * forms the windows directly.
 *
 *   warp  = a strip of 64 image columns (two adjacent columns per lane) walked top to bottom over a segment of rows;
 *   flat  : rows arrive in a per-warp shared-memory ring through the TMA engine (cp.async.bulk, one 288-byte row per lane
 *           and group, completion on an mbarrier), one group of D = R+2 rows ahead of the arithmetic; the four taps of
 *           gradientDot are conflict-free LDS.64;
 *   grad  : four DP4A per pixel, stored once (8-byte stores);
 *   V     : per column the sum of the last K gradient rows (register ring, packed fp32x2 adds);
 *   Q     : horizontal K-window of V by warp shuffles: QA = columns (x+1, x+R], QB = columns (x-R, x-1];
 *   circ  : pp = QA(y+1)  nn = QB(y-R)  pn = -QA(y-R)  np = -QB(y+1), min, exact division by R*R (3 operations: q0 = m*y, e = fma(-q0,d,m), q = fma(e,y,q0) with y = RN(1/d) is the correctly rounded quotient for
 *           all |m| <= 2^24, d = R*R, R <= 24; tests/test_exact_division.py checks every case),
 *           stored once; the last rows stay in registers for the 4-neighbour peak test (one vote per group, almost never taken).
 *
 * Per pixel ~38 instructions, 4 B read + 8 B written.  Every sum is an exact integer in fp32 as long as the reference's own
 * summed-area table stays below 2^22 in magnitude (SAT_EXACT_LIMIT); then any summation order gives the reference's bits.
 * The kernel also emits, per row segment and column, the column sum of gradDot and the largest magnitude it reached on the
 * way; sat_bound_exceeded_g() turns those into a conservative bound of |SAT| over the whole frame.  Frames that leave
 * the bound are redone in the reference's sequential order (k_fallback_frame).
 *
 * Clamped taps (CLAMP_TO_EDGE): with clamped SAT taps a box of satBlobCenter.cl:37-40 is still a window sum in which column 0
 * and row 0 never take part and columns >= w / rows >= h do not exist -- the windows below simply treat those as zero.
 */
#pragma once

#include <cuda.h>

#include "device_util.cuh"

namespace vpk {

constexpr int GC_OH = 4;                /* columns staged to the left and right of a strip's 64 (gradient offset <= 4) */
constexpr int GC_NW = 64 + 2 * GC_OH;   /* words of a staged row that are read */
/* bytes from one staged row to the next = width of the TMA box: 288 (the 72 words) when a slot of D = R+2 such rows is a multiple
 * of 128 bytes (the alignment of a TMA destination), else 384 */
__host__ __device__ constexpr int gc_row_bytes(int R) { return ((R + 2) * GC_NW * 4) % 128 == 0 ? GC_NW * 4 : 384; }
constexpr int GC_WARPS = 4;
constexpr int GC_MAX_R = 12;
constexpr int GC_MAX_OFFSET = GC_OH;

/* columns a strip computes to the left of its first output column: circ(x-1) needs QB = a shuffle of the QA formed
 * R/2 (+1) lanes to the left; a multiple of 4 so that staged rows start on a 16-byte boundary */
__host__ __device__ constexpr int gc_halo_left(int R) { return (((R & 1) ? R + 3 : R + 2) + 3) & ~3; }
/* output columns per strip: circ(x+1) of the last one needs V up to column x+1+R inside the 64 */
__host__ __device__ constexpr int gc_strip_width(int R) { return (62 - R - gc_halo_left(R)) & ~3; }
/* rows of the per-warp ring: three slots of D = R+2 rows and a mirror slot below them (see k_grad_circ) */
__host__ __device__ constexpr int gc_ring_rows(int R) { return 4 * (R + 2); }
__host__ __device__ constexpr size_t gc_smem_bytes(int R) { return (size_t)GC_WARPS * gc_ring_rows(R) * gc_row_bytes(R) + 128; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "WAIT_%=:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra DONE_%=;\n"
	             "bra WAIT_%=;\n"
	             "DONE_%=:\n"
	             "}" ::"r"(smem_u32(bar)), "r"(parity)
	             : "memory");
}
/* a box of D rows of one frame of the flat images -> shared memory in ONE instruction (UTMALDG); coordinates in elements:
 * column, row, frame.  Elements outside the image arrive as zeros (the caller only uses the box where that cannot matter). */
__device__ __forceinline__ void tensor_copy_g2s(void* smem_dst, const void* tmap, int x, int y, int z, uint64_t* bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
	             "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
	             : "memory");
}
/* one row of a flat image -> shared memory through the TMA engine (UBLKCP); completes `bytes` on the mbarrier */
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}

/* true in exactly one lane of a converged warp (ELECT): what follows is issued once, with operands ptxas may keep uniform */
__device__ __forceinline__ bool elect_one()
{
	uint32_t pred;
	asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
	return pred != 0;
}

__device__ __forceinline__ uint2 lds64(const uint32_t* p)
{
	uint2 v;
	asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)));
	return v;
}

/* sum_{j=1..N} P(lane + j), N <= 6: shuffles that do not depend on one another wherever that costs no extra shuffle (the
 * kernel waits on these chains: two dependent steps at most) */
template <int N>
__device__ __forceinline__ float pairs_right(float P)
{
	if constexpr (N == 0)
		return 0.0f;
	const float p1 = __shfl_down_sync(0xffffffffu, P, 1);
	if constexpr (N == 1)
		return p1;
	const float p12 = __fadd_rn(p1, __shfl_down_sync(0xffffffffu, P, 2));
	if constexpr (N == 2)
		return p12;
	if constexpr (N == 3)
		return __fadd_rn(p12, __shfl_down_sync(0xffffffffu, P, 3));
	const float p1234 = __fadd_rn(p12, __shfl_down_sync(0xffffffffu, p12, 2));
	if constexpr (N == 4)
		return p1234;
	if constexpr (N == 5)
		return __fadd_rn(p1234, __shfl_down_sync(0xffffffffu, P, 5));
	return __fadd_rn(p1234, __shfl_down_sync(0xffffffffu, p12, 4));
}

/* The horizontal window of a lane's two columns a = c0, b = c0 + 1 from the per-column sums V (exact integers):
 *   qa = ( sum V[a+2 .. a+R], sum V[b+2 .. b+R] )   = Q(x+1, .) of satBlobCenter.cl's pp / pn
 * The window of nn / np, Q(x-R, .) = sum V[x-R+1 .. x-1], is the qa of the column R+1 to the left (gc_left). */
template <int R>
__device__ __forceinline__ float2 window_right(float2 V)
{
	float2 qa = make_float2(0.0f, 0.0f);
	if constexpr (R > 1) {
		constexpr int M = R / 2;
		const float P = __fadd_rn(V.x, V.y);
		const float a1 = __shfl_down_sync(0xffffffffu, V.x, 1); /* V[a+2] */
		if constexpr (R & 1) { /* R = 2M+1: a+2 .. a+R are M whole lanes */
			qa.x = pairs_right<M>(P);
			qa.y = __fadd_rn(__fsub_rn(qa.x, a1), __shfl_down_sync(0xffffffffu, V.x, M + 1));
		} else { /* R = 2M: M-1 whole lanes and the first column of lane + M */
			qa.x = __fadd_rn(pairs_right<M - 1>(P), M == 1 ? a1 : __shfl_down_sync(0xffffffffu, V.x, M));
			qa.y = __fadd_rn(__fsub_rn(qa.x, a1), __shfl_down_sync(0xffffffffu, V.y, M));
		}
	}
	return qa;
}
/* a per-column value of the columns R+1 to the left of this lane's: (f(a-R-1), f(b-R-1)) */
template <int R>
__device__ __forceinline__ float2 gc_left(float2 v)
{
	if constexpr (R & 1) /* R+1 even: the same column of lane - (R+1)/2 */
		return make_float2(__shfl_up_sync(0xffffffffu, v.x, (R + 1) / 2), __shfl_up_sync(0xffffffffu, v.y, (R + 1) / 2));
	else /* a-R-1 is the second column of lane - (R+2)/2, b-R-1 = a-R the first column of lane - R/2 */
		return make_float2(__shfl_up_sync(0xffffffffu, v.y, (R + 2) / 2), __shfl_up_sync(0xffffffffu, v.x, R / 2));
}

__device__ __forceinline__ uint2 lds64a(uint32_t addr)
{
	uint2 v;
	asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds32a(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
/* p + bytes as ONE instruction (IMAD.WIDE.U32) */
template <class T>
__device__ __forceinline__ T* bump(T* p, unsigned bytes)
{
	return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + bytes);
}

/* Ring of staged flat rows of one warp: three slots of D rows and a mirror slot below them.  Group g (gradient rows t .. t+D-1)
 * owns logical slot g % 3, which holds the D rows it ADDS to what is staged: image rows t+o .. t+o+D-1.  The other rows it reads,
 * t-o .. t+o-1, are the tail of the previous group's slot, which lies directly below in shared memory -- for logical slot 0
 * that is the mirror, a second copy of logical slot 2.  So every tap address is the slot's base plus an immediate, and a group
 * is staged by ONE TMA tensor copy of D rows (two for logical slot 2), one group ahead of the arithmetic.  Groups whose rows
 * leave the image at the top or bottom need the edge row repeated (CLAMP_TO_EDGE), which the tensor copy cannot do: those few
 * are staged row by row (cp.async.bulk with a clamped row); strips at the left/right image edge lane by lane (cp.async). */
template <int R, bool O_ODD>
__global__ void __launch_bounds__(GC_WARPS * 32, R <= 8 ? 4 : 3)
    k_grad_circ(const __grid_constant__ CUtensorMap tmap, const uint32_t* __restrict__ flat, float* __restrict__ grad, float* __restrict__ circ_out, int w, int h,
                int o, int seg_rows, float thr, float min_score, int radius, int need_score, int32_t* __restrict__ counter, int32_t* __restrict__ rowcount,
                uint32_t* __restrict__ masks, int wpr, float* __restrict__ segsum, float* __restrict__ segmax, int32_t* __restrict__ striptot, int n_strips, int n_frames)
{
	constexpr int K = R - 1, D = R + 2;
	constexpr int HL = gc_halo_left(R), SW = gc_strip_width(R);
	constexpr int RB = gc_row_bytes(R), RING = gc_ring_rows(R);
	constexpr float DIV = (float)(R * R);
	constexpr float RCP = 1.0f / DIV;
	extern __shared__ __align__(128) unsigned char gc_smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned char* const ring = gc_smem + (size_t)warp * RING * RB; /* physical slot 0 = the mirror; logical slot j = physical slot j + 1 */
	uint64_t* const bars = reinterpret_cast<uint64_t*>(gc_smem + (size_t)GC_WARPS * RING * RB) + 2 * warp;
	if (lane == 0) {
		mbar_init(bars, 1);
		mbar_init(bars + 1, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	/* blockIdx.x runs over the strips of a PAIR of frames: 26 strips of a 1224-wide image leave two of a frame's 28 warp slots
	 * idle for the lifetime of their CTA, the 52 strips of two frames fill 13 CTAs exactly (the warps of a CTA share nothing) */
	const int task = blockIdx.x * GC_WARPS + warp;
	const int second = task >= n_strips ? 1 : 0;
	const int strip = task - (second ? n_strips : 0);
	const int f = 2 * blockIdx.z + second;
	if (strip >= n_strips || f >= n_frames)
		return;
	const int xs = strip * SW;
	const int xl = xs - HL - GC_OH;           /* image column of staged word 0 */
	const int c0 = xs - HL + 2 * lane;        /* this lane's columns: c0 and c0 + 1 (either both inside the image or both outside: w is even) */
	const int ys = blockIdx.y * seg_rows, ye = min(ys + seg_rows, h);
	const size_t fbase = (size_t)f * w * h;
	const uint32_t* flatf = flat + fbase;
	const bool out_lane = c0 >= xs && c0 < xs + SW && c0 < w;
	/* the 72 staged columns inside the image and rows 16-byte aligned: the TMA engine stages this strip; otherwise (strips at the
	 * left and right image edge, widths that are not a multiple of 4) the lanes fill the ring themselves with clamped columns */
	const bool bulk = xl >= 0 && xl + GC_NW <= w && (w & 3) == 0;
	/* every column of the strip can take part in a window (1 <= c <= w-1): no per-column masks */
	const bool cols_free = xs - HL >= 1 && xs - HL + 63 <= w - 1;
	const bool ea = c0 >= 1 && c0 <= w - 1, eb = c0 + 1 >= 1 && c0 + 1 <= w - 1;
	const uint32_t my = smem_u32(ring) + (uint32_t)(2 * lane + GC_OH) * 4u; /* shared address of column c0 in physical slot 0, row 0 */
	const int ecol[3] = { clampi(xl + lane, 0, w - 1), clampi(xl + lane + 32, 0, w - 1), clampi(xl + lane + 64, 0, w - 1) }; /* edge strips: CLAMP_TO_EDGE in x */
	int nb = 0, ns = 0, npk = 0;
	int32_t* rcf = rowcount + f * h;
	uint32_t* mkf = masks + (size_t)f * h * wpr;
	const unsigned w4 = (unsigned)w * 4u;

	auto publish = [&](int cls, int yy, int x) { /* warp-uniform call */
		if (cls == 3) {
			atomicOr(mkf + (yy * wpr + (x >> 5)), 1u << (x & 31));
			atomicAdd(rcf + yy, 1);
		}
		nb += __popc(__ballot_sync(0xffffffffu, cls == 3));
		ns += __popc(__ballot_sync(0xffffffffu, cls == 2));
		npk += __popc(__ballot_sync(0xffffffffu, cls == 1));
	};

	/* image rows first .. first+D-1 into logical slot `slot` (and into the mirror if that is slot 2); `n` counts the staging
	 * operations of this warp: operation n completes on mbarrier n & 1 */
	auto stage_group = [&](int first, int slot, int n) {
		__syncwarp(); /* every lane has finished reading the slot that is overwritten */
		unsigned char* const dst = ring + (slot + 1) * (D * RB);
		const bool twice = slot == 2;
		if (bulk) {
			uint64_t* bar = bars + (n & 1);
			if (first >= 0 && first + D <= h) { /* the common case: one tensor copy, issued by one lane */
				if (elect_one()) {
					mbar_expect_tx(bar, (uint32_t)(twice ? 2 : 1) * D * RB);
					tensor_copy_g2s(dst, &tmap, xl, first, f, bar);
					if (twice)
						tensor_copy_g2s(ring, &tmap, xl, first, f, bar);
				}
			} else { /* rows above the first / below the last image row repeat it: row by row with the row index clamped */
				if (lane == 0)
					mbar_expect_tx(bar, (uint32_t)(twice ? 2 : 1) * D * GC_NW * 4);
				__syncwarp();
				if (lane < D) {
					const uint32_t* src = flatf + ((size_t)clampi(first + lane, 0, h - 1) * w + xl);
					bulk_copy_g2s(dst + lane * RB, src, GC_NW * 4, bar);
					if (twice)
						bulk_copy_g2s(ring + lane * RB, src, GC_NW * 4, bar);
				}
			}
		} else {
			/* strips at the image edge (or rows that are not 16-byte aligned): 4-byte cp.async with clamped columns, three staged
			 * words per lane and row; asynchronous like the bulk copies, completion through cp.async groups */
#pragma unroll 2
			for (int i = 0; i < D; i++) {
				const uint32_t* src = flatf + (size_t)clampi(first + i, 0, h - 1) * w;
				unsigned char* d = dst + i * RB + lane * 4;
#pragma unroll
				for (int q = 0; q < 3; q++) {
					if (q < 2 || lane < GC_NW - 64) {
						cp_async4(d + q * 128, src + ecol[q]);
						if (twice)
							cp_async4(d - 3 * (D * RB) + q * 128, src + ecol[q]);
					}
				}
			}
			cp_async_commit();
		}
	};
	auto wait_staged = [&](int n) {
		if (bulk) {
			mbar_wait(bars + (n & 1), (uint32_t)(n >> 1) & 1u);
		} else {
			cp_async_wait<0>();
			__syncwarp();
		}
	};

	float2 gq[D], hold[K > 0 ? K : 1], qa[D];
#pragma unroll
	for (int i = 0; i < D; i++)
		gq[i] = qa[i] = make_float2(0.f, 0.f);
	float2 c_prev2 = make_float2(0.f, 0.f), c_prev1 = make_float2(0.f, 0.f); /* the last two circularity rows of the previous group */
	float2 V = make_float2(0.f, 0.f);
	float2 colrun = make_float2(0.f, 0.f), colmax = make_float2(0.f, 0.f);
	int strip_run = 0; /* sum of gradDot over the strip's own columns and the segment's rows so far (warp-uniform, exact) */
	const int t0 = ys - 1 - R;
	const int n_groups = (ye + R - t0 + D) / D; /* whole groups: the extra rows of the last one are computed and never used */
	/* groups g_lo .. g_hi-1 are FAST: gradient rows t .. t+D-1 and output rows t-R .. t-R+D-1 all inside [ys, ye) (hence inside
	 * [1, h-1]), t = t0 + g D: t - R >= ys <=> g D >= 2R + 1, t + D <= ye <=> (g + 1) D <= ye - t0 */
	const int g_lo = cols_free ? (2 * R + 1 + D - 1) / D : 0;
	const int g_hi = cols_free ? (ye - t0) / D : 0;
	/* output pointers of this lane's column pair, advanced row by row: gradDot row tau and circularity row y = tau - R */
	char* const gbase = reinterpret_cast<char*>(grad + fbase + (ptrdiff_t)t0 * w);          /* warp-uniform bases (may lie before the */
	char* const cbase = reinterpret_cast<char*>(circ_out + fbase + (ptrdiff_t)(t0 - R) * w); /* buffer: only owned rows are stored)    */
	unsigned off = (unsigned)c0 * 4u; /* ONE 32-bit byte offset for both stores, advanced by a row per step (<= (seg + 3R + D) rows x w x 4 < 2^32) */

	/* One group = D gradient rows tau = t .. t+D-1, straight-line code.  Step s finishes window row v = tau-K and the
	 * circularity row y = tau-R.  FAST groups (every row owned by the segment both as a gradient row and as an output row,
	 * every column of the strip eligible) carry no range predicates. */
	auto group = [&](auto FAST_C, int g, int slot) {
		constexpr bool FAST = decltype(FAST_C)::value;
		const int t = t0 + g * D;
		/* tap addresses of step 0: the slot starts with image row t+o (the tap below), the centre row is o rows and the tap above
		 * 2o rows further down -- in the previous slot */
		const uint32_t au = my + (uint32_t)((slot + 1) * D) * RB;
		const uint32_t ac = au - (uint32_t)o * RB, ad = au - (uint32_t)(2 * o) * RB;
		const uint32_t al = ac - (uint32_t)o * 4u, ar = ac + (uint32_t)o * 4u;
#pragma unroll
		for (int i = 0; i < K; i++)
			hold[i] = gq[D - K + i];
		constexpr int DD = 4 * D;
		const int y0 = t - R; /* circularity row of step 0 */
		float2 crow[D + 2];   /* circularity rows y0-2 .. y0+D-1 */
		crow[0] = c_prev2;
		crow[1] = c_prev1;
		int strip_row = 0; /* lane s: the strip's running sum after row t+s */
#pragma unroll
		for (int s = 0; s < D; s++) {
			const int tau = t + s;
			/* gradientDot.cl:25-29: taps (x+-o, tau) and (x, tau+-o) of both columns */
			const uint2 U = lds64a(au + s * RB), Dn = lds64a(ad + s * RB);
			uint2 Lf, Rt;
			if constexpr (O_ODD) {
				Lf = make_uint2(lds32a(al + s * RB), lds32a(al + s * RB + 4));
				Rt = make_uint2(lds32a(ar + s * RB), lds32a(ar + s * RB + 4));
			} else {
				Lf = lds64a(al + s * RB);
				Rt = lds64a(ar + s * RB);
			}
			const int g0 = grad_dot_opaque(Rt.x, Lf.x, U.x, Dn.x), g1 = grad_dot_opaque(Rt.y, Lf.y, U.y, Dn.y);
			const float2 gf = add2(make_float2(__int_as_float(g0 + 0x4B400000), __int_as_float(g1 + 0x4B400000)), make_float2(-12582912.0f, -12582912.0f));
			const bool own_row = FAST || (tau >= ys && tau < ye);
			if (out_lane && own_row)
				*reinterpret_cast<float2*>(gbase + off) = gf;
			if (own_row) { /* what the exactness bound of the SAT is made of: the strip's running sum after every row (one warp reduction) ... */
				strip_run += __reduce_add_sync(0xffffffffu, out_lane ? g0 + g1 : 0);
				if (lane == s)
					strip_row = strip_run;
			}
			if (own_row) { /* ... and per column the sum over the segment's rows */
				colrun = add2(colrun, gf);
				colmax.x = fmaxf(colmax.x, fabsf(colrun.x));
				colmax.y = fmaxf(colmax.y, fabsf(colrun.y));
			}
			float2 gw = gf;
			if (!FAST) { /* row 0, column 0 and everything outside the image never take part in a box */
				const bool row_ok = tau >= 1 && tau <= h - 1;
				gw.x = row_ok && ea ? gf.x : 0.0f;
				gw.y = row_ok && eb ? gf.y : 0.0f;
			}
			gq[s] = gw;
			const float2 g_old = s - K >= 0 ? gq[s - K >= 0 ? s - K : 0] : hold[s - K >= 0 ? 0 : s];
			V = add2(V, sub2(gw, g_old)); /* rows tau-K+1 .. tau */
			const int so = (s - K + DD) % D, sq = (s - 2 * R + DD) % D;
			qa[so] = window_right<R>(V); /* window row v = tau-K */
			/* satBlobCenter.cl:37-41 for y = tau-R with A1 = Q(. , y+1) = qa[so] and A2 = Q(. , y-R) = qa[sq]:
			 *   pp = A1(x)   pn = 0 - A2(x)   nn = A2(x-R-1)   np = 0 - A1(x-R-1)        (0 - q keeps the +0 of the reference's last addition)
			 * the two terms of the column R+1 to the left travel as ONE value, min(nn, np), formed where they live */
			const float2 z = make_float2(0.f, 0.f);
			const float2 n1 = sub2(z, qa[so]), n2 = sub2(z, qa[sq]);
			const float2 far = gc_left<R>(make_float2(fminf(qa[sq].x, n1.x), fminf(qa[sq].y, n1.y)));
			const float2 m = make_float2(fminf(fminf(qa[so].x, n2.x), far.x), fminf(fminf(qa[so].y, n2.y), far.y));
			const float2 q0 = mul2(m, make_float2(RCP, RCP));
			unsigned long long e2, c2;
			asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(e2) : "l"(f2_bits(q0)), "l"(f2_bits(make_float2(-DIV, -DIV))), "l"(f2_bits(m)));
			asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c2) : "l"(e2), "l"(f2_bits(make_float2(RCP, RCP))), "l"(f2_bits(q0)));
			const float2 c = bits_f2(c2); /* == m / (R*R), satBlobCenter.cl:41 */
			const int y = y0 + s;
			crow[s + 2] = c;
			if (out_lane && (FAST || (y >= ys && y < ye)))
				*reinterpret_cast<float2*>(cbase + off) = c;
			off += w4;
		}
		c_prev2 = crow[D];
		c_prev1 = crow[D + 1];
		if (lane < D && (FAST || (t + lane >= ys && t + lane < ye)))
			striptot[((size_t)f * n_strips + strip) * h + (t + lane)] = strip_row;
		/* rows classified by this group: yy = y0-1 .. y0+D-2, i.e. crow[1 .. D] */
		float mx = fmaxf(crow[1].x, crow[1].y);
#pragma unroll
		for (int s = 2; s <= D; s++)
			mx = fmaxf(mx, fmaxf(crow[s].x, crow[s].y));
		if (__any_sync(0xffffffffu, out_lane && !(mx < thr))) {
#pragma unroll
			for (int i = 1; i <= D; i++) { /* unrolled: the rows stay where they are, a row without a candidate costs a compare and a vote */
				const int yy = y0 + i - 2;
				const float2 cm = crow[i], up = crow[i - 1], dn = crow[i + 1];
				const bool rows_in = yy >= ys && yy < ye;
				const bool cand_a = out_lane && rows_in && !(cm.x < thr), cand_b = out_lane && rows_in && !(cm.y < thr);
				if (!__any_sync(0xffffffffu, cand_a || cand_b))
					continue;
				const float left_a = __shfl_up_sync(0xffffffffu, cm.y, 1), right_b = __shfl_down_sync(0xffffffffu, cm.x, 1);
				const int cls_a = cand_a ? classify_px(flatf, w, h, c0, yy, radius, thr, min_score, need_score, cm.x, c0 > 0 ? left_a : cm.x, cm.y,
				                                       yy > 0 ? up.x : cm.x, yy < h - 1 ? dn.x : cm.x)
				                         : 0;
				const int cls_b = cand_b ? classify_px(flatf, w, h, c0 + 1, yy, radius, thr, min_score, need_score, cm.y, cm.x, c0 + 1 < w - 1 ? right_b : cm.y,
				                                       yy > 0 ? up.y : cm.y, yy < h - 1 ? dn.y : cm.y)
				                         : 0;
				publish(cls_a, yy, c0);
				publish(cls_b, yy, c0 + 1);
			}
		}
	};

	/* staging operation n brings the rows group n-1 adds; "group -1" are the D rows before t0+o, of which group 0 reads the last 2o */
	stage_group(t0 + o - D, 2, 0);
	stage_group(t0 + o, 0, 1);
	wait_staged(0);
	int slot = 0;
#pragma unroll 1
	for (int g = 0; g < n_groups; g++) {
		const int t = t0 + g * D;
		wait_staged(g + 1);
		const int next = slot == 2 ? 0 : slot + 1;
		if (g + 1 < n_groups)
			stage_group(t + D + o, next, g + 2);
		if (g >= g_lo && g < g_hi)
			group(IntC<1>{}, g, slot);
		else
			group(IntC<0>{}, g, slot);
		slot = next;
	}
	publish_counters(lane, counter + 3 * f, nb, ns, npk);
	if (out_lane) {
		const size_t i = ((size_t)f * gridDim.y + blockIdx.y) * w + c0;
		*reinterpret_cast<float2*>(segsum + i) = colrun;
		*reinterpret_cast<float2*>(segmax + i) = colmax;
	}
}

} // namespace vpk
