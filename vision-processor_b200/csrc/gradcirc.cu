/* gradcirc.cu -- instantiations and launchers of the fused gradient + circularity kernel (gradcirc.cuh). */
#include "gradcirc.cuh"
#include "gradcirc.h"

namespace vpk {

/* batch path: one CTA per (row segment, frame) evaluates the bound and raises the frame's flag */
__global__ void __launch_bounds__(256) k_sat_check_g(GcCheck gc, int w, int h, int* __restrict__ flag)
{
	const int f = blockIdx.y;
	if (sat_bound_exceeded_g(gc, w, h, f, blockIdx.x) && threadIdx.x == 0)
		flag[f] = 2;
}

bool grad_circ_supported(int r, int o) { return r >= 1 && r <= GC_MAX_R && o >= 0 && o <= GC_MAX_OFFSET && 2 * o <= r + 2; }
static_assert(gc_strip_width(GC_MAX_R) >= 32 && gc_strip_width(1) <= 64, "the bound check walks a strip with two columns per lane");

int grad_circ_strip_width(int r) { return gc_strip_width(r); }

#define VP_GC_ALL(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)

int grad_circ_prepare()
{
#define VP_GC_ATTR(RR)                                                                                                         \
	{                                                                                                                          \
		cudaError_t e = cudaFuncSetAttribute(k_grad_circ<RR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes(RR)); \
		if (e == cudaSuccess) e = cudaFuncSetAttribute(k_grad_circ<RR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes(RR)); \
		if (e != cudaSuccess) return (int)e;                                                                                   \
	}
	VP_GC_ALL(VP_GC_ATTR)
#undef VP_GC_ATTR
	return 0;
}

int launch_grad_circ(cudaStream_t stream, int r, const uint32_t* flat, float* grad, float* circ, int w, int h, int o, int seg_rows, int n_frames, float thr,
                     float min_score, int blob_radius, int need_score, int32_t* counter, int32_t* rowcount, uint32_t* masks, int wpr, float* segsum,
                     float* segmax, int32_t* striptot)
{
	const int n_seg = (h + seg_rows - 1) / seg_rows;
#define VP_GC_LAUNCH(RR)                                                                                                       \
	case RR: {                                                                                                                 \
		const int strips = (w + gc_strip_width(RR) - 1) / gc_strip_width(RR);                                                  \
		const dim3 grid((strips + GC_WARPS - 1) / GC_WARPS, n_seg, n_frames);                                                  \
		if (o & 1)                                                                                                             \
			k_grad_circ<RR, true><<<grid, GC_WARPS * 32, gc_smem_bytes(RR), stream>>>(flat, grad, circ, w, h, o, seg_rows, thr, min_score, blob_radius,    \
			                                                                         need_score, counter, rowcount, masks, wpr, segsum, segmax, striptot, strips); \
		else                                                                                                                   \
			k_grad_circ<RR, false><<<grid, GC_WARPS * 32, gc_smem_bytes(RR), stream>>>(flat, grad, circ, w, h, o, seg_rows, thr, min_score, blob_radius,   \
			                                                                          need_score, counter, rowcount, masks, wpr, segsum, segmax, striptot, strips); \
	} break;
	switch (r) {
		VP_GC_ALL(VP_GC_LAUNCH)
	default: return (int)cudaErrorInvalidValue;
	}
#undef VP_GC_LAUNCH
	return (int)cudaGetLastError();
}

int grad_circ_strips(int r, int w) { return (w + gc_strip_width(r) - 1) / gc_strip_width(r); }

GcCheck grad_circ_check(int r, const float* segsum, const float* segmax, const int32_t* striptot, int seg_rows, int w, int h)
{
	GcCheck gc;
	gc.segsum = segsum;
	gc.segmax = segmax;
	gc.striptot = striptot;
	gc.seg_rows = seg_rows;
	gc.n_seg = (h + seg_rows - 1) / seg_rows;
	gc.sw = gc_strip_width(r);
	gc.n_strips = grad_circ_strips(r, w);
	return gc;
}

int launch_sat_check_g(cudaStream_t stream, const GcCheck& gc, int w, int h, int n_frames, int* flag)
{
	k_sat_check_g<<<dim3(gc.n_seg, n_frames), 256, 0, stream>>>(gc, w, h, flag);
	return (int)cudaGetLastError();
}

} // namespace vpk
