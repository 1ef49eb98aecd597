/* gradcirc.cu -- instantiations and launchers of the fused gradient + circularity kernel (gradcirc.cuh). */
#include "gradcirc.cuh"
#include "gradcirc.h"

#include <cstring>
#include <mutex>

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

namespace {
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled g_encode = nullptr;
std::once_flag g_encode_once;

/* the driver's tensor-map encoder through the runtime (libvp_b200.so links cudart only) */
EncodeTiled encode_tiled()
{
	std::call_once(g_encode_once, [] {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			g_encode = (EncodeTiled)fn;
	});
	return g_encode;
}

/* the flat images of one launch as a rank-3 tensor (column, row, frame) of 32-bit pixels, box = one staged slot */
int flat_tensor_map(CUtensorMap* map, const uint32_t* flat, int w, int h, int n_frames, int r)
{
	EncodeTiled enc = encode_tiled();
	if (!enc)
		return (int)cudaErrorNotSupported;
	const cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_frames };
	const cuuint64_t strides[2] = { (cuuint64_t)w * 4, (cuuint64_t)w * h * 4 };
	const cuuint32_t box[3] = { (cuuint32_t)gc_row_bytes(r) / 4, (cuuint32_t)(r + 2), 1 };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)flat, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return rc == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}
} // namespace

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
	alignas(64) CUtensorMap tmap;
	memset(&tmap, 0, sizeof tmap);
	if ((w & 3) == 0) { /* else no strip takes the TMA path (k_grad_circ: `bulk`) and the map is never touched */
		const int rc = flat_tensor_map(&tmap, flat, w, h, n_frames, r);
		if (rc)
			return rc;
	}
#define VP_GC_LAUNCH(RR)                                                                                                       \
	case RR: {                                                                                                                 \
		const int strips = (w + gc_strip_width(RR) - 1) / gc_strip_width(RR);                                                  \
		const dim3 grid(((n_frames > 1 ? 2 : 1) * strips + GC_WARPS - 1) / GC_WARPS, n_seg, (n_frames + 1) / 2); /* a pair of frames per blockIdx.z */ \
		if (o & 1)                                                                                                             \
			k_grad_circ<RR, true><<<grid, GC_WARPS * 32, gc_smem_bytes(RR), stream>>>(tmap, flat, grad, circ, w, h, o, seg_rows, thr, min_score, blob_radius, \
			                                                                         need_score, counter, rowcount, masks, wpr, segsum, segmax, striptot, strips, n_frames); \
		else                                                                                                                   \
			k_grad_circ<RR, false><<<grid, GC_WARPS * 32, gc_smem_bytes(RR), stream>>>(tmap, flat, grad, circ, w, h, o, seg_rows, thr, min_score, blob_radius, \
			                                                                          need_score, counter, rowcount, masks, wpr, segsum, segmax, striptot, strips, n_frames); \
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
