/* gradcirc.h -- host-side entry points of gradcirc.cu (its own translation unit: the twelve radius instantiations of the
 * fused gradient + circularity kernel compile next to, not inside, vp_b200.cu). */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_util.cuh"

namespace vpk {

/* can k_grad_circ run this configuration?  (radius 1..12, gradient offset <= 4; anything else takes the row-sum kernels) */
bool grad_circ_supported(int circle_radius, int grad_offset);
/* output columns per warp for this radius: the grid is ceil(ceil(w / strip) / 4) x n_seg x n_frames */
int grad_circ_strip_width(int circle_radius);
/* one-time opt-in to the dynamic shared memory of every instantiation; returns a cudaError_t */
int grad_circ_prepare();
/* strips (warps) across an image of width w */
int grad_circ_strips(int circle_radius, int w);
/* gradientDot + circularity + peak classification of n_frames dense frames; side outputs for the exactness bound of the
 * summed-area table: segsum/segmax n_frames x n_seg x w floats, striptot n_frames x strips x h int32 */
int launch_grad_circ(cudaStream_t stream, int circle_radius, const uint32_t* flat, float* grad, float* circ, int w, int h, int grad_offset, int seg_rows,
                     int n_frames, float thr, float min_score, int blob_radius, int need_score, int32_t* counter, int32_t* rowcount, uint32_t* masks, int wpr,
                     float* segsum, float* segmax, int32_t* striptot);
/* what the bound check reads (device_util.cuh: GcCheck) */
GcCheck grad_circ_check(int circle_radius, const float* segsum, const float* segmax, const int32_t* striptot, int seg_rows, int w, int h);
/* the bound itself, one CTA per row segment and frame: raises flag[f] = 2 for frames that may have left it */
int launch_sat_check_g(cudaStream_t stream, const GcCheck& gc, int w, int h, int n_frames, int* flag);

} // namespace vpk
