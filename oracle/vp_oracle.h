/*
 * vp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-frame detection kernels of
 * TIGERs-Mannheim/vision-processor (kernel/ *.cl, OpenCL C) with the OpenCL
 * image/sampler semantics written out.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call it.
 * The product (libvp_b200.so) never does.
 *
 * Parity pinning: the reference ships no golden vectors or tests for this
 * path.  The restatement is pinned against the reference's own kernel sources
 * compiled in place through oracle/clemu.h (see oracle/Makefile target
 * _ref/libvp_clref.so, tests/test_oracle_vs_clref.py) and against committed
 * fixtures generated from that build (tests/golden/).
 *
 * Canonical arithmetic (shared with the CUDA kernels): fp32, every operation
 * individually rounded to nearest-even, no FMA contraction, evaluation order as
 * written in the cited reference line.  Compile with -ffp-contract=off.
 */
#ifndef VP_ORACLE_H
#define VP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bayer / raw formats: PixelFormat::{RGGB8,GRBG8,BGR8}  (src/opencl.cpp:29-31) */
enum { VPO_FMT_RGGB = 0, VPO_FMT_GRBG = 1, VPO_FMT_BGR = 2 };

/* What read_imageui + CLK_FILTER_LINEAR means on an integer image (undefined by
 * the OpenCL C spec; resampling.cl:50, quad2nv12.cl:21, quad2rgba.cl:21). */
enum { VPO_SAMPLE_BILINEAR_RTE = 0, VPO_SAMPLE_BILINEAR_TRUNC = 1, VPO_SAMPLE_NEAREST = 2 };

/* resampling.cl:20-27 == Perspective.h:22-29, packed, 72 bytes */
typedef struct __attribute__((packed)) {
	int32_t shape[2];
	float f;
	float p[2];
	float d;
	float r[9];
	float c[3];
} vpo_camera_model;

/* blobList.cl:20-32 == main.cpp:33-41, packed, 22 bytes */
typedef struct __attribute__((packed)) {
	float x, y;
	uint8_t color[3];
	uint8_t center[3];
	float circ;
	float score;
} vpo_match;

/* number of OpenMP threads used by the row-parallel loops (1 = sequential) */
void vpo_set_threads(int n);
int vpo_get_threads(void);

/* raw2quad.cl:21-39.  Bayer: raw is (2*wq) x (2*hq) bytes; BGR: raw is wq x hq x 3,
 * ch[3] is left untouched. */
void vpo_raw2quad(const uint8_t* raw, int fmt, int wq, int hq,
                  uint8_t* ch0, uint8_t* ch1, uint8_t* ch2, uint8_t* ch3);

/* resampling.cl:29-47: field (mm) -> image (quad px).  out[0]=x, out[1]=y */
void vpo_field2image(const vpo_camera_model* m, float fx, float fy, float fz, float* out);

/* resampling.cl:52-99.  planes are wq x hq U8; flat is wf x hf RGBA8 (dRGB). */
void vpo_resampling(const uint8_t* ch0, const uint8_t* ch1, const uint8_t* ch2, const uint8_t* ch3,
                    int fmt, int wq, int hq, uint8_t* flat, int wf, int hf,
                    const vpo_camera_model* m, float max_robot_height, float field_scale,
                    float off_x, float off_y, int sample_mode);

/* gradientDot.cl:22-30 */
void vpo_gradient_dot(const uint8_t* rgba, int w, int h, int offset, float* out);

/* satHorizontal.cl:22-31 / satVertical.cl:22-31 (sequential fp32 running sums) */
void vpo_sat_horizontal(const float* in, int w, int h, float* out);
void vpo_sat_vertical(const float* in, int w, int h, float* out);

/* satBlobCenter.cl:22-42 */
void vpo_circle(const float* sat, int w, int h, int r, float* out);

/* blobList.cl:36-102.  Emits matches in raster order (one of the reference's legal
 * outcomes), keeps the first max_matches; counter[0..2] as the reference. */
void vpo_blob_list(const uint8_t* rgba, const float* circ, int w, int h,
                   vpo_match* matches, int32_t* counter,
                   float circ_threshold, float min_score, int radius, int max_matches);

/* rgba2nv12.cl:22-31, f2nv12.cl:22-26, quad2nv12.cl:23-58, quad2rgba.cl:23-53.
 * NV12 buffers: Y at [0, w*h), interleaved UV at [w*h, w*h*3/2). UV race winner =
 * last writer in sequential raster order. */
void vpo_rgba2nv12(const uint8_t* rgba, int w, int h, uint8_t* out);
void vpo_f2nv12(const float* in, int w, int h, uint8_t* out);
void vpo_quad2nv12(const uint8_t* ch0, const uint8_t* ch1, const uint8_t* ch2, const uint8_t* ch3,
                   int fmt, int wq, int hq, uint8_t* out, int sample_mode);
void vpo_quad2rgba(const uint8_t* ch0, const uint8_t* ch1, const uint8_t* ch2, const uint8_t* ch3,
                   int fmt, int wq, int hq, uint8_t* rgba, int sample_mode);

/* dead kernels named by north_star: blobCenter.cl:29-63, blobScore.cl:23-66 */
void vpo_circularize(const float* in, int w, int h, int min_blob_radius, int max_blob_radius, float* out);
void vpo_blob_score(const uint8_t* rgba, const float* circ, int w, int h,
                    float circ_threshold, int radius, float* out);

/* Whole frame, reference stage order with a full image round trip per stage
 * (Resources.cpp:138-164 + main.cpp:283-289).  Any output pointer may be NULL
 * (scratch is allocated internally).  Returns max|SAT| seen (the 2^24 exactness
 * premise of the parallel scan is checked by the caller). */
typedef struct {
	int fmt, wq, hq, wf, hf;
	vpo_camera_model model;
	float max_robot_height, field_scale, off_x, off_y;
	int grad_offset;      /* (int)ceilf(maxBlobRadius/fieldScale)/3   Resources.cpp:160 */
	int circle_radius;    /* (int)ceilf(minBlobRadius/fieldScale)     Resources.cpp:163 */
	float circ_threshold; /* thresholds.circularity                   main.cpp:289 */
	float min_score;      /* literal 0.0f at the call site            main.cpp:289 */
	int blob_radius;      /* (int)floorf(minBlobRadius/fieldScale)    main.cpp:289 */
	int max_blobs;        /* thresholds.blobs                         main.cpp:289 */
	int sample_mode;
} vpo_params;

float vpo_detect(const uint8_t* raw, const vpo_params* p,
                 uint8_t* flat, float* grad_dot, float* sat, float* circ,
                 vpo_match* matches, int32_t* counter, int with_blob_list);

#ifdef __cplusplus
}
#endif
#endif
