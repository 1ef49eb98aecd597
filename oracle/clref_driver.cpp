/*
 * clref_driver.cpp -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * NDRange driver for the reference's OpenCL kernels compiled in place through clemu.h
 * (oracle/Makefile -> oracle/_ref/libvp_clref.so).  Exposes the same C signatures as
 * vp_oracle.h with the prefix clref_, and launches each kernel with the NDRange and
 * argument order of its reference call site (Resources.cpp:138-186, main.cpp:283-289).
 * Work-items of one row run in x order; rows run in y order on one thread, or split
 * over OpenMP threads when clref_set_threads(n>1) (used only for the CPU baseline).
 */
#define CLEMU_NO_KEYWORDS
#include "clemu.h"
#include "vp_oracle.h"

#include <cstdlib>
#include <cstring>

thread_local clemu_workitem clemu_wi __attribute__((tls_model("initial-exec")));
int clemu_linear_mode = CLEMU_BILINEAR_RTE;

static int g_threads = 1;

/* kernels: one extern "C" symbol per (file, Bayer macro) -- names set by -D in the Makefile */
#define DECL3(name, ...) \
	extern "C" void name##_RGGB(__VA_ARGS__); \
	extern "C" void name##_GRBG(__VA_ARGS__); \
	extern "C" void name##_BGR(__VA_ARGS__);

DECL3(clk_raw2quad, const uchar*, image2d_t, image2d_t, image2d_t, image2d_t)
DECL3(clk_resampling, image2d_t, image2d_t, image2d_t, image2d_t, image2d_t, const vpo_camera_model, const float, const float, const float, const float)
DECL3(clk_quad2nv12, image2d_t, image2d_t, image2d_t, image2d_t, uchar*)
DECL3(clk_quad2rgba, image2d_t, image2d_t, image2d_t, image2d_t, image2d_t)
extern "C" void clk_gradientDot(image2d_t, image2d_t, int);
extern "C" void clk_satHorizontal(image2d_t, image2d_t);
extern "C" void clk_satVertical(image2d_t, image2d_t);
extern "C" void clk_satBlobCenter(image2d_t, image2d_t, int);
extern "C" void clk_blobList(image2d_t, image2d_t, vpo_match*, volatile int*, const float, const float, const int, const int);
extern "C" void clk_rgba2nv12(image2d_t, uchar*);
extern "C" void clk_f2nv12(image2d_t, uchar*);
extern "C" void clk_blobCenter(image2d_t, image2d_t, int, int);
extern "C" void clk_blobScore(image2d_t, image2d_t, image2d_t, const float, const int);

template <typename F>
static void ndrange2(int gx, int gy, F f)
{
#pragma omp parallel for schedule(static) num_threads(g_threads) if(g_threads > 1)
	for (int y = 0; y < gy; y++) {
		clemu_wi.gsize[0] = (size_t)gx; clemu_wi.gsize[1] = (size_t)gy; clemu_wi.gsize[2] = 1;
		clemu_wi.gid[1] = (size_t)y; clemu_wi.gid[2] = 0;
		for (int x = 0; x < gx; x++) {
			clemu_wi.gid[0] = (size_t)x;
			f();
		}
	}
}

template <typename F>
static void ndrange1(int g, F f)
{
#pragma omp parallel for schedule(static) num_threads(g_threads) if(g_threads > 1)
	for (int i = 0; i < g; i++) {
		clemu_wi.gsize[0] = (size_t)g; clemu_wi.gsize[1] = 1; clemu_wi.gsize[2] = 1;
		clemu_wi.gid[0] = (size_t)i; clemu_wi.gid[1] = 0; clemu_wi.gid[2] = 0;
		f();
	}
}

static clemu_image img(const void* d, int w, int h, int type) { return clemu_image{ const_cast<void*>(d), w, h, type }; }

extern "C" {

void clref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
void clref_set_linear_mode(int m) { clemu_linear_mode = m; }

/* Resources.cpp:138-143 */
void clref_raw2quad(const uint8_t* raw, int fmt, int wq, int hq, uint8_t* c0, uint8_t* c1, uint8_t* c2, uint8_t* c3)
{
	clemu_image i0 = img(c0, wq, hq, CLEMU_U8), i1 = img(c1, wq, hq, CLEMU_U8), i2 = img(c2, wq, hq, CLEMU_U8), i3 = img(c3, wq, hq, CLEMU_U8);
	ndrange2(wq, hq, [&] {
		if (fmt == VPO_FMT_RGGB) clk_raw2quad_RGGB(raw, &i0, &i1, &i2, &i3);
		else if (fmt == VPO_FMT_GRBG) clk_raw2quad_GRBG(raw, &i0, &i1, &i2, &i3);
		else clk_raw2quad_BGR(raw, &i0, &i1, &i2, &i3);
	});
}

/* Resources.cpp:159 */
void clref_resampling(const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, const uint8_t* c3, int fmt, int wq, int hq,
                      uint8_t* flat, int wf, int hf, const vpo_camera_model* m, float h, float scale, float offx, float offy, int mode)
{
	clemu_linear_mode = mode;
	clemu_image i0 = img(c0, wq, hq, CLEMU_U8), i1 = img(c1, wq, hq, CLEMU_U8), i2 = img(c2, wq, hq, CLEMU_U8), i3 = img(c3, wq, hq, CLEMU_U8);
	clemu_image o = img(flat, wf, hf, CLEMU_RGBA8);
	ndrange2(wf, hf, [&] {
		if (fmt == VPO_FMT_RGGB) clk_resampling_RGGB(&i0, &i1, &i2, &i3, &o, *m, h, scale, offx, offy);
		else if (fmt == VPO_FMT_GRBG) clk_resampling_GRBG(&i0, &i1, &i2, &i3, &o, *m, h, scale, offx, offy);
		else clk_resampling_BGR(&i0, &i1, &i2, &i3, &o, *m, h, scale, offx, offy);
	});
}

/* Resources.cpp:160 */
void clref_gradient_dot(const uint8_t* rgba, int w, int h, int offset, float* out)
{
	clemu_image i = img(rgba, w, h, CLEMU_RGBA8), o = img(out, w, h, CLEMU_F32);
	ndrange2(w, h, [&] { clk_gradientDot(&i, &o, offset); });
}

/* Resources.cpp:161: NDRange(height) */
void clref_sat_horizontal(const float* in, int w, int h, float* out)
{
	clemu_image i = img(in, w, h, CLEMU_F32), o = img(out, w, h, CLEMU_F32);
	ndrange1(h, [&] { clk_satHorizontal(&i, &o); });
}

/* Resources.cpp:162: NDRange(width) */
void clref_sat_vertical(const float* in, int w, int h, float* out)
{
	clemu_image i = img(in, w, h, CLEMU_F32), o = img(out, w, h, CLEMU_F32);
	ndrange1(w, [&] { clk_satVertical(&i, &o); });
}

/* Resources.cpp:163 */
void clref_circle(const float* sat, int w, int h, int r, float* out)
{
	clemu_image i = img(sat, w, h, CLEMU_F32), o = img(out, w, h, CLEMU_F32);
	ndrange2(w, h, [&] { clk_satBlobCenter(&i, &o, r); });
}

/* main.cpp:289 */
void clref_blob_list(const uint8_t* rgba, const float* circ, int w, int h, vpo_match* matches, int32_t* counter,
                     float thr, float min_score, int radius, int max_matches)
{
	clemu_image i = img(rgba, w, h, CLEMU_RGBA8), c = img(circ, w, h, CLEMU_F32);
	ndrange2(w, h, [&] { clk_blobList(&i, &c, matches, counter, thr, min_score, radius, max_matches); });
}

/* Resources.cpp:172-186 */
void clref_rgba2nv12(const uint8_t* rgba, int w, int h, uint8_t* out)
{
	clemu_image i = img(rgba, w, h, CLEMU_RGBA8);
	ndrange2(w, h, [&] { clk_rgba2nv12(&i, out); });
}

void clref_f2nv12(const float* in, int w, int h, uint8_t* out)
{
	clemu_image i = img(in, w, h, CLEMU_F32);
	ndrange2(w, h, [&] { clk_f2nv12(&i, out); });
}

/* Resources.cpp:166-170 */
void clref_quad2nv12(const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, const uint8_t* c3, int fmt, int wq, int hq, uint8_t* out, int mode)
{
	clemu_linear_mode = mode;
	clemu_image i0 = img(c0, wq, hq, CLEMU_U8), i1 = img(c1, wq, hq, CLEMU_U8), i2 = img(c2, wq, hq, CLEMU_U8), i3 = img(c3, wq, hq, CLEMU_U8);
	ndrange2(wq, hq, [&] {
		if (fmt == VPO_FMT_RGGB) clk_quad2nv12_RGGB(&i0, &i1, &i2, &i3, out);
		else if (fmt == VPO_FMT_GRBG) clk_quad2nv12_GRBG(&i0, &i1, &i2, &i3, out);
		else clk_quad2nv12_BGR(&i0, &i1, &i2, &i3, out);
	});
}

/* Resources.cpp:145-149 */
void clref_quad2rgba(const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, const uint8_t* c3, int fmt, int wq, int hq, uint8_t* rgba, int mode)
{
	clemu_linear_mode = mode;
	clemu_image i0 = img(c0, wq, hq, CLEMU_U8), i1 = img(c1, wq, hq, CLEMU_U8), i2 = img(c2, wq, hq, CLEMU_U8), i3 = img(c3, wq, hq, CLEMU_U8);
	clemu_image o = img(rgba, wq, hq, CLEMU_RGBA8);
	ndrange2(wq, hq, [&] {
		if (fmt == VPO_FMT_RGGB) clk_quad2rgba_RGGB(&i0, &i1, &i2, &i3, &o);
		else if (fmt == VPO_FMT_GRBG) clk_quad2rgba_GRBG(&i0, &i1, &i2, &i3, &o);
		else clk_quad2rgba_BGR(&i0, &i1, &i2, &i3, &o);
	});
}

/* dead kernels */
void clref_circularize(const float* in, int w, int h, int minr, int maxr, float* out)
{
	clemu_image i = img(in, w, h, CLEMU_F32), o = img(out, w, h, CLEMU_F32);
	ndrange2(w, h, [&] { clk_blobCenter(&i, &o, minr, maxr); });
}

void clref_blob_score(const uint8_t* rgba, const float* circ, int w, int h, float thr, int radius, float* out)
{
	clemu_image i = img(rgba, w, h, CLEMU_RGBA8), c = img(circ, w, h, CLEMU_F32), o = img(out, w, h, CLEMU_F32);
	ndrange2(w, h, [&] { clk_blobScore(&i, &c, &o, thr, radius); });
}

/* whole frame in the reference's stage order: Resources.cpp:138-164 + main.cpp:283-289 */
float clref_detect(const uint8_t* raw, const vpo_params* p, uint8_t* flat, float* grad_dot, float* sat, float* circ,
                   vpo_match* matches, int32_t* counter, int with_blob_list)
{
	const size_t nq = (size_t)p->wq * p->hq, nf = (size_t)p->wf * p->hf;
	uint8_t* planes = (uint8_t*)calloc(4 * nq, 1);
	uint8_t* flat_ = flat ? flat : (uint8_t*)malloc(4 * nf);
	float* grad_ = grad_dot ? grad_dot : (float*)malloc(4 * nf);
	float* hor = (float*)malloc(4 * nf);
	float* sat_ = sat ? sat : (float*)malloc(4 * nf);
	float* circ_ = circ ? circ : (float*)malloc(4 * nf);
	clref_raw2quad(raw, p->fmt, p->wq, p->hq, planes, planes + nq, planes + 2 * nq, planes + 3 * nq);
	clref_resampling(planes, planes + nq, planes + 2 * nq, planes + 3 * nq, p->fmt, p->wq, p->hq, flat_, p->wf, p->hf,
	                 &p->model, p->max_robot_height, p->field_scale, p->off_x, p->off_y, p->sample_mode);
	clref_gradient_dot(flat_, p->wf, p->hf, p->grad_offset, grad_);
	clref_sat_horizontal(grad_, p->wf, p->hf, hor);
	clref_sat_vertical(hor, p->wf, p->hf, sat_);
	clref_circle(sat_, p->wf, p->hf, p->circle_radius, circ_);
	if (with_blob_list && matches && counter) {
		counter[0] = counter[1] = counter[2] = 0;
		clref_blob_list(flat_, circ_, p->wf, p->hf, matches, counter, p->circ_threshold, p->min_score, p->blob_radius, p->max_blobs);
	}
	float mx = 0.f;
	for (size_t i = 0; i < nf; i++) {
		const float a = fabsf(sat_[i]), b = fabsf(hor[i]);
		if (a > mx) mx = a;
		if (b > mx) mx = b;
	}
	free(planes); free(hor);
	if (!flat) free(flat_);
	if (!grad_dot) free(grad_);
	if (!sat) free(sat_);
	if (!circ) free(circ_);
	return mx;
}

} /* extern "C" */
