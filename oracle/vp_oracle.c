/*
 * vp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See vp_oracle.h for scope and pinning.  All citations are path:line under the
 * reference tree (TIGERs-Mannheim/vision-processor).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile)
 */
#include "vp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

void vpo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int vpo_get_threads(void) { return g_threads; }

#define VPO_PAR_ROWS _Pragma("omp parallel for schedule(static) num_threads(g_threads) if(g_threads > 1)")

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* OpenCL C 6.12.4 min(x, y): "returns y if y < x, otherwise x" (not fminf: differs for (+0, -0) and NaN) */
static inline float min_cl(float x, float y) { return y < x ? y : x; }

/* ------------------------------------------------------------------ raw2quad */

/* kernel/raw2quad.cl:21-39; NDRange (img.width, img.height) Resources.cpp:142 */
void vpo_raw2quad(const uint8_t* raw, int fmt, int wq, int hq,
                  uint8_t* ch0, uint8_t* ch1, uint8_t* ch2, uint8_t* ch3)
{
	if (fmt == VPO_FMT_BGR) {
		/* raw2quad.cl:23-29: imgpos = 3*(x + y*global_size(0)); channel3 untouched */
		VPO_PAR_ROWS
		for (int y = 0; y < hq; y++)
			for (int x = 0; x < wq; x++) {
				const size_t p = 3 * ((size_t)x + (size_t)y * wq);
				ch0[x + (size_t)y * wq] = raw[p];
				ch1[x + (size_t)y * wq] = raw[p + 1];
				ch2[x + (size_t)y * wq] = raw[p + 2];
			}
		return;
	}
	/* raw2quad.cl:31-37: row_size = 2*global_size(0); imgpos = 2x + 2y*row_size */
	const size_t row = 2 * (size_t)wq;
	VPO_PAR_ROWS
	for (int y = 0; y < hq; y++)
		for (int x = 0; x < wq; x++) {
			const size_t p = 2 * (size_t)x + 2 * (size_t)y * row;
			const size_t o = x + (size_t)y * wq;
			ch0[o] = raw[p];
			ch1[o] = raw[p + 1];
			ch2[o] = raw[p + row];
			ch3[o] = raw[p + 1 + row];
		}
}

/* ---------------------------------------------------------------- resampling */

/* kernel/resampling.cl:29-47, canonical arithmetic of SURVEY section 10 */
void vpo_field2image(const vpo_camera_model* m, float fx, float fy, float fz, float* out)
{
	const float vx = fx - m->c[0];
	const float vy = fy - m->c[1];
	const float vz = fz - m->c[2];
	const float rx = (m->r[0] * vx + m->r[1] * vy) + m->r[2] * vz;
	const float ry = (m->r[3] * vx + m->r[4] * vy) + m->r[5] * vz;
	const float rz = (m->r[6] * vx + m->r[7] * vy) + m->r[8] * vz;
	const float nx = rx / rz;
	const float ny = ry / rz;
	float ux = nx, uy = ny;
	for (int i = 0; i < 8; i++) { /* resampling.cl:40 (8 iterations on the GPU) */
		const float q = ux * ux + uy * uy;
		const float dr = 1.0f + m->d * q;
		ux = nx / dr;
		uy = ny / dr;
	}
	out[0] = m->f * ux + m->p[0];
	out[1] = m->f * uy + m->p[1];
}

/* float -> texel index, saturating, NaN -> lo.  fmaxf/fminf return the non-NaN operand. */
static inline int sat_index(float f, int n)
{
	f = fminf(fmaxf(f, -1.0f), (float)n);
	return clampi((int)f, 0, n - 1);
}

/* read_imageui(plane, LINEAR|UNNORMALIZED|CLAMP_TO_EDGE, (u,v)).x on a w x h U8 plane.
 * OpenCL 1.2 spec 8.2: i0=floor(u-0.5), a=frac(u-0.5), clamp after computing weights. */
static inline uint32_t sample_u8(const uint8_t* T, int w, int h, float u, float v, int mode)
{
	if (mode == VPO_SAMPLE_NEAREST) {
		const int i = sat_index(floorf(u), w);
		const int j = sat_index(floorf(v), h);
		return T[i + (size_t)j * w];
	}
	const float fu = u - 0.5f;
	const float fv = v - 0.5f;
	const float fi = floorf(fu);
	const float fj = floorf(fv);
	const float a = fu - fi;
	const float b = fv - fj;
	const int i0 = sat_index(fi, w);
	const int i1 = sat_index(fi + 1.0f, w);
	const int j0 = sat_index(fj, h);
	const int j1 = sat_index(fj + 1.0f, h);
	const float oma = 1.0f - a;
	const float omb = 1.0f - b;
	const float t00 = (float)T[i0 + (size_t)j0 * w];
	const float t10 = (float)T[i1 + (size_t)j0 * w];
	const float t01 = (float)T[i0 + (size_t)j1 * w];
	const float t11 = (float)T[i1 + (size_t)j1 * w];
	const float val = (((oma * omb) * t00 + (a * omb) * t10) + (oma * b) * t01) + (a * b) * t11;
	if (!(val >= 0.0f))
		return 0; /* negative or NaN */
	if (val >= 255.0f)
		return 255;
	return mode == VPO_SAMPLE_BILINEAR_TRUNC ? (uint32_t)val : (uint32_t)rintf(val);
}

/* the demosaic taps shared by resampling.cl:56-81, quad2nv12.cl:27-51, quad2rgba.cl:27-51.
 * int_coords: the BGR branch of quad2nv12/quad2rgba passes an int2 to a LINEAR sampler
 * (undefined); defined here as the direct texel. */
static inline void demosaic(const uint8_t* const ch[4], int fmt, int wq, int hq,
                            float px, float py, int mode, int bgr_direct, uint32_t rgb[3])
{
	if (fmt == VPO_FMT_BGR) {
		if (bgr_direct) {
			const int i = clampi((int)px, 0, wq - 1), j = clampi((int)py, 0, hq - 1);
			rgb[0] = ch[2][i + (size_t)j * wq];
			rgb[1] = ch[1][i + (size_t)j * wq];
			rgb[2] = ch[0][i + (size_t)j * wq];
		} else {
			rgb[0] = sample_u8(ch[2], wq, hq, px, py, mode);
			rgb[1] = sample_u8(ch[1], wq, hq, px, py, mode);
			rgb[2] = sample_u8(ch[0], wq, hq, px, py, mode);
		}
	} else if (fmt == VPO_FMT_RGGB) {
		rgb[0] = sample_u8(ch[0], wq, hq, px + 0.25f, py + 0.25f, mode);
		rgb[1] = sample_u8(ch[1], wq, hq, px - 0.25f, py + 0.25f, mode) / 2
		       + sample_u8(ch[2], wq, hq, px + 0.25f, py - 0.25f, mode) / 2;
		rgb[2] = sample_u8(ch[3], wq, hq, px - 0.25f, py - 0.25f, mode);
	} else { /* GRBG */
		rgb[0] = sample_u8(ch[1], wq, hq, px - 0.25f, py + 0.25f, mode);
		rgb[1] = sample_u8(ch[0], wq, hq, px + 0.25f, py + 0.25f, mode) / 2
		       + sample_u8(ch[3], wq, hq, px - 0.25f, py - 0.25f, mode) / 2;
		rgb[2] = sample_u8(ch[2], wq, hq, px + 0.25f, py - 0.25f, mode);
	}
}

/* kernel/resampling.cl:52-99; launch Resources.cpp:159 */
void vpo_resampling(const uint8_t* ch0, const uint8_t* ch1, const uint8_t* ch2, const uint8_t* ch3,
                    int fmt, int wq, int hq, uint8_t* flat, int wf, int hf,
                    const vpo_camera_model* m, float max_robot_height, float field_scale,
                    float off_x, float off_y, int sample_mode)
{
	const uint8_t* const ch[4] = { ch0, ch1, ch2, ch3 };
	VPO_PAR_ROWS
	for (int gy = 0; gy < hf; gy++)
		for (int gx = 0; gx < wf; gx++) {
			float pos[2];
			/* resampling.cl:53 */
			vpo_field2image(m, (float)gx * field_scale + off_x, (float)gy * field_scale + off_y,
			                max_robot_height, pos);
			uint32_t c[3];
			demosaic(ch, fmt, wq, hq, pos[0], pos[1], sample_mode, 0, c);
			/* resampling.cl:86-91, uint32 arithmetic (wraps, +510 restores) */
			uint8_t* o = flat + 4 * ((size_t)gx + (size_t)gy * wf);
			o[0] = (uint8_t)((2u * c[0] - c[1] - c[2] + 510u) / 4u);
			o[1] = (uint8_t)((2u * c[1] - c[2] - c[0] + 510u) / 4u);
			o[2] = (uint8_t)((2u * c[2] - c[0] - c[1] + 510u) / 4u);
			o[3] = 255;
		}
}

/* --------------------------------------------------------------- gradientDot */

/* kernel/gradientDot.cl:22-30; NEAREST + CLAMP_TO_EDGE, int coords */
void vpo_gradient_dot(const uint8_t* rgba, int w, int h, int offset, float* out)
{
	VPO_PAR_ROWS
	for (int y = 0; y < h; y++) {
		const int yp = clampi(y + offset, 0, h - 1), yn = clampi(y - offset, 0, h - 1);
		for (int x = 0; x < w; x++) {
			const int xp = clampi(x + offset, 0, w - 1), xn = clampi(x - offset, 0, w - 1);
			const uint8_t* a = rgba + 4 * ((size_t)xp + (size_t)y * w);
			const uint8_t* b = rgba + 4 * ((size_t)xn + (size_t)y * w);
			const uint8_t* c = rgba + 4 * ((size_t)x + (size_t)yp * w);
			const uint8_t* d = rgba + 4 * ((size_t)x + (size_t)yn * w);
			float g[3];
			for (int k = 0; k < 3; k++) {
				const float gx = (float)a[k] - (float)b[k];
				const float gy = (float)c[k] - (float)d[k];
				g[k] = gx * gy;
			}
			out[x + (size_t)y * w] = (g[0] + g[1]) + g[2]; /* gradientDot.cl:29 */
		}
	}
}

/* ------------------------------------------------------------------------ SAT */

/* kernel/satHorizontal.cl:22-31; one work-item per row, NDRange(Hf) Resources.cpp:161 */
void vpo_sat_horizontal(const float* in, int w, int h, float* out)
{
	VPO_PAR_ROWS
	for (int y = 0; y < h; y++) {
		float sum = 0.f;
		for (int x = 0; x < w; x++) {
			sum += in[x + (size_t)y * w];
			out[x + (size_t)y * w] = sum;
		}
	}
}

/* kernel/satVertical.cl:22-31; one work-item per column, NDRange(Wf) Resources.cpp:162.
 * Columns are independent, so the row-major loop nest below is the same arithmetic. */
void vpo_sat_vertical(const float* in, int w, int h, float* out)
{
	const int nt = g_threads;
#pragma omp parallel num_threads(nt) if(nt > 1)
	{
#ifdef _OPENMP
		const int t = omp_get_thread_num(), n = omp_get_num_threads();
#else
		const int t = 0, n = 1;
#endif
		const int x0 = (int)((long long)w * t / n), x1 = (int)((long long)w * (t + 1) / n);
		if (x1 > x0) {
			float* sum = (float*)calloc((size_t)(x1 - x0), sizeof(float));
			for (int y = 0; y < h; y++)
				for (int x = x0; x < x1; x++) {
					sum[x - x0] += in[x + (size_t)y * w];
					out[x + (size_t)y * w] = sum[x - x0];
				}
			free(sum);
		}
	}
}

/* --------------------------------------------------------------------- circle */

/* kernel/satBlobCenter.cl:22-42; launch Resources.cpp:163 */
void vpo_circle(const float* sat, int w, int h, int r, float* out)
{
	const float div = (float)(r * r);
	VPO_PAR_ROWS
	for (int y = 0; y < h; y++) {
		const float* rp = sat + (size_t)clampi(y + r, 0, h - 1) * w; /* dy = +r */
		const float* r1 = sat + (size_t)clampi(y + 1, 0, h - 1) * w; /* dy = +1 */
		const float* m1 = sat + (size_t)clampi(y - 1, 0, h - 1) * w; /* dy = -1 */
		const float* mr = sat + (size_t)clampi(y - r, 0, h - 1) * w; /* dy = -r */
		for (int x = 0; x < w; x++) {
			const int xp = clampi(x + r, 0, w - 1), x1 = clampi(x + 1, 0, w - 1);
			const int xm = clampi(x - 1, 0, w - 1), xr = clampi(x - r, 0, w - 1);
			const float pp = ((rp[xp] - r1[xp]) - rp[x1]) + r1[x1]; /* satBlobCenter.cl:37 */
			const float pn = ((mr[xp] - m1[xp]) - mr[x1]) + m1[x1]; /* :38 */
			const float np = ((rp[xr] - r1[xr]) - rp[xm]) + r1[xm]; /* :39 */
			const float nn = ((mr[xr] - m1[xr]) - mr[xm]) + m1[xm]; /* :40 */
			out[x + (size_t)y * w] = min_cl(min_cl(pp, nn), min_cl(pn, np)) / div; /* :41 */
		}
	}
}

/* ------------------------------------------------------------------- blobList */

typedef struct {
	vpo_match* v;
	int n, cap;
} match_vec;

/* the peak test, disc statistics and record of blobList.cl:38-101 for one pixel.
 * returns 0 = below threshold, 1 = not a peak, 2 = rejected by score, 3 = match */
static inline int blob_at(const uint8_t* rgba, const float* circ, int w, int h, int x, int y,
                          float thr, float min_score, int radius, vpo_match* m, float* score_out)
{
	const float c = circ[x + (size_t)y * w];
	if (c < thr) /* blobList.cl:39 */
		return 0;
	const float cnx = circ[clampi(x - 1, 0, w - 1) + (size_t)y * w];
	const float cpx = circ[clampi(x + 1, 0, w - 1) + (size_t)y * w];
	const float cny = circ[x + (size_t)clampi(y - 1, 0, h - 1) * w];
	const float cpy = circ[x + (size_t)clampi(y + 1, 0, h - 1) * w];
	if (cnx > c || cpx > c || cny > c || cpy > c) /* :47-55 */
		return 1;

	int n = 0;
	uint32_t s1[3] = { 0, 0, 0 }, s2[3] = { 0, 0, 0 };
	const int sq = radius * radius;
	for (int dy = -radius; dy <= radius; dy++) /* :63-72 */
		for (int dx = -radius; dx <= radius; dx++)
			if (dx * dx + dy * dy <= sq) {
				const uint8_t* v = rgba + 4 * ((size_t)clampi(x + dx, 0, w - 1)
				                              + (size_t)clampi(y + dy, 0, h - 1) * w);
				for (int k = 0; k < 3; k++) {
					s1[k] += v[k];
					s2[k] += (uint32_t)v[k] * v[k];
				}
				n++;
			}
	const float fn = (float)n;
	float sd[3];
	for (int k = 0; k < 3; k++) { /* :76, native_sqrt -> correctly rounded sqrt */
		const float f1 = (float)s1[k];
		sd[k] = sqrtf(((float)s2[k] - (f1 * f1) / fn) / fn);
	}
	const float score = c / ((sd[0] + sd[1]) + sd[2]); /* :78 */
	if (score_out)
		*score_out = score;
	if (score < min_score) /* :79 */
		return 2;
	if (m) {
		const uint8_t* ctr = rgba + 4 * ((size_t)x + (size_t)y * w);
		m->x = (float)x + (0.5f * (cnx - cpx)) / ((cnx - 2.0f * c) + cpx); /* :93 */
		m->y = (float)y + (0.5f * (cny - cpy)) / ((cny - 2.0f * c) + cpy); /* :94 */
		for (int k = 0; k < 3; k++) {
			m->color[k] = (uint8_t)(s1[k] / (uint32_t)n); /* :85, uint4 / int */
			m->center[k] = ctr[k];
		}
		m->circ = c;
		m->score = score;
	}
	return 3;
}

/* kernel/blobList.cl:36-102; launch main.cpp:289; counters zeroed main.cpp:283-288 by the
 * caller in the reference -- here the function starts from the values passed in. */
void vpo_blob_list(const uint8_t* rgba, const float* circ, int w, int h,
                   vpo_match* matches, int32_t* counter,
                   float circ_threshold, float min_score, int radius, int max_matches)
{
	match_vec* rows = (match_vec*)calloc((size_t)h, sizeof(match_vec));
	int32_t* rej_score = (int32_t*)calloc((size_t)h, sizeof(int32_t));
	int32_t* rej_peak = (int32_t*)calloc((size_t)h, sizeof(int32_t));
	VPO_PAR_ROWS
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			vpo_match m;
			const int k = blob_at(rgba, circ, w, h, x, y, circ_threshold, min_score, radius, &m, NULL);
			if (k == 1)
				rej_peak[y]++;
			else if (k == 2)
				rej_score[y]++;
			else if (k == 3) {
				match_vec* r = &rows[y];
				if (r->n == r->cap) {
					r->cap = r->cap ? 2 * r->cap : 8;
					r->v = (vpo_match*)realloc(r->v, (size_t)r->cap * sizeof(vpo_match));
				}
				r->v[r->n++] = m;
			}
		}
	for (int y = 0; y < h; y++) { /* sequential raster order == atomic_inc order of a serial run */
		counter[1] += rej_score[y];
		counter[2] += rej_peak[y];
		for (int k = 0; k < rows[y].n; k++) {
			const int i = counter[0]++; /* :87 */
			if (i < max_matches)        /* :88 */
				matches[i] = rows[y].v[k];
		}
		free(rows[y].v);
	}
	free(rows);
	free(rej_score);
	free(rej_peak);
}

/* kernel/blobScore.cl:23-66 (dead in the reference: blob_benchmark.cpp:154-155) */
void vpo_blob_score(const uint8_t* rgba, const float* circ, int w, int h,
                    float circ_threshold, int radius, float* out)
{
	VPO_PAR_ROWS
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			float score = 0.f;
			/* min_score = -inf can never reject: NaN < x and x < -inf are both false */
			const int k = blob_at(rgba, circ, w, h, x, y, circ_threshold, -INFINITY, radius, NULL, &score);
			out[x + (size_t)y * w] = k == 3 ? score : -INFINITY;
		}
}

/* kernel/blobCenter.cl:29-63 (dead: never compiled, absent from Resources.cpp:121-130) */
void vpo_circularize(const float* in, int w, int h, int min_blob_radius, int max_blob_radius, float* out)
{
	(void)min_blob_radius;
	const float sq = ((float)max_blob_radius + 0.5f) * ((float)max_blob_radius + 0.5f);
	VPO_PAR_ROWS
	for (int py = 0; py < h; py++)
		for (int px = 0; px < w; px++) {
			int n = 0;
			float pp = 0.f, pn = 0.f, np = 0.f, nn = 0.f;
			for (int y = 1; y <= max_blob_radius; y++)
				for (int x = 1; x <= max_blob_radius; x++)
					if ((float)(x * x + y * y) <= sq) {
						const int xl = clampi(px - x, 0, w - 1), xr = clampi(px + x, 0, w - 1);
						const int yu = clampi(py + y, 0, h - 1), yd = clampi(py - y, 0, h - 1);
						np += in[xl + (size_t)yu * w];
						pp += in[xr + (size_t)yu * w];
						nn += in[xl + (size_t)yd * w];
						pn += in[xr + (size_t)yd * w];
						n++;
					}
			const float fn = (float)n;
			pp /= fn;
			nn /= fn;
			pn /= fn;
			np /= fn;
			out[px + (size_t)py * w] = min_cl(min_cl(pp, nn), min_cl(-pn, -np));
		}
}

/* ----------------------------------------------------------------------- NV12 */

static inline uint8_t sat_u8_i(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

static inline void nv12_store(uint8_t* out, int w, int wh, int x, int y, uint32_t r, uint32_t g, uint32_t b)
{
	/* rgba2nv12.cl:26-31 / quad2nv12.cl:53-58; Y in uint, UV in int with C truncation */
	const uint32_t yy = (66u * r + 129u * g + 25u * b) / 256u + 16u;
	out[x + (size_t)y * w] = (uint8_t)(yy > 255u ? 255u : yy);
	const size_t uv = (size_t)wh + (size_t)(x / 2) * 2 + (size_t)(y / 2) * w;
	out[uv] = sat_u8_i((-38 * (int)r + -74 * (int)g + 112 * (int)b) / 256 + 128);
	out[uv + 1] = sat_u8_i((112 * (int)r + -94 * (int)g + -18 * (int)b) / 256 + 128);
}

/* kernel/rgba2nv12.cl:22-31; sequential raster order decides the UV race */
void vpo_rgba2nv12(const uint8_t* rgba, int w, int h, uint8_t* out)
{
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			const uint8_t* v = rgba + 4 * ((size_t)x + (size_t)y * w);
			nv12_store(out, w, w * h, x, y, v[0], v[1], v[2]);
		}
}

/* kernel/f2nv12.cl:22-26; convert_uchar_sat(float) = round toward zero, saturate, NaN -> 0 */
void vpo_f2nv12(const float* in, int w, int h, uint8_t* out)
{
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			const float v = in[x + (size_t)y * w] + 127.0f;
			uint8_t o;
			if (!(v > 0.0f))
				o = 0;
			else if (v >= 255.0f)
				o = 255;
			else
				o = (uint8_t)(int)v;
			out[x + (size_t)y * w] = o;
			out[(size_t)w * h + x + (size_t)(y / 2) * w] = 127;
		}
}

/* kernel/quad2nv12.cl:23-58 (integer pos +/- 0.25: no texel-centre offset) */
void vpo_quad2nv12(const uint8_t* ch0, const uint8_t* ch1, const uint8_t* ch2, const uint8_t* ch3,
                   int fmt, int wq, int hq, uint8_t* out, int sample_mode)
{
	const uint8_t* const ch[4] = { ch0, ch1, ch2, ch3 };
	for (int y = 0; y < hq; y++)
		for (int x = 0; x < wq; x++) {
			uint32_t c[3];
			demosaic(ch, fmt, wq, hq, (float)x, (float)y, sample_mode, 1, c);
			nv12_store(out, wq, wq * hq, x, y, c[0], c[1], c[2]);
		}
}

/* kernel/quad2rgba.cl:23-53 */
void vpo_quad2rgba(const uint8_t* ch0, const uint8_t* ch1, const uint8_t* ch2, const uint8_t* ch3,
                   int fmt, int wq, int hq, uint8_t* rgba, int sample_mode)
{
	const uint8_t* const ch[4] = { ch0, ch1, ch2, ch3 };
	VPO_PAR_ROWS
	for (int y = 0; y < hq; y++)
		for (int x = 0; x < wq; x++) {
			uint32_t c[3];
			demosaic(ch, fmt, wq, hq, (float)x, (float)y, sample_mode, 1, c);
			uint8_t* o = rgba + 4 * ((size_t)x + (size_t)y * wq);
			o[0] = (uint8_t)c[0]; /* write_imageui to UNSIGNED_INT8 saturates; max is 255 */
			o[1] = (uint8_t)c[1];
			o[2] = (uint8_t)c[2];
			o[3] = 255;
		}
}

/* ------------------------------------------------------------- whole pipeline */

/* Resources.cpp:138-164 + main.cpp:283-289: stage by stage, a full image per stage */
float vpo_detect(const uint8_t* raw, const vpo_params* p,
                 uint8_t* flat, float* grad_dot, float* sat, float* circ,
                 vpo_match* matches, int32_t* counter, int with_blob_list)
{
	const size_t nq = (size_t)p->wq * p->hq, nf = (size_t)p->wf * p->hf;
	uint8_t* planes = (uint8_t*)calloc(4 * nq, 1);
	uint8_t* flat_ = flat ? flat : (uint8_t*)malloc(4 * nf);
	float* grad_ = grad_dot ? grad_dot : (float*)malloc(4 * nf);
	float* hor = (float*)malloc(4 * nf);
	float* sat_ = sat ? sat : (float*)malloc(4 * nf);
	float* circ_ = circ ? circ : (float*)malloc(4 * nf);

	vpo_raw2quad(raw, p->fmt, p->wq, p->hq, planes, planes + nq, planes + 2 * nq, planes + 3 * nq);
	vpo_resampling(planes, planes + nq, planes + 2 * nq, planes + 3 * nq, p->fmt, p->wq, p->hq,
	               flat_, p->wf, p->hf, &p->model, p->max_robot_height, p->field_scale,
	               p->off_x, p->off_y, p->sample_mode);
	vpo_gradient_dot(flat_, p->wf, p->hf, p->grad_offset, grad_);
	vpo_sat_horizontal(grad_, p->wf, p->hf, hor);
	vpo_sat_vertical(hor, p->wf, p->hf, sat_);
	vpo_circle(sat_, p->wf, p->hf, p->circle_radius, circ_);
	if (with_blob_list && matches && counter) {
		counter[0] = counter[1] = counter[2] = 0; /* main.cpp:283-288 */
		vpo_blob_list(flat_, circ_, p->wf, p->hf, matches, counter, p->circ_threshold,
		              p->min_score, p->blob_radius, p->max_blobs);
	}
	float mx = 0.f;
	for (size_t i = 0; i < nf; i++) {
		const float a = fabsf(sat_[i]);
		const float b = fabsf(hor[i]);
		if (a > mx) mx = a;
		if (b > mx) mx = b;
	}
	free(planes);
	free(hor);
	if (!flat) free(flat_);
	if (!grad_dot) free(grad_);
	if (!sat) free(sat_);
	if (!circ) free(circ_);
	return mx;
}
