/*
 * geometry.cpp -- host-side parameter derivation behind the C ABI.
 *
 * Dependency-free C++ restatement (no Eigen, no protobuf) of the CPU code that turns a camera calibration and the field
 * size into the scalars the detection kernels are launched with:
 *
 *   CameraModel(const SSL_GeometryCameraCalibration&)     src/CameraModel.cpp:80-88
 *   CameraModel::ensureSize / normalizeUndistort          src/CameraModel.cpp:124-141
 *   CameraModel::field2image (10 iterations, CPU twin)    src/CameraModel.cpp:147-157
 *   CameraModel::image2field                              src/CameraModel.cpp:159-172
 *   Perspective::geometryCheck                            src/Perspective.cpp:35-125
 *   Perspective::flat2field / field2flat                  src/Perspective.cpp:127-133
 *   Perspective::getCLCameraModel                         src/Perspective.cpp:136-150
 *   launch scalars                                        src/Resources.cpp:159-163, src/main.cpp:289
 *
 * Arithmetic follows the reference: fp32 throughout, the field-scale sum accumulated sequentially in fp32 in raster
 * order (Perspective.cpp:78-91).  What Eigen does inside (quaternion normalisation, quaternion * vector, the order of
 * the three products of a 3x3 * 3-vector) is written out explicitly here; Eigen is absent from this image, so agreement
 * with a build against Eigen is to the last ulp of those products, not pinned bit for bit.
 *
 * Runs once per geometry change; nothing here is on the per-frame path and nothing here touches the GPU.
 */
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include <limits>

#include "vp_b200.h"

namespace {

struct Vec3 {
	float x, y, z;
};

/* Eigen::Quaternionf::normalize + toRotationMatrix (row-major) */
void quat_to_matrix(float qx, float qy, float qz, float qw, float r[9])
{
	const float n = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
	if (n > 0.f) {
		qx /= n; qy /= n; qz /= n; qw /= n;
	}
	const float tx = 2.f * qx, ty = 2.f * qy, tz = 2.f * qz;
	const float twx = tx * qw, twy = ty * qw, twz = tz * qw;
	const float txx = tx * qx, txy = ty * qx, txz = tz * qx;
	const float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
	r[0] = 1.f - (tyy + tzz); r[1] = txy - twz;         r[2] = txz + twy;
	r[3] = txy + twz;         r[4] = 1.f - (txx + tzz); r[5] = tyz - twx;
	r[6] = txz - twy;         r[7] = tyz + twx;         r[8] = 1.f - (txx + tyy);
}

Vec3 mul(const float r[9], Vec3 v) { return { r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z }; }
Vec3 mul_t(const float r[9], Vec3 v) { return { r[0] * v.x + r[3] * v.y + r[6] * v.z, r[1] * v.x + r[4] * v.y + r[7] * v.z, r[2] * v.x + r[5] * v.y + r[8] * v.z }; }

/* CameraModel::normalizeUndistort, CameraModel.cpp:137-141 */
void normalize_undistort(const vp_camera_model& m, float px, float py, float& nx, float& ny)
{
	nx = (px - m.p[0]) / m.f;
	ny = (py - m.p[1]) / m.f;
	const float s = 1.0f + m.d * (nx * nx + ny * ny);
	nx *= s;
	ny *= s;
}

/* CameraModel::image2field, CameraModel.cpp:159-172; false when the ray does not hit the plane (NaN in the reference) */
bool image2field(const vp_camera_model& m, float px, float py, float height, float out[3])
{
	float nx, ny;
	normalize_undistort(m, px, py, nx, ny);
	const Vec3 ray = mul_t(m.r, { nx, ny, 1.0f }); /* i2f = inverse of the unit quaternion's rotation = its transpose */
	if (ray.z >= 0) {
		out[0] = out[1] = out[2] = std::numeric_limits<float>::quiet_NaN();
		return false;
	}
	const float k = (-m.c[2] + height) / ray.z;
	out[0] = ray.x * k + m.c[0];
	out[1] = ray.y * k + m.c[1];
	out[2] = height;
	return true;
}

float norm2(float ax, float ay) { return std::sqrt(ax * ax + ay * ay); }

} // namespace

extern "C" {

/* CameraModel(const SSL_GeometryCameraCalibration&) + updateDerived + Perspective::getCLCameraModel */
int vp_camera_model_from_calib(const vp_camera_calib* c, vp_camera_model* out)
{
	if (!c || !out)
		return VP_ERR_INVALID;
	if (!(c->focal_length > 0.f) || c->pixel_image_width <= 0 || c->pixel_image_height <= 0)
		return VP_ERR_INVALID;
	vp_camera_model m;
	m.shape[0] = c->pixel_image_width;
	m.shape[1] = c->pixel_image_height;
	m.f = c->focal_length;
	m.p[0] = c->principal_point_x;
	m.p[1] = c->principal_point_y;
	m.d = c->distortion;
	/* f2iOrientation(calib.q3(), calib.q0(), calib.q1(), calib.q2()) = (w, x, y, z), CameraModel.cpp:84 */
	quat_to_matrix(c->q0, c->q1, c->q2, c->q3, m.r);
	/* pos = f2iOrientation.inverse() * -(tx, ty, tz), CameraModel.cpp:86; a derived world position overrides nothing in the
	 * reference (it only suppresses the re-broadcast, Perspective.cpp:45), so it is not read here either */
	const Vec3 p = mul_t(m.r, { -c->tx, -c->ty, -c->tz });
	m.c[0] = p.x; m.c[1] = p.y; m.c[2] = p.z;
	*out = m;
	return VP_OK;
}

/* CameraModel::ensureSize, CameraModel.cpp:124-135 */
int vp_camera_model_ensure_size(vp_camera_model* m, int width, int height)
{
	if (!m || width <= 0 || height <= 0 || m->shape[0] <= 0)
		return VP_ERR_INVALID;
	if (m->shape[0] == width && m->shape[1] == height)
		return VP_OK;
	const float factor = (float)width / (float)m->shape[0];
	m->shape[0] = width;
	m->shape[1] = height;
	m->f *= factor;
	m->p[0] *= factor;
	m->p[1] *= factor;
	return VP_OK;
}

/* CameraModel::field2image, CameraModel.cpp:147-157: the CPU twin of resampling.cl:29-47 with TEN iterations and the
 * translation folded into the affine transform (R*p + R*(-pos)) as Eigen::Affine3f(q) * Translation3f(-pos) does */
int vp_field2image(const vp_camera_model* m, const float field[3], float image[2])
{
	if (!m || !field || !image)
		return VP_ERR_INVALID;
	const Vec3 t = mul(m->r, { -m->c[0], -m->c[1], -m->c[2] });
	const Vec3 rp = mul(m->r, { field[0], field[1], field[2] });
	const Vec3 ray = { rp.x + t.x, rp.y + t.y, rp.z + t.z };
	const float ox = ray.x / ray.z, oy = ray.y / ray.z;
	float nx = ox, ny = oy;
	for (int i = 0; i < 10; i++) {
		const float s = 1 + m->d * (nx * nx + ny * ny);
		nx = ox / s;
		ny = oy / s;
	}
	image[0] = m->f * nx + m->p[0];
	image[1] = m->f * ny + m->p[1];
	return VP_OK;
}

int vp_image2field(const vp_camera_model* m, const float image[2], float height, float field[3])
{
	if (!m || !image || !field)
		return VP_ERR_INVALID;
	image2field(*m, image[0], image[1], height, field);
	return VP_OK;
}

/* Perspective::geometryCheck, Perspective.cpp:66-124 (the part after the calibration lookup) */
int vp_geometry_check(const vp_camera_model* model, const vp_field_size* field, int width, int height, double max_bot_height, float resampling_factor,
                      float geometry_tolerance, vp_geometry* out)
{
	if (!model || !field || !out || width < 2 || height < 2)
		return VP_ERR_INVALID;
	vp_geometry g;
	std::memset(&g, 0, sizeof g);
	g.model = *model;
	int rc = vp_camera_model_ensure_size(&g.model, width, height);
	if (rc)
		return rc;
	const vp_camera_model& m = g.model;
	const float h = (float)max_bot_height;
	const float goal_boundary = field->boundary_width_goal_line >= 0.f ? field->boundary_width_goal_line : field->boundary_width; /* CameraModel.cpp:18-20 */
	constexpr float CENTER_BLOB_RADIUS = 25.0f, SIDE_BLOB_RADIUS = 20.0f; /* pattern.h:55-56 */
	g.min_blob_radius = std::fmin(std::fmin(CENTER_BLOB_RADIUS, SIDE_BLOB_RADIUS), field->ball_radius);
	g.max_blob_radius = std::fmax(std::fmax(CENTER_BLOB_RADIUS, SIDE_BLOB_RADIUS), field->ball_radius);

	/* optimal field scale: mean distance to the right and lower neighbour over the pixels that see the field, :72-91.
	 * One image2field per pixel instead of three: the neighbours' positions are this row's / the next row's own values. */
	float min_scale = std::numeric_limits<float>::max(), max_scale = 0.f, sum = 0.f;
	long long n = 0;
	const float lim_x = field->field_length / 2.f + goal_boundary, lim_y = field->field_width / 2.f + field->boundary_width;
	/* The field position of every pixel (one image2field each) and the neighbour distances dx + dy derived from them are
	 * independent and computed by all host threads; minimum, maximum and count do not depend on the order either.  The SUM stays
	 * one fp32 variable fed in raster order, exactly like the reference's loop -- past 2^23 every dx + dy is absorbed with a
	 * rounding error that depends on that order.  43 ms -> a few ms for 1224x1024 on 16 cores. */
	std::vector<float> pos((size_t)2 * width * height);
	std::vector<float> dist((size_t)width * height); /* dx + dy of pixel (x, y), or -1 where the pixel does not count */
	const unsigned hw = std::thread::hardware_concurrency();
	const int n_threads = (int)std::max(1u, std::min(hw ? hw : 1u, std::min(16u, (unsigned)height / 64u + 1u)));
	auto in_parallel = [&](auto&& body) {
		std::vector<std::thread> pool;
		for (int t = 1; t < n_threads; t++)
			pool.emplace_back(body, t, (int)((long long)height * t / n_threads), (int)((long long)height * (t + 1) / n_threads));
		body(0, 0, (int)((long long)height * 1 / n_threads));
		for (std::thread& t : pool)
			t.join();
	};
	in_parallel([&](int, int y0, int y1) {
		float p[3];
		for (int y = y0; y < y1; y++)
			for (int x = 0; x < width; x++) {
				image2field(m, (float)x, (float)y, h, p);
				pos[((size_t)y * width + x) * 2] = p[0];
				pos[((size_t)y * width + x) * 2 + 1] = p[1];
			}
	});
	std::vector<float> t_min(n_threads, std::numeric_limits<float>::max()), t_max(n_threads, 0.f);
	std::vector<long long> t_n(n_threads, 0);
	in_parallel([&](int t, int y0, int y1) {
		float mn = std::numeric_limits<float>::max(), mx = 0.f;
		long long cnt = 0;
		for (int y = y0; y < std::min(y1, height - 1); y++) {
			const float* cur = &pos[(size_t)y * width * 2];
			const float* nxt = cur + (size_t)width * 2;
			float* d = &dist[(size_t)y * width];
			for (int x = 0; x < width - 1; x++) {
				const float px = cur[2 * x], py = cur[2 * x + 1];
				d[x] = -1.f;
				if (std::fabs(px) < lim_x && std::fabs(py) < lim_y) { /* false for NaN, as in the reference */
					const float dx = norm2(cur[2 * x + 2] - px, cur[2 * x + 3] - py);
					const float dy = norm2(nxt[2 * x] - px, nxt[2 * x + 1] - py);
					mn = std::fmin(mn, std::fmin(dx, dy));
					mx = std::fmax(mx, std::fmax(dx, dy));
					const float v = dx + dy;
					d[x] = v >= 0.f ? v : std::numeric_limits<float>::quiet_NaN(); /* a NaN distance still counts (and poisons the sum) */
					cnt += 2;
				}
			}
		}
		t_min[t] = mn;
		t_max[t] = mx;
		t_n[t] = cnt;
	});
	for (int t = 0; t < n_threads; t++) {
		min_scale = std::fmin(min_scale, t_min[t]);
		max_scale = std::fmax(max_scale, t_max[t]);
		n += t_n[t];
	}
	for (int y = 0; y < height - 1; y++) { /* Perspective.cpp:78-91: one running fp32 sum, raster order */
		const float* d = &dist[(size_t)y * width];
		for (int x = 0; x < width - 1; x++)
			if (!(d[x] < 0.f))
				sum += d[x];
	}
	float tmp[3];
	g.field_scale = sum / (float)n * resampling_factor; /* 0/0 = NaN when no pixel sees the field, as in the reference */
	g.min_field_scale = min_scale;
	g.max_field_scale = max_scale;

	/* visible extent from the image border, :94-105; comparisons with NaN are false, so NaN points are ignored */
	float ext[4];
	image2field(m, 0.f, 0.f, h, tmp);
	ext[0] = ext[1] = tmp[0];
	ext[2] = ext[3] = tmp[1];
	auto update = [&](float ix, float iy) { /* updateExtent, Perspective.cpp:24-33 */
		float p[3];
		image2field(m, ix, iy, h, p);
		if (p[0] < ext[0]) ext[0] = p[0];
		if (p[0] > ext[1]) ext[1] = p[0];
		if (p[1] < ext[2]) ext[2] = p[1];
		if (p[1] > ext[3]) ext[3] = p[1];
	};
	for (int x = 0; x < width; x++) {
		update((float)x, 0.0f);
		update((float)x, (float)height - 1.0f);
	}
	for (int y = 0; y < height; y++) {
		update(0.0f, (float)y);
		update((float)width - 1.0f, (float)y);
	}
	/* clamp to the field, :107-113 */
	const float half_l = field->field_length / 2.0f + goal_boundary + geometry_tolerance;
	const float half_w = field->field_width / 2.0f + field->boundary_width + geometry_tolerance;
	ext[0] = std::fmax(ext[0], -half_l);
	ext[1] = std::fmin(ext[1], half_l);
	ext[2] = std::fmax(ext[2], -half_w);
	ext[3] = std::fmin(ext[3], half_w);
	for (int i = 0; i < 4; i++)
		g.visible_field_extent[i] = ext[i];

	/* flat size: rounded to nearest (ties to even, Eigen rint), then made even "for rtpstreamer", :115-122 */
	int w = (int)std::nearbyint((ext[1] - ext[0]) / g.field_scale);
	int hh = (int)std::nearbyint((ext[3] - ext[2]) / g.field_scale);
	if (w % 2) w++;
	if (hh % 2) hh++;
	g.reprojected_field_size[0] = w;
	g.reprojected_field_size[1] = hh;
	*out = g;
	if (!(g.field_scale > 0.f) || w <= 0 || hh <= 0)
		return VP_ERR_UNSUPPORTED; /* the camera does not see the field: the reference would go on with NaN / empty images */
	return VP_OK;
}

/* Perspective::flat2field / field2flat, Perspective.cpp:127-133 */
int vp_flat2field(const vp_geometry* g, const float flat[2], float field[2])
{
	if (!g || !flat || !field)
		return VP_ERR_INVALID;
	field[0] = flat[0] * g->field_scale + g->visible_field_extent[0];
	field[1] = flat[1] * g->field_scale + g->visible_field_extent[2];
	return VP_OK;
}

int vp_field2flat(const vp_geometry* g, const float field[2], float flat[2])
{
	if (!g || !flat || !field)
		return VP_ERR_INVALID;
	flat[0] = (field[0] - g->visible_field_extent[0]) / g->field_scale;
	flat[1] = (field[1] - g->visible_field_extent[2]) / g->field_scale;
	return VP_OK;
}

/* the scalars of Resources.cpp:159-163 and main.cpp:283-289 for one geometry */
int vp_geometry_params(const vp_geometry* g, int fmt, int wq, int hq, double max_bot_height, float circ_threshold, int max_blobs, int sample_mode,
                       vp_params* out)
{
	if (!g || !out || !(g->field_scale > 0.f))
		return VP_ERR_INVALID;
	vp_params p;
	std::memset(&p, 0, sizeof p);
	p.fmt = fmt;
	p.wq = wq;
	p.hq = hq;
	p.wf = g->reprojected_field_size[0];
	p.hf = g->reprojected_field_size[1];
	p.model = g->model; /* Perspective::getCLCameraModel(): the same 72 bytes */
	p.max_robot_height = (float)max_bot_height;
	p.field_scale = g->field_scale;
	p.off_x = g->visible_field_extent[0];
	p.off_y = g->visible_field_extent[2];
	p.grad_offset = (int)std::ceil(g->max_blob_radius / g->field_scale) / 3; /* integer division, Resources.cpp:160 */
	p.circle_radius = (int)std::ceil(g->min_blob_radius / g->field_scale);   /* Resources.cpp:163 */
	p.circ_threshold = circ_threshold;
	p.min_score = 0.0f;                                                      /* literal at main.cpp:289 */
	p.blob_radius = (int)std::floor(g->min_blob_radius / g->field_scale);    /* main.cpp:289 */
	p.max_blobs = max_blobs;
	p.sample_mode = sample_mode;
	*out = p;
	return VP_OK;
}

} /* extern "C" */
