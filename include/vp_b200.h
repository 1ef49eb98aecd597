/*
 * vp_b200.h -- C ABI of libvp_b200.so, the B200-native (sm_100a CUDA) replacement for the
 * OpenCL layer of TIGERs-Mannheim/vision-processor (src/opencl.h, src/opencl.cpp) and the
 * 13 kernels under kernel/ *.cl.
 *
 * All citations are path:line in the reference tree.  Every entry point returns a vp_status
 * (0 = ok) unless noted; the text of the last failure is available through vp_last_error().
 * The reference turns every OpenCL failure into FATAL = log + exit(1) (src/log.h:21,
 * src/opencl.h:81-83); the C++ shim include/compat/opencl.h restores that behaviour on top
 * of these status codes.
 *
 * Threading: kernels are launched from ONE thread per context (the reference launches only
 * from its main thread, src/main.cpp:262-423); buffers and images may be mapped and unmapped
 * from other threads (src/rtpstreamer.cpp:177, src/snapshotwriter.cpp:52-54).
 *
 * There is no CPU fallback: without a CUDA device vp_ctx_create fails.
 */
#ifndef VP_B200_H
#define VP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VP_API __attribute__((visibility("default")))
#else
#define VP_API
#endif

typedef enum {
	VP_OK = 0,
	VP_ERR_INVALID = 1,     /* bad argument (null handle, size mismatch, unknown format) */
	VP_ERR_CUDA = 2,        /* a CUDA runtime call failed; see vp_last_error */
	VP_ERR_NOMEM = 3,
	VP_ERR_UNSUPPORTED = 4,
	VP_ERR_NO_DEVICE = 5
} vp_status;

/* PixelFormat table, src/opencl.cpp:24-31 / src/opencl.h:30-56.  pixelSize = stride*rowStride. */
typedef enum {
	VP_FMT_RGGB8 = 0, /* stride 2, rowStride 2, "-DRGGB"; width/height count QUADS (spinnakerdriver.cpp:124) */
	VP_FMT_GRBG8 = 1, /* stride 2, rowStride 2, "-DGRBG" */
	VP_FMT_BGR8 = 2,  /* stride 3, rowStride 1, "-DBGR" */
	VP_FMT_RGBA8 = 3, /* stride 4 */
	VP_FMT_U8 = 4,    /* stride 1 */
	VP_FMT_F32 = 5,   /* stride 4 */
	VP_FMT_NV12 = 6   /* stride 1, rowStride 2 (allocated 2*w*h, 1.5*w*h used; opencl.cpp:27) */
} vp_format;

VP_API int vp_format_pixel_size(int fmt); /* PixelFormat::pixelSize(), opencl.h:44 */

/* What read_imageui + CLK_FILTER_LINEAR means on an integer image (left to the OpenCL runtime:
 * resampling.cl:50, quad2nv12.cl:21, quad2rgba.cl:21).  Default = bilinear per the OpenCL 1.2
 * formula in fp32 (no FMA), converted round-to-nearest-even. */
typedef enum { VP_SAMPLE_BILINEAR_RTE = 0, VP_SAMPLE_BILINEAR_TRUNC = 1, VP_SAMPLE_NEAREST = 2 } vp_sample_mode;

/* CLCameraModel: src/Perspective.h:22-29 == kernel/resampling.cl:20-27, packed, 72 bytes, by value */
typedef struct __attribute__((packed)) {
	int32_t shape[2];
	float f;     /* focal length in (quad) px */
	float p[2];  /* principal point */
	float d;     /* k2 distortion */
	float r[9];  /* row-major field->image rotation */
	float c[3];  /* camera position, mm */
} vp_camera_model;

/* CLMatch: src/main.cpp:33-41 == kernel/blobList.cl:20-32, packed, 22 bytes */
typedef struct __attribute__((packed)) {
	float x, y;
	uint8_t color[3];
	uint8_t center[3];
	float circ;
	float score;
} vp_match;

/* Every scalar the seven live kernels receive for one camera geometry
 * (src/Resources.cpp:159-163, src/main.cpp:289). */
typedef struct {
	int32_t fmt;  /* VP_FMT_RGGB8 | VP_FMT_GRBG8 | VP_FMT_BGR8 */
	int32_t wq, hq; /* RawImage width/height: quads for Bayer, pixels for BGR */
	int32_t wf, hf; /* Perspective::reprojectedFieldSize (even, Perspective.cpp:115-122) */
	vp_camera_model model;
	float max_robot_height; /* (float)gcSocket->maxBotHeight */
	float field_scale;
	float off_x, off_y;     /* visibleFieldExtent[0], [2] */
	int32_t grad_offset;    /* (int)ceilf(maxBlobRadius/fieldScale) / 3 */
	int32_t circle_radius;  /* (int)ceilf(minBlobRadius/fieldScale) */
	float circ_threshold;   /* thresholds.circularity */
	float min_score;        /* literal 0.0f at main.cpp:289 */
	int32_t blob_radius;    /* (int)floorf(minBlobRadius/fieldScale) */
	int32_t max_blobs;      /* thresholds.blobs */
	int32_t sample_mode;    /* vp_sample_mode */
} vp_params;

/* ---- host-side parameter derivation (CPU only, no context needed) ----------------------------------
 * The subset of SSL_GeometryCameraCalibration / SSL_GeometryFieldSize (ssl_vision_geometry.proto) the path reads, as plain
 * structs so that no protobuf is needed on this side of the boundary. */
typedef struct {
	int32_t pixel_image_width, pixel_image_height; /* quad pixels for Bayer cameras */
	float focal_length, principal_point_x, principal_point_y, distortion;
	float q0, q1, q2, q3; /* field -> image orientation, (x, y, z, w) as in the protobuf */
	float tx, ty, tz;     /* translation in the camera frame, mm */
} vp_camera_calib;

typedef struct {
	float field_length, field_width;   /* mm */
	float boundary_width;
	float boundary_width_goal_line;    /* < 0: field not present (goalBoundaryWidth(), CameraModel.cpp:18-20) */
	float ball_radius;
} vp_field_size;

/* what Perspective::geometryCheck leaves behind (src/Perspective.h:32-58) */
typedef struct {
	vp_camera_model model;             /* Perspective::getCLCameraModel(), after ensureSize */
	float field_scale;                 /* mm per flat pixel */
	float visible_field_extent[4];     /* xmin, xmax, ymin, ymax, mm */
	int32_t reprojected_field_size[2]; /* even */
	float min_blob_radius, max_blob_radius;
	float min_field_scale, max_field_scale; /* what the reference only logs (Perspective.cpp:92) */
} vp_geometry;

/* CameraModel::CameraModel(const SSL_GeometryCameraCalibration&), src/CameraModel.cpp:80-88 */
VP_API int vp_camera_model_from_calib(const vp_camera_calib* calib, vp_camera_model* out);
/* CameraModel::ensureSize, src/CameraModel.cpp:124-135 */
VP_API int vp_camera_model_ensure_size(vp_camera_model* model, int width, int height);
/* CameraModel::field2image (the 10-iteration CPU twin of the kernel's projection), src/CameraModel.cpp:147-157 */
VP_API int vp_field2image(const vp_camera_model* model, const float field[3], float image[2]);
/* CameraModel::image2field, src/CameraModel.cpp:159-172 (NaN when the ray misses the plane) */
VP_API int vp_image2field(const vp_camera_model* model, const float image[2], float height, float field[3]);
/* Perspective::geometryCheck, src/Perspective.cpp:66-124: field scale, visible extent, flat size, blob radii.
 * VP_ERR_UNSUPPORTED (with *out filled in) when the camera does not see the field. */
VP_API int vp_geometry_check(const vp_camera_model* model, const vp_field_size* field, int width, int height, double max_bot_height,
                             float resampling_factor, float geometry_tolerance, vp_geometry* out);
/* Perspective::flat2field / field2flat, src/Perspective.cpp:127-133 */
VP_API int vp_flat2field(const vp_geometry* g, const float flat[2], float field[2]);
VP_API int vp_field2flat(const vp_geometry* g, const float field[2], float flat[2]);
/* the launch scalars of src/Resources.cpp:159-163 and src/main.cpp:283-289 -> vp_params for the fused entry points */
VP_API int vp_geometry_params(const vp_geometry* g, int fmt, int wq, int hq, double max_bot_height, float circ_threshold, int max_blobs,
                              int sample_mode, vp_params* out);

typedef struct vp_ctx vp_ctx;   /* class OpenCL, opencl.h:69-112: device + in-order queue + pools */
typedef struct vp_buf vp_buf;   /* class CLArray / RawImage storage, opencl.h:154-188 */
typedef struct vp_img vp_img;   /* class CLImage storage, opencl.h:195-212 */

/* ---- context (OpenCL::OpenCL, opencl.cpp:34-50) ------------------------------------------- */
VP_API int vp_ctx_create(int device_ordinal, vp_ctx** out);
VP_API void vp_ctx_destroy(vp_ctx* ctx);
VP_API const char* vp_last_error(const vp_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
VP_API int vp_ctx_sync(vp_ctx* ctx);                  /* OpenCL::wait on everything enqueued so far */
VP_API void* vp_ctx_stream(vp_ctx* ctx);              /* the cudaStream_t all stages are launched on */
VP_API int vp_device_count(void);

/* per-launch profiling, OpenCL::run/printRuntimes/clearEvents (opencl.h:76-96, opencl.cpp:94-105).
 * Disabled by default; when enabled every stage call records a CUDA event pair. */
VP_API int vp_profiling_enable(vp_ctx* ctx, int on);
VP_API int vp_profiling_count(vp_ctx* ctx);
VP_API int vp_profiling_get(vp_ctx* ctx, int i, const char** stage_name, float* ms); /* blocks on the event */
VP_API int vp_profiling_clear(vp_ctx* ctx);

/* ---- buffers (CLArray, opencl.cpp:146-147) -------------------------------------------------
 * Device memory with a pinned host mirror.  map(READ) copies device->host and blocks;
 * unmap after map(WRITE|READWRITE) copies host->device and blocks (opencl.h:115-152).
 * Read maps hand out writable memory (blob_benchmark.cpp:190-191 sorts a read map in place). */
enum { VP_MAP_READ = 1, VP_MAP_WRITE = 2 /* invalidate */, VP_MAP_READWRITE = 3 };
VP_API int vp_buf_alloc(vp_ctx* ctx, size_t bytes, vp_buf** out);                        /* CL_MEM_ALLOC_HOST_PTR */
VP_API int vp_buf_alloc_copy(vp_ctx* ctx, const void* host, size_t bytes, vp_buf** out); /* CL_MEM_COPY_HOST_PTR */
VP_API int vp_buf_retain(vp_buf* buf);  /* RawImage copies share the buffer (opencl.h:170) */
VP_API int vp_buf_release(vp_buf* buf);
VP_API int vp_buf_map(vp_buf* buf, int mode, void** host);
VP_API int vp_buf_unmap(vp_buf* buf);
VP_API size_t vp_buf_size(const vp_buf* buf);
VP_API void* vp_buf_device_ptr(vp_buf* buf);

/* ---- images (CLImage, opencl.cpp:149-158); formats RGBA8, U8, F32; dense pitch -------------- */
VP_API int vp_img_alloc(vp_ctx* ctx, int fmt, int width, int height, vp_img** out);
VP_API int vp_img_retain(vp_img* img);
VP_API int vp_img_release(vp_img* img);
VP_API int vp_img_map(vp_img* img, int mode, void** host, size_t* byte_pitch); /* CLImageMap, opencl.h:215-262 */
VP_API int vp_img_unmap(vp_img* img);
VP_API int vp_img_info(const vp_img* img, int* fmt, int* width, int* height);
VP_API void* vp_img_device_ptr(vp_img* img);

/* ---- stages: one call per reference kernel, asynchronous on the context stream -------------- */
/* kernel/raw2quad.cl:21-39, launch Resources.cpp:138-143 (NDRange wq x hq) */
VP_API int vp_raw2quad(vp_ctx* ctx, const vp_buf* raw, int fmt, int wq, int hq, vp_img* const ch[4]);
/* kernel/resampling.cl:52-99, launch Resources.cpp:159 (NDRange = size of `flat`) */
VP_API int vp_resampling(vp_ctx* ctx, vp_img* const ch[4], int fmt, vp_img* flat, const vp_camera_model* model,
                         float max_robot_height, float field_scale, float off_x, float off_y, int sample_mode);
/* kernel/gradientDot.cl:22-30, launch Resources.cpp:160 */
VP_API int vp_gradient_dot(vp_ctx* ctx, const vp_img* rgba, vp_img* out, int offset);
/* kernel/satHorizontal.cl:22-31 (NDRange height), kernel/satVertical.cl:22-31 (NDRange width); Resources.cpp:161-162 */
VP_API int vp_sat_horizontal(vp_ctx* ctx, const vp_img* in, vp_img* out);
VP_API int vp_sat_vertical(vp_ctx* ctx, const vp_img* in, vp_img* out);
/* kernel/satBlobCenter.cl:22-42, launch Resources.cpp:163 */
VP_API int vp_circle(vp_ctx* ctx, const vp_img* sat, vp_img* out, int radius);
/* kernel/blobList.cl:36-102, launch main.cpp:289.  matches: >= 22*max_matches bytes; counter: 3 x int32,
 * incremented from the values it holds (the caller zeroes it, main.cpp:283-288).  Blobs are emitted in raster
 * order (y-major); on overflow the first max_matches in that order are kept and counter[0] keeps counting. */
VP_API int vp_blob_list(vp_ctx* ctx, const vp_img* rgba, const vp_img* circ, vp_buf* matches, vp_buf* counter,
                        float circ_threshold, float min_score, int radius, int max_matches);
/* kernel/rgba2nv12.cl:22-31, f2nv12.cl:22-26 (Resources.cpp:172-186); quad2nv12.cl:23-58 (Resources.cpp:166-170);
 * quad2rgba.cl:23-53 (Resources.cpp:145-149).  The racing UV writes of the reference are resolved to the
 * bottom-right pixel of each 2x2 block (the last writer of a sequential raster-order run). */
VP_API int vp_rgba2nv12(vp_ctx* ctx, const vp_img* rgba, vp_buf* nv12);
VP_API int vp_f2nv12(vp_ctx* ctx, const vp_img* f32, vp_buf* nv12);
VP_API int vp_quad2nv12(vp_ctx* ctx, vp_img* const ch[4], int fmt, vp_buf* nv12, int sample_mode);
VP_API int vp_quad2rgba(vp_ctx* ctx, vp_img* const ch[4], int fmt, vp_img* rgba, int sample_mode);
/* dead kernels named by north_star: kernel/blobCenter.cl:29-63 (never compiled), kernel/blobScore.cl:23-66
 * (enqueue commented out, blob_benchmark.cpp:154-155) */
VP_API int vp_circularize(vp_ctx* ctx, const vp_img* in, vp_img* out, int min_blob_radius, int max_blob_radius);
VP_API int vp_blob_score(vp_ctx* ctx, const vp_img* rgba, const vp_img* circ, vp_img* out, float circ_threshold, int radius);

/* ---- fused detection: Resources::raw2quad + rgba2blobCenter (Resources.cpp:138-164) + counter reset and
 * blobList (main.cpp:283-289) for `n_frames` frames of one camera geometry, no quad planes materialised.
 *
 * Device-pointer form (used by the batch benchmark and by torch/CUDA callers): all pointers are device
 * memory, frame i at base + i*stride:
 *   d_raw      n_frames x raw_bytes          (raw_bytes = wq*hq*pixelSize(fmt))
 *   d_flat     n_frames x wf*hf*4  RGBA8     (dRGB image `flat`)
 *   d_grad     n_frames x wf*hf    F32       (`gradDot`)
 *   d_circ     n_frames x wf*hf    F32       (`blobCenter`)
 *   d_matches  n_frames x max_blobs*22 bytes
 *   d_counter  n_frames x 3 int32            (zeroed by the call)
 * Alignment contract: d_raw, d_flat, d_grad and d_circ must be 16-byte aligned (the kernels move them as 16-byte vectors;
 * anything from cudaMalloc / vp_buf_alloc / a torch tensor's data_ptr at offset 0 is), d_counter 4-byte, d_matches 2-byte;
 * a misaligned pointer is refused with VP_ERR_INVALID before anything is launched.
 * Asynchronous on the context stream. */
VP_API int vp_detect_batch_device(vp_ctx* ctx, const uint8_t* d_raw, int n_frames, const vp_params* params,
                                  uint8_t* d_flat, float* d_grad, float* d_circ,
                                  vp_match* d_matches, int32_t* d_counter);

/* Host form: raw frames in host memory (pinned for full speed, see vp_host_alloc), blob lists and counters
 * back in host memory; images stay on the device (vp_detect_images).  Blocking.  Internally a ring of
 * device slots so the upload of frame i+1 overlaps the kernels of frame i. */
VP_API int vp_detect_host(vp_ctx* ctx, const uint8_t* h_raw, int n_frames, const vp_params* params,
                          vp_match* h_matches, int32_t* h_counter);
/* device pointers of the images produced for the LAST frame of the most recent vp_detect_host call */
VP_API int vp_detect_images(vp_ctx* ctx, const uint8_t** d_flat, const float** d_grad, const float** d_circ);
/* number of frames of the last fused call whose SAT left the exact-integer range of fp32 (2^24) and were
 * recomputed in the reference's sequential summation order */
VP_API int vp_detect_sat_fallbacks(vp_ctx* ctx, int* n);

/* ---- blob list -> hypothesis hand-off ------------------------------------------------------------
 * struct Match, src/blobs/match.h:22-30 (Eigen::Vector2f pos; Vector3i color, center; float circ, score): 40 bytes */
typedef struct {
	float pos[2];      /* field millimetres: Perspective::flat2field(match.x, match.y), main.cpp:303 */
	int32_t color[3];
	int32_t center[3];
	float circ, score;
} vp_field_match;
/* What src/main.cpp:297-325 does per frame on the CPU (copy into `Match` records, KD-tree insert), for a whole batch on the
 * device: d_out[f][i] = record of blob i (i < min(counter[3f], max_blobs), list order kept), and a uniform grid of
 * cell_mm x cell_mm cells over the visible extent starting at (off_x, off_y) instead of the tree: d_order[f][k] = blob
 * indices sorted by (cell, index), d_cell_start[f][c] = first k of cell c (cells_x*cells_y + 1 entries per frame; cells
 * are row-major, positions outside the grid -- and NaN positions of plateau peaks -- are clamped into it).  A radius
 * search visits the cells the disc overlaps.  max_blobs <= 4096. */
VP_API int vp_blobs_to_field_device(vp_ctx* ctx, const vp_match* d_matches, const int32_t* d_counter, int n_frames, int max_blobs, float field_scale,
                                    float off_x, float off_y, float cell_mm, int cells_x, int cells_y, vp_field_match* d_out, int32_t* d_order,
                                    int32_t* d_cell_start);

/* debug-stream conversions straight from a raw frame / detection images (device pointers), Resources.cpp:166-186 */
VP_API int vp_raw2nv12_device(vp_ctx* ctx, const uint8_t* d_raw, int fmt, int wq, int hq, uint8_t* d_nv12, int sample_mode);
VP_API int vp_raw2rgba_device(vp_ctx* ctx, const uint8_t* d_raw, int fmt, int wq, int hq, uint8_t* d_rgba, int sample_mode);
VP_API int vp_rgba2nv12_device(vp_ctx* ctx, const uint8_t* d_rgba, int w, int h, uint8_t* d_nv12);
VP_API int vp_f2nv12_device(vp_ctx* ctx, const float* d_f32, int w, int h, uint8_t* d_nv12);
/* the same conversions for n dense frames in ONE launch (the four debug views of main.cpp:380-393 over a batch); NV12 frame
 * i is written at d_nv12 + i*nv12_stride, nv12_stride >= 1.5*w*h (the reference allocates 2*w*h, opencl.cpp:27).  The
 * result stays in device memory, where a hardware encoder session can take it (rtpstreamer.cpp:62: h264_nvenc first). */
VP_API int vp_rgba2nv12_batch_device(vp_ctx* ctx, const uint8_t* d_rgba, int n_frames, int w, int h, uint8_t* d_nv12, size_t nv12_stride);
VP_API int vp_f2nv12_batch_device(vp_ctx* ctx, const float* d_f32, int n_frames, int w, int h, uint8_t* d_nv12, size_t nv12_stride);
VP_API int vp_raw2nv12_batch_device(vp_ctx* ctx, const uint8_t* d_raw, int n_frames, int fmt, int wq, int hq, uint8_t* d_nv12, size_t nv12_stride,
                                    int sample_mode);

/* ---- NV12 hand-off to a hardware encoder -------------------------------------------------------------
 * What the reference's encoder thread does with a debug view (src/rtpstreamer.cpp:120-121, 177-181): it maps the NV12
 * RawImage and points libav at it with data[0] = buffer, data[1] = buffer + W*H and linesize[0] = linesize[1] = W; the
 * codec is h264_nvenc when available (rtpstreamer.cpp:62).  The batched conversions above leave exactly that layout in
 * DEVICE memory, one surface per frame, so an NVENC session (or libav with a CUDA hw frame) can take the planes without the
 * 1.9 MB device->host->device round trip per view; this descriptor states the contract instead of leaving it to pointer
 * arithmetic at the call site:
 *   y            luma plane: `height` rows of `width` bytes at pitch_y = width (dense, as libav's linesize expects)
 *   uv           chroma plane at y + width*height: height/2 rows of interleaved (U, V) byte pairs at pitch_uv = width
 *   width,height both even (Perspective.cpp:118-122 rounds the flat size up to even "for rtpstreamer"; quads are even by construction)
 *   bytes_used   width*height*3/2; the reference allocates width*height*2 per buffer (opencl.cpp:27, :132)
 *   aligned16    1 when y, uv and both pitches are multiples of 16 bytes (what NVENC's registered CUDA resources and
 *                cuMemcpy2D-style consumers want); true for every frame of a batch when width % 16 == 0 and
 *                nv12_stride % 16 == 0 and the base comes from cudaMalloc / vp_buf_alloc */
typedef struct {
	uint8_t* y;
	uint8_t* uv;
	int32_t width, height;
	int32_t pitch_y, pitch_uv;
	size_t bytes_used;
	int32_t aligned16;
} vp_nv12_surface;
/* surface of frame `frame` of a batch written by vp_*2nv12_batch_device(..., d_nv12, nv12_stride) (frame 0, any stride, for the
 * single-view calls).  Pure arithmetic, no device access; VP_ERR_INVALID for odd sizes or a stride below 1.5*w*h. */
VP_API int vp_nv12_surface_of(uint8_t* d_nv12, int w, int h, size_t nv12_stride, int frame, vp_nv12_surface* out);

/* blocking copies ordered after everything enqueued on the context stream (tests, tools) */
VP_API int vp_copy_to_host(vp_ctx* ctx, void* host, const void* dev, size_t bytes);
VP_API int vp_copy_to_device(vp_ctx* ctx, void* dev, const void* host, size_t bytes);

/* pinned host memory for frame sources (the camera drivers' user buffers, spinnakerdriver.cpp:120-133).  Frames inside
 * memory from vp_host_alloc (or inside a mapped vp_buf) are known to be pinned without a driver query; a frame the caller
 * pinned by other means is recognised too, at the price of one cudaPointerGetAttributes (~5 us) per one-frame call */
VP_API int vp_host_alloc(size_t bytes, void** out);
VP_API int vp_host_free(void* p);

/* tuning knob: frames per kernel launch group of the fused path (0 = automatic: at most 160 Mpx and 128 frames per launch, as
 * many groups as lanes -- or a multiple of it -- all of about the same size, whole quads of frames) */
VP_API int vp_ctx_set_group(vp_ctx* ctx, int frames_per_group);
/* tuning knob: number of CUDA streams the frame groups of one batch are spread over (1..4, default 3) so that the
 * issue-bound reprojection of one group overlaps the bandwidth-bound scans of another */
VP_API int vp_ctx_set_lanes(vp_ctx* ctx, int lanes);
/* A/B switch, results are bit-identical: 0 = direct-gather reprojection, 1 or 2 (default) = shared-memory staged kernels that
 * keep the frame-invariant weights of a tile in registers over a chunk of frames of the batch (four frames per shared-memory
 * word for chunks of at least four frames) */
VP_API int vp_ctx_set_staged_reproject(vp_ctx* ctx, int on);
/* tuning knob of variant 2: frames of a launch group one CTA processes with the same registers (0 = automatic: 32, fewer
 * when the grid would not fill the GPU) */
VP_API int vp_ctx_set_hoist_chunk(vp_ctx* ctx, int frames);
/* latency knob of vp_detect_host with ONE pinned frame (a camera delivering frame by frame, src/main.cpp:262-289): the
 * upload is cut into `strips` chunks of raw rows (1..16, default 2) and the reprojection of the flat rows a chunk
 * completes runs while the next chunk is still crossing PCIe; 1 = upload the frame, then compute.  Results are
 * bit-identical for every value. */
VP_API int vp_ctx_set_strips(vp_ctx* ctx, int strips);
/* A/B switch (default on), results are bit-identical: from the second one-frame call of vp_detect_host in an unchanged
 * configuration (parameters, buffers, knobs) the enqueued sequence -- chunked upload, strip kernels, download -- is
 * captured once into a CUDA graph and replayed with one launch per frame; frames at other pinned addresses (the
 * camera's buffer ring) only repoint the upload nodes.  Pageable frames take direct launches. */
VP_API int vp_ctx_set_latency_graph(vp_ctx* ctx, int on);
/* one-frame calls of vp_detect_host served by a graph replay so far (tests, tools) */
VP_API uint64_t vp_latency_graph_replays(const vp_ctx* ctx);
/* A/B switch (0 = never, 1 = default: calls of more than two frames, 2 = also the one- and two-frame calls of the latency path;
 * circle radius 1..12, gradient offset <= 4 and <= (radius+2)/2): gradientDot, the box sums of
 * satBlobCenter, circularity and peak classification in ONE kernel that reads the flat image once (no row sums, no SAT) vs
 * gradient + row prefix sums followed by the streaming circularity kernel; results are bit-identical */
VP_API int vp_ctx_set_fused_gradcirc(vp_ctx* ctx, int on);
/* what the most recent fused call (vp_detect_batch_device / vp_detect_host) actually launched, for tests and sweeps that
 * must know which specialised kernel they exercised: plan[0] reprojection kernel (0 direct gather, 1 staged per frame,
 * 2 frame-invariant one frame per word, 4 frame-invariant four frames per word), plan[1] frames per CTA of that kernel,
 * plan[2] frames per launch group, plan[3] streams used, plan[4] circularity flow (0 materialised SAT + unfused circle for
 * radii outside 1..12, 3 row sums + streaming circularity, 4 fused gradient + circularity), plan[5] rows per
 * circularity segment, plan[6], plan[7] reserved */
VP_API int vp_detect_last_plan(const vp_ctx* ctx, int32_t plan[8]);
/* how the geometry of `p` maps onto the staged reprojection (resampling.cl:52-99 on the fused path): stats[0] = 64x16 flat tiles
 * per frame, stats[1] = tiles whose raw footprint fits the staged planes (the rest fall back to the per-pixel gather: strong
 * distortion or tilt), stats[2] = tiles staged with 16-byte vectors, stats[3] = largest number of staged quad rows of a tile.
 * Builds (and caches) the geometry tables like a detection call would; blocks until they are there. */
VP_API int vp_tile_stats(vp_ctx* ctx, const vp_params* p, int32_t stats[4]);

/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
VP_API uint64_t vp_launch_count(const vp_ctx* ctx);
VP_API const char* vp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VP_B200_H */
