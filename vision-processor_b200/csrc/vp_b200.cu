/*
 * vp_b200.cu -- host side of libvp_b200.so: the C ABI declared in include/vp_b200.h.
 *
 * Replaces src/opencl.cpp of the reference (context, in-order queue, host-mappable buffers and
 * images, kernel launches) and the launch sequences of src/Resources.cpp:138-186 and
 * src/main.cpp:283-289.  Plain CUDA runtime; no torch types, no CPU fallback.
 */
#include "kernels.cuh"
#include "gradcirc.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace vpk;

namespace {

thread_local std::string t_last_error;

/* Pinned host ranges handed out by this library (vp_host_alloc, the host mirrors of vp_buf): a frame inside one of them is
 * known to be pinned without asking the driver -- cudaPointerGetAttributes costs several microseconds per call, which the
 * latency path cannot afford on every frame.  Anything else (memory the caller pinned itself) is asked. */
std::mutex g_pinned_mu;
std::vector<std::pair<uintptr_t, size_t>> g_pinned;

void pinned_register(const void* p, size_t bytes)
{
	std::lock_guard<std::mutex> l(g_pinned_mu);
	g_pinned.emplace_back((uintptr_t)p, bytes);
}
void pinned_unregister(const void* p)
{
	std::lock_guard<std::mutex> l(g_pinned_mu);
	for (size_t i = 0; i < g_pinned.size(); i++)
		if (g_pinned[i].first == (uintptr_t)p) {
			g_pinned[i] = g_pinned.back();
			g_pinned.pop_back();
			return;
		}
}
bool is_pinned(const void* host, size_t bytes)
{
	{
		std::lock_guard<std::mutex> l(g_pinned_mu);
		const uintptr_t a = (uintptr_t)host;
		for (const auto& r : g_pinned)
			if (a >= r.first && a + bytes <= r.first + r.second)
				return true;
	}
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost;
}

struct LutEntry {
	uint8_t key[104];
	float2* d = nullptr;
	TileEntry* tiles = nullptr; /* footprint of every 64x16 flat tile (k_tile_table) */
	uint64_t stamp = 0;
	std::vector<int> rows_needed; /* per tile row: raw Bayer rows [0, n) its tiles read (host copy, made on first use by the latency path) */
};

struct ProfEntry {
	const char* name;
	cudaEvent_t start, stop;
};

constexpr int HOST_SLOTS = 3;
constexpr int MAX_STRIPS = 16;

/* Latency path: the raw frame is uploaded in n chunks of rows and the row-local stages of the pipeline (reprojection,
 * gradient + row sums, circularity segments) follow strip by strip, so that only the last strip's work, the exactness
 * check and the record kernel remain once the upload has ended.  ty_end[k] = tile rows of the flat image whose raw rows
 * are on the device after chunk k; uploaded[k] is recorded on the copy stream behind chunk k. */
struct StripPlan {
	int n = 0;
	int raw_end[MAX_STRIPS]; /* chunk k = raw rows [raw_end[k-1], raw_end[k]) */
	int ty_end[MAX_STRIPS];
	cudaEvent_t uploaded[MAX_STRIPS];
};

struct HostSlot {
	uint8_t* raw = nullptr;
	uint8_t* flat = nullptr;
	float* grad = nullptr;
	float* circ = nullptr;
	/* one allocation, one download: [counter 3 x frames][flags frames][pad to 16][matches frames x max_blobs x 22 B] */
	uint8_t* results = nullptr;
	uint8_t* results_host = nullptr; /* pinned mirror (latency path) */
	size_t off_matches = 0;
	int32_t* counter = nullptr; /* = results */
	int* flags = nullptr;       /* = results + 12 * frames */
	vp_match* matches = nullptr; /* = results + off_matches */
	cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
};

/* Everything the enqueued work of a lone frame depends on besides the frame's address: when two consecutive calls agree
 * on all of it, the second one is captured into a CUDA graph and later calls replay it (one cudaGraphLaunch instead of
 * ~6 driver calls per strip: the CPU, not the GPU, is what limits a finely chunked frame otherwise). */
struct LoneFingerprint {
	vp_params params;
	const void* ptr[18];
	int knob[10];
};

struct LoneFrameGraph {
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	LoneFingerprint fp;       /* of the last direct call (have_fp) or of the captured graph (exec) */
	bool have_fp = false;
	const uint8_t* h_raw = nullptr; /* frame address the upload nodes currently point at */
	struct Upload { cudaGraphNode_t node; size_t off, bytes; };
	std::vector<Upload> uploads;
	bool deferred = false;
	uint64_t launches = 0;    /* kernels one replay launches */
	uint64_t replays = 0;
	void reset()
	{
		if (exec) cudaGraphExecDestroy(exec);
		if (graph) cudaGraphDestroy(graph);
		exec = nullptr;
		graph = nullptr;
		uploads.clear();
		h_raw = nullptr;
	}
};

} // namespace

struct vp_ctx {
	static constexpr int MAX_LANES_DECL = 4;
	int device = 0;
	cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
	std::mutex mu;
	std::string err;
	std::atomic<uint64_t> launches{ 0 };
	bool profiling = false;
	std::vector<ProfEntry> prof;

	std::vector<LutEntry> luts;
	uint64_t lut_clock = 0;

	/* scratch of the fused path: per lane (stream) `group` frames of row sums and SAT */
	static constexpr int MAX_LANES = 4;
	cudaStream_t lane_stream[MAX_LANES] = {}; /* [0] aliases `stream` */
	cudaEvent_t lane_done[MAX_LANES] = {};
	cudaEvent_t fork = nullptr;
	int lanes = 3;
	int32_t* rowsum[MAX_LANES] = {};
	float* sat[MAX_LANES] = {};
	size_t scratch_px = 0;
	int32_t* rowcount = nullptr;
	uint32_t* masks = nullptr; /* one bit per pixel: blob pixels of the current call */
	int32_t* first_slot = nullptr;
	int* flag = nullptr;
	size_t rows_cap = 0, frames_cap = 0, mask_words_cap = 0;
	int group = 0; /* 0 = choose from the frame size */
	int staged_reproject = 2; /* 0 direct gather, else shared-memory staged with the frame-invariant weights hoisted */
	int hoist_chunk = 0;      /* frames per CTA of the hoisted kernel; 0 = automatic */
	int sm_count = 148;
	int fused_gc = 1; /* gradient + circularity + classification in one kernel (gradcirc.cuh): 0 never, 1 for calls of more than two frames, 2 always */
	bool gc_attr = false;
	int32_t* striptot[MAX_LANES] = {}; /* per lane: k_grad_circ's per-row strip sums of gradDot (frames of the group x strips x rows) */
	size_t striptot_words = 0;
	float* segsum[MAX_LANES] = {}; /* per lane: column sums of the row sums per (frame of the group, row segment) */
	float* segmax[MAX_LANES] = {};
	size_t seg_words = 0;
	bool hoist_attr = false;
	bool hoist4_attr = false;
	volatile float one = 1.0f; /* handed to kernels that need a 1.0 the compiler cannot fold (add2_opaque) */
	int last_fallbacks = 0;
	int* flag_host = nullptr; /* pinned */
	int32_t last_plan[8] = {}; /* vp_detect_last_plan */

	cudaEvent_t strip_uploaded[MAX_STRIPS] = {};
	cudaEvent_t strip_flat[1] = {}; /* all strips reprojected */
	int strips = 2; /* chunks a lone frame's upload is cut into (vp_detect_host); 1 = upload, then compute */
	cudaEvent_t lone_fork = nullptr;
	bool latency_graph = true; /* replay the lone-frame sequence (chunked upload, strip kernels, download) as one CUDA graph */
	LoneFrameGraph lone;

	HostSlot slots[HOST_SLOTS];
	size_t slot_frames = 0, slot_raw = 0, slot_nf = 0, slot_blobs = 0;
	const uint8_t* last_flat = nullptr;
	const float* last_grad = nullptr;
	const float* last_circ = nullptr;
};

struct vp_buf {
	vp_ctx* ctx;
	void* d = nullptr;
	void* h = nullptr;
	size_t size = 0;
	std::atomic<int> refs{ 1 };
	int mapped = 0;
	std::mutex mu;
};

struct vp_img {
	vp_ctx* ctx;
	int fmt, w, h;
	vp_buf* buf;
	std::atomic<int> refs{ 1 };
};

namespace {

int fail(vp_ctx* ctx, int code, const char* fmt, ...)
{
	char msg[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(msg, sizeof msg, fmt, ap);
	va_end(ap);
	t_last_error = msg;
	if (ctx) {
		std::lock_guard<std::mutex> l(ctx->mu);
		ctx->err = msg;
	}
	return code;
}

#define CK(ctx, call)                                                                                                  \
	do {                                                                                                               \
		cudaError_t e_ = (call);                                                                                       \
		if (e_ != cudaSuccess)                                                                                         \
			return fail(ctx, e_ == cudaErrorMemoryAllocation ? VP_ERR_NOMEM : VP_ERR_CUDA, "%s: %s (%s:%d)", #call,   \
			            cudaGetErrorString(e_), __FILE__, __LINE__);                                                   \
	} while (0)

#define REQUIRE(ctx, cond, ...)                                                                                        \
	do {                                                                                                               \
		if (!(cond))                                                                                                   \
			return fail(ctx, VP_ERR_INVALID, __VA_ARGS__);                                                             \
	} while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

/* RAII-less stage scope: counts the launch and, when profiling, brackets it with events */
struct Stage {
	vp_ctx* c;
	int idx = -1;
	cudaStream_t s;
	Stage(vp_ctx* ctx, const char* name, int n_launches = 1, cudaStream_t on = nullptr): c(ctx), s(on ? on : ctx->stream)
	{
		c->launches += (uint64_t)n_launches;
		if (c->profiling) {
			if (c->prof.size() >= 4096) /* blob_benchmark never clears its events (blob_benchmark.cpp); keep it bounded */
				return;
			ProfEntry e{ name, nullptr, nullptr };
			cudaEventCreate(&e.start);
			cudaEventCreate(&e.stop);
			cudaEventRecord(e.start, s);
			c->prof.push_back(e);
			idx = (int)c->prof.size() - 1;
		}
	}
	~Stage()
	{
		if (idx >= 0)
			cudaEventRecord(c->prof[idx].stop, s);
	}
};

int check_launch(vp_ctx* ctx, const char* what)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return fail(ctx, VP_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
	return VP_OK;
}

bool is_raw_fmt(int fmt) { return fmt == VP_FMT_RGGB8 || fmt == VP_FMT_GRBG8 || fmt == VP_FMT_BGR8; }
bool is_mode(int m) { return m == VP_SAMPLE_BILINEAR_RTE || m == VP_SAMPLE_BILINEAR_TRUNC || m == VP_SAMPLE_NEAREST; }

/* ---- template dispatch over (format, sample mode) ------------------------------------------ */
#define VP_DISPATCH_MODE(FMT, mode, CALL)                                                                              \
	switch (mode) {                                                                                                    \
	case VP_SAMPLE_BILINEAR_RTE: { constexpr int MODE = MODE_RTE; constexpr int FMTC = FMT; CALL; } break;             \
	case VP_SAMPLE_BILINEAR_TRUNC: { constexpr int MODE = MODE_TRUNC; constexpr int FMTC = FMT; CALL; } break;         \
	default: { constexpr int MODE = MODE_NEAREST; constexpr int FMTC = FMT; CALL; } break;                             \
	}

/* coordinate table of a geometry (plus the tile footprints for a wq x hq source), cached per context: a handful of
 * geometries at most, one per camera */
int get_lut(vp_ctx* ctx, const vp_camera_model* m, float height, float scale, float offx, float offy, int wf, int hf, int wq, int hq,
            const float2** out, const TileEntry** tiles_out)
{
	uint8_t key[104];
	memset(key, 0, sizeof key);
	memcpy(key, m, 72);
	memcpy(key + 72, &height, 4);
	memcpy(key + 76, &scale, 4);
	memcpy(key + 80, &offx, 4);
	memcpy(key + 84, &offy, 4);
	memcpy(key + 88, &wf, 4);
	memcpy(key + 92, &hf, 4);
	memcpy(key + 96, &wq, 4);
	memcpy(key + 100, &hq, 4);
	for (LutEntry& e : ctx->luts)
		if (memcmp(e.key, key, sizeof key) == 0) {
			e.stamp = ++ctx->lut_clock;
			*out = e.d;
			if (tiles_out) *tiles_out = e.tiles;
			return VP_OK;
		}
	LutEntry* slot = nullptr;
	if (ctx->luts.size() < 8) {
		ctx->luts.emplace_back();
		slot = &ctx->luts.back();
	} else {
		slot = &ctx->luts[0];
		for (LutEntry& e : ctx->luts)
			if (e.stamp < slot->stamp)
				slot = &e;
		/* the evicted table may still be read by kernels in flight */
		CK(ctx, cudaDeviceSynchronize());
		cudaFree(slot->d);
		cudaFree(slot->tiles);
		slot->d = nullptr;
		slot->tiles = nullptr;
		memset(slot->key, 0xff, sizeof slot->key);
	}
	const int tiles_x = cdiv(wf, FT_W), tiles_y = cdiv(hf, FT_H);
	cudaError_t e = cudaMalloc(&slot->d, sizeof(float2) * (size_t)wf * hf);
	if (e == cudaSuccess)
		e = cudaMalloc(&slot->tiles, sizeof(TileEntry) * (size_t)tiles_x * tiles_y);
	if (e != cudaSuccess) {
		cudaFree(slot->d);
		slot->d = nullptr;
		slot->tiles = nullptr;
		memset(slot->key, 0xff, sizeof slot->key);
		return fail(ctx, VP_ERR_NOMEM, "coordinate table allocation failed: %s", cudaGetErrorString(e));
	}
	memcpy(slot->key, key, sizeof key);
	slot->stamp = ++ctx->lut_clock;
	slot->rows_needed.clear();
	{
		Stage st(ctx, "coord_table", 2);
		dim3 b(32, 8), g(cdiv(wf, 32), cdiv(hf, 8));
		k_coord_table<<<g, b, 0, ctx->stream>>>(slot->d, *m, height, scale, offx, offy, wf, hf);
		k_tile_table<<<dim3(tiles_x, tiles_y), 256, 0, ctx->stream>>>(slot->d, slot->tiles, wf, hf, wq, hq);
	}
	*out = slot->d;
	if (tiles_out) *tiles_out = slot->tiles;
	return check_launch(ctx, "k_coord_table/k_tile_table");
}

template <class Src>
int launch_reproject(vp_ctx* ctx, cudaStream_t stream, const Src& src, size_t frame_stride, int fmt, int mode, const float2* lut, uint32_t* flat, int wq,
                     int hq, int nf, int n_frames)
{
	const dim3 g(cdiv(nf, 256), n_frames);
#define VP_CALL k_reproject<FMTC, MODE, Src><<<g, 256, 0, stream>>>(src, frame_stride, lut, flat, wq, hq, nf)
	if (fmt == VP_FMT_RGGB8) { VP_DISPATCH_MODE(FMT_RGGB, mode, VP_CALL) }
	else if (fmt == VP_FMT_GRBG8) { VP_DISPATCH_MODE(FMT_GRBG, mode, VP_CALL) }
	else { VP_DISPATCH_MODE(FMT_BGR, mode, VP_CALL) }
#undef VP_CALL
	return check_launch(ctx, "k_reproject");
}

template <class Src>
int launch_quad2nv12(vp_ctx* ctx, const Src& src, int fmt, int mode, uint8_t* out, int wq, int hq, int n = 1, size_t src_stride = 0, size_t out_stride = 0)
{
	const dim3 g(cdiv(wq / 2, 256), hq / 2, n);
#define VP_CALL k_quad2nv12<FMTC, MODE, Src><<<g, 256, 0, ctx->stream>>>(src, out, wq, hq, src_stride, out_stride)
	if (fmt == VP_FMT_RGGB8) { VP_DISPATCH_MODE(FMT_RGGB, mode, VP_CALL) }
	else if (fmt == VP_FMT_GRBG8) { VP_DISPATCH_MODE(FMT_GRBG, mode, VP_CALL) }
	else { VP_DISPATCH_MODE(FMT_BGR, mode, VP_CALL) }
#undef VP_CALL
	return check_launch(ctx, "k_quad2nv12");
}

template <class Src>
int launch_quad2rgba(vp_ctx* ctx, const Src& src, int fmt, int mode, uint32_t* out, int wq, int hq)
{
	const dim3 g(cdiv(wq, 256), hq);
#define VP_CALL k_quad2rgba<FMTC, MODE, Src><<<g, 256, 0, ctx->stream>>>(src, out, wq, hq)
	if (fmt == VP_FMT_RGGB8) { VP_DISPATCH_MODE(FMT_RGGB, mode, VP_CALL) }
	else if (fmt == VP_FMT_GRBG8) { VP_DISPATCH_MODE(FMT_GRBG, mode, VP_CALL) }
	else { VP_DISPATCH_MODE(FMT_BGR, mode, VP_CALL) }
#undef VP_CALL
	return check_launch(ctx, "k_quad2rgba");
}

int launch_colscan(vp_ctx* ctx, cudaStream_t stream, const int32_t* rowsum, float* sat, int wf, int hf, int n_frames, int* flag)
{
	const int rpw = cdiv(hf, 32);
	const dim3 g(cdiv(wf, 32), n_frames);
	if (rpw <= 16)
		k_colscan<16><<<g, 1024, 0, stream>>>(rowsum, sat, wf, hf, rpw, flag);
	else if (rpw <= 32)
		k_colscan<32><<<g, 1024, 0, stream>>>(rowsum, sat, wf, hf, rpw, flag);
	else if (rpw <= 48)
		k_colscan<48><<<g, 1024, 0, stream>>>(rowsum, sat, wf, hf, rpw, flag);
	else
		return fail(ctx, VP_ERR_UNSUPPORTED, "flat image height %d exceeds 1536 rows", hf);
	return check_launch(ctx, "k_colscan");
}

/* blobList.cl:79 can only reject when minScore > 0 or circularities may be negative */
int need_score(float thr, float min_score) { return !(min_score <= 0.0f && thr >= 0.0f); }

int ensure_scratch(vp_ctx* ctx, size_t group_px, size_t rows, size_t frames, size_t mask_words)
{
	if (mask_words > ctx->mask_words_cap) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->masks);
		ctx->masks = nullptr;
		ctx->mask_words_cap = 0;
		CK(ctx, cudaMalloc(&ctx->masks, mask_words * 4));
		ctx->mask_words_cap = mask_words;
	}
	if (group_px > ctx->scratch_px) {
		CK(ctx, cudaDeviceSynchronize());
		for (int l = 0; l < vp_ctx::MAX_LANES; l++) {
			cudaFree(ctx->rowsum[l]);
			cudaFree(ctx->sat[l]);
			ctx->rowsum[l] = nullptr;
			ctx->sat[l] = nullptr;
		}
		ctx->scratch_px = 0;
		for (int l = 0; l < vp_ctx::MAX_LANES; l++) {
			CK(ctx, cudaMalloc(&ctx->rowsum[l], group_px * 4 + 256)); /* k_circ_stream_rs reads up to R-1 floats past a row end */
			CK(ctx, cudaMalloc(&ctx->sat[l], group_px * 4 + 256)); /* padded like the row sums */
		}
		ctx->scratch_px = group_px;
	}
	if (rows > ctx->rows_cap || frames > ctx->frames_cap) {
		const size_t r = rows > ctx->rows_cap ? rows : ctx->rows_cap, f = frames > ctx->frames_cap ? frames : ctx->frames_cap;
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->rowcount);
		cudaFree(ctx->first_slot);
		cudaFree(ctx->flag);
		if (ctx->flag_host)
			cudaFreeHost(ctx->flag_host);
		ctx->rowcount = ctx->first_slot = nullptr;
		ctx->flag = nullptr;
		ctx->flag_host = nullptr;
		ctx->rows_cap = ctx->frames_cap = 0;
		CK(ctx, cudaMalloc(&ctx->rowcount, r * 4));
		CK(ctx, cudaMalloc(&ctx->first_slot, f * 4));
		CK(ctx, cudaMalloc(&ctx->flag, f * 4));
		CK(ctx, cudaMallocHost(&ctx->flag_host, f * 4));
		ctx->rows_cap = r;
		ctx->frames_cap = f;
	}
	return VP_OK;
}

int validate_params(vp_ctx* ctx, const vp_params* p)
{
	REQUIRE(ctx, p, "params is null");
	REQUIRE(ctx, is_raw_fmt(p->fmt), "params.fmt %d is not a raw format", p->fmt);
	REQUIRE(ctx, p->wq > 0 && p->hq > 0 && p->wf > 0 && p->hf > 0, "non-positive image size");
	REQUIRE(ctx, (size_t)p->wf * p->hf < (1u << 30) && (size_t)p->wq * p->hq < (1u << 29), "image too large");
	REQUIRE(ctx, is_mode(p->sample_mode), "unknown sample mode %d", p->sample_mode);
	REQUIRE(ctx, p->grad_offset >= 0 && p->circle_radius >= 0 && p->blob_radius >= 0 && p->max_blobs >= 0, "negative radius/offset/max_blobs");
	return VP_OK;
}

size_t raw_frame_bytes(const vp_params* p) { return (size_t)p->wq * p->hq * (size_t)vp_format_pixel_size(p->fmt); }

int launch_peaks_emit(vp_ctx* ctx, cudaStream_t stream, const uint32_t* flat, const float* circ, int w, int h, int n, int radius, int max_matches,
                      const int32_t* first_slot, const int32_t* rowcount, const uint32_t* masks, uint8_t* matches, size_t match_stride,
                      const float* segsum = nullptr, const float* segmax = nullptr, int n_seg = 0, int* flag = nullptr, GcCheck gc = GcCheck())
{
	/* segsum given: more CTAs per frame check the exactness bound of the SAT next to the record warps (see k_peaks_emit):
	 * one per 256 columns for the row-sum flow, one per row segment for the fused gradient + circularity kernel.
	 * Rows per CTA: 8 for a few frames (every row at once), 64 for a batch (a grid of 8-row CTAs took longer to launch than to run) */
	const int rpc = n >= 16 ? 64 : 8;
	k_peaks_emit<<<dim3(cdiv(h, rpc) + (segsum ? (gc.striptot ? gc.n_seg : cdiv(w, 256)) : 0), n), 256, 0, stream>>>(flat, circ, w, h, radius, max_matches, first_slot, rowcount,
	                                                                                                      masks, cdiv(w, 32), matches, match_stride, segsum, segmax, n_seg,
	                                                                                                      flag, gc, rpc);
	return check_launch(ctx, "k_peaks_emit");
}

/* blob list of `n` frames from materialised images (stage API): count + emit */
int launch_blob_list(vp_ctx* ctx, const uint32_t* flat, const float* circ, int w, int h, int n, float thr, float min_score, int radius,
                     int max_matches, int32_t* counter, int32_t* first_slot, int32_t* rowcount, uint32_t* masks, uint8_t* matches,
                     size_t match_stride)
{
	const int ns = need_score(thr, min_score);
	k_peaks_count<<<dim3(cdiv(w, 256), h, n), 256, 0, ctx->stream>>>(flat, circ, w, h, thr, min_score, radius, ns, counter, rowcount, masks, cdiv(w, 32));
	int rc = check_launch(ctx, "k_peaks_count");
	if (rc)
		return rc;
	return launch_peaks_emit(ctx, ctx->stream, flat, circ, w, h, n, radius, max_matches, first_slot, rowcount, masks, matches, match_stride);
}

/* recompute the SAT of one frame in the reference's sequential order (one CTA of 1024 threads) */
__device__ __forceinline__ void sat_fix_frame(const float* __restrict__ grad, float* __restrict__ hor, float* __restrict__ sat, int w, int h, size_t fbase)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int y = warp; y < h; y += 32) { /* satHorizontal.cl:26-31 */
		const size_t base = fbase + (size_t)y * w;
		float sum = 0.f;
		for (int x0 = 0; x0 < w; x0 += 32) {
			const int x = x0 + lane;
			const float val = x < w ? grad[base + x] : 0.f;
			float mine = 0.f;
			const int n = min(32, w - x0);
			for (int k = 0; k < n; k++) {
				sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, val, k));
				if (lane == k)
					mine = sum;
			}
			if (x < w)
				hor[base + x] = mine;
		}
	}
	__syncthreads();
	for (int x = threadIdx.x; x < w; x += 1024) { /* satVertical.cl:26-31 */
		float sum = 0.f;
		for (int y = 0; y < h; y++) {
			sum = __fadd_rn(sum, hor[fbase + (size_t)y * w + x]);
			sat[fbase + (size_t)y * w + x] = sum;
		}
	}
}

/* single-launch fallback: recompute the SAT of flagged frames in the reference's sequential order */
__global__ void __launch_bounds__(1024) k_sat_fix(const float* __restrict__ grad, float* __restrict__ hor, float* __restrict__ sat, int w, int h,
                                                  const int* __restrict__ flag)
{
	if (flag[blockIdx.x] == 0)
		return;
	sat_fix_frame(grad, hor, sat, w, h, (size_t)blockIdx.x * w * h);
}

/* row-sum flow, after the fast pass, one CTA per frame: the exactness bound of the summed-area table from k_circ_stream_rs's
 * per-segment column sums of the row sums, |SAT(x, y)| <= |sum of the segments above| + max |running sum inside the segment|
 * (conservative); raises the flag of a frame that left it (the row scan raised it already if a row sum did) */
__global__ void __launch_bounds__(1024) k_sat_check_rs(const float* __restrict__ segsum, const float* __restrict__ segmax, int n_seg, int w, int* __restrict__ flag)
{
	const int f = blockIdx.x;
	if (flag[f] == 0 && sat_bound_exceeded(segsum, segmax, n_seg, w, f, threadIdx.x, 1024) && threadIdx.x == 0)
		flag[f] = 2;
}

/* Flow of the fused gradient + circularity kernel: the bound has been evaluated (k_sat_check_g / k_peaks_emit).  ONE launch
 * with one CTA per frame, which exits at once for a clean frame.  A flagged frame (never seen on camera images) forgets what the
 * fast pass published and is redone entirely here, in the reference's order: sequential summed-area table (satHorizontal.cl,
 * satVertical.cl), the literal 16-tap circularity with IEEE division (satBlobCenter.cl:37-41) and the peak classification
 * (blobList.cl:38-81).  Slow (about a millisecond for a 1224x1024 frame) and rare; what matters is that clean frames do not
 * pay for it -- the previous flow launched a full-size grid of the streaming kernel just to have every CTA exit. */
__global__ void __launch_bounds__(1024) k_fallback_frame(const uint32_t* __restrict__ flat, const float* __restrict__ grad, float* __restrict__ hor,
                                                         float* __restrict__ sat, float* __restrict__ circ, int w, int h, int r, float thr, float min_score,
                                                         int radius, int need_score, const int* __restrict__ flag, int32_t* __restrict__ counter,
                                                         int32_t* __restrict__ rowcount, uint32_t* __restrict__ masks, int wpr)
{
	const int f = blockIdx.x;
	if (flag[f] == 0)
		return;
	const size_t fbase = (size_t)f * w * h;
	for (int i = threadIdx.x; i < h * wpr; i += 1024)
		masks[(size_t)f * h * wpr + i] = 0u;
	for (int i = threadIdx.x; i < h; i += 1024)
		rowcount[(size_t)f * h + i] = 0;
	if (threadIdx.x < 3)
		counter[3 * f + threadIdx.x] = 0;
	sat_fix_frame(grad, hor, sat, w, h, fbase);
	__syncthreads();
	const float div = (float)(r * r);
	for (int i = threadIdx.x; i < w * h; i += 1024) {
		const int y = i / w, x = i - y * w;
		circ[fbase + i] = circle_px(sat + fbase, w, h, x, y, r, div);
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const float* cf = circ + fbase;
	int nb = 0, ns = 0, np = 0;
	for (int sg = warp; sg < h * wpr; sg += 32) { /* a warp = 32 consecutive pixels of a row */
		const int y = sg / wpr, x = (sg - y * wpr) * 32 + lane;
		int cls = 0;
		if (x < w) {
			const float c = __ldcg(cf + (size_t)y * w + x); /* written above by this CTA: not through the read-only path */
			if (!(c < thr)) {
				const float lf = __ldcg(cf + (size_t)y * w + max(x - 1, 0)), rt = __ldcg(cf + (size_t)y * w + min(x + 1, w - 1));
				const float up = __ldcg(cf + (size_t)max(y - 1, 0) * w + x), dn = __ldcg(cf + (size_t)min(y + 1, h - 1) * w + x);
				cls = classify_px(flat + fbase, w, h, x, y, radius, thr, min_score, need_score, c, lf, rt, up, dn);
			}
		}
		publish_segment(cls, lane, rowcount + (size_t)f * h, masks + (size_t)f * h * wpr, wpr, y, sg - y * wpr, nb, ns, np);
	}
	publish_counters(lane, counter + 3 * f, nb, ns, np);
}

/* one-time opt-in to the dynamic shared memory of the one-frame-per-word reprojection kernel (both call sites ask for exactly
 * HOIST_SMEM) */
int ensure_hoist_attr(vp_ctx* ctx)
{
	if (!ctx->hoist_attr) {
		CK(ctx, cudaFuncSetAttribute(k_reproject_hoist<FMT_RGGB, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HOIST_SMEM));
		CK(ctx, cudaFuncSetAttribute(k_reproject_hoist<FMT_GRBG, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HOIST_SMEM));
		ctx->hoist_attr = true;
	}
	return VP_OK;
}

/* rows per CTA of the streaming circularity kernels: 128 (few halo rows per segment) whenever that already gives the GPU two
 * CTAs per SM; fewer frames or smaller images take 64 or 32 rows and trade halo work for shorter dependent chains and a
 * full GPU (a lone 1224x1024 frame has only 14 x 8 CTAs to offer at 128 rows) */
int circ_seg_rows(const vp_ctx* ctx, int wf, int hf, int n_frames, int r, bool gc)
{
	static const int seg_env = getenv("VP_CIRC_SEG") ? atoi(getenv("VP_CIRC_SEG")) : 0; /* tuning aid */
	if (seg_env > 0)
		return seg_env;
	const int swu = gc ? grad_circ_strip_width(r) : 32 - (r + 2) - 1; /* output columns per warp, see k_grad_circ / k_circ_stream_rs */
	const long long per_row_of_segments = (long long)cdiv(cdiv(wf, swu > 0 ? swu : 1), 4) * n_frames;
	/* the fused gradient + circularity kernel recomputes 2(R + offset) rows and runs four predicated groups of rows per segment:
	 * with a whole batch to spread over the GPU (>= 8 CTAs per SM left) 256-row segments measured 2.5 % faster than 128
	 * (profiles/r02_sweeps.txt).  (A strip's running sum may wrap int32 at 256 rows on an all-saturated frame; such a frame is
	 * flagged by the column terms alone, see sat_bound_exceeded_g.) */
	if (gc && per_row_of_segments * cdiv(hf, 256) >= 8LL * ctx->sm_count)
		return 256;
	for (int seg = 128; seg > 32; seg >>= 1)
		if (per_row_of_segments * cdiv(hf, seg) >= 2LL * ctx->sm_count)
			return seg;
	return 32;
}

int choose_group(vp_ctx* ctx, size_t nf, int n_frames, int lanes)
{
	if (ctx->group > 0)
		return ctx->group < n_frames ? ctx->group : n_frames;
	/* Measured on B200 (profiles/r01_group_sweep.txt, r02_sweeps.txt): a 1.25 Mpx frame is ~4 pixels per resident thread, so kernels
	 * over a few frames are launch- and tail-bound and throughput rises with the group size even after the group's working set has
	 * left the 126 MB L2.  With the round-2 kernels: 8.42 us/frame at groups of 64 frames (80 Mpx), 8.20 at 96, 8.10 at 128, 8.00 at
	 * 256 -- every launch ends in a tail of partly filled SMs and has a ramp at its start, and a group twice as large has half as
	 * many of them per frame.  128 frames (160 Mpx) it is: the scratch of a lane grows with the group.  Three lanes (streams) with
	 * one group each in flight cover the launch gaps and tails of one group with the other groups' kernels. */
	size_t g = (size_t)160 * 1024 * 1024 / nf;
	if (g < 1) g = 1;
	if (g > 128) g = 128;
	/* as many groups as lanes (or a multiple of it when the cap says so), all of about the same size, whole quads of frames
	 * (the reprojection packs four frames into a shared-memory word): 64 frames of 4096x3000 on three lanes are groups of
	 * 24 + 24 + 16, not four groups of 16 of which one lane gets two */
	size_t n_groups = ((size_t)n_frames + g - 1) / g;
	if (n_groups < (size_t)lanes) n_groups = lanes;
	n_groups = (n_groups + lanes - 1) / lanes * lanes;
	g = ((size_t)n_frames + n_groups - 1) / n_groups;
	if (g >= 4) g = (g + 3) & ~(size_t)3;
	return (int)g;
}

} // namespace

/* =================================================================================================
 * exported C ABI
 * =============================================================================================== */
extern "C" {

const char* vp_version(void) { return "vp_b200 0.1 (sm_100a)"; }

int vp_format_pixel_size(int fmt)
{
	switch (fmt) { /* opencl.cpp:24-31: stride * rowStride */
	case VP_FMT_RGGB8:
	case VP_FMT_GRBG8: return 4;
	case VP_FMT_BGR8: return 3;
	case VP_FMT_RGBA8: return 4;
	case VP_FMT_U8: return 1;
	case VP_FMT_F32: return 4;
	case VP_FMT_NV12: return 2;
	default: return 0;
	}
}

int vp_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

const char* vp_last_error(const vp_ctx* ctx)
{
	(void)ctx;
	return t_last_error.c_str();
}

void vp_ctx_destroy(vp_ctx* c);

int vp_ctx_create(int device, vp_ctx** out)
{
	if (!out)
		return fail(nullptr, VP_ERR_INVALID, "out is null");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		return fail(nullptr, VP_ERR_NO_DEVICE, "no CUDA device: %s (there is no CPU fallback)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
	}
	if (device < 0 || device >= n)
		return fail(nullptr, VP_ERR_INVALID, "device ordinal %d out of range [0,%d)", device, n);
	CK(nullptr, cudaSetDevice(device));
	cudaDeviceProp prop;
	CK(nullptr, cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10)
		return fail(nullptr, VP_ERR_UNSUPPORTED, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", device, prop.name, prop.major, prop.minor);
	vp_ctx* c = new vp_ctx();
	c->device = device;
	c->sm_count = prop.multiProcessorCount;
	if (getenv("VP_HOIST_CHUNK")) c->hoist_chunk = atoi(getenv("VP_HOIST_CHUNK")); /* tuning aid */
	if (getenv("VP_STRIPS")) c->strips = std::max(1, std::min(MAX_STRIPS, atoi(getenv("VP_STRIPS")))); /* tuning aid */
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess
	    || cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
		delete c;
		return fail(nullptr, VP_ERR_CUDA, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
	}
	c->lane_stream[0] = c->stream;
	bool ok = cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) == cudaSuccess;
	for (int l = 0; l < vp_ctx::MAX_LANES && ok; l++) {
		if (l > 0)
			ok = ok && cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&c->lane_done[l], cudaEventDisableTiming) == cudaSuccess;
	}
	for (int k = 0; k < MAX_STRIPS && ok; k++)
		ok = ok && cudaEventCreateWithFlags(&c->strip_uploaded[k], cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->strip_flat[0], cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->lone_fork, cudaEventDisableTiming) == cudaSuccess;
	if (!ok) {
		vp_ctx_destroy(c);
		return fail(nullptr, VP_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
	}
	*out = c;
	return VP_OK;
}

static void free_slots(vp_ctx* c)
{
	for (HostSlot& s : c->slots) {
		cudaFree(s.raw); cudaFree(s.flat); cudaFree(s.grad); cudaFree(s.circ); cudaFree(s.results);
		if (s.results_host) cudaFreeHost(s.results_host);
		if (s.uploaded) cudaEventDestroy(s.uploaded);
		if (s.computed) cudaEventDestroy(s.computed);
		if (s.downloaded) cudaEventDestroy(s.downloaded);
		s = HostSlot();
	}
	c->slot_frames = c->slot_raw = c->slot_nf = c->slot_blobs = 0;
}

void vp_ctx_destroy(vp_ctx* c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	cudaStreamSynchronize(c->copy_in);
	cudaStreamSynchronize(c->copy_out);
	for (LutEntry& e : c->luts) {
		cudaFree(e.d);
		cudaFree(e.tiles);
	}
	for (ProfEntry& p : c->prof) {
		cudaEventDestroy(p.start);
		cudaEventDestroy(p.stop);
	}
	cudaFree(c->masks);
	for (int l = 0; l < vp_ctx::MAX_LANES; l++) {
		cudaFree(c->rowsum[l]);
		cudaFree(c->sat[l]);
		cudaFree(c->segsum[l]);
		cudaFree(c->segmax[l]);
		cudaFree(c->striptot[l]);
		if (l > 0 && c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]);
		if (c->lane_done[l]) cudaEventDestroy(c->lane_done[l]);
	}
	if (c->fork) cudaEventDestroy(c->fork);
	for (int k = 0; k < MAX_STRIPS; k++)
		if (c->strip_uploaded[k]) cudaEventDestroy(c->strip_uploaded[k]);
	if (c->strip_flat[0]) cudaEventDestroy(c->strip_flat[0]);
	if (c->lone_fork) cudaEventDestroy(c->lone_fork);
	cudaFree(c->rowcount); cudaFree(c->first_slot); cudaFree(c->flag);
	if (c->flag_host) cudaFreeHost(c->flag_host);
	c->lone.reset();
	free_slots(c);
	cudaStreamDestroy(c->stream);
	cudaStreamDestroy(c->copy_in);
	cudaStreamDestroy(c->copy_out);
	delete c;
}

int vp_ctx_sync(vp_ctx* ctx)
{
	REQUIRE(ctx, ctx, "ctx is null");
	CK(ctx, cudaSetDevice(ctx->device));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return VP_OK;
}

void* vp_ctx_stream(vp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int vp_tile_stats(vp_ctx* ctx, const vp_params* p, int32_t stats[4])
{
	REQUIRE(ctx, ctx && p && stats, "null argument");
	REQUIRE(ctx, p->wq > 0 && p->hq > 0 && p->wf > 0 && p->hf > 0, "empty geometry");
	const float2* lut;
	const TileEntry* tiles;
	int rc = get_lut(ctx, &p->model, p->max_robot_height, p->field_scale, p->off_x, p->off_y, p->wf, p->hf, p->wq, p->hq, &lut, &tiles);
	if (rc != VP_OK)
		return rc;
	const int n = cdiv(p->wf, FT_W) * cdiv(p->hf, FT_H);
	std::vector<TileEntry> host((size_t)n);
	CK(ctx, cudaMemcpyAsync(host.data(), tiles, host.size() * sizeof(TileEntry), cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	stats[0] = n;
	stats[1] = stats[2] = stats[3] = 0;
	for (const TileEntry& e : host) {
		stats[1] += e.flags & 1;
		stats[2] += (e.flags & 3) == 3;
		if ((e.flags & 1) && e.height > stats[3])
			stats[3] = e.height;
	}
	return VP_OK;
}

int vp_detect_last_plan(const vp_ctx* ctx, int32_t plan[8])
{
	if (!ctx || !plan)
		return fail(nullptr, VP_ERR_INVALID, "null argument");
	memcpy(plan, ctx->last_plan, sizeof ctx->last_plan);
	return VP_OK;
}
uint64_t vp_launch_count(const vp_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int vp_ctx_set_group(vp_ctx* ctx, int frames_per_group) /* tuning knob used by the benchmark sweep; 0 = automatic */
{
	REQUIRE(ctx, ctx && frames_per_group >= 0, "bad argument");
	ctx->group = frames_per_group;
	return VP_OK;
}

int vp_ctx_set_staged_reproject(vp_ctx* ctx, int on) /* A/B switch: shared-memory staged (frame-invariant part hoisted) vs direct-gather reprojection */
{
	REQUIRE(ctx, ctx, "ctx is null");
	REQUIRE(ctx, on >= 0 && on <= 2, "variant must be 0 (direct gather) or 1/2 (staged, frame-invariant part hoisted)");
	ctx->staged_reproject = on;
	return VP_OK;
}

int vp_ctx_set_fused_gradcirc(vp_ctx* ctx, int on) /* A/B switch: one gradient + circularity kernel vs row sums + streaming circularity */
{
	REQUIRE(ctx, ctx, "ctx is null");
	REQUIRE(ctx, on >= 0 && on <= 2, "0 = never, 1 = for calls of more than two frames (default), 2 = always");
	ctx->fused_gc = on;
	return VP_OK;
}

int vp_ctx_set_hoist_chunk(vp_ctx* ctx, int frames) /* frames one CTA of the hoisted reprojection keeps its weights for; 0 = automatic */
{
	REQUIRE(ctx, ctx && frames >= 0, "bad argument");
	ctx->hoist_chunk = frames;
	return VP_OK;
}

int vp_ctx_set_strips(vp_ctx* ctx, int strips) /* chunks the upload of a lone frame is cut into (latency path of vp_detect_host) */
{
	REQUIRE(ctx, ctx, "ctx is null");
	REQUIRE(ctx, strips >= 1 && strips <= MAX_STRIPS, "strips must be in [1, %d]", MAX_STRIPS);
	ctx->strips = strips;
	return VP_OK;
}

int vp_ctx_set_latency_graph(vp_ctx* ctx, int on) /* A/B switch: replay the lone-frame sequence of vp_detect_host as a CUDA graph (default on) */
{
	REQUIRE(ctx, ctx, "ctx is null");
	ctx->latency_graph = on != 0;
	return VP_OK;
}

uint64_t vp_latency_graph_replays(const vp_ctx* ctx) { return ctx ? ctx->lone.replays : 0; }

int vp_ctx_set_lanes(vp_ctx* ctx, int lanes) /* concurrent streams the groups of one batch are spread over */
{
	REQUIRE(ctx, ctx && lanes >= 1 && lanes <= vp_ctx::MAX_LANES, "lanes must be in [1,%d]", vp_ctx::MAX_LANES);
	ctx->lanes = lanes;
	return VP_OK;
}

/* ---- profiling -------------------------------------------------------------------------------- */
int vp_profiling_enable(vp_ctx* ctx, int on)
{
	REQUIRE(ctx, ctx, "ctx is null");
	ctx->profiling = on != 0;
	return VP_OK;
}
int vp_profiling_count(vp_ctx* ctx) { return ctx ? (int)ctx->prof.size() : 0; }
int vp_profiling_get(vp_ctx* ctx, int i, const char** name, float* ms)
{
	REQUIRE(ctx, ctx && i >= 0 && i < (int)ctx->prof.size(), "profiling index out of range");
	CK(ctx, cudaSetDevice(ctx->device));
	ProfEntry& p = ctx->prof[i];
	CK(ctx, cudaEventSynchronize(p.stop));
	float t = 0.f;
	CK(ctx, cudaEventElapsedTime(&t, p.start, p.stop));
	if (name) *name = p.name;
	if (ms) *ms = t;
	return VP_OK;
}
int vp_profiling_clear(vp_ctx* ctx)
{
	REQUIRE(ctx, ctx, "ctx is null");
	CK(ctx, cudaSetDevice(ctx->device));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	for (ProfEntry& p : ctx->prof) {
		cudaEventDestroy(p.start);
		cudaEventDestroy(p.stop);
	}
	ctx->prof.clear();
	return VP_OK;
}

/* ---- buffers ---------------------------------------------------------------------------------- */
int vp_buf_alloc(vp_ctx* ctx, size_t bytes, vp_buf** out)
{
	REQUIRE(ctx, ctx && out, "null argument");
	*out = nullptr;
	CK(ctx, cudaSetDevice(ctx->device));
	vp_buf* b = new vp_buf();
	b->ctx = ctx;
	b->size = bytes;
	const size_t alloc = bytes ? bytes : 1;
	cudaError_t e = cudaMalloc(&b->d, alloc);
	if (e == cudaSuccess)
		e = cudaMallocHost(&b->h, alloc);
	if (e != cudaSuccess) {
		cudaFree(b->d);
		delete b;
		return fail(ctx, VP_ERR_NOMEM, "buffer allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
	}
	pinned_register(b->h, alloc);
	*out = b;
	return VP_OK;
}

int vp_buf_alloc_copy(vp_ctx* ctx, const void* host, size_t bytes, vp_buf** out)
{
	REQUIRE(ctx, host || bytes == 0, "host pointer is null");
	int rc = vp_buf_alloc(ctx, bytes, out);
	if (rc)
		return rc;
	if (bytes) {
		memcpy((*out)->h, host, bytes);
		CK(ctx, cudaMemcpyAsync((*out)->d, (*out)->h, bytes, cudaMemcpyHostToDevice, ctx->stream));
		CK(ctx, cudaStreamSynchronize(ctx->stream));
	}
	return VP_OK;
}

int vp_buf_retain(vp_buf* b)
{
	if (!b) return fail(nullptr, VP_ERR_INVALID, "buf is null");
	b->refs++;
	return VP_OK;
}

int vp_buf_release(vp_buf* b)
{
	if (!b) return fail(nullptr, VP_ERR_INVALID, "buf is null");
	if (--b->refs > 0)
		return VP_OK;
	cudaSetDevice(b->ctx->device);
	cudaStreamSynchronize(b->ctx->stream);
	cudaFree(b->d);
	pinned_unregister(b->h);
	cudaFreeHost(b->h);
	delete b;
	return VP_OK;
}

int vp_buf_map(vp_buf* b, int mode, void** host)
{
	if (!b || !host) return fail(nullptr, VP_ERR_INVALID, "null argument");
	vp_ctx* ctx = b->ctx;
	REQUIRE(ctx, mode == VP_MAP_READ || mode == VP_MAP_WRITE || mode == VP_MAP_READWRITE, "bad map mode %d", mode);
	std::lock_guard<std::mutex> l(b->mu);
	REQUIRE(ctx, b->mapped == 0, "buffer is already mapped");
	CK(ctx, cudaSetDevice(ctx->device));
	if (mode & VP_MAP_READ) { /* blocking read of everything enqueued so far (in-order queue, opencl.h:118) */
		CK(ctx, cudaMemcpyAsync(b->h, b->d, b->size, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	b->mapped = mode;
	*host = b->h;
	return VP_OK;
}

int vp_buf_unmap(vp_buf* b)
{
	if (!b) return fail(nullptr, VP_ERR_INVALID, "buf is null");
	vp_ctx* ctx = b->ctx;
	std::lock_guard<std::mutex> l(b->mu);
	REQUIRE(ctx, b->mapped != 0, "buffer is not mapped");
	CK(ctx, cudaSetDevice(ctx->device));
	if (b->mapped & VP_MAP_WRITE) {
		CK(ctx, cudaMemcpyAsync(b->d, b->h, b->size, cudaMemcpyHostToDevice, ctx->stream));
		CK(ctx, cudaStreamSynchronize(ctx->stream)); /* unmap waits, opencl.h:128-132 */
	}
	b->mapped = 0;
	return VP_OK;
}

size_t vp_buf_size(const vp_buf* b) { return b ? b->size : 0; }
void* vp_buf_device_ptr(vp_buf* b) { return b ? b->d : nullptr; }

/* ---- images ----------------------------------------------------------------------------------- */
int vp_img_alloc(vp_ctx* ctx, int fmt, int w, int h, vp_img** out)
{
	REQUIRE(ctx, ctx && out, "null argument");
	*out = nullptr;
	REQUIRE(ctx, fmt == VP_FMT_RGBA8 || fmt == VP_FMT_U8 || fmt == VP_FMT_F32, "format %d is not an image format", fmt);
	REQUIRE(ctx, w >= 0 && h >= 0, "negative image size");
	vp_buf* b = nullptr;
	int rc = vp_buf_alloc(ctx, (size_t)w * h * vp_format_pixel_size(fmt), &b);
	if (rc)
		return rc;
	vp_img* i = new vp_img();
	i->ctx = ctx;
	i->fmt = fmt;
	i->w = w;
	i->h = h;
	i->buf = b;
	*out = i;
	return VP_OK;
}

int vp_img_retain(vp_img* i)
{
	if (!i) return fail(nullptr, VP_ERR_INVALID, "img is null");
	i->refs++;
	return VP_OK;
}

int vp_img_release(vp_img* i)
{
	if (!i) return fail(nullptr, VP_ERR_INVALID, "img is null");
	if (--i->refs > 0)
		return VP_OK;
	vp_buf_release(i->buf);
	delete i;
	return VP_OK;
}

int vp_img_map(vp_img* i, int mode, void** host, size_t* byte_pitch)
{
	if (!i) return fail(nullptr, VP_ERR_INVALID, "img is null");
	if (byte_pitch)
		*byte_pitch = (size_t)i->w * vp_format_pixel_size(i->fmt); /* dense: CLImage::save indexes x + width*y (opencl.cpp:164-168) */
	return vp_buf_map(i->buf, mode, host);
}

int vp_img_unmap(vp_img* i)
{
	if (!i) return fail(nullptr, VP_ERR_INVALID, "img is null");
	return vp_buf_unmap(i->buf);
}

int vp_img_info(const vp_img* i, int* fmt, int* w, int* h)
{
	if (!i) return fail(nullptr, VP_ERR_INVALID, "img is null");
	if (fmt) *fmt = i->fmt;
	if (w) *w = i->w;
	if (h) *h = i->h;
	return VP_OK;
}

void* vp_img_device_ptr(vp_img* i) { return i ? i->buf->d : nullptr; }

/* ---- stages ----------------------------------------------------------------------------------- */
#define IMG_IS(img, f) ((img) && (img)->fmt == (f))
#define SAME_SIZE(a, b) ((a)->w == (b)->w && (a)->h == (b)->h)

static int check_planes(vp_ctx* ctx, vp_img* const ch[4], int fmt, int* wq, int* hq)
{
	REQUIRE(ctx, ch, "channel array is null");
	const int n = fmt == VP_FMT_BGR8 ? 3 : 4;
	for (int c = 0; c < n; c++) {
		REQUIRE(ctx, IMG_IS(ch[c], VP_FMT_U8), "channel %d is not a U8 image", c);
		REQUIRE(ctx, SAME_SIZE(ch[c], ch[0]), "channel %d size differs from channel 0", c);
	}
	*wq = ch[0]->w;
	*hq = ch[0]->h;
	return VP_OK;
}

static SrcPlanes planes_of(vp_img* const ch[4], int fmt)
{
	SrcPlanes s;
	for (int c = 0; c < 4; c++)
		s.ch[c] = (c == 3 && (fmt == VP_FMT_BGR8 || !ch[3])) ? (const uint8_t*)ch[0]->buf->d : (const uint8_t*)ch[c]->buf->d;
	s.w = ch[0]->w;
	return s;
}

int vp_raw2quad(vp_ctx* ctx, const vp_buf* raw, int fmt, int wq, int hq, vp_img* const ch[4])
{
	REQUIRE(ctx, ctx && raw, "null argument");
	REQUIRE(ctx, is_raw_fmt(fmt), "format %d is not a raw format", fmt);
	int w2, h2;
	int rc = check_planes(ctx, ch, fmt, &w2, &h2);
	if (rc) return rc;
	REQUIRE(ctx, w2 == wq && h2 == hq && wq > 0 && hq > 0, "plane size %dx%d does not match %dx%d", w2, h2, wq, hq);
	REQUIRE(ctx, raw->size >= (size_t)wq * hq * vp_format_pixel_size(fmt), "raw buffer too small");
	CK(ctx, cudaSetDevice(ctx->device));
	if (raw->mapped & VP_MAP_WRITE) {
		/* A frame source that keeps its buffers mapped for the life of the driver and lets the camera SDK write into them
		 * (spinnakerdriver.cpp:120-133) never unmaps: take the host mirror as it is now, stream-ordered. */
		CK(ctx, cudaMemcpyAsync(raw->d, raw->h, raw->size, cudaMemcpyHostToDevice, ctx->stream));
	}
	Stage st(ctx, "raw2quad");
	if (fmt == VP_FMT_BGR8) {
		const int n = wq * hq;
		k_raw2quad_bgr<<<cdiv(n, 256), 256, 0, ctx->stream>>>((const uint8_t*)raw->d, (uint8_t*)ch[0]->buf->d, (uint8_t*)ch[1]->buf->d,
		                                                      (uint8_t*)ch[2]->buf->d, n);
	} else {
		k_raw2quad_bayer<<<dim3(cdiv(cdiv(wq, 4), 128), hq), 128, 0, ctx->stream>>>((const uint8_t*)raw->d, (uint8_t*)ch[0]->buf->d,
		                                                                            (uint8_t*)ch[1]->buf->d, (uint8_t*)ch[2]->buf->d,
		                                                                            (uint8_t*)ch[3]->buf->d, wq, hq);
	}
	return check_launch(ctx, "k_raw2quad");
}

int vp_resampling(vp_ctx* ctx, vp_img* const ch[4], int fmt, vp_img* flat, const vp_camera_model* model, float height, float scale,
                  float offx, float offy, int mode)
{
	REQUIRE(ctx, ctx && model, "null argument");
	REQUIRE(ctx, is_raw_fmt(fmt) && is_mode(mode), "bad format or sample mode");
	int wq, hq;
	int rc = check_planes(ctx, ch, fmt, &wq, &hq);
	if (rc) return rc;
	REQUIRE(ctx, IMG_IS(flat, VP_FMT_RGBA8), "flat is not an RGBA8 image");
	if (flat->w == 0 || flat->h == 0)
		return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	const float2* lut;
	rc = get_lut(ctx, model, height, scale, offx, offy, flat->w, flat->h, wq, hq, &lut, nullptr);
	if (rc) return rc;
	Stage st(ctx, "resampling");
	return launch_reproject(ctx, ctx->stream, planes_of(ch, fmt), 0, fmt, mode, lut, (uint32_t*)flat->buf->d, wq, hq, flat->w * flat->h, 1);
}

int vp_gradient_dot(vp_ctx* ctx, const vp_img* in, vp_img* out, int offset)
{
	REQUIRE(ctx, ctx && IMG_IS(in, VP_FMT_RGBA8) && IMG_IS(out, VP_FMT_F32) && SAME_SIZE(in, out), "gradient_dot needs RGBA8 in and F32 out of equal size");
	if (in->w == 0 || in->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "gradientDot");
	k_gradient_dot<<<dim3(cdiv(in->w, 64), cdiv(in->h, 4)), dim3(64, 4), 0, ctx->stream>>>((const uint32_t*)in->buf->d, (float*)out->buf->d, in->w, in->h, offset);
	return check_launch(ctx, "k_gradient_dot");
}

int vp_sat_horizontal(vp_ctx* ctx, const vp_img* in, vp_img* out)
{
	REQUIRE(ctx, ctx && IMG_IS(in, VP_FMT_F32) && IMG_IS(out, VP_FMT_F32) && SAME_SIZE(in, out), "sat_horizontal needs F32 images of equal size");
	if (in->w == 0 || in->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "satHorizontal");
	k_sat_h_seq<<<dim3(cdiv(in->h, 8), 1), 256, 0, ctx->stream>>>((const float*)in->buf->d, (float*)out->buf->d, in->w, in->h, nullptr);
	return check_launch(ctx, "k_sat_h_seq");
}

int vp_sat_vertical(vp_ctx* ctx, const vp_img* in, vp_img* out)
{
	REQUIRE(ctx, ctx && IMG_IS(in, VP_FMT_F32) && IMG_IS(out, VP_FMT_F32) && SAME_SIZE(in, out), "sat_vertical needs F32 images of equal size");
	if (in->w == 0 || in->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "satVertical");
	k_sat_v_seq<<<dim3(cdiv(in->w, 128), 1), 128, 0, ctx->stream>>>((const float*)in->buf->d, (float*)out->buf->d, in->w, in->h, nullptr);
	return check_launch(ctx, "k_sat_v_seq");
}

int vp_circle(vp_ctx* ctx, const vp_img* sat, vp_img* out, int radius)
{
	REQUIRE(ctx, ctx && IMG_IS(sat, VP_FMT_F32) && IMG_IS(out, VP_FMT_F32) && SAME_SIZE(sat, out), "circle needs F32 images of equal size");
	if (sat->w == 0 || sat->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "satBlobCenter");
	k_circle<<<dim3(cdiv(sat->w, 64), cdiv(sat->h, 4), 1), 256, 0, ctx->stream>>>((const float*)sat->buf->d, (float*)out->buf->d, sat->w, sat->h, radius);
	return check_launch(ctx, "k_circle");
}

int vp_blob_list(vp_ctx* ctx, const vp_img* rgba, const vp_img* circ, vp_buf* matches, vp_buf* counter, float thr, float min_score, int radius,
                 int max_matches)
{
	REQUIRE(ctx, ctx && IMG_IS(rgba, VP_FMT_RGBA8) && IMG_IS(circ, VP_FMT_F32) && SAME_SIZE(rgba, circ), "blob_list needs RGBA8 + F32 images of equal size");
	REQUIRE(ctx, matches && counter && counter->size >= 12, "matches/counter buffers missing or counter smaller than 3 ints");
	REQUIRE(ctx, radius >= 0 && max_matches >= 0 && matches->size >= (size_t)max_matches * 22, "matches buffer smaller than max_matches records");
	if (rgba->w == 0 || rgba->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_scratch(ctx, 0, rgba->h, 1, (size_t)rgba->h * cdiv(rgba->w, 32));
	if (rc) return rc;
	Stage st(ctx, "blobList", 3);
	k_peaks_prepare<<<cdiv(rgba->h, 256), 256, 0, ctx->stream>>>((int32_t*)counter->d, ctx->first_slot, ctx->rowcount, rgba->h, 1, 0, nullptr, ctx->masks,
	                                                             rgba->h * cdiv(rgba->w, 32));
	rc = check_launch(ctx, "k_peaks_prepare");
	if (rc) return rc;
	return launch_blob_list(ctx, (const uint32_t*)rgba->buf->d, (const float*)circ->buf->d, rgba->w, rgba->h, 1, thr, min_score, radius, max_matches,
	                        (int32_t*)counter->d, ctx->first_slot, ctx->rowcount, ctx->masks, (uint8_t*)matches->d, 0);
}

static int check_nv12(vp_ctx* ctx, int w, int h, const vp_buf* nv12)
{
	REQUIRE(ctx, nv12, "nv12 buffer is null");
	REQUIRE(ctx, (w % 2) == 0 && (h % 2) == 0, "NV12 needs even dimensions (Perspective.cpp:118-122), got %dx%d", w, h);
	REQUIRE(ctx, nv12->size >= (size_t)w * h * 3 / 2, "nv12 buffer smaller than 1.5*w*h");
	return VP_OK;
}

/* the wide NV12 kernels want 16-byte aligned source rows and 8-byte aligned destination rows */
static bool nv12_wide_ok(const void* in, const void* out, int w, int h, size_t out_stride)
{
	static const bool off = getenv("VP_NV12_WIDE") && atoi(getenv("VP_NV12_WIDE")) == 0; /* A/B aid: the one-thread-per-2x2-block kernels */
	return !off && (w % 8) == 0 && (h % 2) == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 8) == 0 && (out_stride % 8) == 0;
}

static int launch_rgba2nv12(vp_ctx* ctx, const uint8_t* d_rgba, int w, int h, uint8_t* d_nv12, int n, size_t out_stride)
{
	if (nv12_wide_ok(d_rgba, d_nv12, w, h, out_stride)) {
		const long long n_thr = (long long)(w / 8) * (h / 2) * n;
		k_rgba2nv12_wide<<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)d_rgba, d_nv12, w, h, (size_t)w * h, out_stride, n);
	} else
		k_rgba2nv12<<<dim3(cdiv(w / 2, 256), h / 2, n), 256, 0, ctx->stream>>>((const uint32_t*)d_rgba, d_nv12, w, h, (size_t)w * h, out_stride);
	return check_launch(ctx, "k_rgba2nv12");
}

static int launch_f2nv12(vp_ctx* ctx, const float* d_f32, int w, int h, uint8_t* d_nv12, int n, size_t out_stride)
{
	if (nv12_wide_ok(d_f32, d_nv12, w, h, out_stride)) {
		const long long n_thr = (long long)(w / 8) * (h / 2) * n;
		k_f2nv12_wide<<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>(d_f32, d_nv12, w, h, (size_t)w * h, out_stride, n);
	} else
		k_f2nv12<<<dim3(cdiv(w / 2, 256), h / 2, n), 256, 0, ctx->stream>>>(d_f32, d_nv12, w, h, (size_t)w * h, out_stride);
	return check_launch(ctx, "k_f2nv12");
}

int vp_rgba2nv12_device(vp_ctx* ctx, const uint8_t* d_rgba, int w, int h, uint8_t* d_nv12)
{
	REQUIRE(ctx, ctx && d_rgba && d_nv12, "null argument");
	REQUIRE(ctx, w >= 0 && h >= 0 && (w % 2) == 0 && (h % 2) == 0, "NV12 needs even dimensions, got %dx%d", w, h);
	if (w == 0 || h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "rgba2nv12");
	return launch_rgba2nv12(ctx, d_rgba, w, h, d_nv12, 1, 0);
}

int vp_f2nv12_device(vp_ctx* ctx, const float* d_f32, int w, int h, uint8_t* d_nv12)
{
	REQUIRE(ctx, ctx && d_f32 && d_nv12, "null argument");
	REQUIRE(ctx, w >= 0 && h >= 0 && (w % 2) == 0 && (h % 2) == 0, "NV12 needs even dimensions, got %dx%d", w, h);
	if (w == 0 || h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "f2nv12");
	return launch_f2nv12(ctx, d_f32, w, h, d_nv12, 1, 0);
}

int vp_rgba2nv12(vp_ctx* ctx, const vp_img* rgba, vp_buf* nv12)
{
	REQUIRE(ctx, ctx && IMG_IS(rgba, VP_FMT_RGBA8), "rgba2nv12 needs an RGBA8 image");
	int rc = check_nv12(ctx, rgba->w, rgba->h, nv12);
	if (rc) return rc;
	return vp_rgba2nv12_device(ctx, (const uint8_t*)rgba->buf->d, rgba->w, rgba->h, (uint8_t*)nv12->d);
}

int vp_f2nv12(vp_ctx* ctx, const vp_img* f32, vp_buf* nv12)
{
	REQUIRE(ctx, ctx && IMG_IS(f32, VP_FMT_F32), "f2nv12 needs an F32 image");
	int rc = check_nv12(ctx, f32->w, f32->h, nv12);
	if (rc) return rc;
	return vp_f2nv12_device(ctx, (const float*)f32->buf->d, f32->w, f32->h, (uint8_t*)nv12->d);
}

int vp_quad2nv12(vp_ctx* ctx, vp_img* const ch[4], int fmt, vp_buf* nv12, int mode)
{
	REQUIRE(ctx, ctx && is_raw_fmt(fmt) && is_mode(mode), "bad format or sample mode");
	int wq, hq;
	int rc = check_planes(ctx, ch, fmt, &wq, &hq);
	if (rc) return rc;
	rc = check_nv12(ctx, wq, hq, nv12);
	if (rc) return rc;
	if (wq == 0 || hq == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "quad2nv12");
	return launch_quad2nv12(ctx, planes_of(ch, fmt), fmt, mode, (uint8_t*)nv12->d, wq, hq);
}

int vp_quad2rgba(vp_ctx* ctx, vp_img* const ch[4], int fmt, vp_img* rgba, int mode)
{
	REQUIRE(ctx, ctx && is_raw_fmt(fmt) && is_mode(mode), "bad format or sample mode");
	int wq, hq;
	int rc = check_planes(ctx, ch, fmt, &wq, &hq);
	if (rc) return rc;
	REQUIRE(ctx, IMG_IS(rgba, VP_FMT_RGBA8) && rgba->w == wq && rgba->h == hq, "rgba image must be RGBA8 of the plane size");
	if (wq == 0 || hq == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "quad2rgba");
	return launch_quad2rgba(ctx, planes_of(ch, fmt), fmt, mode, (uint32_t*)rgba->buf->d, wq, hq);
}

static int raw_src_launch_nv12(vp_ctx* ctx, const uint8_t* d_raw, int fmt, int wq, int hq, uint8_t* out, int mode, int n = 1, size_t out_stride = 0)
{
	const size_t src_stride = (size_t)wq * hq * (size_t)vp_format_pixel_size(fmt); /* frames of a batch are dense */
	if (fmt == VP_FMT_BGR8) {
		SrcBGR s{ d_raw, wq };
		return launch_quad2nv12(ctx, s, fmt, mode, out, wq, hq, n, src_stride, out_stride);
	}
	if (mode == VP_SAMPLE_BILINEAR_RTE && nv12_wide_ok(d_raw, out, wq, hq, out_stride)) {
		/* default sampling, aligned rows: the integer 16-bit-lane kernel (bit-identical, see k_raw2nv12_wide) */
		const long long n_thr = (long long)(wq / 8) * (hq / 2) * n;
		if (fmt == VP_FMT_RGGB8)
			k_raw2nv12_wide<FMT_RGGB><<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>(d_raw, out, wq, hq, src_stride, out_stride, n);
		else
			k_raw2nv12_wide<FMT_GRBG><<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>(d_raw, out, wq, hq, src_stride, out_stride, n);
		return check_launch(ctx, "k_raw2nv12_wide");
	}
	SrcBayer s{ d_raw, 2 * wq };
	return launch_quad2nv12(ctx, s, fmt, mode, out, wq, hq, n, src_stride, out_stride);
}

/* ---- batched debug-stream conversions: n frames, one launch (the NV12 view stays in device memory, where
 * an NVENC session -- rtpstreamer.cpp:62 prefers h264_nvenc -- can take it without the 1.9 MB read-map per frame) ---- */
static int check_nv12_batch(vp_ctx* ctx, const void* in, const void* out, int n, int w, int h, size_t stride)
{
	REQUIRE(ctx, ctx && in && out, "null argument");
	REQUIRE(ctx, n >= 0 && w >= 0 && h >= 0 && (w % 2) == 0 && (h % 2) == 0, "NV12 needs even dimensions, got %dx%d (n = %d)", w, h, n);
	REQUIRE(ctx, n <= 65535, "at most 65535 frames per call");
	REQUIRE(ctx, stride >= (size_t)w * h * 3 / 2, "NV12 frame stride %zu smaller than 1.5*w*h", stride);
	return VP_OK;
}

int vp_rgba2nv12_batch_device(vp_ctx* ctx, const uint8_t* d_rgba, int n_frames, int w, int h, uint8_t* d_nv12, size_t nv12_stride)
{
	int rc = check_nv12_batch(ctx, d_rgba, d_nv12, n_frames, w, h, nv12_stride);
	if (rc) return rc;
	if (n_frames == 0 || w == 0 || h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "rgba2nv12");
	return launch_rgba2nv12(ctx, d_rgba, w, h, d_nv12, n_frames, nv12_stride);
}

int vp_f2nv12_batch_device(vp_ctx* ctx, const float* d_f32, int n_frames, int w, int h, uint8_t* d_nv12, size_t nv12_stride)
{
	int rc = check_nv12_batch(ctx, d_f32, d_nv12, n_frames, w, h, nv12_stride);
	if (rc) return rc;
	if (n_frames == 0 || w == 0 || h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "f2nv12");
	return launch_f2nv12(ctx, d_f32, w, h, d_nv12, n_frames, nv12_stride);
}

int vp_raw2nv12_batch_device(vp_ctx* ctx, const uint8_t* d_raw, int n_frames, int fmt, int wq, int hq, uint8_t* d_nv12, size_t nv12_stride, int mode)
{
	REQUIRE(ctx, ctx && is_raw_fmt(fmt) && is_mode(mode), "bad format or sample mode");
	int rc = check_nv12_batch(ctx, d_raw, d_nv12, n_frames, wq, hq, nv12_stride);
	if (rc) return rc;
	if (n_frames == 0 || wq == 0 || hq == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "raw2nv12");
	return raw_src_launch_nv12(ctx, d_raw, fmt, wq, hq, d_nv12, mode, n_frames, nv12_stride);
}

int vp_raw2nv12_device(vp_ctx* ctx, const uint8_t* d_raw, int fmt, int wq, int hq, uint8_t* d_nv12, int mode)
{
	REQUIRE(ctx, ctx && d_raw && d_nv12 && is_raw_fmt(fmt) && is_mode(mode), "bad argument");
	REQUIRE(ctx, wq >= 0 && hq >= 0 && (wq % 2) == 0 && (hq % 2) == 0, "NV12 needs even dimensions, got %dx%d", wq, hq);
	if (wq == 0 || hq == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "raw2nv12");
	return raw_src_launch_nv12(ctx, d_raw, fmt, wq, hq, d_nv12, mode);
}

int vp_raw2rgba_device(vp_ctx* ctx, const uint8_t* d_raw, int fmt, int wq, int hq, uint8_t* d_rgba, int mode)
{
	REQUIRE(ctx, ctx && d_raw && d_rgba && is_raw_fmt(fmt) && is_mode(mode), "bad argument");
	REQUIRE(ctx, wq >= 0 && hq >= 0, "negative size");
	if (wq == 0 || hq == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "raw2rgba");
	if (fmt == VP_FMT_BGR8) {
		SrcBGR s{ d_raw, wq };
		return launch_quad2rgba(ctx, s, fmt, mode, (uint32_t*)d_rgba, wq, hq);
	}
	if (mode == VP_SAMPLE_BILINEAR_RTE && nv12_wide_ok(d_raw, d_rgba, wq, 2, 0) && ((uintptr_t)d_rgba % 16) == 0) {
		/* default sampling, aligned rows: the integer 16-bit-lane demosaic (bit-identical, see k_raw2nv12_wide) */
		const long long n_thr = (long long)(wq / 8) * hq;
		if (fmt == VP_FMT_RGGB8)
			k_raw2rgba_wide<FMT_RGGB><<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>(d_raw, (uint32_t*)d_rgba, wq, hq);
		else
			k_raw2rgba_wide<FMT_GRBG><<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>(d_raw, (uint32_t*)d_rgba, wq, hq);
		return check_launch(ctx, "k_raw2rgba_wide");
	}
	SrcBayer s{ d_raw, 2 * wq };
	return launch_quad2rgba(ctx, s, fmt, mode, (uint32_t*)d_rgba, wq, hq);
}

int vp_circularize(vp_ctx* ctx, const vp_img* in, vp_img* out, int minr, int maxr)
{
	(void)minr; /* unused by blobCenter.cl as well */
	REQUIRE(ctx, ctx && IMG_IS(in, VP_FMT_F32) && IMG_IS(out, VP_FMT_F32) && SAME_SIZE(in, out), "circularize needs F32 images of equal size");
	if (in->w == 0 || in->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "blobCenter");
	k_circularize<<<dim3(cdiv(in->w, 64), cdiv(in->h, 4)), 256, 0, ctx->stream>>>((const float*)in->buf->d, (float*)out->buf->d, in->w, in->h, maxr);
	return check_launch(ctx, "k_circularize");
}

int vp_blob_score(vp_ctx* ctx, const vp_img* rgba, const vp_img* circ, vp_img* out, float thr, int radius)
{
	REQUIRE(ctx, ctx && IMG_IS(rgba, VP_FMT_RGBA8) && IMG_IS(circ, VP_FMT_F32) && IMG_IS(out, VP_FMT_F32) && SAME_SIZE(rgba, circ) && SAME_SIZE(rgba, out),
	        "blob_score needs RGBA8 + F32 in and F32 out of equal size");
	REQUIRE(ctx, radius >= 0, "negative radius");
	if (rgba->w == 0 || rgba->h == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "blobScore");
	k_blob_score<<<dim3(cdiv(rgba->w, 256), rgba->h), 256, 0, ctx->stream>>>((const uint32_t*)rgba->buf->d, (const float*)circ->buf->d, (float*)out->buf->d,
	                                                                         rgba->w, rgba->h, thr, radius);
	return check_launch(ctx, "k_blob_score");
}

/* ---- fused detection -------------------------------------------------------------------------- */
/* how vp_detect_host drives the fused path for a few frames (latency path) */
struct DetectOpts {
	const StripPlan* plan = nullptr; /* one frame, still arriving chunk by chunk on the copy stream */
	int* flags = nullptr;            /* exactness flags live here (next to the results the host downloads) instead of ctx->flag */
	bool defer_fallback = false;     /* the SAT bound is checked inside the record kernel and the HOST redoes flagged frames (redo_flagged) */
	mutable bool deferred = false;   /* out: the call really left the fallback to the host (one group, SAT-free flow) */
};
static int detect_batch_impl(vp_ctx* ctx, const uint8_t* d_raw, int n_frames, const vp_params* p, uint8_t* d_flat, float* d_grad, float* d_circ,
                             vp_match* d_matches, int32_t* d_counter, const DetectOpts& opts);

int vp_detect_batch_device(vp_ctx* ctx, const uint8_t* d_raw, int n_frames, const vp_params* p, uint8_t* d_flat, float* d_grad, float* d_circ,
                           vp_match* d_matches, int32_t* d_counter)
{
	return detect_batch_impl(ctx, d_raw, n_frames, p, d_flat, d_grad, d_circ, d_matches, d_counter, DetectOpts());
}

static int detect_batch_impl(vp_ctx* ctx, const uint8_t* d_raw, int n_frames, const vp_params* p, uint8_t* d_flat, float* d_grad, float* d_circ,
                             vp_match* d_matches, int32_t* d_counter, const DetectOpts& opts)
{
	const StripPlan* const plan = opts.plan;
	REQUIRE(ctx, ctx, "ctx is null");
	int rc = validate_params(ctx, p);
	if (rc) return rc;
	REQUIRE(ctx, n_frames >= 0, "negative frame count");
	if (n_frames == 0) return VP_OK;
	REQUIRE(ctx, d_raw && d_flat && d_grad && d_circ && d_counter && (d_matches || p->max_blobs == 0), "null device pointer");
	/* the fused kernels move raw rows, flat pixels, gradients and circularities as 16-byte vectors (cp.async / TMA boxes, uint4
	 * and float4 accesses): a misaligned base would fault inside a kernel, which is a sticky error that takes the context
	 * down, so it is refused here (cudaMalloc and vp_buf memory is 256-byte aligned; frame strides are multiples of 16
	 * whenever the vector paths are taken) */
	REQUIRE(ctx, ((((uintptr_t)d_raw) | ((uintptr_t)d_flat) | ((uintptr_t)d_grad) | ((uintptr_t)d_circ)) & 15u) == 0,
	        "d_raw, d_flat, d_grad and d_circ must be 16-byte aligned");
	REQUIRE(ctx, (((uintptr_t)d_counter) & 3u) == 0 && (((uintptr_t)d_matches) & 1u) == 0, "d_counter must be 4-byte and d_matches 2-byte aligned");
	CK(ctx, cudaSetDevice(ctx->device));
	const int wf = p->wf, hf = p->hf;
	const size_t nf = (size_t)wf * hf;
	const size_t raw_bytes = raw_frame_bytes(p);
	const int lanes_max = ctx->profiling ? 1 : ctx->lanes; /* per-stage event timing is only meaningful on one stream */
	const int G = choose_group(ctx, nf, n_frames, lanes_max);
	const int n_groups = cdiv(n_frames, G);
	const int lanes = n_groups < lanes_max ? n_groups : lanes_max;
	const int wpr = cdiv(wf, 32);
	rc = ensure_scratch(ctx, (size_t)G * nf, (size_t)n_frames * hf, n_frames, (size_t)n_frames * hf * wpr);
	if (rc) return rc;
	const bool fused_circ = p->circle_radius >= 1 && p->circle_radius <= CIRC_STREAM_MAX_R;
	/* three flows after the reprojection: the fused gradient + circularity kernel (default), gradient + row sums followed by the
	 * streaming circularity kernel (A/B switch; also gradient offsets the fused kernel does not stage), and -- for radii outside
	 * the specialised range -- a materialised summed-area table with the unfused circle / count kernels */
	/* one or two frames (the latency path of a camera delivering frame by frame) stay with the row-sum flow: its one-CTA-per-row
	 * gradient kernel and short circularity segments are one memory round trip deep, where the fused kernel walks ~50 rows per
	 * warp behind its TMA pipeline (p50 0.149 against 0.176 ms, profiles/r02_latency.txt); batches take the fused kernel */
	const bool use_gc = fused_circ && (ctx->fused_gc == 2 || (ctx->fused_gc == 1 && n_frames > 2)) && grad_circ_supported(p->circle_radius, p->grad_offset) &&
	                    wf <= 8192 && (wf & 1) == 0;
	const bool rowsums = fused_circ && !use_gc;
	const int seg = circ_seg_rows(ctx, wf, hf, n_frames, p->circle_radius, use_gc);
	const int n_seg = cdiv(hf, seg);
	if (use_gc && !ctx->gc_attr) {
		CK(ctx, (cudaError_t)grad_circ_prepare());
		ctx->gc_attr = true;
	}
	if (use_gc || rowsums) {
		const size_t need = (size_t)G * n_seg * wf;
		if (need > ctx->seg_words) {
			CK(ctx, cudaDeviceSynchronize());
			for (int l = 0; l < vp_ctx::MAX_LANES; l++) {
				cudaFree(ctx->segsum[l]);
				cudaFree(ctx->segmax[l]);
				ctx->segsum[l] = ctx->segmax[l] = nullptr;
			}
			ctx->seg_words = 0;
			for (int l = 0; l < vp_ctx::MAX_LANES; l++) {
				CK(ctx, cudaMalloc(&ctx->segsum[l], need * 4));
				CK(ctx, cudaMalloc(&ctx->segmax[l], need * 4));
			}
			ctx->seg_words = need;
		}
	}
	const int gc_strips = use_gc ? grad_circ_strips(p->circle_radius, wf) : 0;
	if (use_gc) {
		const size_t need_t = (size_t)G * gc_strips * hf;
		if (need_t > ctx->striptot_words) {
			CK(ctx, cudaDeviceSynchronize());
			for (int l = 0; l < vp_ctx::MAX_LANES; l++) {
				cudaFree(ctx->striptot[l]);
				ctx->striptot[l] = nullptr;
			}
			ctx->striptot_words = 0;
			for (int l = 0; l < vp_ctx::MAX_LANES; l++)
				CK(ctx, cudaMalloc(&ctx->striptot[l], need_t * 4));
			ctx->striptot_words = need_t;
		}
	}
	const float2* lut;
	const TileEntry* tiles;
	rc = get_lut(ctx, &p->model, p->max_robot_height, p->field_scale, p->off_x, p->off_y, wf, hf, p->wq, p->hq, &lut, &tiles);
	if (rc) return rc;
	const bool hoisted = ctx->staged_reproject != 0 && p->fmt != VP_FMT_BGR8 && p->sample_mode == VP_SAMPLE_BILINEAR_RTE;
	const int ns = need_score(p->circ_threshold, p->min_score);
	const bool by_strips = plan && plan->n > 1 && n_frames == 1 && hoisted && (use_gc || rowsums) && !ctx->profiling;
	int* const flags = opts.flags ? opts.flags : ctx->flag;
	const bool defer_fallback = opts.defer_fallback && (use_gc || rowsums) && n_groups == 1; /* only these flows have a check to move */
	opts.deferred = defer_fallback;
	int32_t* const plan_out = ctx->last_plan;
	plan_out[0] = hoisted ? 2 : 0;
	plan_out[1] = 1;
	plan_out[2] = G;
	plan_out[3] = lanes;
	plan_out[4] = use_gc ? 4 : rowsums ? 3 : 0;
	plan_out[5] = seg;
	plan_out[6] = plan_out[7] = 0;

	/* scratch of the compaction (blob masks, row counts, counters, flags) cleared for the whole batch up front -- or, when
	 * the batch runs as several groups, group by group at the head of each group's lane, so that clearing overlaps the other
	 * lanes' kernels instead of standing alone before the fork */
	const bool prepare_per_group = n_groups > 1;
	if (!prepare_per_group) {
		Stage st(ctx, "prepare");
		const int n = n_frames * hf;
		/* enough CTAs to clear the blob masks of a lone frame in one pass (the kernel strides over them) */
		const int prep_ctas = std::max(cdiv(n, 256), std::min(cdiv(n * wpr, 256), 4 * ctx->sm_count));
		k_peaks_prepare<<<prep_ctas, 256, 0, ctx->stream>>>(d_counter, ctx->first_slot, ctx->rowcount, n, n_frames, 1, flags, ctx->masks, n * wpr);
		if ((rc = check_launch(ctx, "k_peaks_prepare"))) return rc;
	}
	if (lanes > 1) { /* fork: the other lanes start after everything already enqueued on the context stream */
		CK(ctx, cudaEventRecord(ctx->fork, ctx->stream));
		for (int l = 1; l < lanes; l++)
			CK(ctx, cudaStreamWaitEvent(ctx->lane_stream[l], ctx->fork, 0));
	}
	if (plan && !by_strips) /* configuration without a strip form: wait for the whole frame */
		CK(ctx, cudaStreamWaitEvent(ctx->stream, plan->uploaded[plan->n - 1], 0));
	if (by_strips) {
		/* One frame whose raw rows are still arriving: after chunk k the tile rows < ty_end[k] can be reprojected, on a
		 * stream of their own, so that only the last strip's reprojection is left when the upload ends.  Same kernel,
		 * restricted to tile rows by offsetting its tables and images: every tile is computed exactly once from the same
		 * inputs.  Only the reprojection is cut up: it is the one stage of a lone frame that spans several waves of CTAs
		 * and therefore gets shorter with fewer rows.  The gradient/row-sum and circularity kernels of a single frame are
		 * single-wave and latency-bound -- a strip of them takes as long as the whole frame (measured: pipelining all three
		 * stages strip by strip made the frame slower, profiles/r01_latency.txt). */
		cudaStream_t sA = ctx->lane_stream[1];
		CK(ctx, cudaEventRecord(ctx->fork, ctx->stream)); /* orders the side stream behind earlier work (and is the fork point of a capture) */
		CK(ctx, cudaStreamWaitEvent(sA, ctx->fork, 0));
		const int tiles_x = cdiv(wf, FT_W), tiles_y = cdiv(hf, FT_H);
		uint32_t* flat = (uint32_t*)d_flat;
		if ((rc = ensure_hoist_attr(ctx))) return rc;
		int ty_done = 0;
		for (int k = 0; k < plan->n; k++) {
			CK(ctx, cudaStreamWaitEvent(sA, plan->uploaded[k], 0));
			const int ty_end = k == plan->n - 1 ? tiles_y : std::min(plan->ty_end[k], tiles_y);
			if (ty_end <= ty_done)
				continue;
			Stage st(ctx, "reproject", 1, sA);
			const size_t row0 = (size_t)ty_done * FT_H;
			const dim3 grid(tiles_x, ty_end - ty_done, 1);
			if (p->fmt == VP_FMT_RGGB8)
				k_reproject_hoist<FMT_RGGB, 4><<<grid, 256, HOIST_SMEM, sA>>>(d_raw, raw_bytes, lut + row0 * wf, tiles + (size_t)ty_done * tiles_x, flat + row0 * wf,
				                                                                p->wq, p->hq, wf, hf - (int)row0, 1, 1, ctx->one);
			else
				k_reproject_hoist<FMT_GRBG, 4><<<grid, 256, HOIST_SMEM, sA>>>(d_raw, raw_bytes, lut + row0 * wf, tiles + (size_t)ty_done * tiles_x, flat + row0 * wf,
				                                                                p->wq, p->hq, wf, hf - (int)row0, 1, 1, ctx->one);
			if ((rc = check_launch(ctx, "k_reproject_hoist (strip)"))) return rc;
			ty_done = ty_end;
		}
		CK(ctx, cudaEventRecord(ctx->strip_flat[0], sA));
		CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->strip_flat[0], 0)); /* join: the flat image is complete */
	}
	for (int gi = 0; gi < n_groups; gi++) {
		const int f0 = gi * G;
		const int g = n_frames - f0 < G ? n_frames - f0 : G;
		const int lane = gi % lanes;
		cudaStream_t s = ctx->lane_stream[lane];
		int32_t* rowsum = ctx->rowsum[lane];
		float* sat = ctx->sat[lane];
		uint32_t* flat = (uint32_t*)d_flat + (size_t)f0 * nf;
		float* grad = d_grad + (size_t)f0 * nf;
		float* circ = d_circ + (size_t)f0 * nf;
		const uint8_t* raw = d_raw + (size_t)f0 * raw_bytes;
		int* flag = flags + f0;
		int32_t* counter = d_counter + 3 * (size_t)f0;
		int32_t* rowcount = ctx->rowcount + (size_t)f0 * hf;
		uint32_t* masks = ctx->masks + (size_t)f0 * hf * wpr;
		if (prepare_per_group) {
			Stage st(ctx, "prepare", 1, s);
			const int n = g * hf;
			const int prep_ctas = std::max(cdiv(n, 256), std::min(cdiv(n * wpr, 256), 4 * ctx->sm_count));
			k_peaks_prepare<<<prep_ctas, 256, 0, s>>>(counter, ctx->first_slot + f0, rowcount, n, g, 1, flag, masks, n * wpr);
			if ((rc = check_launch(ctx, "k_peaks_prepare"))) return rc;
		}
		if (!by_strips) {
			Stage st(ctx, "reproject", 1, s);
			if (hoisted) {
				/* one CTA keeps a tile's weights in registers for `chunk` frames; enough CTAs to fill the GPU several times */
				int chunk = ctx->hoist_chunk > 0 ? ctx->hoist_chunk : 32; /* 32 at 64-frame groups: 10.46 against 10.56 us/frame at 16 (profiles/r01_group_sweep.txt) */
				const long long tiles_per_frame = (long long)cdiv(wf, FT_W) * cdiv(hf, FT_H);
				if (ctx->hoist_chunk <= 0) { /* an explicit vp_ctx_set_hoist_chunk is taken as it is (tests, sweeps) */
					while (chunk > 1 && tiles_per_frame * cdiv(g, chunk) < 8LL * 2 * ctx->sm_count) chunk >>= 1;
					/* chunks of about the same size, whole quads: a group of 44 frames is 24 + 20, not 32 + 12 */
					if (chunk >= 4 && g > chunk) chunk = (cdiv(g, cdiv(g, chunk)) + 3) & ~3;
				}
				const dim3 grid(cdiv(wf, FT_W), cdiv(hf, FT_H), cdiv(g, chunk));
				static const int hoist_quads = getenv("VP_HOIST_QUADS") ? atoi(getenv("VP_HOIST_QUADS")) : 1; /* tuning aid / A-B */
				plan_out[1] = chunk;
				if (hoist_quads && chunk >= 4) {
					plan_out[0] = 4;
					if (!ctx->hoist4_attr) {
						CK(ctx, cudaFuncSetAttribute(k_reproject_hoist4<FMT_RGGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HOIST4_SMEM));
						CK(ctx, cudaFuncSetAttribute(k_reproject_hoist4<FMT_GRBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HOIST4_SMEM));
						ctx->hoist4_attr = true;
					}
					if (p->fmt == VP_FMT_RGGB8)
						k_reproject_hoist4<FMT_RGGB><<<grid, 256, HOIST4_SMEM, s>>>(raw, raw_bytes, lut, tiles, flat, p->wq, p->hq, wf, hf, g, chunk);
					else
						k_reproject_hoist4<FMT_GRBG><<<grid, 256, HOIST4_SMEM, s>>>(raw, raw_bytes, lut, tiles, flat, p->wq, p->hq, wf, hf, g, chunk);
				} else {
					if ((rc = ensure_hoist_attr(ctx))) return rc;
					if (p->fmt == VP_FMT_RGGB8)
						k_reproject_hoist<FMT_RGGB, 4><<<grid, 256, HOIST_SMEM, s>>>(raw, raw_bytes, lut, tiles, flat, p->wq, p->hq, wf, hf, g, chunk, ctx->one);
					else
						k_reproject_hoist<FMT_GRBG, 4><<<grid, 256, HOIST_SMEM, s>>>(raw, raw_bytes, lut, tiles, flat, p->wq, p->hq, wf, hf, g, chunk, ctx->one);
				}
				rc = check_launch(ctx, "k_reproject_hoist");
			} else if (p->fmt == VP_FMT_BGR8) {
				SrcBGR src{ raw, p->wq };
				rc = launch_reproject(ctx, s, src, raw_bytes, p->fmt, p->sample_mode, lut, flat, p->wq, p->hq, (int)nf, g);
			} else {
				SrcBayer src{ raw, 2 * p->wq };
				rc = launch_reproject(ctx, s, src, raw_bytes, p->fmt, p->sample_mode, lut, flat, p->wq, p->hq, (int)nf, g);
			}
			if (rc) return rc;
		}
		if (use_gc) {
			/* gradientDot + windows + circularity + classification in one pass over the flat image (gradcirc.cuh) */
			Stage st(ctx, "grad_circ", 1, s);
			CK(ctx, (cudaError_t)launch_grad_circ(s, p->circle_radius, flat, grad, circ, wf, hf, p->grad_offset, seg, g, p->circ_threshold, p->min_score,
			                                      p->blob_radius, ns, counter, rowcount, masks, wpr, ctx->segsum[lane], ctx->segmax[lane], ctx->striptot[lane]));
		} else if (rowsums) {
			{
				Stage st(ctx, "grad_rowscan", 1, s);
				const bool wide = g <= 2 && cdiv(wf, 128) <= ROWWIDE_MAX_WARPS;
				if (wide) /* one or two frames: a CTA per row, the whole row in flight at once */
					k_grad_rowscan_wide<float><<<dim3(hf, g), cdiv(wf, 128) * 32, 0, s>>>(flat, grad, (float*)rowsum, wf, hf, p->grad_offset, flag);
				else
					k_grad_rowscan<float><<<dim3(cdiv(hf, ROWSCAN_WARPS), g), ROWSCAN_WARPS * 32, 0, s>>>(flat, grad, (float*)rowsum, wf, hf, p->grad_offset, flag);
				if ((rc = check_launch(ctx, "k_grad_rowscan"))) return rc;
			}
			Stage st(ctx, "circ_peaks", 1, s);
#define VP_CSR(RR)                                                                                                             \
	case RR: {                                                                                                                 \
		constexpr int SWU = 32 - (RR + 2) - 1;                                                                                 \
		const dim3 grid(cdiv(cdiv(wf, SWU), 4), n_seg, g);                                                                     \
		k_circ_stream_rs<RR><<<grid, 128, 0, s>>>((const float*)rowsum, circ, flat, wf, hf, seg, p->circ_threshold, p->min_score, p->blob_radius, ns, flag, \
		                                          counter, rowcount, masks, wpr, ctx->segsum[lane], ctx->segmax[lane]);         \
	} break;
			switch (p->circle_radius) {
				VP_CSR(1) VP_CSR(2) VP_CSR(3) VP_CSR(4) VP_CSR(5) VP_CSR(6) VP_CSR(7) VP_CSR(8) VP_CSR(9) VP_CSR(10) VP_CSR(11) VP_CSR(12)
			}
#undef VP_CSR
			if ((rc = check_launch(ctx, "k_circ_stream_rs"))) return rc;
		} else {
			/* radius outside the specialised range: materialised SAT (exact int32 scans; flagged frames in sequential order),
			 * unfused circle + count */
			{
				Stage st(ctx, "grad_rowscan", 1, s);
				k_grad_rowscan<int32_t><<<dim3(cdiv(hf, ROWSCAN_WARPS), g), ROWSCAN_WARPS * 32, 0, s>>>(flat, grad, rowsum, wf, hf, p->grad_offset, flag);
				if ((rc = check_launch(ctx, "k_grad_rowscan"))) return rc;
			}
			{
				Stage st(ctx, "colscan", 2, s);
				if ((rc = launch_colscan(ctx, s, rowsum, sat, wf, hf, g, flag))) return rc;
				k_sat_fix<<<g, 1024, 0, s>>>(grad, (float*)rowsum, sat, wf, hf, flag);
				if ((rc = check_launch(ctx, "k_sat_fix"))) return rc;
			}
			Stage st(ctx, "circle+count", 2, s);
			k_circle<<<dim3(cdiv(wf, 64), cdiv(hf, 4), g), 256, 0, s>>>(sat, circ, wf, hf, p->circle_radius);
			if ((rc = check_launch(ctx, "k_circle"))) return rc;
			k_peaks_count<<<dim3(cdiv(wf, 256), hf, g), 256, 0, s>>>(flat, circ, wf, hf, p->circ_threshold, p->min_score, p->blob_radius, ns, counter, rowcount,
			                                                        masks, wpr);
			if ((rc = check_launch(ctx, "k_peaks_count"))) return rc;
		}
		if ((use_gc || rowsums) && !defer_fallback) {
			/* the exactness bound of the summed-area table, checked after the fact; a frame that left it (or whose row sums did)
			 * is redone in the reference's sequential order by ONE more launch that exits at once for every other frame */
			Stage st(ctx, "sat_check", 2, s);
			if (use_gc)
				CK(ctx, (cudaError_t)launch_sat_check_g(s, grad_circ_check(p->circle_radius, ctx->segsum[lane], ctx->segmax[lane], ctx->striptot[lane], seg, wf, hf),
				                                        wf, hf, g, flag));
			else
				k_sat_check_rs<<<g, 1024, 0, s>>>(ctx->segsum[lane], ctx->segmax[lane], n_seg, wf, flag);
			k_fallback_frame<<<g, 1024, 0, s>>>(flat, grad, (float*)rowsum, sat, circ, wf, hf, p->circle_radius, p->circ_threshold, p->min_score, p->blob_radius,
			                                    ns, flag, counter, rowcount, masks, wpr);
			if ((rc = check_launch(ctx, "sat_check/fallback"))) return rc;
		}
		{
			Stage st(ctx, "peaks_emit", 1, s);
			if (defer_fallback) {
				GcCheck gc;
				if (use_gc)
					gc = grad_circ_check(p->circle_radius, ctx->segsum[lane], ctx->segmax[lane], ctx->striptot[lane], seg, wf, hf);
				rc = launch_peaks_emit(ctx, s, flat, circ, wf, hf, g, p->blob_radius, p->max_blobs, ctx->first_slot + f0, rowcount, masks,
				                       (uint8_t*)d_matches + (size_t)f0 * p->max_blobs * 22, (size_t)p->max_blobs * 22, ctx->segsum[lane], ctx->segmax[lane], n_seg,
				                       flag, gc);
			}
			else
				rc = launch_peaks_emit(ctx, s, flat, circ, wf, hf, g, p->blob_radius, p->max_blobs, ctx->first_slot + f0, rowcount, masks,
				                       (uint8_t*)d_matches + (size_t)f0 * p->max_blobs * 22, (size_t)p->max_blobs * 22);
			if (rc) return rc;
		}
	}
	for (int l = 1; l < lanes; l++) { /* join */
		CK(ctx, cudaEventRecord(ctx->lane_done[l], ctx->lane_stream[l]));
		CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_done[l], 0));
	}
	if (!opts.flags) {
		CK(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, ctx->stream));
		ctx->last_fallbacks = -n_frames; /* negative: flag_host holds n flags not summed yet */
	}
	return VP_OK;
}

/* Second half of a deferred fallback (latency path): the host has seen a raised flag among `n_frames` frames.  Those
 * frames are redone in the reference's sequential order -- forget what the fast pass published, rebuild the SAT, literal
 * circularity -- and the records of all frames are written again (idempotent for the clean ones). */
static int redo_flagged(vp_ctx* ctx, int n_frames, const vp_params* p, uint8_t* d_flat, float* d_grad, float* d_circ, vp_match* d_matches,
                        int32_t* d_counter, int* flags)
{
	const int wf = p->wf, hf = p->hf, wpr = cdiv(wf, 32);
	const int ns = need_score(p->circ_threshold, p->min_score);
	cudaStream_t s = ctx->stream;
	uint32_t* flat = (uint32_t*)d_flat;
	float* sat = ctx->sat[0];
	int rc;
	/* the flags are final: the bound was evaluated next to the record kernel */
	Stage st(ctx, "sat_check", 2, s);
	k_fallback_frame<<<n_frames, 1024, 0, s>>>(flat, d_grad, (float*)ctx->rowsum[0], sat, d_circ, wf, hf, p->circle_radius, p->circ_threshold, p->min_score,
	                                           p->blob_radius, ns, flags, d_counter, ctx->rowcount, ctx->masks, wpr);
	if ((rc = check_launch(ctx, "sat_check/fallback (redo)"))) return rc;
	return launch_peaks_emit(ctx, s, flat, d_circ, wf, hf, n_frames, p->blob_radius, p->max_blobs, ctx->first_slot, ctx->rowcount, ctx->masks,
	                         (uint8_t*)d_matches, (size_t)p->max_blobs * 22);
}

int vp_blobs_to_field_device(vp_ctx* ctx, const vp_match* d_matches, const int32_t* d_counter, int n_frames, int max_blobs, float field_scale, float off_x,
                             float off_y, float cell_mm, int cells_x, int cells_y, vp_field_match* d_out, int32_t* d_order, int32_t* d_cell_start)
{
	REQUIRE(ctx, ctx && d_counter && d_out && d_order && d_cell_start && (d_matches || max_blobs == 0), "null argument");
	REQUIRE(ctx, n_frames >= 0 && max_blobs >= 0 && max_blobs <= FB_MAX, "max_blobs must be in [0, %d]", FB_MAX);
	REQUIRE(ctx, cell_mm > 0.f && cells_x > 0 && cells_y > 0 && (long long)cells_x * cells_y <= (1 << 19), "bad grid (cell size must be positive, at most 2^19 cells)");
	if (n_frames == 0) return VP_OK;
	CK(ctx, cudaSetDevice(ctx->device));
	Stage st(ctx, "blobs2field");
	k_blobs_to_field<<<n_frames, 256, 0, ctx->stream>>>((const uint8_t*)d_matches, (size_t)max_blobs * 22, d_counter, max_blobs, field_scale, off_x, off_y,
	                                                    cell_mm, cells_x, cells_y, d_out, d_order, d_cell_start);
	return check_launch(ctx, "k_blobs_to_field");
}

int vp_detect_sat_fallbacks(vp_ctx* ctx, int* n)
{
	REQUIRE(ctx, ctx && n, "null argument");
	CK(ctx, cudaSetDevice(ctx->device));
	if (ctx->last_fallbacks < 0) {
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		int cnt = 0;
		for (int i = 0; i < -ctx->last_fallbacks; i++)
			cnt += ctx->flag_host[i] != 0;
		ctx->last_fallbacks = cnt;
	}
	*n = ctx->last_fallbacks;
	return VP_OK;
}

/* Upload schedule of a lone frame for this geometry: tile rows are handed out in `strips` shares, and chunk k ends
 * at the last raw row the tiles of share k read (known per geometry from the tile table, copied to the host once).  A
 * camera whose flat rows do not advance with the raw rows (rolled by 90 or 180 degrees) gets a plan whose first chunk is
 * most of the frame: still correct, just without the overlap.  plan->n = 0 when there is nothing to gain. */
static int make_strip_plan(vp_ctx* ctx, const vp_params* p, StripPlan* plan)
{
	plan->n = 0;
	if (ctx->staged_reproject != 2 || p->sample_mode != VP_SAMPLE_BILINEAR_RTE || ctx->profiling)
		return VP_OK;
	const int wf = p->wf, hf = p->hf, H = 2 * p->hq;
	const int tiles_x = cdiv(wf, FT_W), tiles_y = cdiv(hf, FT_H);
	const float2* lut;
	const TileEntry* tiles;
	int rc = get_lut(ctx, &p->model, p->max_robot_height, p->field_scale, p->off_x, p->off_y, wf, hf, p->wq, p->hq, &lut, &tiles);
	if (rc) return rc;
	LutEntry* entry = nullptr;
	for (LutEntry& e : ctx->luts)
		if (e.tiles == tiles)
			entry = &e;
	if (!entry)
		return VP_OK;
	if (entry->rows_needed.empty()) {
		std::vector<TileEntry> host((size_t)tiles_x * tiles_y);
		CK(ctx, cudaMemcpyAsync(host.data(), tiles, host.size() * sizeof(TileEntry), cudaMemcpyDeviceToHost, ctx->stream));
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		entry->rows_needed.assign(tiles_y, 0);
		for (int ty = 0; ty < tiles_y; ty++) {
			int need = 0;
			for (int tx = 0; tx < tiles_x; tx++) {
				const TileEntry& e = host[(size_t)ty * tiles_x + tx];
				int rows = H; /* direct gather: footprint not known */
				if (e.flags & 1) {
					const int last = std::min(p->hq - 1, std::max(0, e.jb + e.height - 1)); /* staged quad rows are clamped into the image */
					rows = 2 * last + 2;
				}
				need = std::max(need, rows);
			}
			entry->rows_needed[ty] = need;
		}
	}
	const int n = std::min(ctx->strips, tiles_y);
	if (n < 2)
		return VP_OK;
	/* the last strip is what remains to be reprojected when the upload has ended: one wave of CTAs (2 per SM), the
	 * shortest the kernel gets; the strips before it share the other tile rows evenly */
	static const int last_env = getenv("VP_STRIP_LAST") ? atoi(getenv("VP_STRIP_LAST")) : 0; /* tuning aid: tile rows of the last strip */
	int last_rows = last_env > 0 ? last_env : std::max(1, 2 * ctx->sm_count / tiles_x);
	if (last_rows > tiles_y / n)
		last_rows = tiles_y / n; /* never larger than an even share */
	const int head_rows = tiles_y - last_rows;
	int ty0 = 0, need = 0;
	for (int k = 0; k < n; k++) {
		const int ty1 = k == n - 1 ? tiles_y : (int)((long long)head_rows * (k + 1) / (n - 1));
		for (int ty = ty0; ty < ty1; ty++)
			need = std::max(need, entry->rows_needed[ty]);
		plan->ty_end[k] = ty1;
		plan->raw_end[k] = k == n - 1 ? H : std::min(need, H);
		plan->uploaded[k] = ctx->strip_uploaded[k];
		ty0 = ty1;
	}
	/* nothing to overlap when the first share already needs (almost) the whole frame: a camera rolled against the flat
	 * image's row order.  One upload, one launch then. */
	if ((long long)plan->raw_end[0] * 10 >= (long long)H * 9)
		return VP_OK;
	plan->n = n;
	return VP_OK;
}

static int ensure_slots(vp_ctx* ctx, size_t frames, size_t raw_bytes, size_t nf, size_t blobs)
{
	if (frames <= ctx->slot_frames && raw_bytes <= ctx->slot_raw && nf <= ctx->slot_nf && blobs <= ctx->slot_blobs)
		return VP_OK;
	/* every capacity only ever grows: alternating call shapes (few frames with many blobs, many frames with few) settle on
	 * the envelope after one reallocation each instead of reallocating -- three stream syncs, fifteen allocations and a
	 * dropped lone-frame graph -- on every call */
	frames = std::max(frames, ctx->slot_frames);
	raw_bytes = std::max(raw_bytes, ctx->slot_raw);
	nf = std::max(nf, ctx->slot_nf);
	blobs = std::max(blobs, ctx->slot_blobs);
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->copy_in));
	CK(ctx, cudaStreamSynchronize(ctx->copy_out));
	free_slots(ctx);
	for (HostSlot& s : ctx->slots) {
		CK(ctx, cudaMalloc(&s.raw, frames * raw_bytes));
		CK(ctx, cudaMalloc(&s.flat, frames * nf * 4));
		CK(ctx, cudaMalloc(&s.grad, frames * nf * 4));
		CK(ctx, cudaMalloc(&s.circ, frames * nf * 4));
		s.off_matches = (frames * 16 + 15) / 16 * 16;
		const size_t res_bytes = s.off_matches + frames * (blobs ? blobs : 1) * 22;
		CK(ctx, cudaMalloc(&s.results, res_bytes));
		CK(ctx, cudaMallocHost(&s.results_host, res_bytes));
		s.counter = (int32_t*)s.results;
		s.flags = (int*)(s.results + frames * 12);
		s.matches = (vp_match*)(s.results + s.off_matches);
		CK(ctx, cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
		CK(ctx, cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming));
		CK(ctx, cudaEventCreateWithFlags(&s.downloaded, cudaEventDisableTiming));
	}
	ctx->slot_frames = frames;
	ctx->slot_raw = raw_bytes;
	ctx->slot_nf = nf;
	ctx->slot_blobs = blobs;
	return VP_OK;
}

/* upload (whole or in the plan's chunks on the copy stream), fused path, one download of results + counters + flags;
 * nothing here blocks.  `forked`: the copy stream has to be ordered behind the compute stream first (stream capture). */
static int enqueue_lone_frames(vp_ctx* ctx, HostSlot& s, const uint8_t* h_raw, int n_frames, const vp_params* p, const DetectOpts& opts, size_t raw_bytes,
                               size_t res_bytes, bool forked)
{
	if (opts.plan) {
		const StripPlan& plan = *opts.plan;
		const size_t row_bytes = raw_bytes / (size_t)(2 * p->hq);
		if (forked) {
			CK(ctx, cudaEventRecord(ctx->lone_fork, ctx->stream));
			CK(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->lone_fork, 0));
		}
		int done = 0;
		for (int k = 0; k < plan.n; k++) {
			const int end = plan.raw_end[k];
			if (end > done)
				CK(ctx, cudaMemcpyAsync(s.raw + (size_t)done * row_bytes, h_raw + (size_t)done * row_bytes, (size_t)(end - done) * row_bytes,
				                        cudaMemcpyHostToDevice, ctx->copy_in));
			CK(ctx, cudaEventRecord(plan.uploaded[k], ctx->copy_in));
			done = std::max(done, end);
		}
	} else {
		/* everything in order on one stream, no cross-stream hops */
		CK(ctx, cudaMemcpyAsync(s.raw, h_raw, (size_t)n_frames * raw_bytes, cudaMemcpyHostToDevice, ctx->stream));
	}
	int rc = detect_batch_impl(ctx, s.raw, n_frames, p, s.flat, s.grad, s.circ, s.matches, s.counter, opts);
	if (rc) return rc;
	CK(ctx, cudaMemcpyAsync(s.results_host, s.results, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
	return VP_OK;
}

static void lone_fingerprint(vp_ctx* ctx, const HostSlot& s, const vp_params* p, LoneFingerprint* fp)
{
	memset(fp, 0, sizeof *fp);
	fp->params = *p;
	const float2* lut = nullptr;
	const TileEntry* tiles = nullptr;
	for (LutEntry& e : ctx->luts) { /* the table this geometry would use, if it exists already (no side effects here) */
		uint8_t key[104];
		memset(key, 0, sizeof key);
		memcpy(key, &p->model, 72);
		memcpy(key + 72, &p->max_robot_height, 4);
		memcpy(key + 76, &p->field_scale, 4);
		memcpy(key + 80, &p->off_x, 4);
		memcpy(key + 84, &p->off_y, 4);
		memcpy(key + 88, &p->wf, 4);
		memcpy(key + 92, &p->hf, 4);
		memcpy(key + 96, &p->wq, 4);
		memcpy(key + 100, &p->hq, 4);
		if (memcmp(e.key, key, sizeof key) == 0) {
			lut = e.d;
			tiles = e.tiles;
		}
	}
	const void* ptrs[18] = { s.raw, s.flat, s.grad, s.circ, s.results, s.results_host, ctx->rowsum[0], ctx->sat[0], ctx->segsum[0], ctx->segmax[0],
		                     ctx->rowcount, ctx->masks, ctx->first_slot, ctx->flag, lut, tiles, ctx->striptot[0], nullptr };
	memcpy(fp->ptr, ptrs, sizeof ptrs);
	const int knobs[10] = { ctx->staged_reproject, 0, 0, 0, ctx->hoist_chunk, ctx->group, ctx->lanes, ctx->strips,
		                    ctx->profiling, ctx->fused_gc };
	memcpy(fp->knob, knobs, sizeof knobs);
}

/* capture one lone-frame enqueue (copy stream forked from and joined back into the compute stream) and instantiate it;
 * on any failure the capture is abandoned and the caller carries on with direct launches */
static int capture_lone_frame(vp_ctx* ctx, HostSlot& s, const uint8_t* h_raw, const vp_params* p, const DetectOpts& opts, size_t raw_bytes, size_t res_bytes)
{
	LoneFrameGraph& lg = ctx->lone;
	lg.reset();
	const uint64_t launches0 = ctx->launches.load();
	if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
		cudaGetLastError();
		lg.have_fp = false;
		return VP_OK;
	}
	const int rc = enqueue_lone_frames(ctx, s, h_raw, 1, p, opts, raw_bytes, res_bytes, true);
	cudaGraph_t graph = nullptr;
	const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
	const uint64_t launched = ctx->launches.load() - launches0;
	ctx->launches -= launched; /* nothing ran */
	LoneFingerprint after;
	lone_fingerprint(ctx, s, p, &after);
	if (rc != VP_OK || e != cudaSuccess || !graph || memcmp(&after, &lg.fp, sizeof after) != 0) {
		cudaGetLastError();
		if (graph) cudaGraphDestroy(graph);
		lg.have_fp = false; /* do not try again until two direct calls agree once more */
		return VP_OK;
	}
	cudaGraphExec_t exec = nullptr;
	if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
		cudaGetLastError();
		cudaGraphDestroy(graph);
		lg.have_fp = false;
		return VP_OK;
	}
	size_t n_nodes = 0;
	cudaGraphGetNodes(graph, nullptr, &n_nodes);
	std::vector<cudaGraphNode_t> nodes(n_nodes);
	if (n_nodes)
		cudaGraphGetNodes(graph, nodes.data(), &n_nodes);
	for (cudaGraphNode_t nd : nodes) {
		cudaGraphNodeType t;
		if (cudaGraphNodeGetType(nd, &t) != cudaSuccess || t != cudaGraphNodeTypeMemcpy)
			continue;
		cudaMemcpy3DParms mp;
		if (cudaGraphMemcpyNodeGetParams(nd, &mp) != cudaSuccess)
			continue;
		const uint8_t* src = (const uint8_t*)mp.srcPtr.ptr;
		if (src >= h_raw && src < h_raw + raw_bytes)
			lg.uploads.push_back({ nd, (size_t)(src - h_raw), mp.extent.width });
	}
	cudaGetLastError();
	size_t covered = 0;
	for (const LoneFrameGraph::Upload& u : lg.uploads)
		covered += u.bytes;
	if (covered != raw_bytes) { /* the upload nodes were not all recognised: this graph could not be repointed at another frame */
		cudaGraphExecDestroy(exec);
		cudaGraphDestroy(graph);
		lg.uploads.clear();
		lg.have_fp = false;
		return VP_OK;
	}
	lg.graph = graph;
	lg.exec = exec;
	lg.h_raw = h_raw;
	lg.deferred = opts.deferred;
	lg.launches = launched;
	return VP_OK;
}

int vp_detect_host(vp_ctx* ctx, const uint8_t* h_raw, int n_frames, const vp_params* p, vp_match* h_matches, int32_t* h_counter)
{
	REQUIRE(ctx, ctx, "ctx is null");
	int rc = validate_params(ctx, p);
	if (rc) return rc;
	REQUIRE(ctx, n_frames >= 0, "negative frame count");
	if (n_frames == 0) return VP_OK;
	REQUIRE(ctx, h_raw && h_counter && (h_matches || p->max_blobs == 0), "null host pointer");
	CK(ctx, cudaSetDevice(ctx->device));
	const size_t nf = (size_t)p->wf * p->hf, raw_bytes = raw_frame_bytes(p), blobs = (size_t)p->max_blobs;
	const int chunk = n_frames < 4 ? n_frames : 4; /* frames per upload: amortises launches, keeps latency low */
	rc = ensure_slots(ctx, (size_t)chunk, raw_bytes, nf, blobs);
	if (rc) return rc;
	if (n_frames <= chunk) {
		/* latency path (a camera delivering one frame at a time) */
		HostSlot& s = ctx->slots[0];
		StripPlan plan;
		const bool pinned = n_frames == 1 && is_pinned(h_raw, raw_bytes);
		if (pinned && ctx->strips > 1 && (p->fmt == VP_FMT_RGGB8 || p->fmt == VP_FMT_GRBG8)) {
			/* the upload (5 MB, ~100 us of PCIe) is the longest step of a lone frame: cut it into chunks of raw rows on the copy
			 * stream and reproject the flat rows a chunk completes while the next one is still on its way (a pageable frame
			 * is staged by the driver inside cudaMemcpyAsync: nothing to overlap with) */
			rc = make_strip_plan(ctx, p, &plan);
			if (rc) return rc;
		}
		/* results, counters and exactness flags come back in ONE download; the SAT bound is checked next to the record
		 * kernel and a flagged frame (never seen on camera images) costs a second round trip instead of two idle launches
		 * on every frame */
		DetectOpts opts;
		opts.flags = s.flags;
		opts.defer_fallback = true;
		if (plan.n > 1)
			opts.plan = &plan;
		const size_t res_bytes = s.off_matches + (size_t)n_frames * blobs * 22;

		LoneFrameGraph& lg = ctx->lone;
		bool replayed = false;
		const bool graph_ok = ctx->latency_graph && n_frames == 1 && !ctx->profiling;
		LoneFingerprint fp;
		if (graph_ok) {
			lone_fingerprint(ctx, s, p, &fp);
			if (lg.exec && memcmp(&fp, &lg.fp, sizeof fp) != 0)
				lg.reset();
			if (!lg.exec && lg.have_fp && memcmp(&fp, &lg.fp, sizeof fp) == 0 && pinned) {
				/* second call in this configuration: everything is allocated, so the same enqueue can be captured */
				rc = capture_lone_frame(ctx, s, h_raw, p, opts, raw_bytes, res_bytes);
				if (rc) return rc;
			}
			if (lg.exec && pinned) {
				bool ok = true;
				if (h_raw != lg.h_raw) { /* another buffer of the camera's ring: repoint the upload nodes */
					for (const LoneFrameGraph::Upload& u : lg.uploads)
						ok = ok && cudaGraphExecMemcpyNodeSetParams1D(lg.exec, u.node, s.raw + u.off, h_raw + u.off, u.bytes, cudaMemcpyHostToDevice) == cudaSuccess;
					lg.h_raw = h_raw;
				}
				if (ok) {
					CK(ctx, cudaGraphLaunch(lg.exec, ctx->stream));
					ctx->launches += lg.launches;
					lg.replays++;
					opts.deferred = lg.deferred;
					replayed = true;
				} else { /* could not be repointed: back to direct launches */
					cudaGetLastError();
					lg.reset();
					lg.have_fp = false;
				}
			}
		} else if (lg.exec || lg.have_fp) {
			lg.reset();
			lg.have_fp = false;
		}
		if (!replayed) {
			rc = enqueue_lone_frames(ctx, s, h_raw, n_frames, p, opts, raw_bytes, res_bytes, false);
			if (rc) return rc;
			if (graph_ok && !lg.exec) { /* fingerprint AFTER the call: its first-use allocations are part of it */
				lone_fingerprint(ctx, s, p, &lg.fp);
				lg.have_fp = true;
			}
		}
		ctx->last_flat = s.flat + (size_t)(n_frames - 1) * nf * 4;
		ctx->last_grad = s.grad + (size_t)(n_frames - 1) * nf;
		ctx->last_circ = s.circ + (size_t)(n_frames - 1) * nf;
		const int* const hflags = (const int*)(s.results_host + ((const uint8_t*)s.flags - s.results));
		for (int round = 0;; round++) {
			CK(ctx, cudaStreamSynchronize(ctx->stream));
			int flagged = 0;
			for (int i = 0; i < n_frames; i++)
				flagged += hflags[i] != 0;
			ctx->last_fallbacks = flagged;
			if (!flagged || !opts.deferred || round == 1)
				break;
			rc = redo_flagged(ctx, n_frames, p, s.flat, s.grad, s.circ, s.matches, s.counter, s.flags);
			if (rc) return rc;
			CK(ctx, cudaMemcpyAsync(s.results_host, s.results, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
		}
		memcpy(h_counter, s.results_host, (size_t)n_frames * 12);
		for (int i = 0; i < n_frames && blobs; i++) { /* the valid records only; the rest of the caller's array keeps its contents */
			const int32_t cnt = ((const int32_t*)s.results_host)[3 * i];
			const size_t n_rec = cnt < 0 ? 0 : std::min((size_t)cnt, blobs);
			memcpy((uint8_t*)h_matches + (size_t)i * blobs * 22, s.results_host + s.off_matches + (size_t)i * blobs * 22, n_rec * 22);
		}
		return VP_OK;
	}
	int k = 0;
	for (int f0 = 0; f0 < n_frames; f0 += chunk, k++) {
		const int g = n_frames - f0 < chunk ? n_frames - f0 : chunk;
		HostSlot& s = ctx->slots[k % HOST_SLOTS];
		if (k >= HOST_SLOTS) { /* the slot's previous results must have left before it is overwritten */
			CK(ctx, cudaStreamWaitEvent(ctx->copy_in, s.downloaded, 0));
		}
		CK(ctx, cudaMemcpyAsync(s.raw, h_raw + (size_t)f0 * raw_bytes, (size_t)g * raw_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
		CK(ctx, cudaEventRecord(s.uploaded, ctx->copy_in));
		CK(ctx, cudaStreamWaitEvent(ctx->stream, s.uploaded, 0));
		rc = vp_detect_batch_device(ctx, s.raw, g, p, s.flat, s.grad, s.circ, s.matches, s.counter);
		if (rc) return rc;
		CK(ctx, cudaEventRecord(s.computed, ctx->stream));
		CK(ctx, cudaStreamWaitEvent(ctx->copy_out, s.computed, 0));
		if (blobs)
			CK(ctx, cudaMemcpyAsync((uint8_t*)h_matches + (size_t)f0 * blobs * 22, s.matches, (size_t)g * blobs * 22, cudaMemcpyDeviceToHost, ctx->copy_out));
		CK(ctx, cudaMemcpyAsync(h_counter + 3 * (size_t)f0, s.counter, (size_t)g * 12, cudaMemcpyDeviceToHost, ctx->copy_out));
		CK(ctx, cudaEventRecord(s.downloaded, ctx->copy_out));
		ctx->last_flat = s.flat + (size_t)(g - 1) * nf * 4;
		ctx->last_grad = s.grad + (size_t)(g - 1) * nf;
		ctx->last_circ = s.circ + (size_t)(g - 1) * nf;
	}
	CK(ctx, cudaStreamSynchronize(ctx->copy_out));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return VP_OK;
}

int vp_detect_images(vp_ctx* ctx, const uint8_t** d_flat, const float** d_grad, const float** d_circ)
{
	REQUIRE(ctx, ctx && ctx->last_flat, "no vp_detect_host call has completed on this context");
	if (d_flat) *d_flat = ctx->last_flat;
	if (d_grad) *d_grad = ctx->last_grad;
	if (d_circ) *d_circ = ctx->last_circ;
	return VP_OK;
}

int vp_nv12_surface_of(uint8_t* d_nv12, int w, int h, size_t nv12_stride, int frame, vp_nv12_surface* out)
{
	if (!d_nv12 || !out || w <= 0 || h <= 0 || (w & 1) || (h & 1) || frame < 0)
		return fail(nullptr, VP_ERR_INVALID, "NV12 surface: null pointer, negative frame or odd size %dx%d", w, h);
	const size_t used = (size_t)w * h * 3 / 2;
	if (frame > 0 && nv12_stride < used)
		return fail(nullptr, VP_ERR_INVALID, "NV12 surface: stride %zu is smaller than a frame (%zu bytes)", nv12_stride, used);
	out->y = d_nv12 + (size_t)frame * nv12_stride;
	out->uv = out->y + (size_t)w * h; /* rtpstreamer.cpp:121: data[1] = buffer + width*height */
	out->width = w;
	out->height = h;
	out->pitch_y = out->pitch_uv = w; /* rtpstreamer.cpp:120: linesize = width for both planes */
	out->bytes_used = used;
	out->aligned16 = (((uintptr_t)out->y | (uintptr_t)out->uv | (uintptr_t)w) & 15u) == 0;
	return VP_OK;
}

int vp_copy_to_host(vp_ctx* ctx, void* host, const void* dev, size_t bytes)
{
	REQUIRE(ctx, ctx && (bytes == 0 || (host && dev)), "null argument");
	CK(ctx, cudaSetDevice(ctx->device));
	if (bytes)
		CK(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return VP_OK;
}

int vp_copy_to_device(vp_ctx* ctx, void* dev, const void* host, size_t bytes)
{
	REQUIRE(ctx, ctx && (bytes == 0 || (host && dev)), "null argument");
	CK(ctx, cudaSetDevice(ctx->device));
	if (bytes)
		CK(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	return VP_OK;
}

int vp_host_alloc(size_t bytes, void** out)
{
	if (!out) return fail(nullptr, VP_ERR_INVALID, "out is null");
	cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
	if (e != cudaSuccess)
		return fail(nullptr, VP_ERR_NOMEM, "pinned allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
	pinned_register(*out, bytes ? bytes : 1);
	return VP_OK;
}

int vp_host_free(void* p)
{
	if (p)
		pinned_unregister(p);
	if (p && cudaFreeHost(p) != cudaSuccess)
		return fail(nullptr, VP_ERR_CUDA, "cudaFreeHost failed");
	return VP_OK;
}

} /* extern "C" */
