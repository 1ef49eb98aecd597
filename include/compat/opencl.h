/*
 * compat/opencl.h -- drop-in replacement for src/opencl.h + src/opencl.cpp of
 * TIGERs-Mannheim/vision-processor on top of the C ABI of libvp_b200.so (include/vp_b200.h).
 *
 * Put include/compat first on the include path (so that `#include "opencl.h"` and
 * `#include "cl_kernels.h"` resolve here), drop src/opencl.cpp and the kernel .incbin step from
 * the build and link libvp_b200.so.  The call sites listed below then compile unchanged:
 *
 *   src/Resources.cpp:70,121-130,138-186   OpenCL(), compile(), acquire(), acquireNV12(), run(), await()
 *   src/main.cpp:253,257-258,283-317,372   blobList launch, CLArray, CLMap<int>/<CLMatch>, clearEvents()
 *   src/driver/ (all)                       RawImage ctors, ->write<uint8_t>(), persistent maps
 *   src/rtpstreamer.cpp:177-181             nv12->read<uint8_t>() from the encoder thread
 *   src/snapshotwriter.cpp:52-54            image->read<RGBA>() from the writer thread
 *   src/blob_benchmark.cpp:162,190-191      blobCenter->read<float>(), rowPitch arithmetic
 *
 * What is emulated of the OpenCL C++ bindings is only what those call sites touch: cl::Kernel (a stage id),
 * cl::NDRange, cl::EnqueueArgs, cl::Event, cl::Buffer, cl::Image2D.  `compile(kernel_<name>_cl, "-DRGGB")`
 * maps the embedded-source symbol (now a short tag, compat/cl_kernels.h) to a stage id; `run/await` forward
 * the type-erased arguments, in the kernel's declaration order, to the matching vp_* entry point.
 * Every failure is FATAL (log + exit(1)) like in the reference (src/log.h:21).
 *
 * Header-only; C++17.
 */
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
#include <algorithm>

#include "vp_b200.h"

#if __has_include("log.h")
#include "log.h"
#else
#define LOG(data) std::cout << "[" << __FILE__ << ":" << __LINE__ << "] " << data << std::endl
#define WARN(data) std::cerr << "[" << __FILE__ << ":" << __LINE__ << "] " << data << std::endl
#define FATAL(data) { WARN(data); exit(1); }
#endif

#if __has_include(<opencv2/core/mat.hpp>)
#include <opencv2/core/mat.hpp>
#define VP_COMPAT_HAVE_OPENCV 1
#else
#define VP_COMPAT_HAVE_OPENCV 0
#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32FC1 5
#endif
#endif

typedef int32_t cl_int;
typedef uint8_t cl_uchar;
#ifndef CL_SUCCESS
#define CL_SUCCESS 0
#define CL_MAP_READ 1
#define CL_MAP_WRITE 2
#define CL_MAP_WRITE_INVALIDATE_REGION 4
#endif

namespace vpcompat {

inline vp_ctx*& default_ctx()
{
	static vp_ctx* ctx = nullptr;
	return ctx;
}

inline void check(int rc, const char* what)
{
	if (rc != VP_OK)
		FATAL(what << " error " << rc << ": " << vp_last_error(default_ctx()));
}

enum Stage { NONE, RAW2QUAD, RESAMPLING, GRADIENT_DOT, SAT_HORIZONTAL, SAT_VERTICAL, SAT_BLOB_CENTER, BLOB_LIST, RGBA2NV12, F2NV12, QUAD2NV12, QUAD2RGBA,
             BLOB_CENTER, BLOB_SCORE };

struct Arg { /* one type-erased kernel argument */
	enum Kind { IMG, BUF, INT, FLOAT, BYTES } kind;
	vp_img* img = nullptr;
	vp_buf* buf = nullptr;
	int i = 0;
	float f = 0.f;
	unsigned char bytes[72] = {};
};

inline int map_mode(int cl_mode) { return cl_mode == CL_MAP_READ ? VP_MAP_READ : (cl_mode == CL_MAP_WRITE ? VP_MAP_READWRITE : VP_MAP_WRITE); }

} // namespace vpcompat

namespace cl {

typedef size_t size_type;

struct ImageFormat {
	int order, type;
};
#ifndef CL_RGBA
#define CL_RGBA 1
#define CL_R 2
#define CL_RGB 3
#define CL_UNSIGNED_INT8 8
#define CL_FLOAT 32
#endif

/* a device buffer: shared handle (RawImage copies share the underlying buffer, opencl.h:170) */
class Buffer {
public:
	Buffer() = default;
	explicit Buffer(vp_buf* b): h(b, [](vp_buf* p) { if (p) vp_buf_release(p); }) {}
	vp_buf* operator()() const { return h.get(); }
private:
	std::shared_ptr<vp_buf> h;
};

class Image2D {
public:
	Image2D() = default;
	explicit Image2D(vp_img* i): h(i, [](vp_img* p) { if (p) vp_img_release(p); }) {}
	vp_img* operator()() const { return h.get(); }
private:
	std::shared_ptr<vp_img> h;
};

class Kernel {
public:
	Kernel() = default;
	Kernel(vpcompat::Stage s, int fmt): stage(s), fmt(fmt) {}
	vpcompat::Stage stage = vpcompat::NONE;
	int fmt = VP_FMT_RGGB8; /* Bayer order selected by the "-D..." build option (PixelFormat::kernelOptions) */
};

class NDRange {
public:
	NDRange(size_type x = 0, size_type y = 1, size_type z = 1): size{ x, y, z } {}
	size_type size[3];
};

class Event { /* the queue is in-order: waiting for an event == waiting for the stream (Resources.cpp:159-163 chains are redundant) */
public:
	int wait() const { return vp_ctx_sync(vpcompat::default_ctx()) == VP_OK ? CL_SUCCESS : -1; }
	int index = -1; /* profiling slot, see OpenCL::printRuntimes */
};

class EnqueueArgs {
public:
	EnqueueArgs(NDRange global): global(global) {}
	EnqueueArgs(const Event&, NDRange global): global(global) {}
	NDRange global;
};

} // namespace cl


class PixelFormat { /* src/opencl.h:30-56, src/opencl.cpp:24-31 */
public:
	static const PixelFormat RGBA8, U8, F32, NV12, RGGB8, GRBG8, BGR8;

	[[nodiscard]] int pixelSize() const { return stride * rowStride; }

	const int stride;
	const int rowStride;
	const bool color;
	const int cvType;
	const cl::ImageFormat clFormat;
	const char* kernelOptions;
	const int vpFormat;

private:
	PixelFormat(int stride, int rowStride, bool color, int cvType, cl::ImageFormat clFormat, const char* kernelOptions, int vpFormat)
		: stride(stride), rowStride(rowStride), color(color), cvType(cvType), clFormat(clFormat), kernelOptions(kernelOptions), vpFormat(vpFormat) {}
};

inline const PixelFormat PixelFormat::RGBA8(4, 1, true, CV_8UC4, { CL_RGBA, CL_UNSIGNED_INT8 }, "", VP_FMT_RGBA8);
inline const PixelFormat PixelFormat::U8(1, 1, false, CV_8UC1, { CL_R, CL_UNSIGNED_INT8 }, "", VP_FMT_U8);
inline const PixelFormat PixelFormat::F32(4, 1, false, CV_32FC1, { CL_R, CL_FLOAT }, "", VP_FMT_F32);
inline const PixelFormat PixelFormat::NV12(1, 2, true, CV_8UC1, { CL_R, CL_UNSIGNED_INT8 }, "", VP_FMT_NV12);
inline const PixelFormat PixelFormat::RGGB8(2, 2, true, CV_8UC1, { CL_R, CL_UNSIGNED_INT8 }, "-DRGGB", VP_FMT_RGGB8);
inline const PixelFormat PixelFormat::GRBG8(2, 2, true, CV_8UC1, { CL_R, CL_UNSIGNED_INT8 }, "-DGRBG", VP_FMT_GRBG8);
inline const PixelFormat PixelFormat::BGR8(3, 1, true, CV_8UC3, { CL_RGB, CL_UNSIGNED_INT8 }, "-DBGR", VP_FMT_BGR8);


typedef struct __attribute__((packed)) RGBA {
	cl_uchar r, g, b, a;
} RGBA;

class CLImage;
class RawImage;


template<typename T>
class CLMap { /* src/opencl.h:115-152: blocking map, unmap (and write-back) on destruction */
public:
	explicit CLMap(const cl::Buffer& buffer, int size, int clRWType): buffer(buffer) {
		(void)size;
		void* p = nullptr;
		vpcompat::check(vp_buf_map(buffer(), vpcompat::map_mode(clRWType), &p), "Enqueue map buffer");
		map = (T*)p;
	}
	~CLMap() {
		if (unmoved)
			vpcompat::check(vp_buf_unmap(buffer()), "Enqueue unmap buffer");
	}
	CLMap(CLMap&& other) noexcept: buffer(other.buffer), map(other.map) { other.unmoved = false; }
	CLMap(const CLMap&) = delete;
	CLMap& operator=(const CLMap&) = delete;
	T*& operator*() { return map; }
	T* operator->() { return map; }
	T& operator[](int i) { return map[i]; }
	const T* const& operator*() const { return map; }
	const T* operator->() const { return map; }
	const T& operator[](int i) const { return map[i]; }

private:
	const cl::Buffer buffer;
	T* map;
	bool unmoved = true;
};

class CLArray { /* src/opencl.h:154-165, src/opencl.cpp:146-147 */
public:
	explicit CLArray(int size): buffer(alloc(nullptr, size)), size(size) {}
	CLArray(void* data, int size): buffer(alloc(data, size)), size(size) {}

	template<typename T> CLMap<T> read() const { return CLMap<T>(buffer, size, CL_MAP_READ); }
	template<typename T> CLMap<T> write() { return CLMap<T>(buffer, size, CL_MAP_WRITE_INVALIDATE_REGION); }
	template<typename T> CLMap<T> readWrite() { return CLMap<T>(buffer, size, CL_MAP_WRITE); }

	const cl::Buffer buffer;
	const int size;

private:
	static cl::Buffer alloc(void* data, int size) {
		vp_buf* b = nullptr;
		if (vpcompat::default_ctx() == nullptr)
			FATAL("Error during image allocation: no OpenCL() context exists yet");
		int rc = data ? vp_buf_alloc_copy(vpcompat::default_ctx(), data, (size_t)size, &b) : vp_buf_alloc(vpcompat::default_ctx(), (size_t)size, &b);
		if (rc != VP_OK)
			FATAL("Error during image allocation: " << rc << " " << vp_last_error(vpcompat::default_ctx()));
		return cl::Buffer(b);
	}
};

class RawImage : public CLArray { /* src/opencl.h:168-188: six constructors plus the sharing copy */
	struct Dims {
		const PixelFormat* fmt;
		int w, h;
		int bytes() const { return w * h * fmt->pixelSize(); }
	};
	RawImage(CLArray storage, Dims d, double ts, std::string label)
		: CLArray(std::move(storage)), format(d.fmt), width(d.w), height(d.h), timestamp(ts), name(std::move(label)) {}

public:
	RawImage(const RawImage& other) = default; /* shares the buffer */
	/* fresh storage */
	RawImage(const PixelFormat* format, int width, int height)
		: RawImage(CLArray(Dims{ format, width, height }.bytes()), Dims{ format, width, height }, 0.0, std::string()) {}
	RawImage(const PixelFormat* format, int width, int height, std::string name)
		: RawImage(CLArray(Dims{ format, width, height }.bytes()), Dims{ format, width, height }, 0.0, std::move(name)) {}
	RawImage(const PixelFormat* format, int width, int height, double timestamp)
		: RawImage(CLArray(Dims{ format, width, height }.bytes()), Dims{ format, width, height }, timestamp, std::string()) {}
	/* adopt existing storage */
	RawImage(CLArray array, const PixelFormat* format, int width, int height, std::string name)
		: RawImage(std::move(array), Dims{ format, width, height }, 0.0, std::move(name)) {}
	/* copy from host memory (drivers whose buffers cannot be registered, mvimpactdriver.cpp:24) */
	RawImage(const PixelFormat* format, int width, int height, unsigned char* data)
		: RawImage(CLArray(data, Dims{ format, width, height }.bytes()), Dims{ format, width, height }, 0.0, std::string()) {}
	RawImage(const PixelFormat* format, int width, int height, double timestamp, unsigned char* data)
		: RawImage(CLArray(data, Dims{ format, width, height }.bytes()), Dims{ format, width, height }, timestamp, std::string()) {}
	virtual ~RawImage() = default;

	const PixelFormat* format;
	const int width;
	const int height;
	double timestamp = 0; /* 0 = not available */
	const std::string name;
};

template<typename T> class CLImageMap;

class CLImage { /* src/opencl.h:195-212, src/opencl.cpp:149-158 */
public:
	explicit CLImage(const PixelFormat* format): format(format), width(0), height(0) {}
	CLImage(const PixelFormat* format, int width, int height, std::string name): image(alloc(format, width, height)), format(format), width(width), height(height), name(std::move(name)) {}

	template<typename T> CLImageMap<T> read() const { return CLImageMap<T>(*this, CL_MAP_READ); }
	template<typename T> CLImageMap<T> write() { return CLImageMap<T>(*this, CL_MAP_WRITE_INVALIDATE_REGION); }
	template<typename T> CLImageMap<T> readWrite() { return CLImageMap<T>(*this, CL_MAP_WRITE); }

	void save(const std::string& suffix, float factor = 1.0f, float offset = 0.0f) const; /* debug PNGs need OpenCV (opencl.cpp:161-179) */

	cl::Image2D image;
	const PixelFormat* format;
	int width;
	int height;
	std::string name;

private:
	static cl::Image2D alloc(const PixelFormat* format, int width, int height) {
		vp_img* i = nullptr;
		if (vpcompat::default_ctx() == nullptr)
			FATAL("Image creation error: no OpenCL() context exists yet");
		int rc = vp_img_alloc(vpcompat::default_ctx(), format->vpFormat, width, height, &i);
		if (rc != VP_OK)
			FATAL("Image creation error: " << rc << " " << width << "," << height << " " << (format == &PixelFormat::RGBA8));
		return cl::Image2D(i);
	}
};

template<typename T>
class CLImageMap { /* src/opencl.h:215-262; the pitch is always dense (CLImage::save indexes x + width*y) */
public:
	explicit CLImageMap(const CLImage& image, int clRWType): image(image.image) {
		void* p = nullptr;
		vpcompat::check(vp_img_map(image.image(), vpcompat::map_mode(clRWType), &p, &bytePitch), "Enqueue map image");
		map = (T*)p;
		rowPitch = bytePitch / sizeof(T);
#if VP_COMPAT_HAVE_OPENCV
		cv = ::cv::Mat(image.height, image.width, image.format->cvType, map, bytePitch);
#endif
	}
	~CLImageMap() {
		if (unmoved)
			vpcompat::check(vp_img_unmap(image()), "Enqueue unmap image");
	}
	CLImageMap(CLImageMap&& other) noexcept: bytePitch(other.bytePitch), rowPitch(other.rowPitch),
#if VP_COMPAT_HAVE_OPENCV
		cv(other.cv),
#endif
		image(other.image), map(other.map) { other.unmoved = false; }
	CLImageMap(const CLImageMap&) = delete;
	CLImageMap& operator=(const CLImageMap&) = delete;
	T*& operator*() { return map; }
	T* operator->() { return map; }
	T& operator[](int i) { return map[i]; }
	T& operator()(int x, int y) { return map[x + y * rowPitch]; }
	const T& operator()(int x, int y) const { return map[x + y * rowPitch]; }
	const T* const& operator*() const { return map; }
	const T* operator->() const { return map; }
	const T& operator[](int i) const { return map[i]; }

	size_t bytePitch;
	size_t rowPitch;
#if VP_COMPAT_HAVE_OPENCV
	cv::Mat cv;
#endif

private:
	const cl::Image2D image;
	T* map;
	bool unmoved = true;
};

inline void CLImage::save(const std::string& suffix, float factor, float offset) const {
#if VP_COMPAT_HAVE_OPENCV && __has_include(<opencv2/imgcodecs.hpp>)
	/* same conversions as src/opencl.cpp:161-179; compiled only where OpenCV exists */
	extern void vp_compat_save_image(const CLImage&, const std::string&, float, float);
	vp_compat_save_image(*this, suffix, factor, offset);
#else
	(void)factor; (void)offset;
	WARN("CLImage::save(" << name << suffix << ") skipped: built without OpenCV");
#endif
}


class OpenCL { /* src/opencl.h:69-112 */
public:
	OpenCL() {
		int rc = vp_ctx_create(0, &ctx);
		if (rc != VP_OK)
			FATAL("No GPU devices found. Check CUDA installation! (" << vp_last_error(nullptr) << ")");
		vpcompat::default_ctx() = ctx;
		vp_profiling_enable(ctx, 1); /* CL_QUEUE_PROFILING_ENABLE, opencl.cpp:48 */
		LOG("Using device: CUDA » B200 (libvp_b200 " << vp_version() << ")");
	}
	~OpenCL() {
		pool.clear();
		nv12pool.clear();
		if (vpcompat::default_ctx() == ctx)
			vpcompat::default_ctx() = nullptr;
		vp_ctx_destroy(ctx);
	}
	OpenCL(const OpenCL&) = delete;
	OpenCL& operator=(const OpenCL&) = delete;

	/* `code` is one of the tags of compat/cl_kernels.h; `options` carries the Bayer order ("-DRGGB" | "-DGRBG" | "-DBGR") */
	cl::Kernel compile(const char* code, const std::string& options = "") {
		static const std::pair<const char*, vpcompat::Stage> tags[] = {
			{ "vp_b200:raw2quad", vpcompat::RAW2QUAD }, { "vp_b200:resampling", vpcompat::RESAMPLING }, { "vp_b200:gradientDot", vpcompat::GRADIENT_DOT },
			{ "vp_b200:satHorizontal", vpcompat::SAT_HORIZONTAL }, { "vp_b200:satVertical", vpcompat::SAT_VERTICAL },
			{ "vp_b200:satBlobCenter", vpcompat::SAT_BLOB_CENTER }, { "vp_b200:blobList", vpcompat::BLOB_LIST }, { "vp_b200:rgba2nv12", vpcompat::RGBA2NV12 },
			{ "vp_b200:f2nv12", vpcompat::F2NV12 }, { "vp_b200:quad2nv12", vpcompat::QUAD2NV12 }, { "vp_b200:quad2rgba", vpcompat::QUAD2RGBA },
			{ "vp_b200:blobCenter", vpcompat::BLOB_CENTER }, { "vp_b200:blobScore", vpcompat::BLOB_SCORE },
		};
		int fmt = VP_FMT_RGGB8;
		if (options.find("-DGRBG") != std::string::npos) fmt = VP_FMT_GRBG8;
		else if (options.find("-DBGR") != std::string::npos) fmt = VP_FMT_BGR8;
		for (const auto& t : tags)
			if (std::strcmp(code, t.first) == 0)
				return cl::Kernel(t.second, fmt);
		FATAL("[OpenCL] Error during kernel compilation: libvp_b200 has no stage for this source (OpenCL C is not compiled): " << std::string(code).substr(0, 60));
	}

	template<typename... Ts>
	cl::Event run(cl::Kernel kernel, const cl::EnqueueArgs& args, Ts... ts) {
		std::vector<vpcompat::Arg> a;
		a.reserve(sizeof...(Ts));
		(a.push_back(erase(ts)), ...);
		cl::Event event;
		event.index = vp_profiling_count(ctx);
		int error = dispatch(kernel, args, a);
		if (error != VP_OK)
			FATAL("Enqueue kernel error: " << error << " " << vp_last_error(ctx));
		events.push_back(event);
		if (events.size() > 4096) /* blob_benchmark never clears its events */
			clearEvents();
		return event;
	}

	template<typename... Ts>
	void await(cl::Kernel kernel, const cl::EnqueueArgs& args, Ts... ts) {
		wait(run(kernel, args, std::forward<Ts>(ts)...));
	}

	static void wait(const cl::Event& event) {
		if (event.wait() != CL_SUCCESS)
			FATAL("Error during kernel execution: " << vp_last_error(vpcompat::default_ctx()));
	}

	void printRuntimes() { /* src/opencl.cpp:94-101 */
		std::cout << std::fixed;
		std::cout.precision(2);
		const int n = vp_profiling_count(ctx);
		for (int i = 0; i < n; i++) {
			float ms = 0.f;
			const char* name = nullptr;
			if (vp_profiling_get(ctx, i, &name, &ms) == VP_OK)
				std::cout << ms << "ms ";
		}
		std::cout << std::endl;
	}

	void clearEvents() {
		events.clear();
		vp_profiling_clear(ctx);
	}

	/* Pools, src/opencl.cpp:108-135: an entry is free exactly when the pool holds its only reference, so a consumer thread
	 * that still keeps a frame (encoder, snapshot writer) blocks its reuse; pools grow on demand and never shrink. */
	std::shared_ptr<CLImage> acquire(const PixelFormat* format, int width, int height, const std::string& name) {
		std::shared_ptr<CLImage> img = reuse(pool[format], width, height, [&] { return std::make_shared<CLImage>(format, width, height, name); });
		img->name = name; /* a reused image is renamed (opencl.cpp:115) */
		return img;
	}

	std::shared_ptr<RawImage> acquireNV12(int width, int height) {
		return reuse(nv12pool, width, height, [&] { return std::make_shared<RawImage>(&PixelFormat::NV12, width, height); });
	}

	vp_ctx* handle() const { return ctx; }

private:
	template<typename P, typename Make>
	static std::shared_ptr<P> reuse(std::vector<std::shared_ptr<P>>& entries, int width, int height, Make make) {
		for (const std::shared_ptr<P>& e : entries)
			if (e.use_count() == 1 && e->width == width && e->height == height)
				return e;
		entries.push_back(make());
		return entries.back();
	}

	static vpcompat::Arg erase(const cl::Image2D& i) { vpcompat::Arg a; a.kind = vpcompat::Arg::IMG; a.img = i(); return a; }
	static vpcompat::Arg erase(const cl::Buffer& b) { vpcompat::Arg a; a.kind = vpcompat::Arg::BUF; a.buf = b(); return a; }
	static vpcompat::Arg erase(int v) { vpcompat::Arg a; a.kind = vpcompat::Arg::INT; a.i = v; return a; }
	static vpcompat::Arg erase(float v) { vpcompat::Arg a; a.kind = vpcompat::Arg::FLOAT; a.f = v; return a; }
	static vpcompat::Arg erase(double v) { return erase((float)v); }
	template<typename S, typename = std::enable_if_t<std::is_class_v<S> && std::is_trivially_copyable_v<S> && sizeof(S) == 72>>
	static vpcompat::Arg erase(const S& model) { /* CLCameraModel by value, Perspective.h:22-29 */
		vpcompat::Arg a; a.kind = vpcompat::Arg::BYTES; std::memcpy(a.bytes, &model, 72); return a;
	}

	static bool kinds(const std::vector<vpcompat::Arg>& a, const char* sig) { /* I image, B buffer, i int, f float, M model */
		if (a.size() != std::strlen(sig)) return false;
		for (size_t k = 0; k < a.size(); k++) {
			const vpcompat::Arg::Kind want = sig[k] == 'I' ? vpcompat::Arg::IMG : sig[k] == 'B' ? vpcompat::Arg::BUF : sig[k] == 'i' ? vpcompat::Arg::INT
			                               : sig[k] == 'f' ? vpcompat::Arg::FLOAT : vpcompat::Arg::BYTES;
			if (a[k].kind != want) return false;
		}
		return true;
	}

	int dispatch(const cl::Kernel& k, const cl::EnqueueArgs& args, const std::vector<vpcompat::Arg>& a) {
		using namespace vpcompat;
		const int gx = (int)args.global.size[0], gy = (int)args.global.size[1];
		switch (k.stage) {
		case RAW2QUAD: { /* raw2quad.cl:21: (img, ch0..3), NDRange (wq, hq) */
			if (!kinds(a, "BIIII")) break;
			vp_img* ch[4] = { a[1].img, a[2].img, a[3].img, a[4].img };
			return vp_raw2quad(ctx, a[0].buf, k.fmt, gx, gy, ch);
		}
		case RESAMPLING: { /* resampling.cl:52: (ch0..3, out, model, maxRobotHeight, fieldScale, offX, offY) */
			if (!kinds(a, "IIIIIMffff")) break;
			vp_img* ch[4] = { a[0].img, a[1].img, a[2].img, a[3].img };
			vp_camera_model m;
			std::memcpy(&m, a[5].bytes, 72);
			return vp_resampling(ctx, ch, k.fmt, a[4].img, &m, a[6].f, a[7].f, a[8].f, a[9].f, VP_SAMPLE_BILINEAR_RTE);
		}
		case GRADIENT_DOT: if (!kinds(a, "IIi")) break; return vp_gradient_dot(ctx, a[0].img, a[1].img, a[2].i);
		case SAT_HORIZONTAL: if (!kinds(a, "II")) break; return vp_sat_horizontal(ctx, a[0].img, a[1].img);
		case SAT_VERTICAL: if (!kinds(a, "II")) break; return vp_sat_vertical(ctx, a[0].img, a[1].img);
		case SAT_BLOB_CENTER: if (!kinds(a, "IIi")) break; return vp_circle(ctx, a[0].img, a[1].img, a[2].i);
		case BLOB_LIST: /* blobList.cl:36: (img, circ, matches, counter, circThreshold, minScore, radius, maxMatches) */
			if (!kinds(a, "IIBBffii")) break;
			return vp_blob_list(ctx, a[0].img, a[1].img, a[2].buf, a[3].buf, a[4].f, a[5].f, a[6].i, a[7].i);
		case RGBA2NV12: if (!kinds(a, "IB")) break; return vp_rgba2nv12(ctx, a[0].img, a[1].buf);
		case F2NV12: if (!kinds(a, "IB")) break; return vp_f2nv12(ctx, a[0].img, a[1].buf);
		case QUAD2NV12: {
			if (!kinds(a, "IIIIB")) break;
			vp_img* ch[4] = { a[0].img, a[1].img, a[2].img, a[3].img };
			return vp_quad2nv12(ctx, ch, k.fmt, a[4].buf, VP_SAMPLE_BILINEAR_RTE);
		}
		case QUAD2RGBA: {
			if (!kinds(a, "IIIII")) break;
			vp_img* ch[4] = { a[0].img, a[1].img, a[2].img, a[3].img };
			return vp_quad2rgba(ctx, ch, k.fmt, a[4].img, VP_SAMPLE_BILINEAR_RTE);
		}
		case BLOB_CENTER: if (!kinds(a, "IIii")) break; return vp_circularize(ctx, a[0].img, a[1].img, a[2].i, a[3].i);
		case BLOB_SCORE: if (!kinds(a, "IIIfi")) break; return vp_blob_score(ctx, a[0].img, a[1].img, a[2].img, a[3].f, a[4].i);
		default: break;
		}
		FATAL("Enqueue kernel error: argument list does not match the kernel signature (stage " << (int)k.stage << ", " << a.size() << " args)");
	}

	vp_ctx* ctx = nullptr;
	std::map<const PixelFormat*, std::vector<std::shared_ptr<CLImage>>> pool;
	std::vector<std::shared_ptr<RawImage>> nv12pool;
	std::vector<cl::Event> events;
};
