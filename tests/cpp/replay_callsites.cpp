/*
 * replay_callsites.cpp -- call-site replay test of the drop-in shim (include/compat/opencl.h).
 *
 * The reference's main.cpp / Resources.cpp cannot be compiled here (OpenCV, Eigen, yaml-cpp, protobuf, libav are
 * absent), so this program exercises the shim exactly the way those files do -- same classes, same member calls, same
 * argument lists in the same order -- and dumps every result for comparison with the CPU oracle (tests/test_gpu_shim.py).
 * The patterns replayed (path:line in the reference tree):
 *   Resources.cpp:70,121-130   OpenCL(), compile() of ten kernels with camera->format().kernelOptions
 *   Resources.cpp:138-143      raw2quad: acquire 4 U8 images, await(raw2quadKernel, NDRange(w,h), img.buffer, ch0..3)
 *   Resources.cpp:151-164      rgba2blobCenter: five acquires, four run() chained through events, one await()
 *   main.cpp:253,257-258       compile(kernel_blobList_cl), CLArray matchArray / counter
 *   main.cpp:283-317           counter reset through a write map, blobList await, read maps of counter and matches
 *   Resources.cpp:145-149,166-186  quad2rgba, streamQuad, streamImage (RGBA8 and F32)
 *   rtpstreamer.cpp:177-181    the encoder thread maps an NV12 RawImage it holds a shared_ptr to
 *   blob_benchmark.cpp:162,190-191  read map of blobCenter, nth_element in place over rowPitch*height
 *   opencvdriver.cpp:57-66 / mvimpactdriver.cpp:24  frame written through a map / copy-in constructor
 *   cameradriver.h:35-47 / main.cpp:262-267   frames pulled through CameraDriver::readImage() until it returns nullptr
 *                                             (include/compat/syntheticdriver.h: the synthetic Bayer driver of SURVEY 8 f4)
 *
 * usage: replay_callsites <in.bin> <out.bin>     in.bin = vp_params + raw frame
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "opencl.h"
#include "cl_kernels.h"
#include "syntheticdriver.h"

typedef struct __attribute__((packed)) { /* Perspective.h:22-29 */
	int shape[2];
	float f;
	float p[2];
	float d;
	float r[9];
	float c[3];
} CLCameraModel;

typedef struct __attribute__((packed)) { /* main.cpp:33-41 */
	float x, y;
	struct __attribute__((packed)) { cl_uchar r, g, b; } color, center;
	float circ;
	float score;
} CLMatch;
static_assert(sizeof(CLCameraModel) == 72 && sizeof(CLMatch) == 22, "packed layouts of the reference");

/* the slice of class Resources this path uses (Resources.h:98-114) */
struct MiniResources {
	std::shared_ptr<OpenCL> openCl;
	const PixelFormat* cameraFormat;
	cl::Kernel raw2quadKernel, resampling, gradientDot, satHorizontal, satVertical, satBlobCenter, quad2rgbaKernel, quad2nv12, rgba2nv12, f2nv12;
	vp_params geo;

	MiniResources(const vp_params& p): geo(p) {
		openCl = std::make_shared<OpenCL>();
		cameraFormat = p.fmt == VP_FMT_RGGB8 ? &PixelFormat::RGGB8 : p.fmt == VP_FMT_GRBG8 ? &PixelFormat::GRBG8 : &PixelFormat::BGR8;
		raw2quadKernel = openCl->compile(kernel_raw2quad_cl, cameraFormat->kernelOptions);
		resampling = openCl->compile(kernel_resampling_cl, cameraFormat->kernelOptions);
		gradientDot = openCl->compile(kernel_gradientDot_cl);
		satHorizontal = openCl->compile(kernel_satHorizontal_cl);
		satVertical = openCl->compile(kernel_satVertical_cl);
		satBlobCenter = openCl->compile(kernel_satBlobCenter_cl);
		quad2rgbaKernel = openCl->compile(kernel_quad2rgba_cl, cameraFormat->kernelOptions);
		quad2nv12 = openCl->compile(kernel_quad2nv12_cl, cameraFormat->kernelOptions);
		rgba2nv12 = openCl->compile(kernel_rgba2nv12_cl);
		f2nv12 = openCl->compile(kernel_f2nv12_cl);
	}

	void raw2quad(const RawImage& img, std::shared_ptr<CLImage>* channels) {
		for (int i = 0; i < 4; i++)
			channels[i] = openCl->acquire(&PixelFormat::U8, img.width, img.height, img.name);
		openCl->await(raw2quadKernel, cl::EnqueueArgs(cl::NDRange(img.width, img.height)), img.buffer, channels[0]->image, channels[1]->image,
		              channels[2]->image, channels[3]->image);
	}

	std::shared_ptr<CLImage> quad2rgba(std::shared_ptr<CLImage>* channels) {
		std::shared_ptr<CLImage> rgba = openCl->acquire(&PixelFormat::RGBA8, channels[0]->width, channels[0]->height, channels[0]->name);
		openCl->await(quad2rgbaKernel, cl::EnqueueArgs(cl::NDRange(channels[0]->width, channels[0]->height)), channels[0]->image, channels[1]->image,
		              channels[2]->image, channels[3]->image, rgba->image);
		return rgba;
	}

	void rgba2blobCenter(const std::shared_ptr<CLImage>* channels, std::shared_ptr<CLImage>& flat, std::shared_ptr<CLImage>& gradDot,
	                     std::shared_ptr<CLImage>& blobCenter) {
		const int wf = geo.wf, hf = geo.hf;
		cl::NDRange visibleFieldRange(wf, hf);
		flat = openCl->acquire(&PixelFormat::RGBA8, wf, hf, channels[0]->name);
		gradDot = openCl->acquire(&PixelFormat::F32, wf, hf, channels[0]->name);
		std::shared_ptr<CLImage> gradDotHor = openCl->acquire(&PixelFormat::F32, wf, hf, channels[0]->name);
		std::shared_ptr<CLImage> gradDotSat = openCl->acquire(&PixelFormat::F32, wf, hf, channels[0]->name);
		blobCenter = openCl->acquire(&PixelFormat::F32, wf, hf, channels[0]->name);

		CLCameraModel model;
		std::memcpy(&model, &geo.model, 72);
		cl::Event e1 = openCl->run(resampling, cl::EnqueueArgs(visibleFieldRange), channels[0]->image, channels[1]->image, channels[2]->image,
		                           channels[3]->image, flat->image, model, (float)geo.max_robot_height, geo.field_scale, geo.off_x, geo.off_y);
		cl::Event e2 = openCl->run(gradientDot, cl::EnqueueArgs(e1, visibleFieldRange), flat->image, gradDot->image, (int)geo.grad_offset);
		cl::Event e3 = openCl->run(satHorizontal, cl::EnqueueArgs(e2, cl::NDRange(hf)), gradDot->image, gradDotHor->image);
		cl::Event e4 = openCl->run(satVertical, cl::EnqueueArgs(e3, cl::NDRange(wf)), gradDotHor->image, gradDotSat->image);
		openCl->await(satBlobCenter, cl::EnqueueArgs(e4, visibleFieldRange), gradDotSat->image, blobCenter->image, (int)geo.circle_radius);
	}

	std::shared_ptr<RawImage> streamQuad(std::shared_ptr<CLImage>* channels) {
		std::shared_ptr<RawImage> nv12 = openCl->acquireNV12(channels[0]->width, channels[0]->height);
		openCl->await(quad2nv12, cl::EnqueueArgs(cl::NDRange(channels[0]->width, channels[0]->height)), channels[0]->image, channels[1]->image,
		              channels[2]->image, channels[3]->image, nv12->buffer);
		return nv12;
	}

	std::shared_ptr<RawImage> streamImage(CLImage& img) {
		cl::Kernel kernel = img.format == &PixelFormat::RGBA8 ? rgba2nv12 : f2nv12;
		std::shared_ptr<RawImage> nv12 = openCl->acquireNV12(img.width, img.height);
		openCl->await(kernel, cl::EnqueueArgs(cl::NDRange(img.width, img.height)), img.image, nv12->buffer);
		return nv12;
	}
};

static void put(FILE* f, const void* p, size_t n) {
	if (fwrite(p, 1, n, f) != n) FATAL("short write");
}

/* what RTPStreamer::encoderRun does with a frame: map it on another thread and read w*h*3/2 bytes (rtpstreamer.cpp:177-181) */
static std::vector<uint8_t> encoder_thread_reads(std::shared_ptr<RawImage> nv12) {
	std::vector<uint8_t> out((size_t)nv12->width * nv12->height * 3 / 2);
	std::thread t([&] {
		CLMap<uint8_t> map = nv12->read<uint8_t>();
		std::copy(*map, *map + out.size(), out.begin());
	});
	t.join();
	return out;
}

int main(int argc, char** argv) {
	if (argc != 3) FATAL("usage: replay_callsites <in.bin> <out.bin>");
	FILE* in = fopen(argv[1], "rb");
	if (!in) FATAL("cannot open " << argv[1]);
	vp_params p;
	if (fread(&p, sizeof p, 1, in) != 1) FATAL("short params");
	const size_t raw_bytes = (size_t)p.wq * p.hq * (p.fmt == VP_FMT_BGR8 ? 3 : 4);
	std::vector<uint8_t> frame(raw_bytes);
	if (fread(frame.data(), 1, raw_bytes, in) != raw_bytes) FATAL("short frame");
	fclose(in);

	MiniResources r(p);
	cl::Kernel blobList = r.openCl->compile(kernel_blobList_cl);
	const int maxBlobs = p.max_blobs;
	CLArray matchArray(sizeof(CLMatch) * maxBlobs);
	CLArray counter(sizeof(cl_int) * 3);
	FILE* out = fopen(argv[2], "wb");
	if (!out) FATAL("cannot open " << argv[2]);

	/* the camera: two copies of the frame behind the CameraDriver interface */
	std::vector<unsigned char> two(frame.begin(), frame.end());
	two.insert(two.end(), frame.begin(), frame.end());
	std::unique_ptr<CameraDriver> camera = std::make_unique<SyntheticDriver>(std::move(two), r.cameraFormat, p.wq, p.hq, 60.0);
	if (camera->format().pixelSize() != r.cameraFormat->pixelSize() || camera->expectedFrametime() <= 0.0) FATAL("driver format");

	for (int frameId = 1; frameId <= 3; frameId++) { /* three frames: pooled images are reused (use_count()==1) */
		/* frame 1: pulled from the camera driver like main.cpp:262-267; frame 2: copy-in constructor; frame 3: the OpenCV driver's
		 * way, written into a mapped buffer */
		std::shared_ptr<RawImage> img;
		if (frameId == 1) {
			img = camera->readImage();
			if (!img || img->width != p.wq || img->height != p.hq || img->timestamp != camera->getTime()) FATAL("driver frame");
		} else if (frameId == 2) {
			img = std::make_shared<RawImage>(r.cameraFormat, p.wq, p.hq, (double)frameId, frame.data());
		} else {
			img = std::make_shared<RawImage>(r.cameraFormat, p.wq, p.hq, (double)frameId);
			CLMap<uint8_t> map = img->write<uint8_t>();
			std::memcpy(*map, frame.data(), raw_bytes);
		}
		RawImage shared_copy(*img); /* SpinnakerImage keeps a copy that shares the buffer (opencl.h:170) */

		std::shared_ptr<CLImage> channels[4];
		r.raw2quad(shared_copy, channels);
		std::shared_ptr<CLImage> flat, gradDot, blobCenter;
		r.rgba2blobCenter(channels, flat, gradDot, blobCenter);
		{
			CLMap<int> counterMap = counter.write<int>();
			counterMap[0] = 0;
			counterMap[1] = 0;
			counterMap[2] = 0;
		}
		r.openCl->await(blobList, cl::EnqueueArgs(cl::NDRange(p.wf, p.hf)), flat->image, blobCenter->image, matchArray.buffer, counter.buffer,
		                (float)p.circ_threshold, (float)0.0f, (int)p.blob_radius, maxBlobs);

		std::vector<CLMatch> matches;
		int counters[3];
		{
			CLMap<int> counterMap = counter.read<int>();
			CLMap<CLMatch> matchMap = matchArray.read<CLMatch>();
			const int matchAmount = std::min(maxBlobs, counterMap[0]);
			for (int i = 0; i < matchAmount; i++)
				matches.push_back(matchMap[i]);
			for (int i = 0; i < 3; i++) counters[i] = counterMap[i];
		}

		std::shared_ptr<RawImage> nv12quad = r.streamQuad(channels);
		std::vector<uint8_t> nv12_quad = encoder_thread_reads(nv12quad);
		std::vector<uint8_t> nv12_flat = encoder_thread_reads(r.streamImage(*flat));
		std::vector<uint8_t> nv12_grad = encoder_thread_reads(r.streamImage(*gradDot));
		std::shared_ptr<CLImage> rgba = r.quad2rgba(channels);

		float percentile;
		{ /* blob_benchmark.cpp:190-191 sorts a READ map in place */
			CLImageMap<float> blobMap = blobCenter->read<float>();
			std::vector<float> copy(*blobMap, *blobMap + blobMap.rowPitch * p.hf);
			std::nth_element(*blobMap, *blobMap + (int)(blobMap.rowPitch * p.hf * 0.99f), *blobMap + blobMap.rowPitch * p.hf);
			percentile = (*blobMap)[(int)(blobMap.rowPitch * p.hf * 0.99f)];
			std::copy(copy.begin(), copy.end(), *blobMap); /* restore what the dump below reads back from the device */
		}
		if (frameId == 3) {
			const size_t nf = (size_t)p.wf * p.hf, nq = (size_t)p.wq * p.hq;
			{ CLImageMap<RGBA> m = flat->read<RGBA>(); if (m.bytePitch != (size_t)p.wf * 4) FATAL("pitch"); put(out, *m, nf * 4); }
			{ CLImageMap<float> m = gradDot->read<float>(); put(out, *m, nf * 4); }
			{ CLImageMap<float> m = blobCenter->read<float>(); put(out, *m, nf * 4); }
			put(out, counters, 12);
			const int n = (int)matches.size();
			put(out, &n, 4);
			put(out, matches.data(), 22 * (size_t)n);
			put(out, nv12_flat.data(), nv12_flat.size());
			put(out, nv12_grad.data(), nv12_grad.size());
			put(out, nv12_quad.data(), nv12_quad.size());
			{ CLImageMap<RGBA> m = rgba->read<RGBA>(); put(out, *m, nq * 4); }
			put(out, &percentile, 4);
			(void)nq;
		}
		if (frameId == 3)
			r.openCl->printRuntimes(); /* main.cpp:363-366 (BENCHMARK) */
		r.openCl->clearEvents(); /* main.cpp:372 */
	}
	if (!camera->readImage() || camera->readImage()) FATAL("the driver must hand out its second frame and then end the stream");
	fclose(out);
	LOG("replay ok");
	return 0;
}
