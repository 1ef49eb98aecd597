/*
 * syntheticdriver.h -- a CameraDriver that replays pre-rendered synthetic Bayer frames.
 *
 * The reference can only be driven from image files through its OpenCV driver, i.e. with BGR8 frames
 * (src/driver/opencvdriver.cpp:57-66, python/dataset.py:92); BASELINE config 1 asks for blob_benchmark on a BayerRG8 frame,
 * which needs a driver the reference does not have.  This is that driver, written against the interface of
 * src/driver/cameradriver.h:35-47 (readImage / format / expectedFrametime / getTime) on top of the drop-in shim: frames
 * rendered by vpb200.synth (raw bytes, frame after frame in one file or one memory block) come back as
 * RawImage{RGGB8|GRBG8|BGR8, width, height} -- width and height in QUADS for the Bayer formats, like the Spinnaker and
 * mvIMPACT drivers hand them out (spinnakerdriver.cpp:124, mvimpactdriver.cpp:24) -- through the copy-in constructor
 * (opencl.h:168-188).  getTime() is frame index x frame time: "Bound to the driver for reproducibility during testing with
 * files" (cameradriver.h:45-46).  readImage() returns nullptr after the last frame, which ends the loops of main.cpp:266-267
 * and blob_benchmark.cpp:138-140.  The ground truth that goes with the frames is written by
 * vpb200.synth.write_ground_truth_yaml in the schema src/GroundTruth.cpp:23-92 parses.
 *
 * In a build of the reference add this header next to the other drivers and construct it in Resources::openCamera
 * (Resources.cpp:39-68) for `cam.driver: SYNTHETIC`; the class declaration below is only used when the reference's own
 * driver/cameradriver.h is not on the include path (this repository's tests).
 */
#pragma once

#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "opencl.h"

#if __has_include("driver/cameradriver.h")
#include "driver/cameradriver.h"
#else
/* src/driver/cameradriver.h:35-47 */
class CameraDriver {
public:
	virtual ~CameraDriver() = default;
	virtual std::shared_ptr<RawImage> readImage() = 0;
	virtual const PixelFormat format() = 0;
	virtual double expectedFrametime() = 0;
	virtual double getTime() { return 0.0; }
};
#endif

class SyntheticDriver : public CameraDriver {
public:
	/* frames: n_frames x (width*height*pixelSize) bytes, back to back; width/height in quads for Bayer formats */
	SyntheticDriver(std::vector<unsigned char> frames, const PixelFormat* fmt, int width, int height, double fps = 60.0)
		: data(std::move(frames)), fmt(fmt), width(width), height(height), frametime(1.0 / fps) {
		if (frameBytes() == 0 || data.size() % frameBytes() != 0)
			FATAL("SyntheticDriver: " << data.size() << " bytes are not a whole number of " << width << "x" << height << " frames");
	}
	SyntheticDriver(const std::string& path, const PixelFormat* fmt, int width, int height, double fps = 60.0)
		: SyntheticDriver(load(path), fmt, width, height, fps) {}

	std::shared_ptr<RawImage> readImage() override {
		if ((size_t)next * frameBytes() >= data.size())
			return nullptr;
		unsigned char* frame = data.data() + (size_t)next * frameBytes();
		next++;
		return std::make_shared<RawImage>(fmt, width, height, getTime(), frame);
	}
	const PixelFormat format() override { return *fmt; }
	double expectedFrametime() override { return frametime; }
	double getTime() override { return next * frametime; } /* deterministic: the capture time of the frame last handed out */
	int framesLeft() const { return (int)(data.size() / frameBytes()) - next; }

private:
	size_t frameBytes() const { return (size_t)width * height * fmt->pixelSize(); }
	static std::vector<unsigned char> load(const std::string& path) {
		FILE* f = fopen(path.c_str(), "rb");
		if (!f)
			FATAL("SyntheticDriver: cannot open " << path);
		std::vector<unsigned char> bytes;
		unsigned char buf[1 << 16];
		size_t n;
		while ((n = fread(buf, 1, sizeof buf, f)) > 0)
			bytes.insert(bytes.end(), buf, buf + n);
		fclose(f);
		return bytes;
	}

	std::vector<unsigned char> data;
	const PixelFormat* fmt;
	int width, height;
	double frametime;
	int next = 0;
};
