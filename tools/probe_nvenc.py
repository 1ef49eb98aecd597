#!/usr/bin/env python
"""Report-only probe: is a hardware H.264 encoder reachable on this box the way the reference reaches it
(src/rtpstreamer.cpp:62: avcodec_find_encoder_by_name("h264_nvenc") first)?  Prints what it finds; changes nothing."""
import ctypes.util
import shutil
import subprocess

print("libnvidia-encode:", ctypes.util.find_library("nvidia-encode") or "not found by the loader")
r = subprocess.run("ldconfig -p | grep -i -E 'nvidia-encode|nvcuvid|libavcodec' || true", shell=True, capture_output=True, text=True)
print("ldconfig:", r.stdout.strip() or "no nvidia-encode / nvcuvid / libavcodec entries")
ff = shutil.which("ffmpeg")
print("ffmpeg:", ff or "not installed")
if ff:
    r = subprocess.run([ff, "-hide_banner", "-encoders"], capture_output=True, text=True)
    print("ffmpeg nvenc encoders:", [l.strip() for l in r.stdout.splitlines() if "nvenc" in l] or "none")
r = subprocess.run("nvidia-smi --query-gpu=name,encoder.stats.sessionCount --format=csv,noheader || true", shell=True, capture_output=True, text=True)
print("nvidia-smi:", r.stdout.strip() or r.stderr.strip())
