/*
 * compat/cl_kernels.h -- replaces the generated cl_kernels.h of the reference build (CMakeLists.txt:62-79 embeds every
 * kernel/ .cl file with .incbin as `kernel_<name>_cl`).  The symbols keep their names so that
 * `openCl->compile(kernel_blobList_cl)` (src/main.cpp:253, src/Resources.cpp:121-130, src/blob_benchmark.cpp:117)
 * compiles unchanged; their content is a stage tag that compat/opencl.h maps to the CUDA entry point.
 */
#pragma once

inline constexpr char kernel_raw2quad_cl[] = "vp_b200:raw2quad";
inline constexpr char kernel_resampling_cl[] = "vp_b200:resampling";
inline constexpr char kernel_gradientDot_cl[] = "vp_b200:gradientDot";
inline constexpr char kernel_satHorizontal_cl[] = "vp_b200:satHorizontal";
inline constexpr char kernel_satVertical_cl[] = "vp_b200:satVertical";
inline constexpr char kernel_satBlobCenter_cl[] = "vp_b200:satBlobCenter";
inline constexpr char kernel_blobList_cl[] = "vp_b200:blobList";
inline constexpr char kernel_rgba2nv12_cl[] = "vp_b200:rgba2nv12";
inline constexpr char kernel_f2nv12_cl[] = "vp_b200:f2nv12";
inline constexpr char kernel_quad2nv12_cl[] = "vp_b200:quad2nv12";
inline constexpr char kernel_quad2rgba_cl[] = "vp_b200:quad2rgba";
inline constexpr char kernel_blobCenter_cl[] = "vp_b200:blobCenter"; /* dead in the reference (never compiled) */
inline constexpr char kernel_blobScore_cl[] = "vp_b200:blobScore";   /* compiled by blob_benchmark.cpp:117, never enqueued */
