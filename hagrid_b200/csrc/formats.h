// On-disk formats of the path:
//   .rays   the reference's ray files: 6 little-endian float32 per ray (org, dir), count = file size / 24;
//           tmin / tmax are supplied by the caller (load_rays, src/main.cpp:277-300)
//   .hgrid  grid cache of this library (the reference never serialises a grid): header + entries + cells
//           (or small cells) + references, so a large build can be reused
#pragma once

#include <string>

#include "build.h"
#include "hgb_types.h"
#include "mem_manager.h"

namespace hagrid {

/// Number of rays in a .rays file, -1 when it cannot be opened.
long long rays_file_count(const std::string& path);

/// Reads a .rays file straight to the device: the 24-byte records are uploaded as they are (a quarter less
/// PCIe traffic than 32-byte rays) and expanded on the device to Ray{org, tmin, dir, tmax}. `rays` is a device
/// buffer of rays_file_count() elements. Returns the number of rays, -1 on I/O errors.
long long load_rays_to_device(MemManager& mem, const std::string& path, float tmin, float tmax, Ray* rays);

/// Writes rays (device) as a .rays file (org and dir only, like the reference's format).
bool save_rays_from_device(MemManager& mem, const std::string& path, const Ray* rays, long long count);

/// Headless stand-in for the SDL window of the reference's viewer (src/main.cpp:558-625): writes a frame of
/// BGRA words (what update_surface / render_frame produce) as a binary PPM (P6, RGB).
bool save_image_ppm(const std::string& path, const unsigned char* bgra, int width, int height);

bool save_grid(MemManager& mem, const std::string& path, const Grid& grid, std::string& error);
/// Arrays come from `mem` like those of build_grid; `grid`'s previous arrays must have been freed by the caller.
bool load_grid(MemManager& mem, const std::string& path, Grid& grid, std::string& error);

/// Device half of load_rays_to_device (ray_records.cu)
void expand_ray_records(const float* dev_records, long long count, float tmin, float tmax, Ray* rays);
void pack_ray_records(const Ray* rays, long long count, float* dev_records);

} // namespace hagrid
