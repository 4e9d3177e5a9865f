// Host side of the on-disk formats (see formats.h).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime_api.h>

#include "formats.h"

namespace hagrid {

namespace {

struct File {
    std::FILE* fp;
    File(const std::string& path, const char* mode) : fp(std::fopen(path.c_str(), mode)) {}
    ~File() { if (fp) std::fclose(fp); }
    bool read(void* dst, size_t bytes) { return bytes == 0 || std::fread(dst, 1, bytes, fp) == bytes; }
    bool write(const void* src, size_t bytes) { return bytes == 0 || std::fwrite(src, 1, bytes, fp) == bytes; }
};

long long file_size(std::FILE* fp) {
    std::fseek(fp, 0, SEEK_END);
    const long long n = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    return n;
}

constexpr char kGridMagic[8] = {'H', 'G', 'R', 'I', 'D', '0', '0', '1'};

struct GridHeader {                 // little-endian, 104 bytes, followed by the offsets
    char    magic[8];
    float   bbox_min[3], bbox_max[3];
    int32_t dims[3];
    int32_t shift, num_cells, num_entries, num_refs, compressed, num_offsets;
    int32_t reserved[9];
};
static_assert(sizeof(GridHeader) == 104, "grid file header layout");

} // namespace

long long rays_file_count(const std::string& path) {
    File f(path, "rb");
    if (!f.fp) return -1;
    return file_size(f.fp) / (long long)(sizeof(float) * 6);          // src/main.cpp:282
}

long long load_rays_to_device(MemManager& mem, const std::string& path, float tmin, float tmax, Ray* rays) {
    File f(path, "rb");
    if (!f.fp) return -1;
    const long long count = file_size(f.fp) / (long long)(sizeof(float) * 6);
    if (count <= 0) return count;
    // staged through page-locked memory in slices so that disk reads and uploads of a large file overlap
    const long long slice = 1ll << 20;                                 // rays per slice (24 MB)
    float* dev_records = mem.alloc<float>(size_t(6 * count));
    float* pinned[2] = {nullptr, nullptr};
    cudaEvent_t done[2];
    for (int i = 0; i < 2; i++) {
        if (cudaMallocHost(reinterpret_cast<void**>(&pinned[i]), size_t(std::min(slice, count)) * 24) != cudaSuccess) { cudaGetLastError(); pinned[i] = nullptr; }
        cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    bool ok = pinned[0] && pinned[1];
    int slot = 0;
    for (long long begin = 0; ok && begin < count; begin += slice, slot ^= 1) {
        const long long n = std::min(slice, count - begin);
        cudaEventSynchronize(done[slot]);                              // the slice buffer is free again
        ok = f.read(pinned[slot], size_t(n) * 24);
        if (ok) {
            cudaMemcpyAsync(dev_records + 6 * begin, pinned[slot], size_t(n) * 24, cudaMemcpyHostToDevice, 0);
            cudaEventRecord(done[slot], 0);
        }
    }
    if (ok) expand_ray_records(dev_records, count, tmin, tmax, rays);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) { if (pinned[i]) cudaFreeHost(pinned[i]); cudaEventDestroy(done[i]); }
    mem.free(dev_records);
    return ok ? count : -1;
}

bool save_rays_from_device(MemManager& mem, const std::string& path, const Ray* rays, long long count) {
    File f(path, "wb");
    if (!f.fp) return false;
    if (count <= 0) return true;
    float* dev_records = mem.alloc<float>(size_t(6 * count));
    pack_ray_records(rays, count, dev_records);
    std::vector<float> host(size_t(6 * count));
    mem.copy<Copy::DEV_TO_HST>(host.data(), dev_records, host.size());
    mem.free(dev_records);
    return f.write(host.data(), host.size() * sizeof(float));
}

bool save_image_ppm(const std::string& path, const unsigned char* bgra, int width, int height) {
    File f(path, "wb");
    if (!f.fp) return false;
    char head[64];
    const int len = std::snprintf(head, sizeof(head), "P6\n%d %d\n255\n", width, height);
    std::vector<unsigned char> rgb(size_t(width) * height * 3);
    for (size_t i = 0; i < size_t(width) * height; i++) {
        rgb[3 * i + 0] = bgra[4 * i + 2];
        rgb[3 * i + 1] = bgra[4 * i + 1];
        rgb[3 * i + 2] = bgra[4 * i + 0];
    }
    return f.write(head, size_t(len)) && f.write(rgb.data(), rgb.size());
}

bool save_grid(MemManager& mem, const std::string& path, const Grid& grid, std::string& error) {
    if (!grid.entries || (!grid.cells && !grid.small_cells)) { error = "no grid to save"; return false; }
    File f(path, "wb");
    if (!f.fp) { error = "cannot create " + path; return false; }
    GridHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, kGridMagic, 8);
    h.bbox_min[0] = grid.bbox.min.x; h.bbox_min[1] = grid.bbox.min.y; h.bbox_min[2] = grid.bbox.min.z;
    h.bbox_max[0] = grid.bbox.max.x; h.bbox_max[1] = grid.bbox.max.y; h.bbox_max[2] = grid.bbox.max.z;
    h.dims[0] = grid.dims.x; h.dims[1] = grid.dims.y; h.dims[2] = grid.dims.z;
    h.shift = grid.shift; h.num_cells = grid.num_cells; h.num_entries = grid.num_entries; h.num_refs = grid.num_refs;
    h.compressed = grid.small_cells ? 1 : 0;
    h.num_offsets = int32_t(grid.offsets.size());
    bool ok = f.write(&h, sizeof(h)) && f.write(grid.offsets.data(), sizeof(int) * grid.offsets.size());
    auto dump = [&](const void* dev, size_t bytes) {
        if (!ok || bytes == 0) return;
        std::vector<char> host(bytes);
        mem.copy<Copy::DEV_TO_HST>(host.data(), static_cast<const char*>(dev), bytes);
        ok = f.write(host.data(), bytes);
    };
    dump(grid.entries, sizeof(Entry) * size_t(grid.num_entries));
    if (grid.small_cells) dump(grid.small_cells, sizeof(SmallCell) * size_t(grid.num_cells));
    else                  dump(grid.cells, sizeof(Cell) * size_t(grid.num_cells));
    dump(grid.ref_ids, sizeof(int) * size_t(grid.num_refs));
    if (!ok) error = "write error on " + path;
    return ok;
}

bool load_grid(MemManager& mem, const std::string& path, Grid& grid, std::string& error) {
    File f(path, "rb");
    if (!f.fp) { error = "cannot open " + path; return false; }
    const long long size = file_size(f.fp);
    GridHeader h;
    if (!f.read(&h, sizeof(h)) || std::memcmp(h.magic, kGridMagic, 8) != 0) { error = "not a grid file: " + path; return false; }
    if (h.num_cells < 0 || h.num_entries < 0 || h.num_refs < 0 || h.num_offsets < 0 || h.num_offsets > 64 || h.shift < 0 || h.shift > 30) {
        error = "corrupt grid header"; return false;
    }
    const size_t cell_bytes = (h.compressed ? sizeof(SmallCell) : sizeof(Cell)) * size_t(h.num_cells);
    const long long want = (long long)sizeof(h) + 4ll * h.num_offsets + 4ll * h.num_entries + (long long)cell_bytes + 4ll * h.num_refs;
    if (size != want) { error = "grid file is truncated or has trailing data"; return false; }
    std::vector<int> offsets(size_t(h.num_offsets));
    if (!f.read(offsets.data(), sizeof(int) * offsets.size())) { error = "read error"; return false; }

    // Everything is read and checked on the host before anything is allocated: a file whose indices point outside its
    // own arrays would send the traversal out of bounds (or round in circles through the voxel map)
    long long top = 1;
    for (int k = 0; k < 3; k++) {
        if (h.dims[k] <= 0 || ((long long)h.dims[k] << h.shift) > (1ll << 30) || !(h.bbox_max[k] > h.bbox_min[k])) { error = "corrupt grid header"; return false; }
        top *= h.dims[k];
    }
    if (top > h.num_entries || h.num_cells <= 0) { error = "corrupt grid header"; return false; }
    for (size_t i = 0; i < offsets.size(); i++)
        if (offsets[i] < 0 || offsets[i] > h.num_entries || (i > 0 && offsets[i] < offsets[i - 1])) { error = "corrupt grid file: level offsets"; return false; }
    std::vector<uint32_t> host_entries(size_t(h.num_entries));
    std::vector<char> host_cells(cell_bytes);
    std::vector<int> host_refs(size_t(h.num_refs));
    if (!f.read(host_entries.data(), 4 * host_entries.size()) || !f.read(host_cells.data(), cell_bytes) ||
        !f.read(host_refs.data(), 4 * host_refs.size())) { error = "read error"; return false; }
    for (size_t i = 0; i < host_entries.size(); i++) {
        const uint32_t log_dim = host_entries[i] & 3u, begin = host_entries[i] >> 2;
        // a leaf names a cell; an inner node names a block of (2^log_dim)^3 entries that lies behind it (levels only point down)
        const bool ok_entry = log_dim == 0 ? begin < uint32_t(h.num_cells)
                                           : begin > i && (unsigned long long)begin + (1ull << (3 * log_dim)) <= (unsigned long long)h.num_entries;
        if (!ok_entry) { error = "corrupt grid file: voxel map"; return false; }
    }
    for (int c = 0; c < h.num_cells; c++) {
        bool ok_cell;
        if (h.compressed) {
            const SmallCell& cell = reinterpret_cast<const SmallCell*>(host_cells.data())[c];
            ok_cell = cell.begin >= -1 && cell.begin < h.num_refs;
        } else {
            const Cell& cell = reinterpret_cast<const Cell*>(host_cells.data())[c];
            ok_cell = cell.begin >= 0 && cell.begin <= cell.end && cell.end <= h.num_refs;
        }
        if (!ok_cell) { error = "corrupt grid file: cells"; return false; }
    }
    if (h.compressed && h.num_refs > 0 && host_refs.back() >= 0) { error = "corrupt grid file: the last reference list has no sentinel"; return false; }
    auto upload = [&](const void* host, size_t bytes) -> char* {
        char* dev = mem.alloc<char>(bytes ? bytes : 16);
        if (bytes) mem.copy<Copy::HST_TO_DEV>(dev, static_cast<const char*>(host), bytes);
        return dev;
    };
    char* entries = upload(host_entries.data(), 4 * host_entries.size());
    char* cells = upload(host_cells.data(), cell_bytes);
    char* refs = upload(host_refs.data(), 4 * host_refs.size());
    grid.entries = reinterpret_cast<Entry*>(entries);
    grid.cells = h.compressed ? nullptr : reinterpret_cast<Cell*>(cells);
    grid.small_cells = h.compressed ? reinterpret_cast<SmallCell*>(cells) : nullptr;
    grid.ref_ids = reinterpret_cast<int*>(refs);
    grid.bbox = BBox(vec3(h.bbox_min[0], h.bbox_min[1], h.bbox_min[2]), vec3(h.bbox_max[0], h.bbox_max[1], h.bbox_max[2]));
    grid.dims = ivec3(h.dims[0], h.dims[1], h.dims[2]);
    grid.shift = h.shift; grid.num_cells = h.num_cells; grid.num_entries = h.num_entries; grid.num_refs = h.num_refs;
    grid.offsets = offsets;
    return true;
}

} // namespace hagrid
