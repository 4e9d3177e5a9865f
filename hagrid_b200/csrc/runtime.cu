// Runtime support of the path: the device buffer pool behind MemManager
// (interface of src/mem_manager.h:34-119) and the device timer profile()
// (src/profile.cu:5-18).
#include <algorithm>
#include <iostream>

#include "mem_manager.h"
#include "runtime.h"

namespace hagrid {

std::atomic<unsigned long long> g_kernel_launches{0};
unsigned long long kernel_launch_count() { return g_kernel_launches.load(std::memory_order_relaxed); }

namespace {

/// Stream-ordered allocation on the legacy default stream: the pool keeps
/// released memory (unlimited release threshold), so growing and shrinking
/// slots costs no cudaMalloc/cudaFree device synchronisation after warm-up.
void prepare_pool() {
    static bool done[64] = {};
    int dev = 0;
    HGB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    HGB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long threshold = ~0ull;
    HGB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    done[dev] = true;
}

void release(Slot& slot) {
    if (slot.ptr) HGB_CUDA(cudaFreeAsync(slot.ptr, 0));
    slot.ptr = nullptr;
    slot.size = 0;
}

} // namespace

MemManager::~MemManager() {
    // Idle and handed-out slots alike: the manager owns them all. Errors are ignored (a manager with static
    // storage duration may be destroyed after the CUDA runtime has shut down).
    for (Slot& slot : slots_) {
        if (slot.ptr) cudaFreeAsync(slot.ptr, 0);
        slot.ptr = nullptr;
        slot.size = 0;
    }
    cudaGetLastError();
}

void MemManager::alloc_slot(Slot& slot, size_t bytes) {
    if (slot.in_use) {
        std::cerr << "MemManager: slot handed out twice" << std::endl;
        std::abort();
    }
    if (slot.size < bytes || !slot.ptr) {
        prepare_pool();
        const size_t old = slot.size;
        release(slot);
        // keep mode: one eighth of slack, so that the slightly larger request of the next build still fits.
        // Zero-byte requests still get a distinct address so that free() can track them.
        const size_t capacity = keep_ ? bytes + bytes / 8 : bytes;
        HGB_CUDA(cudaMallocAsync(&slot.ptr, std::max<size_t>(capacity, 16), 0));
        slot.size = capacity;
        usage_ += capacity - old;
        max_usage_ = std::max(max_usage_, usage_);
    }
    slot.in_use = true;
    // Unlike the reference (src/mem_manager.cu:36-44) keep mode does not hand an idle slot back at every new
    // high-water mark: that made a steady sequence of builds re-allocate a buffer per build.
}

void MemManager::free_slot(Slot& slot) {
    slot.in_use = false;
    if (!keep_) {
        usage_ -= slot.size;
        release(slot);
    }
}

void MemManager::copy_dev_to_dev(void* dst, const void* src, size_t bytes) {
    HGB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
}

void MemManager::copy_hst_to_dev(void* dst, const void* src, size_t bytes) {
    HGB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
}

void MemManager::copy_dev_to_hst(void* dst, const void* src, size_t bytes) {
    HGB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
}

void MemManager::zero_dev(void* ptr, size_t bytes) { HGB_CUDA(cudaMemsetAsync(ptr, 0x00, bytes, 0)); }
void MemManager::one_dev(void* ptr, size_t bytes)  { HGB_CUDA(cudaMemsetAsync(ptr, 0xFF, bytes, 0)); }

void MemManager::debug_slots() const {
    size_t total = 0;
    std::cout << "SLOTS: " << std::endl;
    for (const Slot& slot : slots_) {
        std::cout << (slot.in_use ? "[X] " : "[ ] ") << double(slot.size) / (1024.0 * 1024.0) << "MB" << std::endl;
        total += slot.size;
    }
    std::cout << double(total) / (1024.0 * 1024.0) << "MB total" << std::endl;
}

int sm_count() {
    static std::atomic<int> counts[64];
    int dev = 0;
    HGB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return 148;
    int n = counts[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        HGB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        counts[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

/// Hands memory the stream-ordered pool retains back to the driver (after a scene has been destroyed)
void trim_device_pool() {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess || cudaMemPoolTrimTo(pool, 0) != cudaSuccess)
        cudaGetLastError();
}

float profile(std::function<void()> work) {
    cudaEvent_t begin, end;
    HGB_CUDA(cudaEventCreate(&begin));
    HGB_CUDA(cudaEventCreate(&end));
    HGB_CUDA(cudaEventRecord(begin, 0));
    work();
    HGB_CUDA(cudaEventRecord(end, 0));
    HGB_CUDA(cudaEventSynchronize(end));
    float ms = 0.0f;
    HGB_CUDA(cudaEventElapsedTime(&ms, begin, end));
    HGB_CUDA(cudaEventDestroy(begin));
    HGB_CUDA(cudaEventDestroy(end));
    return ms;
}

} // namespace hagrid
