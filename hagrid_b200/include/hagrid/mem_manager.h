// Device buffer pool used by the construction pipeline and by the caller
// (same public interface and object layout as src/mem_manager.h:21-119, so
// a front end compiled against either header links against this library).
//
// Slots are recycled best-fit: alloc() picks the smallest idle slot that is large
// enough and grows the largest idle one (with some slack) when none is. In `keep`
// mode freed slots retain their memory (= --keep-alive, fast rebuilds) and are
// never given back while the manager lives — a B200 has 180 GB; otherwise free()
// returns the memory. B200-first difference: slot memory comes from
// the device's stream-ordered pool (cudaMallocAsync on the legacy default
// stream with an unlimited release threshold), so grow/shrink never forces a
// device-wide synchronisation the way cudaMalloc/cudaFree do.
#ifndef MEM_MANAGER_H
#define MEM_MANAGER_H

#include <cassert>
#include <cstddef>
#include <iostream>
#include <limits>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace hagrid {

/// Direction of a MemManager::copy
enum class Copy { HST_TO_DEV, DEV_TO_HST, DEV_TO_DEV };

/// One pooled device buffer
struct Slot {
    void*  ptr;
    size_t size;
    bool   in_use;
    Slot() : ptr(nullptr), size(0), in_use(false) {}
};

class MemManager {
public:
    explicit MemManager(bool keep = false) : usage_(0), max_usage_(0), keep_(keep) {}

    /// Returns the memory of every slot to the device's pool (the reference has no destructor and leaks what keep
    /// mode retained, src/mem_manager.h:34-42). Non-virtual, defined in the library: the object layout is unchanged.
    /// The device the buffers live on must be current.
    ~MemManager();
    MemManager(const MemManager&) = delete;
    MemManager& operator=(const MemManager&) = delete;

    /// Device buffer of n elements of T (uninitialised)
    template <typename T>
    HOST T* alloc(size_t n) {
        const size_t bytes = n * sizeof(T);
        // smallest idle slot that already holds `bytes`; failing that the largest idle one (it is grown).
        // Capacities only ever grow, so a repeated build finds every buffer it needs after a few rounds and
        // stops talking to the driver (the reference picks the slot of closest size, src/mem_manager.h:46-70,
        // which keeps re-growing slots that were a little too small).
        int best = -1, largest = -1;
        for (size_t i = 0; i < slots_.size(); i++) {
            if (slots_[i].in_use) continue;
            const size_t have = slots_[i].size;
            if (have >= bytes && (best < 0 || have < slots_[best].size)) best = int(i);
            if (largest < 0 || have > slots_[largest].size) largest = int(i);
        }
        if (best < 0) best = largest;
        if (best < 0) {
            best = int(slots_.size());
            slots_.emplace_back();
        }
        alloc_slot(slots_[best], bytes);
        tracker_[slots_[best].ptr] = best;
        return static_cast<T*>(slots_[best].ptr);
    }

    /// Releases a buffer obtained from alloc(); nullptr is ignored
    template <typename T>
    HOST void free(T* ptr) {
        if (!ptr) return;
        auto it = tracker_.find(const_cast<void*>(static_cast<const void*>(ptr)));
        assert(it != tracker_.end() && "pointer does not belong to this MemManager");
        if (it == tracker_.end()) return;
        free_slot(slots_[it->second]);
        tracker_.erase(it);
    }

    /// Blocking copy of n elements
    template <Copy dir, typename T>
    HOST void copy(T* dst, const T* src, size_t n) {
        const size_t bytes = n * sizeof(T);
        if (dir == Copy::HST_TO_DEV)      copy_hst_to_dev(dst, src, bytes);
        else if (dir == Copy::DEV_TO_HST) copy_dev_to_hst(dst, src, bytes);
        else                              copy_dev_to_dev(dst, src, bytes);
    }

    /// Sets every byte of n elements to 0x00 / 0xFF
    template <typename T> HOST void zero(T* ptr, size_t n) { zero_dev(ptr, n * sizeof(T)); }
    template <typename T> HOST void one(T* ptr, size_t n)  { one_dev(ptr, n * sizeof(T)); }

    void debug_slots() const;

    size_t usage() const { return usage_; }
    size_t max_usage() const { return max_usage_; }

private:
    HOST void alloc_slot(Slot&, size_t);
    HOST void free_slot(Slot&);
    HOST void copy_dev_to_dev(void*, const void*, size_t);
    HOST void copy_dev_to_hst(void*, const void*, size_t);
    HOST void copy_hst_to_dev(void*, const void*, size_t);
    HOST void zero_dev(void*, size_t);
    HOST void one_dev(void*, size_t);

    // Member order is part of the binary interface (src/mem_manager.h:113-116).
    std::unordered_map<void*, int> tracker_;
    std::vector<Slot> slots_;
    size_t usage_, max_usage_;
    bool keep_;
};

} // namespace hagrid
#endif
