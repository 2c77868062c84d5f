// slab_pool.h — storage of the sketches themselves (host-side bookkeeping; no CUDA in this header).
//
// Sketch arrays live as long as their handles, so they cannot sit in the grow-only scratch arena; the stream-ordered pool
// (cudaMallocAsync) turned out to need 20-150 ms per 100 MB batch while it grows (against ~1-100 ms for a 1 GB cudaMalloc).
// So: slabs from a raw allocator (cudaMalloc in the library; malloc in tests/test_slab_pool_host.py), 64 MB doubling to
// 1 GB, or the request if larger; bump allocation inside the current slab; one live-count per slab.  A slab whose count
// returns to zero is reused from the start; freeing the most recent allocation rolls the bump pointer back, which is what
// the query-sketch-per-call pattern of Database.query produces.  All users of the memory are ordered on the context's
// stream, so reuse needs no event.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

namespace skb {

struct Slab { char* base = nullptr; size_t cap = 0, used = 0; uint32_t live = 0; };

// RawAlloc: void* operator()(size_t bytes) -> nullptr on failure;  RawFree: void operator()(void*)
template <typename RawAlloc, typename RawFree>
struct SlabPoolT {
    static constexpr size_t GRANULE = 512;                  // every allocation is a multiple of this (and so aligned)
    static constexpr size_t FIRST_SLAB = (size_t)64 << 20, MAX_SLAB = (size_t)1 << 30;
    std::mutex mu;
    std::vector<std::unique_ptr<Slab>> slabs;
    Slab* cur = nullptr;
    size_t next_cap = FIRST_SLAB;
    RawAlloc raw_alloc;
    RawFree raw_free;

    static size_t round_up(size_t bytes) { return (bytes + GRANULE - 1) & ~(GRANULE - 1); }

    void* alloc(size_t bytes, Slab** owner) {
        bytes = round_up(bytes);
        std::lock_guard<std::mutex> lock(mu);
        if (!cur || cur->used + bytes > cur->cap) {
            Slab* pick = nullptr;                            // smallest idle slab that fits, else a new one
            for (auto& sl : slabs) if (sl->live == 0 && sl->cap >= bytes && (!pick || sl->cap < pick->cap)) pick = sl.get();
            if (!pick) {
                std::unique_ptr<Slab> sl(new Slab);
                sl->cap = std::max(next_cap, bytes);
                sl->base = (char*)raw_alloc(sl->cap);
                if (!sl->base && sl->cap > bytes) { sl->cap = bytes; sl->base = (char*)raw_alloc(sl->cap); }
                if (!sl->base) return nullptr;
                next_cap = std::min(next_cap * 2, MAX_SLAB);
                pick = sl.get();
                slabs.push_back(std::move(sl));
            }
            pick->used = 0;
            cur = pick;
        }
        void* p = cur->base + cur->used;
        cur->used += bytes; cur->live++;
        *owner = cur;
        return p;
    }
    void free(Slab* sl, void* p, size_t bytes) {
        bytes = round_up(bytes);
        std::lock_guard<std::mutex> lock(mu);
        if ((char*)p + bytes == sl->base + sl->used) sl->used -= bytes;      // most recent allocation: roll back
        if (--sl->live == 0) sl->used = 0;                                     // idle slab: reusable from the start
    }
    void destroy() {
        for (auto& sl : slabs) if (sl->base) raw_free(sl->base);
        slabs.clear(); cur = nullptr;
    }
    size_t reserved_bytes() const { size_t s = 0; for (auto& sl : slabs) s += sl->cap; return s; }
};

}  // namespace skb
