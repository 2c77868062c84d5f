// Host build of pyskani_b200/csrc/slab_pool.h with malloc as the raw allocator (tests/test_slab_pool_host.py).
#include <cstdlib>
#include <map>
#include "../../pyskani_b200/csrc/slab_pool.h"

namespace {
struct CountingAlloc {
    static long calls, live_bytes, fail_above;
    void* operator()(size_t n) const {
        if (fail_above >= 0 && (long)n > fail_above) return nullptr;
        calls++; live_bytes += (long)n;
        return std::aligned_alloc(512, (n + 511) / 512 * 512);      // cudaMalloc returns at least 512-byte aligned blocks
    }
};
long CountingAlloc::calls = 0, CountingAlloc::live_bytes = 0, CountingAlloc::fail_above = -1;
struct CountingFree { void operator()(void* p) const { std::free(p); } };
using Pool = skb::SlabPoolT<CountingAlloc, CountingFree>;
struct Handle { skb::Slab* slab; void* p; size_t bytes; };
}  // namespace

extern "C" {
void* pool_new() { CountingAlloc::calls = 0; CountingAlloc::live_bytes = 0; CountingAlloc::fail_above = -1; return new Pool(); }
void pool_delete(void* h) { auto* p = (Pool*)h; p->destroy(); delete p; }
void pool_fail_above(long bytes) { CountingAlloc::fail_above = bytes; }
long pool_raw_calls() { return CountingAlloc::calls; }
unsigned long long pool_reserved(void* h) { return ((Pool*)h)->reserved_bytes(); }
unsigned long long pool_slabs(void* h) { return ((Pool*)h)->slabs.size(); }
// returns an opaque handle (NULL on failure); *addr receives the address, *slab_index the slab it came from
void* pool_alloc(void* h, unsigned long long bytes, unsigned long long* addr, long* slab_index) {
    auto* pool = (Pool*)h;
    skb::Slab* sl = nullptr;
    void* p = pool->alloc((size_t)bytes, &sl);
    if (!p) return nullptr;
    *addr = (unsigned long long)(uintptr_t)p;
    *slab_index = -1;
    for (size_t i = 0; i < pool->slabs.size(); i++) if (pool->slabs[i].get() == sl) *slab_index = (long)i;
    return new Handle{sl, p, (size_t)bytes};
}
void pool_free(void* h, void* handle) {
    auto* hd = (Handle*)handle;
    ((Pool*)h)->free(hd->slab, hd->p, hd->bytes);
    delete hd;
}
unsigned long long slab_used(void* h, long i) { return ((Pool*)h)->slabs[(size_t)i]->used; }
unsigned long long slab_cap(void* h, long i) { return ((Pool*)h)->slabs[(size_t)i]->cap; }
unsigned slab_live(void* h, long i) { return ((Pool*)h)->slabs[(size_t)i]->live; }
}
