// host_pack.h — ingest compaction on the host side of skb_sketch_batch (CUDA-free header).
//
// The ASCII contigs a caller hands to Database.sketch / Database.query (reference lib.rs:140-185 borrows them as &[u8])
// reach a B200 over PCIe at ~55 GB/s, which is 5x slower than the seeding kernel consumes them.  For large host batches
// the library therefore shrinks part of the bytes BEFORE the link: worker threads translate 16 bases into one 2-bit word
// (the very word layout kmer_bits.cuh::pack16 produces on the device: A=0 C=1 G=2 T=3, any other byte 0, first base in
// the most significant pair) straight into pinned staging memory, and the copy engine moves a quarter of the bytes, while
// the remaining chunks travel as plain ASCII by DMA at the same time (skb_api.cu: IngestPipeline).
// This is a change of transport encoding only: hashing, thresholds, ordering, index build, screen, chaining and ANI all
// stay on the device, and the device-side pack of the ASCII path (skb_sketch_batch_device, small calls) is unchanged.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace skb {

// words[i] = bases 16 i .. 16 i + 15 of src (missing bases of the last word encode as 0).  n_bases may be 0.
// Results are globally visible when the call returns (the streaming stores are fenced).
// stream = true: streaming (non-temporal) stores, for output that is written once and read by the copy engine much later;
// stream = false: ordinary stores, for output blocks that are reused and should stay in the cache.
void host_pack_bases(const uint8_t* src, size_t n_bases, uint32_t* words, bool stream = true);
// which implementation host_pack_bases dispatches to on this CPU: "avx512vbmi", "avx2" or "scalar"
const char* host_pack_isa();
// CPUs this process may run on (sched_getaffinity), at least 1
unsigned host_cpu_count();

// A fixed team of threads; launch() runs fn(worker) once on every member, wait() joins the round.
// One round at a time (the context's mutex serialises callers).
class HostTeam {
public:
    explicit HostTeam(unsigned n_threads, std::function<void()> thread_init = nullptr);
    ~HostTeam();
    unsigned size() const { return (unsigned)workers_.size(); }
    void launch(std::function<void(unsigned)> fn);
    void wait();
private:
    void run(unsigned id);
    std::vector<std::thread> workers_;
    std::function<void()> thread_init_;
    std::function<void(unsigned)> fn_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    uint64_t epoch_ = 0;
    unsigned pending_ = 0;
    bool stop_ = false;
};

}  // namespace skb
