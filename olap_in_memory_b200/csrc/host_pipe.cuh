// Host <-> device copies of PAGEABLE host memory at PCIe speed (in-memory.js:30-46: `get data` / `set data`
// hand typed arrays across the boundary; a Node addon sees them as plain pageable memory).
//
// cudaMemcpy of pageable memory goes through the driver's own staging buffer on ONE thread (8-12 GB/s, and the
// first touch of a fresh destination buffer is paid on that thread too).  Here a transfer is cut into chunks that
// travel through a ring of pinned buffers: the PCIe copy of chunk k is asynchronous on the library's stream while
// the host-side memcpy of chunk k+1 (upload) or k-1 (download) runs on a small pool of worker threads, each
// taking a slice of the chunk.  Pinned (cudaHostAlloc / olap_host_alloc) and small buffers take the direct copy.
#pragma once

#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace olap {

class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool* pool = new CopyPool();  // never destroyed: the workers end with the process (and a forked
        return *pool;                            // child, which has none of them, must not try to join them)
    }
    int threads() const { return n_threads_; }
    // memcpy split into one slice per thread (the caller takes the first slice); returns when all are done
    void copy(void* dst, const void* src, size_t bytes) {
        int parts = (int)std::min<size_t>((size_t)n_threads_, std::max<size_t>(1, bytes >> 20));  // >= 1 MiB per slice
        if (getpid() != pid_) parts = 1;  // a forked child has no worker threads: copy on the calling thread
        if (parts <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = static_cast<char*>(dst);
            src_ = static_cast<const char*>(src);
            bytes_ = bytes;
            slice_ = slice;
            parts_ = parts;
            next_ = 1;
            pending_ = parts - 1;
            ++generation_;
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(slice, bytes));
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        const char* e = getenv("OLAP_COPY_THREADS");
        int n = e ? atoi(e) : 0;
        if (n <= 0) n = (int)std::min<unsigned>(8u, std::max<unsigned>(1u, std::thread::hardware_concurrency() / 2));  // 16 vCPUs: 4 threads 24-30 GB/s, 8 threads 26-39
        n_threads_ = std::max(1, std::min(n, 32));
        pid_ = getpid();
        for (int t = 1; t < n_threads_; ++t) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    void run() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || (generation_ != seen && next_ < parts_); });
            if (stop_) return;
            // take slices of this generation until none is left
            while (next_ < parts_) {
                const int part = next_++;
                const size_t off = (size_t)part * slice_;
                const size_t n = off < bytes_ ? std::min(slice_, bytes_ - off) : 0;
                char* d = dst_ + off;
                const char* s = src_ + off;
                lk.unlock();
                if (n) memcpy(d, s, n);
                lk.lock();
                if (--pending_ == 0) done_.notify_one();
            }
            seen = generation_;
        }
    }
    int n_threads_ = 1;
    pid_t pid_ = 0;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0;
    int parts_ = 0, next_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

struct HostPipe {
    static constexpr int kSlots = 4;
    size_t chunk = 0;
    char* pin[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool busy[kSlots] = {false, false, false, false};
};
inline HostPipe g_pipe;

inline int host_pipe_ready() {
    if (g_pipe.chunk) return OLAP_OK;
    const char* e = getenv("OLAP_PIPE_CHUNK_MB");
    const size_t mb = e && atoi(e) > 0 ? (size_t)atoi(e) : 16;
    for (int k = 0; k < HostPipe::kSlots; ++k) {
        OLAP_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_pipe.pin[k]), mb << 20, cudaHostAllocDefault));
        OLAP_CUDA(cudaEventCreateWithFlags(&g_pipe.ev[k], cudaEventDisableTiming));
    }
    g_pipe.chunk = mb << 20;
    return OLAP_OK;
}

// Is this host pointer plain pageable memory, and the transfer large enough to be worth the ring?
inline bool wants_pipe(const void* host, size_t bytes) {
    static const int knob = [] { const char* e = getenv("OLAP_HOST_PIPE"); return e ? atoi(e) : 1; }();
    if (!knob || bytes < ((size_t)8 << 20)) return false;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// host -> device.  Asynchronous towards the device (ordered on g.stream); the host buffer is free on return.
inline int copy_h2d(void* dev, const void* host, size_t bytes) {
    if (!wants_pipe(host, bytes)) {
        OLAP_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, g.stream));
        return OLAP_OK;
    }
    OLAP_TRY(host_pipe_ready());
    CopyPool& pool = CopyPool::get();
    const size_t chunk = g_pipe.chunk;
    int slot = 0;
    for (size_t off = 0; off < bytes; off += chunk, slot = (slot + 1) % HostPipe::kSlots) {
        const size_t n = std::min(chunk, bytes - off);
        if (g_pipe.busy[slot]) OLAP_CUDA(cudaEventSynchronize(g_pipe.ev[slot]));  // its previous copy has left the buffer
        pool.copy(g_pipe.pin[slot], static_cast<const char*>(host) + off, n);
        OLAP_CUDA(cudaMemcpyAsync(static_cast<char*>(dev) + off, g_pipe.pin[slot], n, cudaMemcpyHostToDevice, g.stream));
        OLAP_CUDA(cudaEventRecord(g_pipe.ev[slot], g.stream));
        g_pipe.busy[slot] = true;
    }
    return OLAP_OK;
}

// device -> host; complete (data in `host`) on return.
inline int copy_d2h(void* host, const void* dev, size_t bytes) {
    if (!wants_pipe(host, bytes)) {
        OLAP_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, g.stream));
        OLAP_CUDA(cudaStreamSynchronize(g.stream));
        return OLAP_OK;
    }
    OLAP_TRY(host_pipe_ready());
    CopyPool& pool = CopyPool::get();
    const size_t chunk = g_pipe.chunk;
    const int64_t n_chunks = (int64_t)((bytes + chunk - 1) / chunk);
    constexpr int K = HostPipe::kSlots;
    for (int k = 0; k < K; ++k)
        if (g_pipe.busy[k]) { OLAP_CUDA(cudaEventSynchronize(g_pipe.ev[k])); g_pipe.busy[k] = false; }
    for (int64_t i = 0; i < n_chunks + K - 1; ++i) {
        if (i < n_chunks) {  // chunk i leaves the device into slot i % K (drained K - 1 iterations later)
            const size_t off = (size_t)i * chunk, n = std::min(chunk, bytes - off);
            OLAP_CUDA(cudaMemcpyAsync(g_pipe.pin[i % K], static_cast<const char*>(dev) + off, n, cudaMemcpyDeviceToHost, g.stream));
            OLAP_CUDA(cudaEventRecord(g_pipe.ev[i % K], g.stream));
        }
        const int64_t j = i - (K - 1);
        if (j >= 0) {
            const size_t off = (size_t)j * chunk, n = std::min(chunk, bytes - off);
            OLAP_CUDA(cudaEventSynchronize(g_pipe.ev[j % K]));
            pool.copy(static_cast<char*>(host) + off, g_pipe.pin[j % K], n);
        }
    }
    return OLAP_OK;
}

}  // namespace olap
