// Host transport at the reference's real seam: stages.main / limiter.limit take PAGEABLE float64 numpy
// arrays (what soundfile hands matchering/core.py:53-62) and return float64 arrays.
//
// Pageable memory cannot be DMA'd at link speed (the driver stages it through one internal bounce
// buffer, ~10 GB/s), float64 doubles the bytes, and cudaHostRegister of a 127 MB array costs more than
// the whole job.  So the library moves the data itself:
//
//   upload    P worker threads narrow (float64 -> float32) or copy the caller's array, slice by slice,
//             into a ring of pinned chunks; the calling thread issues one cudaMemcpyAsync per finished
//             chunk and recycles a chunk when its copy event has fired.  Conversion of chunk k+1 overlaps
//             the DMA of chunk k; half the bytes cross the link.
//   download  into pinned float64 memory (mgb_host_alloc): the device widens, ONE DMA, no host pass.
//             into anything else: float32 chunks through the ring, widened / copied by the workers.
//
// The worker pool is persistent (threads sleep on a condition variable between transfers and spin on
// atomics inside one).  Nothing here touches sample VALUES except the float64 <-> float32 conversion
// the reference's caller would otherwise pay for on the device.
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "kernels.cuh"

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define MGB_CPU_RELAX() _mm_pause()
#else
#define MGB_CPU_RELAX() std::this_thread::yield()
#endif

#ifdef MGB_EMULATE
// The test emulator has no DMA engine: a copy is a memcpy at issue time and every event has fired.
// The ring, the worker threads and their hand-shakes below run exactly as on the device build.
#include <map>
typedef int cudaEvent_t;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = 1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static std::mutex g_pinned_mutex;
static std::map<const unsigned char*, size_t> g_pinned_blocks;  // what mgb_host_alloc / the ring handed out
static inline cudaError_t cudaHostAlloc(void** p, size_t n, int) {
    *p = aligned_alloc(256, (n + 255) / 256 * 256);
    std::lock_guard<std::mutex> lk(g_pinned_mutex);
    g_pinned_blocks[(const unsigned char*)*p] = n;
    return *p ? cudaSuccess : 1;
}
static inline cudaError_t cudaFreeHost(void* p) {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mutex);
        g_pinned_blocks.erase((const unsigned char*)p);
    }
    free(p);
    return cudaSuccess;
}
#endif

namespace mgb {
namespace {

class WorkerPool {
public:
    explicit WorkerPool(int n) : n_(n) {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    // every worker runs fn(worker_index) once; `lead` runs on the calling thread meanwhile
    void run(const std::function<void(int)>& fn, const std::function<void()>& lead) {
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &fn;
            pending_ = n_;
            ++generation_;
        }
        cv_.notify_all();
        lead();
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int index) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                job = job_;
            }
            (*job)(index);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* job_ = nullptr;
    unsigned long long generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

inline void spin_until(const std::function<bool()>& ready) {
    int spins = 0;
    while (!ready()) {
        if (++spins < 4096)
            MGB_CPU_RELAX();
        else
            std::this_thread::yield();  // an oversubscribed host must not burn its time slices here
    }
}

// the two conversions (auto-vectorised: cvtpd2ps / cvtps2pd)
inline void narrow_plain(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}
#if defined(__x86_64__)
// Non-temporal variant: the staging chunk is written once and read once by the DMA engine; a streaming store
// skips the read-for-ownership of every destination line (a third of the ring's memory traffic).  dst is 64-byte
// aligned by construction (slices start on 16-sample boundaries of a pinned allocation).
__attribute__((target("avx2"))) inline void narrow_stream(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    int64_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
        for (; i + 8 <= n; i += 8) {
            const __m128 lo = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i));
            const __m128 hi = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4));
            _mm256_stream_ps(dst + i, _mm256_set_m128(hi, lo));
        }
        _mm_sfence();
    }
    for (; i < n; ++i) dst[i] = (float)src[i];
}
static const bool g_stream_stores = [] {
    const char* e = getenv("MGB_HOST_NT");
    return e && atoi(e) != 0 && __builtin_cpu_supports("avx2");
}();
inline void narrow(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    if (g_stream_stores) narrow_stream(src, dst, n);
    else narrow_plain(src, dst, n);
}
#else
inline void narrow(const double* __restrict__ src, float* __restrict__ dst, int64_t n) { narrow_plain(src, dst, n); }
#endif
inline void widen(const float* __restrict__ src, double* __restrict__ dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = (double)src[i];
}

}  // namespace
}  // namespace mgb

struct mgb_host_io {
    mgb::WorkerPool* pool = nullptr;
    int64_t chunk = 0;  // samples per ring chunk
    int ring = 0;
    float* staging = nullptr;  // pinned, ring * chunk floats
    std::vector<cudaEvent_t> events;
};

using namespace mgb;

namespace {

#ifdef MGB_EMULATE
bool is_pinned(const void* p) {
    std::lock_guard<std::mutex> lk(g_pinned_mutex);
    auto it = g_pinned_blocks.upper_bound((const unsigned char*)p);
    if (it == g_pinned_blocks.begin()) return false;
    --it;
    return (const unsigned char*)p < it->first + it->second;
}
#else
bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
#endif

// host (float32 or float64, any memory) -> device float32
int upload(mgb_host_io* io, const void* h_src, int src_width, float* d_dst, int64_t samples, cudaStream_t st) {
    MGB_REQUIRE(src_width == 4 || src_width == 8, MGB_ERR_INVALID, "host array must be float32 or float64");
    if (samples == 0) return MGB_OK;
    if (src_width == 4 && is_pinned(h_src)) {  // nothing to convert, DMA-able as it is
        if (cudaMemcpyAsync(d_dst, h_src, (size_t)samples * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return cuda_status("H2D");
        return MGB_OK;
    }
    const int64_t chunk = io->chunk;
    const int64_t nchunks = (samples + chunk - 1) / chunk;
    const int P = io->pool->size();
    const int ring = io->ring;
    std::vector<std::atomic<int>> ready(nchunks);
    for (auto& r : ready) r.store(0, std::memory_order_relaxed);
    std::atomic<int64_t> released{0};  // chunks whose DMA has finished (their ring slot is free again)
    std::atomic<int> failed{0};
    auto work = [&](int p) {
        for (int64_t k = 0; k < nchunks; ++k) {
            if (k >= ring) spin_until([&] { return released.load(std::memory_order_acquire) > k - ring || failed.load(); });
            if (failed.load()) return;
            const int64_t base = k * chunk;
            const int64_t len = (samples - base < chunk) ? samples - base : chunk;
            // slices on 64-byte boundaries of the destination
            const int64_t per = ((len + P - 1) / P + 15) / 16 * 16;
            const int64_t lo = (int64_t)p * per, hi = lo + per < len ? lo + per : len;
            float* dst = io->staging + (k % ring) * chunk;
            if (lo < hi) {
                if (src_width == 8) narrow((const double*)h_src + base + lo, dst + lo, hi - lo);
                else memcpy(dst + lo, (const float*)h_src + base + lo, (size_t)(hi - lo) * 4);
            }
            ready[k].fetch_add(1, std::memory_order_release);
        }
    };
    int rc = MGB_OK;
    auto lead = [&]() {
        int64_t issued = 0, freed = 0;
        std::vector<int64_t> group_end(nchunks);  // chunk k left the ring when the copy that ends with chunk group_end[k] has
        while (issued < nchunks) {
            if (ready[issued].load(std::memory_order_acquire) == P) {
                // every finished chunk that follows in the ring without wrapping goes into the same copy: when the workers
                // are ahead of this thread (large transfers) one launch moves several chunks
                int64_t last = issued;
                while (last + 1 < nchunks && (last + 1) % ring != 0 && ready[last + 1].load(std::memory_order_acquire) == P) ++last;
                const int64_t base = issued * chunk;
                const int64_t end = (last + 1) * chunk < samples ? (last + 1) * chunk : samples;
                const int slot = (int)(issued % ring);
                if (cudaMemcpyAsync(d_dst + base, io->staging + slot * chunk, (size_t)(end - base) * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                    cudaEventRecord(io->events[last % ring], st) != cudaSuccess) {
                    rc = cuda_status("H2D chunk");
                    failed.store(1);
                    return;
                }
                for (int64_t k = issued; k <= last; ++k) group_end[k] = last;
                issued = last + 1;
            } else if (freed < issued && cudaEventQuery(io->events[group_end[freed] % ring]) == cudaSuccess) {
                freed = group_end[freed] + 1;
                released.store(freed, std::memory_order_release);
            } else {
                MGB_CPU_RELAX();
            }
        }
        // the ring is reused by the next transfer: its last copies must have left it
        if (cudaEventSynchronize(io->events[(nchunks - 1) % ring]) != cudaSuccess) rc = cuda_status("H2D drain");
    };
    io->pool->run(work, lead);
    return rc;
}

// device float32 -> host (float32 or float64)
int download(mgb_host_io* io, const float* d_src, void* h_dst, int dst_width, int64_t samples, double* d_wide,
             cudaStream_t st) {
    MGB_REQUIRE(dst_width == 4 || dst_width == 8, MGB_ERR_INVALID, "host array must be float32 or float64");
    if (samples == 0) return MGB_OK;
    // float64 results into pinned memory: widened on the device and copied by ONE DMA (2.4 ms for a 3-minute
    // track at 54 GB/s).  Float32 chunks through the ring, widened by the workers, move half the bytes over the
    // link and measured 1.5-1.8 ms for that track when the host was otherwise idle -- but 50 % SLOWER on the
    // one-hour limiter buffer (2.5 GB: the widening competes with itself for the socket's memory bandwidth), so
    // the DMA route, which does not depend on host threads at all, is the default; MGB_DOWNLOAD_RING=1 switches.
    static const bool prefer_ring = getenv("MGB_DOWNLOAD_RING") && atoi(getenv("MGB_DOWNLOAD_RING")) != 0;
    const bool direct = is_pinned(h_dst) && (dst_width == 4 || (d_wide && !prefer_ring));
    if (direct) {
        const void* src = d_src;
        if (dst_width == 8) {
            MGB_TRY(launch_convert_f32_f64(d_src, d_wide, samples, st));
            src = d_wide;
        }
        if (cudaMemcpyAsync(h_dst, src, (size_t)samples * dst_width, cudaMemcpyDeviceToHost, st) != cudaSuccess) return cuda_status("D2H");
        if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_status("D2H sync");
        return MGB_OK;
    }
    const int64_t chunk = io->chunk;
    const int64_t nchunks = (samples + chunk - 1) / chunk;
    const int P = io->pool->size();
    const int ring = io->ring;
    std::vector<std::atomic<int>> consumed(nchunks);
    for (auto& c : consumed) c.store(0, std::memory_order_relaxed);
    std::atomic<int64_t> arrived{0};
    std::atomic<int> failed{0};
    auto work = [&](int p) {
        for (int64_t k = 0; k < nchunks; ++k) {
            spin_until([&] { return arrived.load(std::memory_order_acquire) > k || failed.load(); });
            if (failed.load()) return;
            const int64_t base = k * chunk;
            const int64_t len = (samples - base < chunk) ? samples - base : chunk;
            const int64_t per = ((len + P - 1) / P + 15) / 16 * 16;
            const int64_t lo = (int64_t)p * per, hi = lo + per < len ? lo + per : len;
            const float* src = io->staging + (k % ring) * chunk;
            if (lo < hi) {
                if (dst_width == 8) widen(src + lo, (double*)h_dst + base + lo, hi - lo);
                else memcpy((float*)h_dst + base + lo, src + lo, (size_t)(hi - lo) * 4);
            }
            consumed[k].fetch_add(1, std::memory_order_release);
        }
    };
    int rc = MGB_OK;
    auto lead = [&]() {
        int64_t issued = 0, landed = 0;
        while (landed < nchunks) {
            const bool slot_free = issued < ring || consumed[issued - ring].load(std::memory_order_acquire) == P;
            if (issued < nchunks && slot_free) {
                const int64_t base = issued * chunk;
                const int64_t len = (samples - base < chunk) ? samples - base : chunk;
                const int slot = (int)(issued % ring);
                if (cudaMemcpyAsync(io->staging + slot * chunk, d_src + base, (size_t)len * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaEventRecord(io->events[slot], st) != cudaSuccess) {
                    rc = cuda_status("D2H chunk");
                    failed.store(1);
                    return;
                }
                ++issued;
            } else if (landed < issued && cudaEventQuery(io->events[landed % ring]) == cudaSuccess) {
                arrived.store(++landed, std::memory_order_release);
            } else {
                MGB_CPU_RELAX();
            }
        }
    };
    io->pool->run(work, lead);  // returns when every worker has consumed every chunk
    return rc;
}

}  // namespace

extern "C" {

int mgb_host_io_create(int32_t threads, int64_t chunk_samples, int32_t ring, mgb_host_io** out) {
    MGB_REQUIRE(out != nullptr, MGB_ERR_INVALID, "host_io: NULL argument");
    if (threads <= 0) {
        // Sixteen workers keep up with the link (each narrows 8-10 GB/s of source, the link takes 108 GB/s of
        // it; 32 were no faster on the 128-thread B200 host) -- fewer when the process may not use that many cores:
        // the affinity mask, and the cgroup's CPU quota (the B200 boxes grant 16 cores per GPU; workers that
        // spin past the quota get the whole process throttled).
        int usable = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) usable = CPU_COUNT(&set);
        long long quota = 0, period = 0;
        if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2: "<quota|max> <period>"
            if (fscanf(f, "%lld %lld", &quota, &period) != 2) quota = period = 0;
            fclose(f);
        } else if (FILE* q = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {  // cgroup v1
            FILE* p = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
            if (fscanf(q, "%lld", &quota) != 1) quota = 0;
            if (!p || fscanf(p, "%lld", &period) != 1) period = 0;
            fclose(q);
            if (p) fclose(p);
        }
        if (quota > 0 && period > 0) {
            const int cores = (int)(quota / period);
            if (cores >= 1 && cores < usable) usable = cores;
        }
#endif
        threads = usable >= 20 ? 16 : (usable > 5 ? usable - 4 : (usable > 1 ? usable - 1 : 1));
    }
    // A ring of 16 chunks of 256 KB: small enough to stay in the cores' caches between the workers' stores and the
    // DMA engine's reads.  Measured with four ranks on one socket (tools/gpu_n4_sweep.sh): 7.8 ms per
    // stages.main call against 14.3 ms with twelve 1 MB chunks, whose stores and re-reads went through DRAM and
    // took a third of the socket's memory bandwidth; chunks of 64 KB cost more in copy launches than they save.
    if (chunk_samples <= 0) chunk_samples = 1 << 16;
    if (ring <= 0) ring = 16;
    MGB_REQUIRE(threads <= 256 && ring <= 64 && chunk_samples % 16 == 0, MGB_ERR_INVALID, "host_io: bad geometry");
    mgb_host_io* io = new mgb_host_io();
    io->chunk = chunk_samples;
    io->ring = ring;
    if (cudaHostAlloc((void**)&io->staging, (size_t)ring * chunk_samples * 4, cudaHostAllocPortable) != cudaSuccess) {
        delete io;
        return cuda_status("host_io: pinned staging ring");
    }
    io->events.resize(ring);
    for (auto& e : io->events)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            mgb_host_io_destroy(io);
            return cuda_status("host_io: events");
        }
    io->pool = new WorkerPool(threads);
    *out = io;
    return MGB_OK;
}

int mgb_host_io_destroy(mgb_host_io* io) {
    if (!io) return MGB_OK;
    delete io->pool;
    for (auto e : io->events)
        if (e) cudaEventDestroy(e);
    if (io->staging) cudaFreeHost(io->staging);
    delete io;
    return MGB_OK;
}

int mgb_host_io_threads(const mgb_host_io* io) { return io && io->pool ? io->pool->size() : 0; }

void* mgb_host_alloc(int64_t bytes) {
    if (bytes <= 0) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        cuda_status("host_alloc");
        return nullptr;
    }
    return p;
}

void mgb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int mgb_host_upload(mgb_host_io* io, const void* h_src, int32_t src_width, float* d_dst, int64_t samples, void* stream) {
    MGB_REQUIRE(io && h_src && d_dst && samples >= 0, MGB_ERR_INVALID, "host_upload: bad arguments");
    return upload(io, h_src, src_width, d_dst, samples, (cudaStream_t)stream);
}

int mgb_host_download(mgb_host_io* io, const float* d_src, void* h_dst, int32_t dst_width, int64_t samples, double* d_wide,
                      void* stream) {
    MGB_REQUIRE(io && h_dst && d_src && samples >= 0, MGB_ERR_INVALID, "host_download: bad arguments");
    return download(io, d_src, h_dst, dst_width, samples, d_wide, (cudaStream_t)stream);
}

int mgb_stages_main_host(mgb_host_io* io, const mgb_plan* plan, const mgb_track_layout* L, const void* h_target,
                         const void* h_reference, int32_t in_width, void* h_out_limited, void* h_out_no_limiter,
                         void* h_out_normalized, int32_t out_width, const mgb_host_buffers* dev,
                         mgb_track_state* h_state_out, void* stream) {
    MGB_REQUIRE(io && plan && L && h_target && h_reference && dev, MGB_ERR_INVALID, "stages_main_host: NULL argument");
    MGB_REQUIRE(h_out_limited || h_out_no_limiter || h_out_normalized, MGB_ERR_INVALID, "no output requested");
    MGB_REQUIRE(dev->d_target_lr && dev->d_reference_lr && dev->d_result_lr && dev->d_out_lr && dev->d_workspace && dev->d_state,
                MGB_ERR_INVALID, "stages_main_host: device staging buffer missing");
    cudaStream_t st = (cudaStream_t)stream;
    MGB_TRY(upload(io, h_target, in_width, dev->d_target_lr, L->target_frames * 2, st));
    MGB_TRY(upload(io, h_reference, in_width, dev->d_reference_lr, L->reference_frames * 2, st));
    MGB_TRY(mgb_match_levels(plan, L, dev->d_target_lr, dev->d_reference_lr, dev->d_workspace, dev->d_state, stream));
    MGB_TRY(mgb_match_frequencies(plan, L, dev->d_target_lr, dev->d_result_lr, nullptr, dev->d_workspace, dev->d_state, stream));
    MGB_TRY(mgb_correct_levels(plan, L, dev->d_workspace, dev->d_state, stream));
    void* outs[3] = {h_out_limited, h_out_no_limiter, h_out_normalized};
    for (int k = 0; k < 3; ++k) {
        if (!outs[k]) continue;
        MGB_TRY(mgb_finalize(plan, L, dev->d_result_lr, k == 0 ? dev->d_out_lr : nullptr, k == 1 ? dev->d_out_lr : nullptr,
                             k == 2 ? dev->d_out_lr : nullptr, dev->d_workspace, dev->d_state, stream));
        MGB_TRY(download(io, dev->d_out_lr, outs[k], out_width, L->target_frames * 2, dev->d_wide, st));
    }
    if (h_state_out && cudaMemcpyAsync(h_state_out, dev->d_state, sizeof(mgb_track_state), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return cuda_status("D2H state");
    if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_status("stream sync");
    return MGB_OK;
}

int mgb_limit_host(mgb_host_io* io, const mgb_limiter_params* params, const void* h_in, int32_t in_width, void* h_out,
                   int32_t out_width, int64_t frames, float* d_in_lr, float* d_out_lr, double* d_wide, void* d_workspace,
                   int64_t workspace_bytes, int32_t* d_engaged, int32_t* h_engaged_out, void* stream) {
    MGB_REQUIRE(io && params && h_in && h_out && d_in_lr && d_out_lr && d_engaged && h_engaged_out, MGB_ERR_INVALID,
                "limit_host: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    MGB_TRY(upload(io, h_in, in_width, d_in_lr, frames * 2, st));
    MGB_TRY(mgb_limit(params, d_in_lr, d_out_lr, frames, d_workspace, workspace_bytes, d_engaged, stream));
    if (cudaMemcpyAsync(h_engaged_out, d_engaged, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return cuda_status("D2H flag");
    if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_status("stream sync");
    if (*h_engaged_out == 0) return MGB_OK;  // hyrax.py:83-85: the caller hands its input back untouched
    return download(io, d_out_lr, h_out, out_width, frames * 2, d_wide, st);
}

}  // extern "C"
