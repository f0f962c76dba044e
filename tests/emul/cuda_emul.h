// TEST INFRASTRUCTURE ONLY -- a tiny CUDA-on-CPU emulator.
//
// The build container has nvcc but no GPU, and a GPU run costs minutes of a small budget, so the
// `not gpu` tests compile the SAME kernel sources (matchering_b200/csrc/*.cu) for the host with
// this header force-included (-DMGB_EMULATE) and run them block by block: every CUDA thread of a
// block is a fiber on one OS thread, __syncthreads()/__shfl_*_sync() are cooperative barriers,
// `__shared__` is a thread_local static.  Blocks run in blockIdx order, so decoupled look-back
// kernels always find their predecessors finished.  It checks kernel LOGIC (indexing, barriers,
// scan carries) against the oracle; it proves nothing about performance and is never loaded by
// the product package (matchering_b200/_native.py only ever loads the nvcc-built library).
#pragma once
#ifndef MGB_EMULATE
#error "cuda_emul.h is only for -DMGB_EMULATE host builds"
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// ----------------------------------------------------------------------------- keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) alignas(n)
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(4) short2 { short x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emul {

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
    unsigned tid = 0;
};

struct WarpState {
    uint64_t vals[32];
    unsigned arrived = 0, gen = 0;
    unsigned ballot_acc = 0;
};

struct BlockState {
    unsigned nthreads = 0, alive = 0;
    unsigned arrived = 0, gen = 0;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    std::function<void()> body;
    void* sched_sp = nullptr;
    unsigned current = 0;
    std::vector<unsigned char> dyn_smem;
};

BlockState& block();
void yield();
void launch(dim3 grid, dim3 blockdim, size_t smem_bytes, const std::function<void()>& body);
unsigned char* dyn_smem();

}  // namespace emul

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;

// ----------------------------------------------------------------------------- barriers
static inline void __syncthreads() {
    emul::BlockState& b = emul::block();
    unsigned gen = b.gen;
    if (++b.arrived >= b.alive) {
        b.arrived = 0;
        b.gen++;
    } else {
        while (b.gen == gen) emul::yield();
    }
}

namespace emul {
static inline void warp_barrier(WarpState& w, unsigned expected) {
    unsigned gen = w.gen;
    if (++w.arrived >= expected) {
        w.arrived = 0;
        w.gen++;
    } else {
        while (w.gen == gen) emul::yield();
    }
}
static inline unsigned lane_id() { return threadIdx.x & 31u; }
static inline WarpState& my_warp() { return block().warps[threadIdx.x >> 5]; }
static inline unsigned warp_width() {
    // number of live lanes in this warp (last warp of a block may be partial)
    unsigned base = threadIdx.x & ~31u;
    unsigned n = block().nthreads - base;
    return n > 32 ? 32 : n;
}
template <typename T>
static inline T shfl_generic(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    WarpState& w = my_warp();
    unsigned expected = std::min<unsigned>(__builtin_popcount(mask), warp_width());
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.vals[lane_id()] = raw;
    warp_barrier(w, expected);
    uint64_t got = w.vals[(unsigned)src_lane & 31u];
    warp_barrier(w, expected);
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
}  // namespace emul

static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    emul::warp_barrier(emul::my_warp(), std::min<unsigned>(__builtin_popcount(mask), emul::warp_width()));
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int lane = (int)emul::lane_id();
    int base = lane & ~(width - 1);
    return emul::shfl_generic(mask, v, base + (src & (width - 1)));
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = (int)emul::lane_id();
    int base = lane & ~(width - 1);
    int src = lane - (int)delta;
    if (src < base) src = lane;
    return emul::shfl_generic(mask, v, src);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = (int)emul::lane_id();
    int base = lane & ~(width - 1);
    int src = lane + (int)delta;
    if (src >= base + width) src = lane;
    return emul::shfl_generic(mask, v, src);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    int lane = (int)emul::lane_id();
    (void)width;
    return emul::shfl_generic(mask, v, lane ^ lanemask);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned bit = pred ? (1u << emul::lane_id()) : 0u;
    unsigned acc = 0;
    for (int l = 0; l < 32; ++l) {
        unsigned b = __shfl_sync(mask, bit, l);
        if ((mask >> l) & 1u) acc |= (b & (1u << l));
    }
    return acc;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
    unsigned live = mask;
    unsigned w = emul::warp_width();
    if (w < 32) live &= ((1u << w) - 1u);
    return (__ballot_sync(mask, pred) & live) == live;
}

// ----------------------------------------------------------------------------- memory / atomics
static inline void __threadfence() { __sync_synchronize(); }
static inline void __threadfence_block() { __sync_synchronize(); }
static inline void __nanosleep(unsigned) { emul::yield(); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcg(const T* p) { return *(const volatile T*)p; }

static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
static inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }
static inline int atomicCAS(int* p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }

// ----------------------------------------------------------------------------- math intrinsics
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __double2float_rn(double d) { return (float)d; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline void sincospi(double x, double* s, double* c) {
    // exact quadrant reduction so that multiples of 1/2 give exact 0/+-1 like CUDA's sincospi
    double r = std::fmod(x, 2.0);
    if (r < 0) r += 2.0;
    int q = (int)std::floor(r * 2.0 + 0.5);  // nearest multiple of 1/2
    double t = r - 0.5 * q;
    double st = std::sin(M_PI * t), ct = std::cos(M_PI * t);
    switch (q & 3) {
        case 0: *s = st; *c = ct; break;
        case 1: *s = ct; *c = -st; break;
        case 2: *s = -st; *c = -ct; break;
        default: *s = -ct; *c = st; break;
    }
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
using std::fabs;
using std::fmax;
using std::fmin;
using std::sqrt;
using std::max;
using std::min;
static inline float fabsf_(float x) { return std::fabs(x); }
