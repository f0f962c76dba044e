// Fixed-frame mixed-radix FFT in shared memory (Stockham autosort, radix 2/4/8/16 butterflies in
// registers).  One frame of N complex points lives in two padded shared arrays (re / im); each
// pass is: every thread gathers the R inputs of its butterflies into registers, barrier, twiddle +
// radix-R DFT in registers, scatter in place, barrier.  The first pass may read from and the last
// pass may write to caller-supplied functors (e.g. the TMA landing buffer, a magnitude
// accumulator), so frames never make an extra trip through shared memory.
//
// Twiddles come from a table computed once per FFT size in double precision (fft_twiddle_kernel)
// and rounded to the working type: table layout is the concatenation over passes of [R-1][NS]
// entries w = exp(-2*pi*i * r*k / (NS*R)), r = 1..R-1, k = 0..NS-1, so that consecutive lanes read
// consecutive entries.  The inverse transform conjugates on load.
#pragma once
#include "common.cuh"

namespace mgb {

// padded index of the split (re / im) layout, see SplitPlanes below
__device__ __host__ __forceinline__ constexpr int fft_pad(int i) { return i + (i >> 5); }
__device__ __host__ constexpr int fft_padded_size(int n) { return n + (n >> 5) + 1; }

// ------------------------------------------------------------------------------------------------
// compile-time roots of unity for the in-register butterflies (multiples of 2*pi/16, and of 2*pi/32)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ constexpr T cos16(int k) {
    // cos(2*pi*k/16), k = 0..8
    return k == 0 ? T(1) : k == 1 ? T(0.92387953251128673848) : k == 2 ? T(0.70710678118654752440)
         : k == 3 ? T(0.38268343236508977173) : k == 4 ? T(0) : k == 5 ? T(-0.38268343236508977173)
         : k == 6 ? T(-0.70710678118654752440) : k == 7 ? T(-0.92387953251128673848) : T(-1);
}
template <typename T>
__device__ __forceinline__ constexpr T sin16(int k) {
    // sin(2*pi*k/16), k = 0..8
    return k == 0 ? T(0) : k == 1 ? T(0.38268343236508977173) : k == 2 ? T(0.70710678118654752440)
         : k == 3 ? T(0.92387953251128673848) : k == 4 ? T(1) : k == 5 ? T(0.92387953251128673848)
         : k == 6 ? T(0.70710678118654752440) : k == 7 ? T(0.38268343236508977173) : T(0);
}

// the odd 32nd roots (radix-32 butterflies of the wide convolution kernel); even ones are the 16th roots
template <typename T>
__device__ __forceinline__ constexpr T cos32(int k) {
    // cos(2*pi*k/32), k = 0..16
    return (k & 1) == 0 ? cos16<T>(k / 2)
         : k == 1 ? T(0.98078528040323044913) : k == 3 ? T(0.83146961230254523708) : k == 5 ? T(0.55557023301960222474)
         : k == 7 ? T(0.19509032201612826785) : k == 9 ? T(-0.19509032201612826785) : k == 11 ? T(-0.55557023301960222474)
         : k == 13 ? T(-0.83146961230254523708) : T(-0.98078528040323044913);
}
template <typename T>
__device__ __forceinline__ constexpr T sin32(int k) {
    // sin(2*pi*k/32), k = 0..16
    return (k & 1) == 0 ? sin16<T>(k / 2)
         : k == 1 ? T(0.19509032201612826785) : k == 3 ? T(0.55557023301960222474) : k == 5 ? T(0.83146961230254523708)
         : k == 7 ? T(0.98078528040323044913) : k == 9 ? T(0.98078528040323044913) : k == 11 ? T(0.83146961230254523708)
         : k == 13 ? T(0.55557023301960222474) : T(0.19509032201612826785);
}

// multiply by exp(DIR * -2*pi*i * q / R), q < R/2, with the trivial cases folded away
template <int R, int Q, int DIR, typename T>
__device__ __forceinline__ cpx<T> mul_root(cpx<T> a) {
    constexpr int K = Q * (32 / R);  // index into the 32nd roots, 0..15
    if constexpr (K == 0) {
        return a;
    } else if constexpr (K == 8) {
        // forward: * (-i) ; inverse: * (+i)
        return DIR > 0 ? cpx<T>{a.y, -a.x} : cpx<T>{-a.y, a.x};
    } else if constexpr (K == 4) {
        constexpr T c = cos16<T>(2);
        return DIR > 0 ? cpx<T>{c * (a.x + a.y), c * (a.y - a.x)} : cpx<T>{c * (a.x - a.y), c * (a.y + a.x)};
    } else if constexpr (K == 12) {
        constexpr T c = cos16<T>(2);
        return DIR > 0 ? cpx<T>{c * (a.y - a.x), -c * (a.x + a.y)} : cpx<T>{-c * (a.x + a.y), c * (a.x - a.y)};
    } else {
        constexpr T wr = cos32<T>(K);
        constexpr T wi = DIR > 0 ? -sin32<T>(K) : sin32<T>(K);
        return cpx<T>{a.x * wr - a.y * wi, a.x * wi + a.y * wr};
    }
}

// radix-R DFT of v[0..R) in registers, natural order in and out (recursive decimation in time)
template <int R, int DIR, typename T>
struct Dft {
    template <int Q>
    static __device__ __forceinline__ void combine(cpx<T>* v, const cpx<T>* e, const cpx<T>* o) {
        if constexpr (Q < R / 2) {
            cpx<T> t = mul_root<R, Q, DIR, T>(o[Q]);
            v[Q] = cadd(e[Q], t);
            v[Q + R / 2] = csub(e[Q], t);
            combine<Q + 1>(v, e, o);
        }
    }
    static __device__ __forceinline__ void run(cpx<T>* v) {
        cpx<T> e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) {
            e[i] = v[2 * i];
            o[i] = v[2 * i + 1];
        }
        Dft<R / 2, DIR, T>::run(e);
        Dft<R / 2, DIR, T>::run(o);
        combine<0>(v, e, o);
    }
};
template <int DIR, typename T>
struct Dft<2, DIR, T> {
    static __device__ __forceinline__ void run(cpx<T>* v) {
        cpx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <int DIR, typename T>
struct Dft<1, DIR, T> {
    static __device__ __forceinline__ void run(cpx<T>*) {}
};

// Where a frame lives in shared memory between passes.
// SplitPlanes: two padded arrays (re / im), one extra word per 32 -- used by the float64 transforms.
// PackedPlanes: one padded float2 array, one extra element per 16 -- used by the float32 kernels:
// a point is one 8-byte access instead of two 4-byte ones, which halves the shared-memory
// instructions of kernels that are bound by instruction issue.  Both paddings keep the stride-R
// scatters of the early passes and the unit-stride gathers conflict-free.
template <typename T>
struct SplitPlanes {
    T* re;
    T* im;
    static constexpr int kPadEvery = 32;  // one extra element per kPadEvery points
    static __device__ __host__ constexpr int pad(int i) { return i + (i >> 5); }
    __device__ __forceinline__ cpx<T> load_at(int a) const { return cpx<T>{re[a], im[a]}; }  // a = pad(i)
    __device__ __forceinline__ void store_at(int a, cpx<T> v) const {
        re[a] = v.x;
        im[a] = v.y;
    }
    static __device__ __host__ constexpr int elems(int n) { return n + (n >> 5) + 1; }
    static __device__ __host__ constexpr size_t bytes(int n) { return 2 * (size_t)elems(n) * sizeof(T); }
    __device__ __forceinline__ cpx<T> load(int i) const {
        const int a = pad(i);
        return cpx<T>{re[a], im[a]};
    }
    __device__ __forceinline__ void store(int i, cpx<T> v) const {
        const int a = pad(i);
        re[a] = v.x;
        im[a] = v.y;
    }
};
struct PackedPlanes {
    float2* z;
    static constexpr int kPadEvery = 16;
    static __device__ __host__ constexpr int pad(int i) { return i + (i >> 4); }
    __device__ __forceinline__ cpx<float> load_at(int a) const {  // a = pad(i)
        const float2 v = z[a];
        return cpx<float>{v.x, v.y};
    }
    __device__ __forceinline__ void store_at(int a, cpx<float> v) const { z[a] = make_float2(v.x, v.y); }
    static __device__ __host__ constexpr int elems(int n) { return n + (n >> 4) + 1; }
    static __device__ __host__ constexpr size_t bytes(int n) { return (size_t)elems(n) * sizeof(float2); }
    __device__ __forceinline__ cpx<float> load(int i) const {
        const float2 v = z[pad(i)];
        return cpx<float>{v.x, v.y};
    }
    __device__ __forceinline__ void store(int i, cpx<float> v) const { z[pad(i)] = make_float2(v.x, v.y); }
};
template <typename P>
struct PlaneLoad {
    P p;
    __device__ __forceinline__ auto operator()(int i) const { return p.load(i); }
};
template <typename P>
struct PlaneStore {
    P p;
    template <typename V>
    __device__ __forceinline__ void operator()(int i, V v) const { p.store(i, v); }
};

// Strided accesses of a pass, with the padding taken out of the per-access arithmetic.  pad(i) = i + i/Q, and
// for a stride D that is a multiple of Q, pad(i + r*D) = pad(i) + r*(D + D/Q) exactly (the low bits of i never
// carry into a multiple of Q): one padded base per butterfly, the rest is an immediate offset of the LDS / STS.
// Left to itself the compiler forms every index and pads it again: three integer instructions per access
// (LOP3, LEA.HI, LEA), ~700 of the 16384-point convolution frame's ~3000 instructions per thread.
// Functors that are not planes (landing buffers, accumulators) keep the plain call.
template <typename Fn> struct PlaneOf { static constexpr int every = 0; };
template <typename P> struct PlaneOf<PlaneLoad<P>> { static constexpr int every = P::kPadEvery; };
template <typename P> struct PlaneOf<PlaneStore<P>> { static constexpr int every = P::kPadEvery; };

// v[r] = load(i0 + r*D), r < R
template <int R, int D, typename T, typename Load>
__device__ __forceinline__ void strided_load(Load load, int i0, cpx<T>* v) {
    constexpr int Q = PlaneOf<Load>::every;
    if constexpr (Q > 0 && D % (Q > 0 ? Q : 1) == 0) {
        const int a0 = decltype(load.p)::pad(i0);
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = load.p.load_at(a0 + r * (D + D / (Q > 0 ? Q : 1)));
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = load(i0 + r * D);
    }
}
// store(i0 + q*D, v[q]), q < R.  D == 1 needs i0 to be a multiple of R (R dividing the padding period).
template <int R, int D, typename T, typename Store>
__device__ __forceinline__ void strided_store(Store store, int i0, const cpx<T>* v) {
    constexpr int Q = PlaneOf<Store>::every;
    if constexpr (Q > 0 && (D % (Q > 0 ? Q : 1) == 0 || (D == 1 && (Q > 0 ? Q : 1) % R == 0))) {
        const int a0 = decltype(store.p)::pad(i0);
        constexpr int step = D == 1 ? 1 : D + D / (Q > 0 ? Q : 1);
#pragma unroll
        for (int q = 0; q < R; ++q) store.p.store_at(a0 + q * step, v[q]);
    } else {
#pragma unroll
        for (int q = 0; q < R; ++q) store(i0 + q * D, v[q]);
    }
}

// ------------------------------------------------------------------------------------------------
// one Stockham pass
// ------------------------------------------------------------------------------------------------
// Twiddles of one butterfly, w[r] = w1^r for r = 1..R-1.  The float kernels fetch only w1 (and w4
// for radix 16) from the table and build the other powers in registers (products at most four
// deep, ~2e-7 relative): a radix-16 butterfly then costs 2 table reads instead of 15, which
// matters because the table does not fit in what is left of L1 next to the frame.  The double
// kernels (FIR design) read every power from the table.
template <int R, int NS, int DIR, bool CHAIN, typename T>
__device__ __forceinline__ void load_twiddles(const cpx<T>* __restrict__ tw, int k, cpx<T>* w) {
    if constexpr (!CHAIN || R == 2) {
#pragma unroll
        for (int r = 1; r < R; ++r) {
            w[r] = tw[(r - 1) * NS + k];
            if constexpr (DIR < 0) w[r].y = -w[r].y;
        }
    } else {
        w[1] = tw[k];
        if constexpr (DIR < 0) w[1].y = -w[1].y;
        w[2] = cmul(w[1], w[1]);
        w[3] = cmul(w[1], w[2]);
        if constexpr (R >= 8) {
            if constexpr (R >= 16) {
                w[4] = tw[3 * NS + k];
                if constexpr (DIR < 0) w[4].y = -w[4].y;
            } else {
                w[4] = cmul(w[2], w[2]);
            }
            w[5] = cmul(w[4], w[1]);
            w[6] = cmul(w[4], w[2]);
            w[7] = cmul(w[4], w[3]);
        }
        if constexpr (R >= 16) {
            w[8] = cmul(w[4], w[4]);
#pragma unroll
            for (int r = 9; r < 16; ++r) w[r] = cmul(w[8], w[r - 8]);
        }
        if constexpr (R == 32) {
            w[16] = tw[15 * NS + k];
            if constexpr (DIR < 0) w[16].y = -w[16].y;
#pragma unroll
            for (int r = 17; r < 32; ++r) w[r] = cmul(w[16], w[r - 16]);
        }
    }
}

template <int N, int R, int NS, int DIR, int THREADS, typename T, bool CHAIN, typename Load, typename Store>
__device__ __forceinline__ void fft_pass(const cpx<T>* __restrict__ tw, Load load, Store store, bool barrier_between) {
    constexpr int NB = N / R;
    static_assert(NB % THREADS == 0 || NB < THREADS, "butterflies must tile the block");
    constexpr int PER = (NB >= THREADS) ? NB / THREADS : 1;
    const int tid = threadIdx.x;
    const bool active = (NB >= THREADS) || (tid < NB);
    cpx<T> v[PER][R];
    if (active) {
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int j = tid + p * THREADS;
            strided_load<R, NB>(load, j, v[p]);
        }
    }
    if (barrier_between) __syncthreads();
    if (active) {
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int j = tid + p * THREADS;
            const int k = j & (NS - 1);
            if constexpr (NS > 1) {
                if constexpr (CHAIN && R > 2) {
                    cpx<T> w[R];
                    load_twiddles<R, NS, DIR, true, T>(tw, k, w);
#pragma unroll
                    for (int r = 1; r < R; ++r) v[p][r] = cmul(v[p][r], w[r]);
                } else {
#pragma unroll
                    for (int r = 1; r < R; ++r) {
                        cpx<T> w = tw[(r - 1) * NS + k];
                        if constexpr (DIR < 0) w.y = -w.y;
                        v[p][r] = cmul(v[p][r], w);
                    }
                }
            }
            Dft<R, DIR, T>::run(v[p]);
            const int j0 = (j - k) * R + k;  // (NS == 1: k = 0, j0 = j*R)
            strided_store<R, NS>(store, j0, v[p]);
        }
    }
}

// Radix schedule of an N-point transform.  Sizes outside this list are rejected by the C ABI.
template <int N> struct Radices;
template <> struct Radices<256>   { static constexpr int n = 2; static constexpr int r[4] = {16, 16, 1, 1}; };
template <> struct Radices<512>   { static constexpr int n = 3; static constexpr int r[4] = {16, 8, 4, 1}; };  // radix 16 first: 16 points per thread, like every other size
template <> struct Radices<1024>  { static constexpr int n = 3; static constexpr int r[4] = {16, 8, 8, 1}; };
template <> struct Radices<2048>  { static constexpr int n = 3; static constexpr int r[4] = {16, 16, 8, 1}; };
template <> struct Radices<4096>  { static constexpr int n = 3; static constexpr int r[4] = {16, 16, 16, 1}; };
template <> struct Radices<8192>  { static constexpr int n = 4; static constexpr int r[4] = {16, 8, 8, 8}; };
template <> struct Radices<16384> { static constexpr int n = 4; static constexpr int r[4] = {16, 16, 8, 8}; };
template <> struct Radices<32768> { static constexpr int n = 4; static constexpr int r[4] = {16, 16, 16, 8}; };  // fft_size 16384's convolution frame (planes in global memory)


// First pass of the transform: `first` supplies the input points (logical index -> value), results
// go to the planes.  `in_place` says whether `first` reads the planes itself (then a barrier
// separates the gathers from the scatters).  No trailing barrier.
template <int N, int DIR, int THREADS, typename T, typename Planes, typename First>
__device__ __forceinline__ void fft_first_pass(Planes pl, const cpx<T>* __restrict__ tw, First first, bool in_place) {
    fft_pass<N, Radices<N>::r[0], 1, DIR, THREADS, T, false>(tw, first, PlaneStore<Planes>{pl}, in_place);  // NS = 1: no twiddles
}

// First pass when the caller already holds the thread's R = N/THREADS input points in registers
// (v[r] = point tid + r*THREADS), e.g. because it had to look at them before transforming.
template <int N, int DIR, int THREADS, typename T, typename Planes>
__device__ __forceinline__ void fft_first_pass_regs(Planes pl, cpx<T>* v, bool barrier_before_store) {
    constexpr int R = Radices<N>::r[0];
    static_assert(N / R == THREADS, "one butterfly per thread in the first pass");
    if (barrier_before_store) __syncthreads();
    Dft<R, DIR, T>::run(v);
    strided_store<R, 1>(PlaneStore<Planes>{pl}, threadIdx.x * R, v);  // NS = 1
}

// Remaining passes, in place on the planes; the caller has put a barrier after the first pass.
// `last` consumes the output points in natural order (`last_in_place`: it writes the planes).
// On return all `last` stores have been ISSUED (no trailing barrier).
template <int N, int DIR, int THREADS, typename T, bool CHAIN = false, typename Planes, typename Last>
__device__ __forceinline__ void fft_remaining(Planes pl, const cpx<T>* __restrict__ tw, Last last, bool last_in_place) {
    using Rd = Radices<N>;
    constexpr int R0 = Rd::r[0], R1 = Rd::r[1], R2 = Rd::r[2], R3 = Rd::r[3];
    constexpr int NP = Rd::n;
    PlaneLoad<Planes> sl{pl};
    PlaneStore<Planes> ss{pl};
    static_assert(NP >= 2 && NP <= 4, "2..4 passes supported");
    constexpr int o1 = 0;  // pass 0 has NS = 1: no twiddles stored
    if constexpr (NP == 2) {
        fft_pass<N, R1, R0, DIR, THREADS, T, CHAIN>(tw + o1, sl, last, last_in_place);
    } else {
        fft_pass<N, R1, R0, DIR, THREADS, T, CHAIN>(tw + o1, sl, ss, true);
        __syncthreads();
        constexpr int o2 = o1 + (R1 - 1) * R0;
        if constexpr (NP == 3) {
            fft_pass<N, R2, R0 * R1, DIR, THREADS, T, CHAIN>(tw + o2, sl, last, last_in_place);
        } else {
            fft_pass<N, R2, R0 * R1, DIR, THREADS, T, CHAIN>(tw + o2, sl, ss, true);
            __syncthreads();
            constexpr int o3 = o2 + (R2 - 1) * R0 * R1;
            fft_pass<N, R3, R0 * R1 * R2, DIR, THREADS, T, CHAIN>(tw + o3, sl, last, last_in_place);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pieces for kernels that keep the ends of a transform in registers (convolve.cu): the passes between
// the first and the last one, and a single butterfly without its scatter.
// ------------------------------------------------------------------------------------------------
// Inverse schedules of the fused convolution: they START with the radix-8 pass whose inputs are exactly
// the outputs of the forward schedule's last radix-8 pass (points j + r*N/8), and end with a radix-8
// pass of two butterflies per thread whose outputs are contiguous across the block.
template <int N> struct InverseRadices { static constexpr bool fused = false; };
template <> struct InverseRadices<8192>  { static constexpr bool fused = true; static constexpr int n = 4; static constexpr int r[4] = {8, 16, 8, 8}; };
template <> struct InverseRadices<16384> { static constexpr bool fused = true; static constexpr int n = 4; static constexpr int r[4] = {8, 16, 16, 8}; };

template <typename Rd>
__host__ __device__ constexpr int fft_schedule_twiddles() {
    int total = 0, ns = Rd::r[0];
    for (int p = 1; p < Rd::n; ++p) {
        total += (Rd::r[p] - 1) * ns;
        ns *= Rd::r[p];
    }
    return total;
}
// NS (transform length already built) of the schedule's last pass, and the offset of that pass's
// twiddles inside the schedule's table
template <typename Rd>
__host__ __device__ constexpr int fft_last_pass_ns() {
    int ns = 1;
    for (int p = 0; p < Rd::n - 1; ++p) ns *= Rd::r[p];
    return ns;
}
template <typename Rd>
__host__ __device__ constexpr int fft_last_pass_twiddles() {
    return fft_schedule_twiddles<Rd>() - (Rd::r[Rd::n - 1] - 1) * fft_last_pass_ns<Rd>();
}

// Passes 1 .. n-2 of schedule Rd, in place, a barrier after each (the caller has put one after pass 0).
template <int N, int DIR, int THREADS, typename T, bool CHAIN, typename Rd, typename Planes>
__device__ __forceinline__ void fft_middle(Planes pl, const cpx<T>* __restrict__ tw) {
    constexpr int R0 = Rd::r[0], R1 = Rd::r[1], R2 = Rd::r[2];
    PlaneLoad<Planes> sl{pl};
    PlaneStore<Planes> ss{pl};
    static_assert(Rd::n >= 2 && Rd::n <= 4, "2..4 passes supported");
    if constexpr (Rd::n >= 3) {
        fft_pass<N, R1, R0, DIR, THREADS, T, CHAIN>(tw, sl, ss, true);
        __syncthreads();
    }
    if constexpr (Rd::n >= 4) {
        fft_pass<N, R2, R0 * R1, DIR, THREADS, T, CHAIN>(tw + (R1 - 1) * R0, sl, ss, true);
        __syncthreads();
    }
}

// The R inputs of butterfly j of a pass with NB = N/R butterflies.
template <int R, int NB, typename T, typename Load>
__device__ __forceinline__ void fft_gather(Load load, int j, cpx<T>* v) {
    strided_load<R, NB>(load, j, v);
}
// Twiddle (index k = j mod NS) and radix-R DFT of one gathered butterfly; outputs belong at
// (j - k)*R + k + q*NS.
template <int R, int NS, int DIR, bool CHAIN, typename T>
__device__ __forceinline__ void fft_butterfly(const cpx<T>* __restrict__ tw, int k, cpx<T>* v) {
    if constexpr (NS > 1) {
        if constexpr (CHAIN && R > 2) {
            cpx<T> w[R];
            load_twiddles<R, NS, DIR, true, T>(tw, k, w);
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
        } else {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                cpx<T> w = tw[(r - 1) * NS + k];
                if constexpr (DIR < 0) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
    }
    Dft<R, DIR, T>::run(v);
}

// Whole transform: first pass, barrier, remaining passes.
template <int N, int DIR, int THREADS, typename T, bool CHAIN = false, typename Planes, typename First, typename Last>
__device__ __forceinline__ void fft_run(Planes pl, const cpx<T>* __restrict__ tw, First first, Last last,
                                        bool first_in_place, bool last_in_place) {
    fft_first_pass<N, DIR, THREADS, T>(pl, tw, first, first_in_place);
    __syncthreads();
    fft_remaining<N, DIR, THREADS, T, CHAIN>(pl, tw, last, last_in_place);
}

// Fill the twiddle table of a transform with the given radix schedule (double-precision sincospi,
// rounded to T).
template <typename T>
__global__ void fft_twiddle_kernel(cpx<T>* __restrict__ tw, int npass, int r0, int r1, int r2, int r3) {
    const int radix[4] = {r0, r1, r2, r3};
    int ns = radix[0];
    int base = 0;
    for (int p = 1; p < npass; ++p) {
        const int R = radix[p];
        const int count = (R - 1) * ns;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
            const int r = i / ns + 1, k = i % ns;
            double s, c;
            sincospi(-2.0 * (double)(r * k) / (double)(ns * R), &s, &c);  // exp(-2*pi*i*r*k/(ns*R))
            tw[base + i] = cpx<T>{(T)c, (T)s};
        }
        base += count;
        ns *= R;
    }
}

}  // namespace mgb
