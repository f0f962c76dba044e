// K6 -- the Hyrax brickwall limiter as one chained-scan kernel.
//
// Replaces (reference file:line):
//   dsp.rectify, dsp.flip, dsp.max_mix            matchering/dsp.py:113-125
//   limiter.__sliding_window_fast                 matchering/limiter/hyrax.py:32-40
//   limiter.__process_attack                      matchering/limiter/hyrax.py:43-53
//     (scipy.signal.filtfilt of a one-pole: odd extension by 6, steady-state initial state)
//   limiter.__process_release                     matchering/limiter/hyrax.py:56-75
//     (two scipy.signal.lfilter first-order Butterworth sections, zero initial state)
//   limiter.limit                                 matchering/limiter/hyrax.py:78-99
//   the result scaling dsp.amplify(result, final_amplitude_coefficient)   stages.py:203
//
// One CTA masters one chunk of kLimiterCore samples, chunks are handed out in order by a ticket:
//   g      = 1 - thr/max(|L|,|R|,thr)                       (float64 math, kept as float32)
//   A      = centred running max of g over +-reach          (log-step doubling in shared memory)
//   g_att  = backward(forward(A)) one-pole, float64 state   (blocked scan; the pole is fast, so a
//            `warmup`-sample halo on both sides replaces cross-chunk carries to < 1e-10)
//   H      = trailing running max of A over `hold`
//   hold   = IIR1(H), rel = IIR1(max(H, hold))              (float64 blocked scans whose carries
//            cross chunks by decoupled look-back: aggregate first, inclusive when known)
//   out    = x * (1 - max(g, g_att, hold, rel)) * post_gain
// HBM traffic: 8 B/frame read (+ the halo, which hits L2) and 8 B/frame written.
#include "kernels.cuh"

namespace mgb {

namespace {

constexpr int NT = kLimiterThreads;
constexpr int CORE_EPT = kLimiterCoreEpt;
constexpr int LC = kLimiterCore;
constexpr int SPAN_EPT_MAX = kLimiterSpanEptMax;

// Exclusive carry of the recurrence y = u + p*y_prev across the block: given each thread's local
// end value B (zero initial state over its elements), returns the state just before the thread's
// first element when the state before the block's first element is c0.
// One barrier: warps scan their own 32 values with shuffles, publish the warp totals, and then EVERY
// warp scans the (at most 32) totals itself instead of waiting for one warp to do it.
// scratch: >= 32 doubles, and must not be the buffer the previous call used (callers alternate
// between two), which is what makes the leading "scratch is free again" barrier unnecessary.
// Every thread of the block must call.
__device__ __forceinline__ double scan_carry(double B, const ScanPow* t, double c0, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = B;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t->ql[d] * up;
    }
    if (lane == 31) scratch[warp] = v;
    __syncthreads();
    double w = lane < NT / 32 ? scratch[lane] : 0.0;
#pragma unroll
    for (int d = 1; d < NT / 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t->qw[d] * up;
    }
    // state at the end of the previous warp (zero block carry): inclusive total of warps 0..warp-1
    const double warp_carry = __shfl_sync(0xffffffffu, w, (warp + 31) & 31);
    double prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) prev = 0.0;
    return prev + t->ql[lane] * ((warp > 0 ? warp_carry : 0.0) + t->qw[warp] * c0);
}

// The same for a recurrence that runs from the block's LAST element to its first (thread NT-1 first):
// B is the thread's local end value after its elements were taken in descending order, *c0 the state
// just beyond the block's last element (read after the barrier, so the caller may have written it to
// shared memory just before the call); returns the state just beyond the thread's last element.
__device__ __forceinline__ double scan_carry_rev(double B, const ScanPow* t, const double* c0, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = 31 - lane, rw = NT / 32 - 1 - warp;  // ranks in processing order
    double v = B;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_down_sync(0xffffffffu, v, d);
        if (rl >= d) v += t->ql[d] * up;
    }
    if (lane == 0) scratch[rw] = v;
    __syncthreads();
    double w = lane < NT / 32 ? scratch[lane] : 0.0;
#pragma unroll
    for (int d = 1; d < NT / 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t->qw[d] * up;
    }
    const double warp_carry = __shfl_sync(0xffffffffu, w, (rw + 31) & 31);
    double prev = __shfl_down_sync(0xffffffffu, v, 1);
    if (rl == 0) prev = 0.0;
    return prev + t->ql[rl] * ((rw > 0 ? warp_carry : 0.0) + t->qw[rw] * (*c0));
}

// 16-byte publish / poll of a LookbackWord through L2 (st.cg / ld.cg: coherent device-wide).
__device__ __forceinline__ void publish(LookbackWord* w, double v, int status) {
#ifdef MGB_EMULATE
    w->value = v;
    w->status = status;
#else
    asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"(__double_as_longlong(v)), "l"((long long)status) : "memory");
#endif
}
__device__ __forceinline__ int poll(const LookbackWord* w, double* v) {
#ifdef MGB_EMULATE
    *v = w->value;
    return (int)w->status;
#else
    long long a, b;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(w) : "memory");
    *v = __longlong_as_double(a);
    return (int)b;
#endif
}

// Carry into chunk `chunk` of a first-order recurrence whose per-chunk multiplier is P = pc[1]
// (decoupled look-back).  One warp inspects 32 predecessors at a time: every lane polls one
// predecessor's word, the window is cut at the nearest predecessor whose INCLUSIVE state is known,
// each lane weighs its value by P^distance, and one warp reduction at the very end adds them up.
// The walk stops when whatever lies further back weighs less than 1e-9 (all values are gains in
// [0, 1], so that bounds the absolute error of the carry).  Called by all 32 lanes of one warp.
__device__ __forceinline__ double lookback(LookbackSlot* slots, int chunk, bool release, const double* pc) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0, mult = 1.0;
    for (int base = chunk - 1; base >= 0 && mult > 1e-9; base -= 32) {
        const int j = base - lane;
        int st = 2;  // before the first chunk: inclusive state 0 (lfilter starts from rest)
        double val = 0.0;
        if (j >= 0) {
            const LookbackWord* w = release ? &slots[j].rel : &slots[j].hold;
            while ((st = poll(w, &val)) == 0) __nanosleep(20);
        }
        const unsigned inclusive = __ballot_sync(0xffffffffu, st == 2);
        const int first = inclusive ? __ffs((int)inclusive) - 1 : 32;
        if (lane <= first) acc += mult * pc[lane] * val;
        if (inclusive) break;
        mult *= pc[32];
    }
    return warp_sum(acc);
}

struct LimiterGeom {
    int reach, hold, warm, left;  // left = max(warm, hold): left halo of the envelope A
    int ept;                      // span elements per thread (odd)
    int span;                     // samples of g the chunk touches
    int filt;                     // samples the attack filter needs to run over (LC + left + warm)
    int publish_inclusive;        // 0: chunks publish aggregates only (test switch: every look-back then walks to the cut-off)
    int shared_core;              // both windows are wide enough for the per-thread shared-core evaluation
    int margin;                   // zeros kept on both sides of G so that window reads need no bounds test (multiple of 4)
};

// powers of the three poles, computed once per parameter set (not per CTA: pow() is slow)
__global__ void limiter_tables_kernel(mgb_limiter_params lp, int span_ept, ScanPow* tables) {
    const int i = threadIdx.x;
    for (int f = 0; f < 3; ++f) {
        const double p = f == 0 ? lp.attack_c : (f == 1 ? -lp.hold_a1 : -lp.release_a1);
        const int ept = f == 0 ? span_ept : CORE_EPT;
        ScanPow* t = tables + f;
        if (i < SPAN_EPT_MAX + 2) t->pe[i] = pow(p, (double)i);
        if (i < 33) t->ql[i] = pow(p, (double)(ept * i));
        if (i < 17) t->qw[i] = pow(p, (double)(ept * 32 * i));
        if (i < 33) t->pc[i] = pow(p, (double)LC * (double)i);
    }
}

template <int EPT>
__global__ void __launch_bounds__(NT, 2)
limiter_kernel(mgb_limiter_params lp, LimiterGeom gm, const float2* __restrict__ in, float2* __restrict__ out,
               long long frames, const double* __restrict__ pre_gain, const double* __restrict__ post_gain,
               const int* __restrict__ engaged, int* __restrict__ ticket, LookbackSlot* __restrict__ slots,
               const ScanPow* __restrict__ tables) {
    constexpr int CAP = EPT * NT;
    MGB_DYN_SMEM(smem);
    double* Fd = reinterpret_cast<double*>(smem);                 // [CAP] float64 work plane
    float* Aenv = reinterpret_cast<float*>(smem) + CAP;           // [CAP] attack envelope, aliases Fd's upper half
    // [margin zeros][CAP hard-clip gain, later max(g, g_att)][margin zeros]: window samples outside the span read as 0
    float* G = reinterpret_cast<float*>(smem + (size_t)CAP * 8) + gm.margin;
    float* Wk = G + CAP + gm.margin;                              // [CAP] suffix maxima, later the hold envelope
    __shared__ float blockmax[NT];
    __shared__ ScanPow pw3[3];
    __shared__ double scratch_a[32], scratch_b[32];  // scan_carry alternates between them
    __shared__ double bcast[2];
    __shared__ double warp_edge[NT / 32];
    __shared__ int chunk_s;
    const ScanPow* pow_att = &pw3[0];
    const ScanPow* pow_hold = &pw3[1];
    const ScanPow* pow_rel = &pw3[2];

    const int tid = threadIdx.x;
    const double pre = pre_gain ? *pre_gain : 1.0;
    const double post = post_gain ? *post_gain : 1.0;
    const double scale = pre * post;

    if (tid == 0) chunk_s = atomicAdd(ticket, 1);
    {
        const double* src = reinterpret_cast<const double*>(tables);
        double* dst = reinterpret_cast<double*>(pw3);
        for (int i = tid; i < (int)(3 * sizeof(ScanPow) / sizeof(double)); i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const int chunk = chunk_s;
    const long long s0 = (long long)chunk * LC;
    const int core_n = (int)((s0 + LC < frames) ? LC : frames - s0);

    if (engaged && *engaged == 0) {  // hyrax.py:83-85: the limiter is not needed
        for (int k = tid; k < core_n; k += NT) {
            const float2 v = in[s0 + k];
            out[s0 + k] = make_float2((float)((double)v.x * pre * post), (float)((double)v.y * pre * post));
        }
        return;
    }

    const int reach = gm.reach, hold = gm.hold, HL = gm.left, FL = gm.filt;
    const long long ga = s0 - HL - reach;  // sample at span index 0
    const int cidx = HL + reach;           // span index of the chunk's first sample
    // span indices that fall inside the signal: [vlo, vhi)
    const int vlo = ga < 0 ? (int)(-ga) : 0;
    const int vhi = (frames - ga < (long long)gm.span) ? (int)(frames - ga) : gm.span;
    const double thr = lp.threshold;

    // ---- P1: hard-clip gain g = 1 - thr/max(|L|,|R|,thr) over the span (dsp.py:117-121, hyrax.py:87)
    for (int i = tid; i < gm.margin; i += NT) {
        G[-1 - i] = 0.0f;
        G[CAP + i] = 0.0f;
    }
    {
        const float2* base = in + ga;
        float2 v[EPT];  // all of the thread's loads are issued before the first use: one DRAM latency, not EPT
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * NT;
            v[k] = (i >= vlo && i < vhi) ? __ldg(base + i) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            // g = 1 - thr/a = (a - thr)/a: the difference in float64 (it decides which frames are touched at
            // all, and cancels when a is close to thr), the quotient in float32 -- g is kept as float32 anyway
            const double a = (double)fmaxf(fabsf(v[k].x), fabsf(v[k].y)) * pre;
            const double over = a - thr;
            float g = 0.0f;
            if (over > 0.0) g = __fdiv_rn((float)over, (float)a);
            G[tid + k * NT] = g;
        }
    }
    __syncthreads();

    // ---- P2: both running maxima of g ----------------------------------------------------------------
    //   A[n] = max g[n-reach .. n+reach]                                 (hyrax.py:35-37)
    //   H[n] = max A[n-hold+1 .. n] = max g[n-hold+1-reach .. n+reach]   (hyrax.py:38-40)
    // Every thread owns EPT consecutive samples: their prefix and suffix maxima inside the block
    // (PF, SF) and the block maximum.  A window [l, r] is max(SF[l], whole blocks between, PF[r]).
    float hc[CORE_EPT + 1];  // H at span indices cidx + tid*CORE_EPT - 1 + e
    {
        float* PF = reinterpret_cast<float*>(smem);  // [CAP] lower half of Fd's bytes (Aenv is the upper half)
        float* SF = Wk;                              // [CAP]
        float x[EPT];
        const int base = tid * EPT;
#pragma unroll
        for (int e = 0; e < EPT; ++e) x[e] = G[base + e];
        float run = 0.0f;  // g >= 0
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            run = fmaxf(run, x[e]);
            PF[base + e] = run;
        }
        blockmax[tid] = run;
        run = 0.0f;
#pragma unroll
        for (int e = EPT - 1; e >= 0; --e) {
            run = fmaxf(run, x[e]);
            SF[base + e] = run;
        }
        __syncthreads();
        auto whole_blocks = [&](int a, int b) -> float {  // blocks a..b inclusive (a run of at most ~a dozen)
            float m = 0.0f;
            for (int q = a; q <= b; ++q) m = fmaxf(m, blockmax[q]);
            return m;
        };
        auto window = [&](int l, int r, int br, float pr) -> float {
            const int bl = l / EPT;
            if (bl == br) {  // shorter than a block: scan it
                float m = 0.0f;
                for (int i = l; i <= r; ++i) m = fmaxf(m, G[i]);
                return m;
            }
            float m = fmaxf(SF[l], pr);
            if (br - bl > 1) m = fmaxf(m, whole_blocks(bl + 1, br - 1));
            return m;
        };
        auto gat = [&](int i) -> float { return G[i]; };  // (outside the span: the zero margins; g >= 0, so 0 is neutral)
        // The thread's consecutive windows of one kind share most of their samples: for the EPT windows
        // [b+e-reach, b+e+reach] (b = base) the core [b+EPT-1-reach, b+reach] does not depend on e; what
        // lies left and right of it are EPT-1 samples each, combined by running maxima in registers.
        // One table query per thread and kind instead of one per sample.
        if (gm.shared_core) {
            float la[EPT], rt[EPT];
#pragma unroll
            for (int e = 0; e < EPT - 1; ++e) {
                la[e] = gat(base - reach + e);
                rt[e + 1] = gat(base + reach + 1 + e);
            }
            la[EPT - 1] = rt[0] = 0.0f;
#pragma unroll
            for (int e = EPT - 2; e >= 0; --e) la[e] = fmaxf(la[e], la[e + 1]);  // suffix maxima of the left part
#pragma unroll
            for (int e = 1; e < EPT; ++e) rt[e] = fmaxf(rt[e], rt[e - 1]);       // prefix maxima of the right part
            const int r = min(base + reach, CAP - 1);
            const float core = window(max(base + EPT - 1 - reach, 0), r, r / EPT, PF[r]);
#pragma unroll
            for (int e = 0; e < EPT; ++e) Aenv[base + e] = fmaxf(core, fmaxf(la[e], rt[e]));
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = base + e;
                const int r = min(i + reach, CAP - 1);
                Aenv[i] = window(max(i - reach, 0), r, r / EPT, PF[r]);
            }
        }
        // H for the thread's own CORE_EPT core samples and the one before them (the filters below run
        // over the core in this very mapping, so H never goes through shared memory for another thread):
        // windows [bh+e-reach-hold+1, bh+e+reach], e = 0..CORE_EPT
        const int bh = cidx + tid * CORE_EPT - 1;
        if (gm.shared_core) {
            float lh[CORE_EPT + 1], rh[CORE_EPT + 1];
#pragma unroll
            for (int e = 0; e < CORE_EPT; ++e) {
                lh[e] = gat(bh - reach - hold + 1 + e);
                rh[e + 1] = gat(bh + reach + 1 + e);
            }
            lh[CORE_EPT] = rh[0] = 0.0f;
#pragma unroll
            for (int e = CORE_EPT - 1; e >= 0; --e) lh[e] = fmaxf(lh[e], lh[e + 1]);
#pragma unroll
            for (int e = 1; e <= CORE_EPT; ++e) rh[e] = fmaxf(rh[e], rh[e - 1]);
            const int r = min(bh + reach, CAP - 1);
            const float core = window(max(bh + CORE_EPT + 1 - reach - hold, 0), r, r / EPT, PF[r]);
#pragma unroll
            for (int e = 0; e <= CORE_EPT; ++e) hc[e] = fmaxf(core, fmaxf(lh[e], rh[e]));
        } else {
#pragma unroll
            for (int e = 0; e <= CORE_EPT; ++e) {
                const int r = min(bh + e + reach, CAP - 1);
                hc[e] = window(max(bh + e - reach - hold + 1, 0), r, r / EPT, PF[r]);
            }
        }
        // lfilter starts from rest: the envelope before the first sample is 0, not a window maximum
#pragma unroll
        for (int e = 0; e <= CORE_EPT; ++e)
            if (bh + e < vlo) hc[e] = 0.0f;
    }

    // ---- P3: hold_out = lfilter(butter(1, f_hold), H), zero-carry pass (hyrax.py:61-66) -------------
    // The chunk's aggregate is published now; the carry from the previous chunks is only needed after
    // the attack filter below, which gives the predecessors time to publish theirs.
    LookbackSlot* slot = slots + chunk;
    double hold_y[CORE_EPT];
    double prev_last;  // the previous thread's last hold sample (zero carry into the chunk)
    float* Hown = Wk + tid * (CORE_EPT + 1);  // the thread's own H samples, parked until the release filter
    {
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            const double u = lp.hold_b0 * (double)hc[e + 1] + lp.hold_b1 * (double)hc[e];
            acc = u - lp.hold_a1 * acc;
            hold_y[e] = acc;
        }
        // (the barrier inside also ends P2: every thread is done with PF, SF and the sparse table)
        const double carry = scan_carry(acc, pow_hold, 0.0, scratch_a);
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) hold_y[e] += pow_hold->pe[e + 1] * carry;
        if (tid == NT - 1) publish(&slot->hold, hold_y[CORE_EPT - 1], 1);
        prev_last = __shfl_up_sync(0xffffffffu, hold_y[CORE_EPT - 1], 1);
        if ((tid & 31) == 31) warp_edge[tid >> 5] = hold_y[CORE_EPT - 1];  // read by the next warp's lane 0 in P5
#pragma unroll
        for (int e = 0; e <= CORE_EPT; ++e) Hown[e] = hc[e];  // SF (= Wk) is free now; only this thread reads these back
    }

    // ---- P4: g_att = filtfilt one-pole over A (hyrax.py:48-51) -------------------------------------
    // scipy's filtfilt runs over the odd extension by 6 samples with the steady-state initial state;
    // holding the extension's end value constant further out reproduces that state exactly.  Only
    // chunks that touch an end of the signal see the extension.
    const double c = lp.attack_c;
    const bool edge_l = vlo > reach, edge_r = vhi < reach + FL;
    if (edge_l || edge_r) {
        const int i0 = vlo, iL = vhi - 1;  // span indices of samples 0 and frames-1
        float fix[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * NT;
            fix[k] = Aenv[i];
            if (i < vlo) {
                const int d = (vlo - i < 6) ? vlo - i : 6;
                fix[k] = 2.0f * Aenv[i0] - Aenv[i0 + d];
            } else if (i >= vhi) {
                const int d = (i - iL < 6) ? i - iL : 6;
                fix[k] = 2.0f * Aenv[iL] - Aenv[iL - d];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < EPT; ++k) Aenv[tid + k * NT] = fix[k];
        __syncthreads();
    }
    {
        // forward over the thread's own samples, then backward over the same samples in the same registers:
        // the backward scan runs across the block in descending thread order, so the forward result never
        // goes through shared memory
        double y[EPT];
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            acc = (1.0 - c) * (double)Aenv[tid * EPT + e] + c * acc;
            y[e] = acc;
        }
        const double c0 = (double)Aenv[0];  // state before the first element: steady state
        const double carry = scan_carry(acc, pow_att, c0, scratch_b);  // barrier inside: Aenv fully read
#pragma unroll
        for (int e = 0; e < EPT; ++e) y[e] += pow_att->pe[e + 1] * carry;
        if (edge_r) {
            // past the extension's last sample (frames+5) the backward pass sees that value held
            const int ilast = (int)(frames + 5 - ga);
            if (ilast < CAP - 1) {
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if (tid * EPT + e == ilast) bcast[1] = y[e];
                __syncthreads();
                const double held = bcast[1];
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if (tid * EPT + e > ilast) y[e] = held;
                __syncthreads();  // bcast[1] is written again just below
            }
        }
        if (tid == NT - 1) bcast[1] = y[EPT - 1];  // the backward pass starts from the forward pass's last value
        acc = 0.0;
#pragma unroll
        for (int e = EPT - 1; e >= 0; --e) {
            acc = (1.0 - c) * y[e] + c * acc;
            y[e] = acc;
        }
        const double back = scan_carry_rev(acc, pow_att, &bcast[1], scratch_a);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = tid * EPT + e;
            if (i >= cidx && i < cidx + LC) G[i] = fmaxf(G[i], (float)(y[e] + pow_att->pe[EPT - e] * back));
        }
    }

    // ---- P5: hold carry from the previous chunks (decoupled look-back), finish hold_out -------------
    if (tid < 32) {
        const double cin = lookback(slots, chunk, false, pow_hold->pc);
        if (tid == 0) bcast[0] = cin;
    }
    __syncthreads();  // also: every thread is done reading Fd as the attack filter's plane
    const double hold_cin = bcast[0];
    double hold_prev;  // hold_out just before the thread's first sample
    {
        const double lead = pow_hold->ql[tid & 31] * pow_hold->qw[tid >> 5];  // pole^(tid*CORE_EPT)
        if ((tid & 31) == 0 && tid > 0) prev_last = warp_edge[(tid >> 5) - 1];
        hold_prev = tid == 0 ? hold_cin : prev_last + lead * hold_cin;
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) hold_y[e] += pow_hold->pe[e + 1] * lead * hold_cin;
        if (tid == NT - 1 && gm.publish_inclusive) publish(&slot->hold, hold_y[CORE_EPT - 1], 2);
    }

    // ---- P6: release_out = lfilter(butter(1, f_rel), max(H, hold_out)) (hyrax.py:68-73) -------------
    {
        double rel_y[CORE_EPT];
        double acc = 0.0;
        double prev_in = fmax((double)Hown[0], hold_prev);
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            const double cur = fmax((double)Hown[e + 1], hold_y[e]);
            const double u = lp.release_b0 * cur + lp.release_b1 * prev_in;
            prev_in = cur;
            acc = u - lp.release_a1 * acc;
            rel_y[e] = acc;
        }
        const double carry = scan_carry(acc, pow_rel, 0.0, scratch_b);
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) rel_y[e] += pow_rel->pe[e + 1] * carry;
        if (tid == NT - 1) publish(&slot->rel, rel_y[CORE_EPT - 1], 1);
        if (tid < 32) {
            const double cr = lookback(slots, chunk, true, pow_rel->pc);
            if (tid == 0) bcast[1] = cr;
        }
        __syncthreads();
        const double cin = bcast[1];
        const double lead = pow_rel->ql[tid & 31] * pow_rel->qw[tid >> 5];
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            rel_y[e] += pow_rel->pe[e + 1] * lead * cin;
            const int i = cidx + tid * CORE_EPT + e;
            const double g_rel = fmax(hold_y[e], rel_y[e]);            // hyrax.py:75
#ifdef MGB_LIM_DEBUG
            Fd[tid * CORE_EPT + e] = hold_y[e];
            G[i] = (float)rel_y[e];
#else
            Fd[tid * CORE_EPT + e] = 1.0 - fmax((double)G[i], g_rel);  // hyrax.py:97
#endif
        }
        if (tid == NT - 1 && gm.publish_inclusive) publish(&slot->rel, rel_y[CORE_EPT - 1], 2);
    }
#ifdef MGB_LIM_DEBUG
    __syncthreads();
#else
    __syncwarp();  // a warp applies the gains of its own 32*CORE_EPT consecutive samples: no block barrier
#endif

    // ---- P7: apply (hyrax.py:99, stages.py:203) -----------------------------------------------------
    // The gains sit in shared memory in the filters' mapping (thread t: samples t*CORE_EPT ..); lane l of
    // warp w now takes samples w*32*CORE_EPT + l + 32 q, so that every global access is 32 neighbours.
    {
        const int wbase = (tid >> 5) * (32 * CORE_EPT) + (tid & 31);
        float2 v[CORE_EPT];
#pragma unroll
        for (int q = 0; q < CORE_EPT; ++q) {
            const int k = wbase + q * 32;
            v[q] = k < core_n ? __ldg(in + s0 + k) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int q = 0; q < CORE_EPT; ++q) {
            const int k = wbase + q * 32;
            if (k < core_n) {
#ifdef MGB_LIM_DEBUG
                out[s0 + k] = make_float2((float)Fd[k], G[cidx + k]);
#else
                // (in float64 to the end: a frame the hard clip brings to the threshold must round to it exactly)
                const double gain = Fd[k] * scale;
                out[s0 + k] = make_float2((float)((double)v[q].x * gain), (float)((double)v[q].y * gain));
#endif
            }
        }
    }
}

// engaged = not all(isclose(rectified, 1.0)) with numpy's defaults rtol=1e-5, atol=1e-8
__global__ void limiter_engaged_kernel(const float* peak_bits, const double* pre_gain, double threshold, int* engaged) {
    const double pre = pre_gain ? *pre_gain : 1.0;
    const double peak = (double)(*peak_bits) * pre;
    const double r = fmax(peak, threshold) / threshold;
    *engaged = (fabs(r - 1.0) <= 1e-8 + 1e-5 * 1.0) ? 0 : 1;
}

int limiter_geometry(const mgb_limiter_params& lp, LimiterGeom* g) {
    MGB_REQUIRE(lp.reach >= 1 && lp.hold >= 3 && lp.warmup >= 8, MGB_ERR_INVALID, "limiter: bad window sizes");
    MGB_REQUIRE(lp.attack_c > 0.0 && lp.attack_c < 1.0, MGB_ERR_INVALID, "limiter: attack pole out of (0,1)");
    MGB_REQUIRE(lp.threshold > 0.0, MGB_ERR_INVALID, "limiter: threshold must be positive");
    g->reach = lp.reach;
    g->hold = lp.hold;
    g->warm = lp.warmup;
    g->left = lp.warmup > lp.hold ? lp.warmup : lp.hold;
    g->filt = LC + g->left + g->warm;
    g->span = g->filt + 2 * g->reach;
    int ept = (g->span + NT - 1) / NT;
    if (!(ept & 1)) ept += 1;  // odd stride: the blocked scans read shared memory conflict-free
    if (ept < 11) ept = 11;
    g->ept = ept;
    g->publish_inclusive = g_lookback_inclusive;
    // the cores [b+ept-1-reach, b+reach] and [b+ept-reach-hold, b+reach] must not be empty
    g->shared_core = (2 * lp.reach >= ept - 1 && 2 * lp.reach + lp.hold >= ept) ? 1 : 0;
    g->margin = (lp.reach + lp.hold + ept + 3) / 4 * 4;  // the furthest a window part reaches outside [0, CAP)
    MGB_REQUIRE(ept <= SPAN_EPT_MAX, MGB_ERR_UNSUPPORTED,
                "limiter: halo of %d samples exceeds the kernel's span", g->span - LC);
    return MGB_OK;
}

}  // namespace

int64_t limiter_lookback_bytes(int64_t frames) {
    const int64_t chunks = (frames + LC - 1) / LC;
    return (chunks * (int64_t)sizeof(LookbackSlot) + 255) / 256 * 256;
}

int launch_limiter_tables(const mgb_limiter_params& lp, ScanPow* tables, cudaStream_t stream) {
    LimiterGeom g;
    MGB_TRY(limiter_geometry(lp, &g));
    return launch("limiter_tables_kernel", limiter_tables_kernel, dim3(1), dim3(64), 0, stream, lp, g.ept, tables);
}

int launch_limiter(const mgb_limiter_params& lp, const float2* in, float2* out, int64_t frames, const double* pre_gain,
                   const double* post_gain, const int* engaged, int* ticket, LookbackSlot* lookback,
                   const ScanPow* tables, cudaStream_t stream) {
    LimiterGeom g;
    MGB_TRY(limiter_geometry(lp, &g));
    MGB_REQUIRE(frames > 6, MGB_ERR_INVALID, "limiter: the input must be longer than filtfilt's padlen (6)");
    MGB_REQUIRE(tables != nullptr, MGB_ERR_INVALID, "limiter: pole tables missing");
    const int64_t chunks = (frames + LC - 1) / LC;
    const size_t smem = (size_t)g.ept * NT * 16 + (size_t)g.margin * 8;
#define MGB_LIMITER_CASE(E)                                                                                       \
    case E:                                                                                                       \
        return launch("limiter_kernel", limiter_kernel<E>, dim3((unsigned)chunks), dim3(NT), smem, stream, lp, g, in, \
                      out, (long long)frames, pre_gain, post_gain, engaged, ticket, lookback, tables);
    switch (g.ept) {
        MGB_LIMITER_CASE(11)
        MGB_LIMITER_CASE(13)
        MGB_LIMITER_CASE(15)
        MGB_LIMITER_CASE(17)
        default: break;
    }
#undef MGB_LIMITER_CASE
    set_error("limiter: no kernel for %d span samples per thread", g.ept);
    return MGB_ERR_UNSUPPORTED;
}

int launch_limiter_engaged(const float* peak_bits, const double* pre_gain, double threshold, int* engaged,
                           cudaStream_t stream) {
    return launch("limiter_engaged_kernel", limiter_engaged_kernel, dim3(1), dim3(1), 0, stream, peak_bits, pre_gain,
                  threshold, engaged);
}

}  // namespace mgb
