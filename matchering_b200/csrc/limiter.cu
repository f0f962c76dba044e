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

// powers of a pole for a blocked scan with `ept` elements per thread
struct ScanPow {
    double pe[SPAN_EPT_MAX + 2];  // p^k
    double ql[33];                // q^k, q = p^ept
    double qw[17];                // Q^k, Q = q^32
};

__device__ __forceinline__ void scanpow_init(ScanPow* t, double p, int ept) {
    const int i = threadIdx.x;
    if (i < SPAN_EPT_MAX + 2) t->pe[i] = pow(p, (double)i);
    if (i < 33) t->ql[i] = pow(p, (double)(ept * i));
    if (i < 17) t->qw[i] = pow(p, (double)(ept * 32 * i));
}

// Exclusive carry of the recurrence y = u + p*y_prev across the block: given each thread's local
// end value B (zero initial state over its `ept` elements), returns the state just before the
// thread's first element when the state before the block's first element is c0.
// scratch: >= 32 doubles.  Contains barriers: every thread of the block must call.
__device__ __forceinline__ double scan_carry(double B, const ScanPow* t, double c0, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = B;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t->ql[d] * up;
    }
    __syncthreads();  // scratch reuse
    if (lane == 31) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double w = lane < NT / 32 ? scratch[lane] : 0.0;
#pragma unroll
        for (int d = 1; d < NT / 32; d <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t->qw[d] * up;
        }
        if (lane < NT / 32) scratch[lane] = w;
    }
    __syncthreads();
    const double warp_carry = warp > 0 ? scratch[warp - 1] : 0.0;
    double prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) prev = 0.0;
    return prev + t->ql[lane] * (warp_carry + t->qw[warp] * c0);
}

// a[i + shift] = max(a[i .. i+win-1]) for every i with the window inside [0, len); other
// elements end up unspecified.  In place, log2(win) doubling steps.  Contains barriers.
__device__ __forceinline__ void sliding_max(float* a, int len, int win, int shift, int ept) {
    const int tid = threadIdx.x;
    float v[SPAN_EPT_MAX];
    int pw = 1;
    while (pw * 2 <= win) {
#pragma unroll
        for (int k = 0; k < SPAN_EPT_MAX; ++k) {
            const int i = tid + k * NT;
            if (k < ept && i < len) v[k] = (i + pw < len) ? fmaxf(a[i], a[i + pw]) : a[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SPAN_EPT_MAX; ++k) {
            const int i = tid + k * NT;
            if (k < ept && i < len) a[i] = v[k];
        }
        __syncthreads();
        pw *= 2;
    }
    const int rem = win - pw;
#pragma unroll
    for (int k = 0; k < SPAN_EPT_MAX; ++k) {
        const int i = tid + k * NT;
        if (k < ept && i < len) v[k] = (i + rem < len) ? fmaxf(a[i], a[i + rem]) : a[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SPAN_EPT_MAX; ++k) {
        const int i = tid + k * NT;
        if (k < ept && i + shift < len) a[i + shift] = v[k];
    }
    __syncthreads();
}

__device__ __forceinline__ void publish(double* value, int* flag, double v, int state) {
    *(volatile double*)value = v;
    __threadfence();
    *(volatile int*)flag = state;
}

// Carry into chunk `chunk` of a first-order recurrence whose per-chunk multiplier is `pchunk`:
// walk back over the predecessors' published values until one is inclusive (or the weight of
// anything older is below double precision).  Called by one thread.
__device__ __forceinline__ double lookback(LookbackSlot* slots, int chunk, bool release, double pchunk) {
    double carry = 0.0, mult = 1.0;
    for (int j = chunk - 1; j >= 0; --j) {
        LookbackSlot* s = slots + j;
        int* flag = release ? &s->rel_flag : &s->hold_flag;
        int f;
        while ((f = *(volatile int*)flag) == 0) __nanosleep(40);
        __threadfence();
        if (f == 2) {
            carry += mult * *(volatile double*)(release ? &s->rel_inc : &s->hold_inc);
            break;
        }
        carry += mult * *(volatile double*)(release ? &s->rel_agg : &s->hold_agg);
        mult *= pchunk;
        if (mult < 1e-20) break;
    }
    return carry;
}

struct LimiterGeom {
    int reach, hold, warm, left;  // left = max(warm, hold): left halo of the envelope A
    int ept;                      // span elements per thread (odd)
    int span;                     // samples of g the chunk touches
    int filt;                     // samples the attack filter runs over (LC + left + warm)
};

__global__ void __launch_bounds__(NT)
limiter_kernel(mgb_limiter_params lp, LimiterGeom gm, const float2* __restrict__ in, float2* __restrict__ out,
               long long frames, const double* __restrict__ pre_gain, const double* __restrict__ post_gain,
               const int* __restrict__ engaged, int* __restrict__ ticket, LookbackSlot* __restrict__ slots) {
    MGB_DYN_SMEM(smem);
    const int cap = gm.ept * NT;
    double* Fd = reinterpret_cast<double*>(smem);                 // [cap] float64 work plane
    float* G = reinterpret_cast<float*>(smem + (size_t)cap * 8);  // [cap] hard-clip gain, later max(g, g_att)
    float* A = G + cap;                                           // [cap] envelope, later the hold envelope
    __shared__ ScanPow pow_att, pow_hold, pow_rel;
    __shared__ double scratch[32];
    __shared__ double bcast[2];
    __shared__ int chunk_s;

    const int tid = threadIdx.x;
    const double pre = pre_gain ? *pre_gain : 1.0;
    const double post = post_gain ? *post_gain : 1.0;

    if (tid == 0) chunk_s = atomicAdd(ticket, 1);
    scanpow_init(&pow_att, lp.attack_c, gm.ept);
    scanpow_init(&pow_hold, -lp.hold_a1, CORE_EPT);
    scanpow_init(&pow_rel, -lp.release_a1, CORE_EPT);
    __syncthreads();
    const int chunk = chunk_s;
    const long long s0 = (long long)chunk * LC;
    const long long e0 = (s0 + LC < frames) ? s0 + LC : frames;

    if (engaged && *engaged == 0) {  // hyrax.py:83-85: the limiter is not needed
        for (long long n = s0 + tid; n < e0; n += NT) {
            const float2 v = in[n];
            out[n] = make_float2((float)((double)v.x * pre * post), (float)((double)v.y * pre * post));
        }
        return;
    }

    const int reach = gm.reach, hold = gm.hold, W = gm.warm, HL = gm.left, ept = gm.ept;
    const int SP = gm.span, FL = gm.filt;
    const long long ga = s0 - HL - reach;  // sample at span index 0
    const int cidx = HL + reach;           // span index of the chunk's first sample
    const double thr = lp.threshold;

    // ---- P0: hard-clip gain g over the span (dsp.rectify + flip, hyrax.py:82-87) ------------------
    for (int i = tid; i < cap; i += NT) {
        const long long n = ga + i;
        float g = 0.0f;
        if (i < SP && n >= 0 && n < frames) {
            const float2 v = in[n];
            const double a = fmax(fabs((double)v.x), fabs((double)v.y)) * pre;
            const double r = fmax(a, thr) / thr;
            g = (float)(1.0 - 1.0 / r);
        }
        G[i] = g;
        A[i] = g;
    }
    __syncthreads();

    // ---- P1: attack envelope A[n] = max g[n-reach .. n+reach] (hyrax.py:35-37) --------------------
    sliding_max(A, cap, 2 * reach + 1, reach, ept);

    // ---- P2: g_att = filtfilt one-pole over A (hyrax.py:48-51) ------------------------------------
    // extended signal of scipy's filtfilt: odd reflection of 6 samples at both ends; constant
    // beyond, which leaves the steady-state initial condition untouched.
    const double c = lp.attack_c;
    auto env = [&](long long n) -> double {  // n inside the signal
        return (double)A[(int)(n - ga)];
    };
    auto ext = [&](long long n) -> double {
        if (n < 0) {
            const long long k = (-n < 6) ? -n : 6;
            return 2.0 * env(0) - env(k);
        }
        if (n >= frames) {
            const long long k = (n - (frames - 1) < 6) ? n - (frames - 1) : 6;
            return 2.0 * env(frames - 1) - env(frames - 1 - k);
        }
        return env(n);
    };
    auto clamp_idx = [&](int i) { return i < reach ? reach : (i >= reach + FL ? reach + FL - 1 : i); };
    {
        double y[SPAN_EPT_MAX];
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < SPAN_EPT_MAX; ++e) {
            if (e < ept) {
                const int i = clamp_idx(tid * ept + e);
                acc = (1.0 - c) * ext(ga + i) + c * acc;
                y[e] = acc;
            }
        }
        const double c0 = ext(ga + reach);  // state before the first element: steady state
        const double carry = scan_carry(acc, &pow_att, c0, scratch);
#pragma unroll
        for (int e = 0; e < SPAN_EPT_MAX; ++e)
            if (e < ept) Fd[tid * ept + e] = y[e] + pow_att.pe[e + 1] * carry;
    }
    __syncthreads();
    {
        // backward pass over the forward output; beyond the extension's end (sample frames+5) the
        // input is held at that last value, which is the reversed filter's steady-state start.
        const long long last_ext = frames + 5;
        auto fwd = [&](int i) -> double {
            i = clamp_idx(i);
            const long long n = ga + i;
            if (n > last_ext) i = (int)(last_ext - ga);
            return Fd[i];
        };
        double y[SPAN_EPT_MAX];
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < SPAN_EPT_MAX; ++e) {
            if (e < ept) {
                const int i = cap - 1 - (tid * ept + e);
                acc = (1.0 - c) * fwd(i) + c * acc;
                y[e] = acc;
            }
        }
        const double c0 = fwd(cap - 1);
        const double carry = scan_carry(acc, &pow_att, c0, scratch);
        // keep max(g, g_att) for the chunk's own samples
#pragma unroll
        for (int e = 0; e < SPAN_EPT_MAX; ++e) {
            if (e < ept) {
                const int i = cap - 1 - (tid * ept + e);
                if (i >= cidx && i < cidx + LC) G[i] = fmaxf(G[i], (float)(y[e] + pow_att.pe[e + 1] * carry));
            }
        }
    }
    __syncthreads();

    // ---- P3: hold envelope H[n] = max A[n-hold+1 .. n], samples before the signal count as 0 -------
    for (int i = tid; i < cap; i += NT)
        if (ga + i < 0) A[i] = 0.0f;
    __syncthreads();
    sliding_max(A, cap, hold, hold - 1, ept);

    // ---- P4: hold_out = lfilter(butter(1, f_hold), H) (hyrax.py:61-66) ----------------------------
    LookbackSlot* slot = slots + chunk;
    double hold_y[CORE_EPT];
    {
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            const int i = cidx + tid * CORE_EPT + e;
            const double u = lp.hold_b0 * (double)A[i] + lp.hold_b1 * (double)A[i - 1];
            acc = u - lp.hold_a1 * acc;
            hold_y[e] = acc;
        }
        const double carry = scan_carry(acc, &pow_hold, 0.0, scratch);
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) hold_y[e] += pow_hold.pe[e + 1] * carry;
        if (tid == NT - 1) publish(&slot->hold_agg, &slot->hold_flag, hold_y[CORE_EPT - 1], 1);
        if (tid == 0) {
            const double cin = lookback(slots, chunk, false, pow_hold.qw[16]);
            bcast[0] = cin;
        }
        __syncthreads();
        const double cin = bcast[0];
        const double lead = pow_hold.ql[tid & 31] * pow_hold.qw[tid >> 5];  // pole^(tid*CORE_EPT)
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            hold_y[e] += pow_hold.pe[e + 1] * lead * cin;
            Fd[tid * CORE_EPT + e] = hold_y[e];
        }
        if (tid == NT - 1) publish(&slot->hold_inc, &slot->hold_flag, hold_y[CORE_EPT - 1], 2);
    }
    __syncthreads();

    // ---- P5: release_out = lfilter(butter(1, f_rel), max(H, hold_out)) (hyrax.py:68-73) -----------
    {
        const double hold_cin = bcast[0];  // hold_out just before the chunk
        double rel_y[CORE_EPT];
        double acc = 0.0;
        double prev_in;
        {
            const int i = cidx + tid * CORE_EPT - 1;
            const double hprev = tid == 0 ? hold_cin : Fd[tid * CORE_EPT - 1];
            prev_in = fmax((double)A[i], hprev);
        }
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            const int i = cidx + tid * CORE_EPT + e;
            const double cur = fmax((double)A[i], hold_y[e]);
            const double u = lp.release_b0 * cur + lp.release_b1 * prev_in;
            prev_in = cur;
            acc = u - lp.release_a1 * acc;
            rel_y[e] = acc;
        }
        const double carry = scan_carry(acc, &pow_rel, 0.0, scratch);
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) rel_y[e] += pow_rel.pe[e + 1] * carry;
        if (tid == NT - 1) publish(&slot->rel_agg, &slot->rel_flag, rel_y[CORE_EPT - 1], 1);
        if (tid == 0) bcast[1] = lookback(slots, chunk, true, pow_rel.qw[16]);
        __syncthreads();
        const double cin = bcast[1];
        const double lead = pow_rel.ql[tid & 31] * pow_rel.qw[tid >> 5];
#pragma unroll
        for (int e = 0; e < CORE_EPT; ++e) {
            rel_y[e] += pow_rel.pe[e + 1] * lead * cin;
            const int i = cidx + tid * CORE_EPT + e;
            const double g_rel = fmax(hold_y[e], rel_y[e]);          // hyrax.py:75
            Fd[tid * CORE_EPT + e] = 1.0 - fmax((double)G[i], g_rel);  // hyrax.py:97
        }
        if (tid == NT - 1) publish(&slot->rel_inc, &slot->rel_flag, rel_y[CORE_EPT - 1], 2);
    }
    __syncthreads();

    // ---- P6: apply (hyrax.py:99, stages.py:203) ----------------------------------------------------
    for (int k = tid; k < (int)(e0 - s0); k += NT) {
        const float2 v = in[s0 + k];
        const double gain = Fd[k];
        out[s0 + k] = make_float2((float)((double)v.x * pre * gain * post), (float)((double)v.y * pre * gain * post));
    }
}

// engaged = not all(isclose(rectified, 1.0)) with numpy's defaults rtol=1e-5, atol=1e-8
__global__ void limiter_engaged_kernel(const float* peak_bits, const double* pre_gain, double threshold, int* engaged) {
    const double pre = pre_gain ? *pre_gain : 1.0;
    const double peak = (double)(*peak_bits) * pre;
    const double r = fmax(peak, threshold) / threshold;
    *engaged = (fabs(r - 1.0) <= 1e-8 + 1e-5 * 1.0) ? 0 : 1;
}

}  // namespace

int64_t limiter_lookback_bytes(int64_t frames) {
    const int64_t chunks = (frames + LC - 1) / LC;
    return (chunks * (int64_t)sizeof(LookbackSlot) + 255) / 256 * 256;
}

static int limiter_geometry(const mgb_limiter_params& lp, LimiterGeom* g) {
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
    g->ept = ept;
    MGB_REQUIRE(ept <= SPAN_EPT_MAX, MGB_ERR_UNSUPPORTED,
                "limiter: halo of %d samples exceeds the kernel's span", g->span - LC);
    return MGB_OK;
}

int launch_limiter(const mgb_limiter_params& lp, const float2* in, float2* out, int64_t frames, const double* pre_gain,
                   const double* post_gain, const int* engaged, int* ticket, LookbackSlot* lookback,
                   cudaStream_t stream) {
    LimiterGeom g;
    MGB_TRY(limiter_geometry(lp, &g));
    MGB_REQUIRE(frames > 6, MGB_ERR_INVALID, "limiter: the input must be longer than filtfilt's padlen (6)");
    const int64_t chunks = (frames + LC - 1) / LC;
    const size_t smem = (size_t)g.ept * NT * (8 + 4 + 4);
    return launch("limiter_kernel", limiter_kernel, dim3((unsigned)chunks), dim3(NT), smem, stream, lp, g, in, out,
                  (long long)frames, pre_gain, post_gain, engaged, ticket, lookback);
}

int launch_limiter_engaged(const float* peak_bits, const double* pre_gain, double threshold, int* engaged,
                           cudaStream_t stream) {
    return launch("limiter_engaged_kernel", limiter_engaged_kernel, dim3(1), dim3(1), 0, stream, peak_bits, pre_gain,
                  threshold, engaged);
}

}  // namespace mgb
