// K1/K2 -- one pass over a signal for everything stages.__match_levels and __average_fft need.
//
// Replaces (reference file:line):
//   dsp.normalize's peak search            matchering/dsp.py:97
//   dsp.lr_to_ms                           matchering/dsp.py:57-64
//   dsp.unfold + dsp.batch_rms             matchering/dsp.py:71-86   (per-piece sum of mid^2)
//   match_frequencies.__average_fft        matchering/stage_helpers/match_frequencies.py:30-42
//
// Work item = (piece p, slot s): a contiguous run of the piece's F-sample frames.  Each frame is
// brought into shared memory by one TMA bulk copy (cp.async.bulk, raw interleaved L/R), turned
// into z = mid + i*side while it is gathered for the first FFT pass (sum(mid^2) accumulated in float64 and
// max|x| are accumulated on the way), transformed with ONE complex F-point FFT, and split into
// |rfft(mid)| and |rfft(side)| from the pair Z[k], Z[F-k].  The magnitudes are summed over the
// slot's frames in registers and written once per work item; which pieces count ("loudest") is
// only known after every piece's RMS is, so the selection and the mean happen later, in design.cu,
// on these per-piece partial sums (|rfft| is positively homogeneous, so the level-matching gain
// and the reference normalisation are applied there too).
#include "fft.cuh"
#include "kernels.cuh"

namespace mgb {

namespace {

// fft_size 16384: frame and landing buffer do not fit one SM together (131 + 139 KB), so the frame's samples are
// read straight from global memory while they are gathered for the first pass (no bulk copy, no overlap of the next
// frame's load: a rare Config, correct first).
template <int F>
struct AnalyzeSmem {
    static constexpr bool kDirect = F > 8192;
    static constexpr int kRawBytes = kDirect ? 0 : ((F + 2) * 8 + 15) / 16 * 16;
    static constexpr int kPlaneBytes = (int)((PackedPlanes::bytes(F) + 15) / 16 * 16);
    static constexpr int kBytes = kRawBytes + kPlaneBytes + 16 /*barrier*/ + 32 * 8 + 96 * 4 + 32;
};

struct AnalyzeFirst {
    const float2* raw;    // frame start inside the landing buffer
    const float2* fixup;  // global address of the one sample the bulk copy could not cover (or null)
    int fix_index;
    __device__ __forceinline__ float2 sample(int i) const {
        float2 v = raw[i];
        if (i == fix_index) v = *fixup;
        return v;
    }
};

template <int F, bool CHAIN>
__global__ void __launch_bounds__(F / 16)
analyze_kernel(const float2* __restrict__ x, long long frames, long long piece, int divisions, int slots,
               const cpx<float>* __restrict__ tw, float* __restrict__ spec_part, double* __restrict__ sumsq_part,
               float* __restrict__ absmax_part, int use_tma) {
    constexpr int THREADS = F / 16;
    constexpr int HB = F / 2 + 1;
    constexpr int BINS = (HB + THREADS - 1) / THREADS;
    using L = AnalyzeSmem<F>;
    MGB_DYN_SMEM(smem);
    float2* raw = reinterpret_cast<float2*>(smem);
    const PackedPlanes planes{reinterpret_cast<float2*>(smem + L::kRawBytes)};
    unsigned char* tail = smem + L::kRawBytes + L::kPlaneBytes;
    tail += (16 - (reinterpret_cast<uintptr_t>(tail) & 15)) & 15;
    TmaBarrier* bar = reinterpret_cast<TmaBarrier*>(tail);
    double* red_d = reinterpret_cast<double*>(tail + 16);
    float* red_f = reinterpret_cast<float*>(tail + 16 + 32 * 8);
    unsigned* red_u = reinterpret_cast<unsigned*>(red_f + 32);  // [4] two alternating slots of block_max2

    const int tid = threadIdx.x;
    const int slot = blockIdx.x;
    const int p = blockIdx.y;
    const long long base = (long long)p * piece;
    const int frames_per_piece = (int)(piece / F);
    const int f_lo = (int)((long long)slot * frames_per_piece / slots);
    const int f_hi = (int)((long long)(slot + 1) * frames_per_piece / slots);

    double sumsq = 0.0;
    float peak = 0.0f;
    float acc_mid[BINS], acc_side[BINS];
#pragma unroll
    for (int b = 0; b < BINS; ++b) acc_mid[b] = acc_side[b] = 0.0f;

    if constexpr (L::kDirect) use_tma = 0;
    if (use_tma && tid == 0) tma_barrier_init(bar);
    if (tid == 0) red_u[0] = red_u[1] = red_u[2] = red_u[3] = 0u;
    __syncthreads();

    // Bulk copies need 16-byte aligned global addresses; a frame may start on an odd sample, so
    // the copy starts one sample early and the frame is read at offset `off` in the buffer.
    auto frame_start = [&](int f) { return base + (long long)f * F; };
    auto issue = [&](int f) {
        const long long start = frame_start(f);
        const long long start_al = start & ~1LL;
        const int off = (int)(start - start_al);
        int count = (F + off + 1) & ~1;
        if (start_al + count > frames) count -= 2;  // never read past the end of the signal
        tma_load_1d(raw, x + start_al, (uint32_t)count * 8u, bar);
    };
    if (use_tma && tid == 0 && f_lo < f_hi) issue(f_lo);

    for (int f = f_lo; f < f_hi; ++f) {
        const long long start = frame_start(f);
        int off = 0;
        AnalyzeFirst first;
        first.fixup = nullptr;
        first.fix_index = -1;
        if (use_tma) {
            off = (int)(start & 1LL);
            if (off == 1 && start + F == frames) {  // the copy was shortened by two samples
                first.fix_index = F - 1;
                first.fixup = x + start + (F - 1);
            }
            tma_barrier_wait(bar, (uint32_t)(f - f_lo));
        } else if constexpr (!L::kDirect) {
            for (int i = tid; i < F; i += THREADS) raw[i] = x[start + i];
            __syncthreads();
        }
        first.raw = L::kDirect ? x + start : raw + off;
        // the thread's 16 input points: mid / side, sum(mid^2) in float64, the peak, and the frame's
        // channel balance before the two channels share a transform (see balance_factor)
        cpx<float> z[F / THREADS];
        float max_mid = 0.0f, max_side = 0.0f;
        auto take = [&](int r, float2 v) {
            // mid and side from L and R directly, one rounding each, exactly as the convolution forms them
            z[r].x = (v.x + v.y) * 0.5f;
            z[r].y = (v.x - v.y) * 0.5f;
            sumsq = fma((double)z[r].x, (double)z[r].x, sumsq);
            peak = fmaxf(peak, fmaxf(fabsf(v.x), fabsf(v.y)));
            max_mid = fmaxf(max_mid, fabsf(z[r].x));
            max_side = fmaxf(max_side, fabsf(z[r].y));
        };
        if (first.fix_index < 0) {  // (all frames but possibly the signal's last one)
#pragma unroll
            for (int r = 0; r < F / THREADS; ++r) take(r, first.raw[tid + r * THREADS]);
        } else {
#pragma unroll
            for (int r = 0; r < F / THREADS; ++r) take(r, first.sample(tid + r * THREADS));
        }
        unsigned* slot = red_u + 2 * ((f - f_lo) & 1);
        if (tid == 0) red_u[2 * ((f - f_lo + 1) & 1)] = red_u[2 * ((f - f_lo + 1) & 1) + 1] = 0u;  // next frame's slot
        block_max2(max_mid, max_side, slot);
        const float g_side = balance_factor(max_mid, max_side);
        const float inv_g = 0.5f / g_side;  // exact (power of two); folds the 1/2 of the spectrum split
        const bool side_silent = max_side == 0.0f;
#pragma unroll
        for (int r = 0; r < F / THREADS; ++r) z[r].y *= g_side;
        fft_first_pass_regs<F, +1, THREADS, float>(planes, z, /*barrier_before_store=*/false);
        __syncthreads();  // planes written, landing buffer consumed by every thread
        if (use_tma && tid == 0 && f + 1 < f_hi) {
            fence_proxy_async();
            issue(f + 1);  // overlaps the remaining passes of this frame
        }
        fft_remaining<F, +1, THREADS, float, CHAIN>(planes, tw, PlaneStore<PackedPlanes>{planes}, /*last_in_place=*/true);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < BINS; ++b) {
            const int k = tid + b * THREADS;
            if (k < HB) {
                const int kn = (F - k) & (F - 1);
                const cpx<float> zk = planes.load(k), zn = planes.load(kn);
                const float zr = zk.x, zi = zk.y, nr = zn.x, ni = zn.y;
                // rfft(mid)[k] = (Z[k] + conj(Z[F-k]))/2 ; rfft(side)[k] = (Z[k] - conj(Z[F-k]))/(2i)
                const float mr = zr + nr, mi = zi - ni;
                const float sr = zi + ni, si = nr - zr;
                acc_mid[b] += 0.5f * sqrt_approx(mr * mr + mi * mi);
                if (!side_silent) acc_side[b] += inv_g * sqrt_approx(sr * sr + si * si);
            }
        }
        __syncthreads();  // planes free for the next frame
    }

    // the piece's tail beyond its last whole frame counts for the RMS and the peak, not the spectrum
    if (slot == slots - 1) {
        const long long tail_lo = base + (long long)frames_per_piece * F;
        const long long tail_hi = base + piece;
        for (long long n = tail_lo + tid; n < tail_hi; n += THREADS) {
            const float2 v = x[n];
            const double mid = ((double)v.x + (double)v.y) * 0.5;
            sumsq += mid * mid;
            peak = fmaxf(peak, fmaxf(fabsf(v.x), fabsf(v.y)));
        }
    }
    const double block_sq = block_sum(sumsq, red_d);
    const float block_pk = block_max(peak, red_f);
    const long long item = (long long)p * slots + slot;
    if (tid == 0) {
        sumsq_part[item] = block_sq;
        absmax_part[item] = block_pk;
    }
    float* out = spec_part + item * 2 * HB;
#pragma unroll
    for (int b = 0; b < BINS; ++b) {
        const int k = tid + b * THREADS;
        if (k < HB) {
            out[k] = acc_mid[b];
            out[HB + k] = acc_side[b];
        }
    }
    // samples past piece*divisions (fewer than `divisions` of them) only matter for the peak
    if (p == 0 && slot == 0) {
        float tail_pk = 0.0f;
        for (long long n = piece * divisions + tid; n < frames; n += THREADS) {
            const float2 v = x[n];
            tail_pk = fmaxf(tail_pk, fmaxf(fabsf(v.x), fabsf(v.y)));
        }
        tail_pk = block_max(tail_pk, red_f);
        if (tid == 0) absmax_part[(long long)divisions * slots] = tail_pk;
    }
}

template <int F>
int launch_analyze_t(const mgb_plan& plan, const float2* x, int64_t frames, int64_t piece, int divisions, int slots,
                     float* spec_part, double* sumsq_part, float* absmax_part, cudaStream_t stream) {
    // option "analyze_chain": twiddle powers built in registers from two table reads per radix-16 butterfly
    // instead of fifteen reads (what the convolution's transforms do)
    auto kernel = g_analyze_chain ? analyze_kernel<F, true> : analyze_kernel<F, false>;
    return launch("analyze_kernel", kernel, dim3(slots, divisions), dim3(F / 16), AnalyzeSmem<F>::kBytes,
                  stream, x, (long long)frames, (long long)piece, divisions, slots,
                  (const cpx<float>*)plan.d_tw_f32_F, spec_part, sumsq_part, absmax_part, g_use_tma);
}

}  // namespace

int launch_analyze(const mgb_plan& plan, const float2* x, int64_t frames, int64_t piece, int divisions, int slots,
                   float* spec_part, double* sumsq_part, float* absmax_part, cudaStream_t stream) {
    switch (plan.fft_size) {
        case 512: return launch_analyze_t<512>(plan, x, frames, piece, divisions, slots, spec_part, sumsq_part, absmax_part, stream);
        case 1024: return launch_analyze_t<1024>(plan, x, frames, piece, divisions, slots, spec_part, sumsq_part, absmax_part, stream);
        case 2048: return launch_analyze_t<2048>(plan, x, frames, piece, divisions, slots, spec_part, sumsq_part, absmax_part, stream);
        case 4096: return launch_analyze_t<4096>(plan, x, frames, piece, divisions, slots, spec_part, sumsq_part, absmax_part, stream);
        case 8192: return launch_analyze_t<8192>(plan, x, frames, piece, divisions, slots, spec_part, sumsq_part, absmax_part, stream);
        case 16384: return launch_analyze_t<16384>(plan, x, frames, piece, divisions, slots, spec_part, sumsq_part, absmax_part, stream);
        default: break;
    }
    set_error("analyze: fft_size %d has no kernel", plan.fft_size);
    return MGB_ERR_UNSUPPORTED;
}

}  // namespace mgb
