// "Last block finishes the job" tails: the tiny serial steps that follow a grid-wide reduction
// (level statistics after the analysis pass, the RMS-correction coefficient after a per-piece sum,
// the final scalars) run in whichever CTA of the reducing kernel retires last, instead of in
// single-CTA kernels of their own -- each of those cost a launch gap of several microseconds on a
// pipeline whose whole device time is a few hundred.
#pragma once
#include "common.cuh"

namespace mgb {

// True in exactly one CTA of the grid: the last one to get here.  Every thread must call (one
// barrier inside).  All global writes the tail depends on must precede the call in program order.
__device__ __forceinline__ bool block_is_last(int* counter, int total_blocks, int* smem_flag) {
    __threadfence();  // this thread's partial results are visible device-wide before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(counter, 1);
        *smem_flag = (t == total_blocks - 1);
        if (*smem_flag) __threadfence();
    }
    __syncthreads();
    return *smem_flag != 0;
}

// ---- level statistics (match_levels.py:29-44, 62-71, 93-111) -------------------------------------
struct LevelsArgs {
    double* sumsq_t;  // [div_t][slots_t] per-(piece, slot) sums of mid^2; slot 0 is overwritten by the RMS
    double* sumsq_r;
    const float* absmax_r;  // [div_r*slots_r + 1]
    unsigned char* mask_t;
    unsigned char* mask_r;
    mgb_track_state* state;
    long long piece_t, piece_r;
    int div_t, slots_t, div_r, slots_r;
    double threshold, eps;
    int enabled;
};

static __device__ __noinline__ void levels_block(const LevelsArgs& a, double* red_d, float* red_f) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    float pk = 0.0f;
    for (int i = tid; i < a.div_r * a.slots_r + 1; i += nthr) pk = fmaxf(pk, __ldcg(a.absmax_r + i));
    pk = block_max(pk, red_f);
    const double peak = (double)pk;
    double coef = 1.0;
    if (peak < a.threshold) coef = fmax(a.eps, peak / a.threshold);  // dsp.py:96-99, normalize_clipped=False

    double match[2];
    int loud[2];
    for (int sig = 0; sig < 2; ++sig) {
        double* part = sig == 0 ? a.sumsq_t : a.sumsq_r;
        const int div = sig == 0 ? a.div_t : a.div_r;
        const int slots = sig == 0 ? a.slots_t : a.slots_r;
        const double piece = (double)(sig == 0 ? a.piece_t : a.piece_r);
        unsigned char* mask = sig == 0 ? a.mask_t : a.mask_r;
        double acc = 0.0;
        for (int p = tid; p < div; p += nthr) {
            double s = 0.0;
            for (int k = 0; k < slots; ++k) s += __ldcg(part + (long long)p * slots + k);
            const double r = sqrt(s / piece);  // dsp.py:86
            part[(long long)p * slots] = r;    // slot 0 now holds the piece's RMS
            acc += r * r;
        }
        const double total = block_sum(acc, red_d);
        const double avg = sqrt(total / (double)div);  // rms(rmses), match_levels.py:101
        double accm = 0.0, cnt = 0.0;
        for (int p = tid; p < div; p += nthr) {
            const double r = part[(long long)p * slots];
            const bool m = r >= avg;  // match_levels.py:65
            mask[p] = m ? 1 : 0;
            if (m) {
                accm += r * r;
                cnt += 1.0;
            }
        }
        const double tm = block_sum(accm, red_d);
        const double tc = block_sum(cnt, red_d);
        match[sig] = sqrt(tm / tc);
        loud[sig] = (int)tc;
    }
    if (tid == 0) {
        mgb_track_state* state = a.state;
        const double ref_match = match[1] / coef;  // the reference measures the normalised reference
        state->reference_peak = peak;
        state->final_amplitude_coef = coef;
        state->target_match_rms = match[0];
        state->reference_match_rms = ref_match;
        state->rms_coefficient = ref_match / fmax(a.eps, match[0]);
        state->gain = 1.0;
        state->result_peak = 0.0;
        state->normalize_coef = 1.0;
        state->target_loud_pieces = loud[0];
        state->reference_loud_pieces = loud[1];
        state->limiter_engaged = 1;
        state->conv_peak_bits = 0.0f;
        state->steps_done = 0;
        for (int i = 0; i < MGB_MAX_CORRECTION_STEPS; ++i) state->correction[i] = 1.0;
    }
}

// ---- one RMS-correction step's coefficient (stages.py:153-168) and the final scalars ---------------
struct CorrectionArgs {
    const double* sums;  // [divisions] sums of clip(mid*gain)^2 of this step
    mgb_track_state* state;
    long long piece;
    int divisions;
    int step;
    int update;    // compute this step's coefficient
    int finalize;  // this was the last step: result peak, limiter early-out flag, normalisation coefficient
    double eps, threshold;
};

static __device__ __noinline__ void correction_block(const CorrectionArgs& a, double* red) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    mgb_track_state* state = a.state;
    if (a.update) {
        double acc = 0.0;
        for (int p = tid; p < a.divisions; p += nthr) acc += __ldcg(a.sums + p) / (double)a.piece;  // rms^2
        const double avg = sqrt(block_sum(acc, red) / (double)a.divisions);
        double accm = 0.0, cnt = 0.0;
        for (int p = tid; p < a.divisions; p += nthr) {
            const double r = sqrt(__ldcg(a.sums + p) / (double)a.piece);
            if (r >= avg) {
                accm += r * r;
                cnt += 1.0;
            }
        }
        const double tm = block_sum(accm, red);
        const double tc = block_sum(cnt, red);
        if (tid == 0) {
            const double match = sqrt(tm / tc);
            const double c = state->reference_match_rms / fmax(a.eps, match);
            state->correction[a.step] = c;
            state->gain *= c;
            state->steps_done = a.step + 1;
        }
    }
    if (a.finalize && tid == 0) {
        const double peak = (double)__ldcg(&state->conv_peak_bits) * state->gain;
        state->result_peak = peak;
        state->normalize_coef = fmax(a.eps, peak / a.threshold);  // dsp.py:99 with normalize_clipped=True
        const double r = fmax(peak, a.threshold) / a.threshold;   // dsp.py:117-121 at the loudest frame
        state->limiter_engaged = (fabs(r - 1.0) <= 1e-8 + 1e-5) ? 0 : 1;  // np.isclose defaults, hyrax.py:83
    }
}

}  // namespace mgb
