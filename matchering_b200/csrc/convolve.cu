// K4 -- overlap-save FFT convolution of the target's mid/side with the two matching FIRs.
//
// Replaces (reference file:line):
//   match_frequencies.convolve        stage_helpers/match_frequencies.py:104-119
//     = scipy.signal.fftconvolve(x, fir, "same") per channel: full linear convolution sliced at
//       (F-1)//2 = F/2-1, edges zero-padded (scipy/signal/_signaltools.py fftconvolve/_centered)
//   dsp.lr_to_ms / dsp.ms_to_lr       dsp.py:57-68   (fused into the load and the store)
//   dsp.amplify by the RMS coefficient match_levels.py:114-131 (folded into the FIR spectra)
//   the first RMS-correction step's clip + per-piece sum of squares   stages.py:151-159
//   the peak that dsp.normalize / the limiter's early-out need        dsp.py:97, hyrax.py:83
//
// Framing: the FIR has F taps with fir[0] == 0 (symmetric Hann is zero at index 0), so taps
// 1..F-1 matter and output sample n needs inputs n-F/2 .. n+F/2-2.  A frame of N = 2F inputs
// starting at n0 - F/2 therefore yields the F outputs n0 .. n0+F-1 (circular index i = F-1 ..
// 2F-2), with n0 a multiple of F: every frame start is 16-byte aligned for the TMA bulk copy.
// mid and side ride one complex transform: z = mid + i*side, Y[k] = Hm[k] M[k] + i Hs[k] S[k]
// with M, S recovered from Z[k], Z[N-k]; real part of the inverse is the mid output, imaginary
// part the side output.  The 64 KB landing buffer of the bulk copy is reused in place as the two
// FFT planes.
#include "fft.cuh"
#include "kernels.cuh"

namespace mgb {

namespace {

template <int F>
struct ConvSmem {
    static constexpr int N = 2 * F;
    static constexpr int kPlaneBytes = (int)((PackedPlanes::bytes(N) + 15) / 16 * 16);
    static constexpr int kBytes = kPlaneBytes + 16 + 32 * 8 + 32 * 8 + 64 * 4 + 32;
};

struct ConvFirst {
    const float2* raw;  // landing buffer, index 0 = sample `origin`
    int lo, hi;         // indices of the buffer that lie inside the signal: [lo, hi)
    const float2* fixup;
    int fix_index;
    __device__ __forceinline__ float2 sample(int i) const {
        float2 v = make_float2(0.0f, 0.0f);
        if (i >= lo && i < hi) {
            v = raw[i];
            if (i == fix_index) v = *fixup;
        }
        return v;
    }
};

template <int F, bool CHAIN>
__global__ void __launch_bounds__(F / 8, (F <= 4096 ? 2 : 1))
convolve_kernel(const float2* __restrict__ x, long long frames, long long piece, int divisions,
                const cpx<float>* __restrict__ tw, const float2* __restrict__ h_mid,
                const float2* __restrict__ h_side, float2* __restrict__ result, float* __restrict__ mid_plane,
                double* __restrict__ piece_sums, mgb_track_state* __restrict__ state, int use_tma) {
    constexpr int N = 2 * F;
    constexpr int THREADS = N / 16;
    using L = ConvSmem<F>;
    MGB_DYN_SMEM(smem);
    const PackedPlanes planes{reinterpret_cast<float2*>(smem)};
    float2* raw = reinterpret_cast<float2*>(smem);  // the landing buffer IS the frame's storage: unpadded
                                                    // until the first pass has gathered it, padded after
    unsigned char* tail = smem + L::kPlaneBytes;
    tail += (16 - (reinterpret_cast<uintptr_t>(tail) & 15)) & 15;
    TmaBarrier* bar = reinterpret_cast<TmaBarrier*>(tail);
    double* red_a = reinterpret_cast<double*>(tail + 16);
    double* red_b = red_a + 32;
    float* red_f = reinterpret_cast<float*>(red_b + 32);
    unsigned* red_u = reinterpret_cast<unsigned*>(red_f + 32);  // [2] slot of block_max2

    const int tid = threadIdx.x;
    const long long n0 = (long long)blockIdx.x * F;
    const long long origin = n0 - F / 2;

    // ---- load the frame's 2F input samples (clipped to the signal) -----------------------------
    const long long lo = origin < 0 ? 0 : origin;
    const long long hi = (origin + N < frames) ? origin + N : frames;  // exclusive
    ConvFirst first;
    first.raw = raw;
    first.lo = (int)(lo - origin);
    first.hi = (int)(hi - origin);
    first.fixup = nullptr;
    first.fix_index = -1;
    if (tid == 0) red_u[0] = red_u[1] = 0u;
    if (use_tma) {
        if (tid == 0) tma_barrier_init(bar);
        __syncthreads();
        // lo - origin is 0 or F/2 (multiple of 2): source and destination stay 16-byte aligned
        long long count = hi - lo;
        if (count & 1) {  // odd tail: the last sample comes through a plain load
            first.fix_index = (int)(hi - 1 - origin);
            first.fixup = x + (hi - 1);
            count -= 1;
        }
        if (tid == 0) tma_load_1d(raw + (lo - origin), x + lo, (uint32_t)(count * 8), bar);
        tma_barrier_wait(bar, 0);
    } else {
        for (long long n = lo + tid; n < hi; n += THREADS) raw[n - origin] = x[n];
        __syncthreads();
    }

    // ---- the thread's 16 input points, mid / side, and the frame's channel balance (balance_factor) -----
    cpx<float> z[N / THREADS];
    float max_mid = 0.0f, max_side = 0.0f;
#pragma unroll
    for (int r = 0; r < N / THREADS; ++r) {
        const float2 v = first.sample(tid + r * THREADS);
        // each channel from L and R directly, one rounding each: (L-R)/2 is the reference's mid - R, and
        // forming it from the already rounded mid would put mid's rounding error into a quiet side
        z[r].x = (v.x + v.y) * 0.5f;
        z[r].y = (v.x - v.y) * 0.5f;
        max_mid = fmaxf(max_mid, fabsf(z[r].x));
        max_side = fmaxf(max_side, fabsf(z[r].y));
    }
    block_max2(max_mid, max_side, red_u);  // its barrier also ends everybody's reads of the landing buffer
    const float g_side = balance_factor(max_mid, max_side);  // >= 1 when the side is the quiet one, < 1 otherwise
    const float inv_g = 1.0f / g_side;                         // exact: a power of two
    const bool mid_silent = max_mid == 0.0f, side_silent = max_side == 0.0f;
#pragma unroll
    for (int r = 0; r < N / THREADS; ++r) z[r].y *= g_side;

    // ---- forward transform of z = mid + i*g*side ---------------------------------------------------
    fft_first_pass_regs<N, +1, THREADS, float>(planes, z, /*barrier_before_store=*/false);
    __syncthreads();
    fft_remaining<N, +1, THREADS, float, CHAIN>(planes, tw, PlaneStore<PackedPlanes>{planes}, true);
    __syncthreads();

    // ---- apply both FIR spectra on the pair (k, N-k) ---------------------------------------------
    for (int k = tid; k <= F; k += THREADS) {
        const int kn = (N - k) & (N - 1);
        const cpx<float> zk = planes.load(k), zn = planes.load(kn);
        const float zr = zk.x, zi = zk.y, nr = zn.x, ni = zn.y;
        // M = (Z[k] + conj Z[N-k])/2 ; g S = (Z[k] - conj Z[N-k])/(2i)
        const float mr = 0.5f * (zr + nr), mi = 0.5f * (zi - ni);
        const float sr = 0.5f * (zi + ni) * inv_g, si = 0.5f * (nr - zr) * inv_g;
        const float2 hm = h_mid[k], hs = h_side[k];
        // a channel that is exactly silent in this frame stays exactly silent, as in the reference
        const float pmr = mid_silent ? 0.0f : hm.x * mr - hm.y * mi, pmi = mid_silent ? 0.0f : hm.x * mi + hm.y * mr;
        const float psr = side_silent ? 0.0f : hs.x * sr - hs.y * si, psi = side_silent ? 0.0f : hs.x * si + hs.y * sr;
        // Y[k] = Pm + i Ps ; Y[N-k] = conj(Pm) + i conj(Ps)
        planes.store(k, cpx<float>{pmr - psi, pmi + psr});
        if (kn != k) planes.store(kn, cpx<float>{pmr + psi, psr - pmi});
    }
    __syncthreads();

    // ---- inverse transform, in place ------------------------------------------------------------------
    fft_first_pass<N, -1, THREADS, float>(planes, tw, PlaneLoad<PackedPlanes>{planes}, /*in_place=*/true);
    __syncthreads();
    fft_remaining<N, -1, THREADS, float, CHAIN>(planes, tw, PlaneStore<PackedPlanes>{planes}, /*last_in_place=*/true);
    __syncthreads();

    // ---- epilogue: circular index F-1+o is output sample n0+o; coalesced stores, mid/side -> L/R
    // (dsp.ms_to_lr), the first RMS-correction step's sum of clip(mid)^2 per piece, the peak ---------
    const int valid = (int)((frames - n0 < F) ? frames - n0 : F);
    const long long counted = piece * divisions;  // samples that enter the piece RMS (dsp.unfold)
    const long long pa = n0 / piece;
    const long long brel = (pa + 1) * piece - n0;  // first output of the next piece
    const int boundary = (int)(brel < F ? brel : F);
    const int count_to = (int)((counted - n0 < 0) ? 0 : (counted - n0 < F ? counted - n0 : F));
    double sq_a = 0.0, sq_b = 0.0;
    float peak = 0.0f;
    float2* res = result + n0;
    float* midp = mid_plane + n0;
#pragma unroll
    for (int k = 0; k < F / THREADS; ++k) {
        const int o = tid + k * THREADS;
        if (o < valid) {
            const cpx<float> y = planes.load(F - 1 + o);
            // (the inverse transform leaves rounding dust in a channel whose spectrum was exactly zero)
            const float m = mid_silent ? 0.0f : y.x, sd = side_silent ? 0.0f : y.y;
            const float l = m + sd, r = m - sd;
            res[o] = make_float2(l, r);
            midp[o] = m;
            peak = fmaxf(peak, fmaxf(fabsf(l), fabsf(r)));
            if (o < count_to) {
                const float cf = fminf(1.0f, fmaxf(-1.0f, m));  // dsp.clip
                const double c2 = (double)cf * (double)cf;
                if (o < boundary) sq_a += c2; else sq_b += c2;
            }
        }
    }
    const double ta = block_sum(sq_a, red_a);
    const double tb = block_sum(sq_b, red_b);
    const float pk = block_max(peak, red_f);
    if (tid == 0) {
        if (pa < divisions && ta != 0.0) atomicAdd(&piece_sums[pa], ta);
        if (pa + 1 < divisions && tb != 0.0) atomicAdd(&piece_sums[pa + 1], tb);
        atomic_max_nonneg(&state->conv_peak_bits, pk);
    }
}

template <int F>
int launch_convolve_t(const mgb_plan& plan, const mgb_track_layout& layout, const float2* target, float2* result,
                      const Workspace& ws, mgb_track_state* state, cudaStream_t stream) {
    const long long T = layout.target_frames;
    const unsigned nframes = (unsigned)((T + F - 1) / F);
    auto kernel = g_twiddle_chain ? convolve_kernel<F, true> : convolve_kernel<F, false>;
    return launch("convolve_kernel", kernel, dim3(nframes), dim3(F / 8), ConvSmem<F>::kBytes, stream, target, T,
                  (long long)layout.target_piece, layout.target_divisions, (const cpx<float>*)plan.d_tw_f32_2F,
                  (const float2*)ws.h_mid, (const float2*)ws.h_side, result, ws.mid_plane, ws.piece_sums, state,
                  g_use_tma);
}

}  // namespace

int launch_convolve(const mgb_plan& plan, const mgb_track_layout& layout, const float2* target, float2* result,
                    const Workspace& ws, mgb_track_state* state, cudaStream_t stream) {
    if (layout.target_piece < plan.fft_size) {
        set_error("convolve: piece (%lld) shorter than fft_size", (long long)layout.target_piece);
        return MGB_ERR_UNSUPPORTED;
    }
    switch (plan.fft_size) {
        case 1024: return launch_convolve_t<1024>(plan, layout, target, result, ws, state, stream);
        case 2048: return launch_convolve_t<2048>(plan, layout, target, result, ws, state, stream);
        case 4096: return launch_convolve_t<4096>(plan, layout, target, result, ws, state, stream);
        case 8192: return launch_convolve_t<8192>(plan, layout, target, result, ws, state, stream);
        default: break;
    }
    set_error("convolve: fft_size %d has no kernel", plan.fft_size);
    return MGB_ERR_UNSUPPORTED;
}

}  // namespace mgb
