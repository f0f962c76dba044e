// Band-limited resampling to Config.internal_sample_rate on the device.
//
// Replaces (reference file:line):
//   checker.__check_sample_rate        matchering/checker.py:30-44
//     = resampy.resample(array, sample_rate, required_sample_rate, axis=0), filter "kaiser_best"
//       (resampy/core.py resample, resampy/interpn.py _resample_loop)
//
// Output frame t sits at input time t / ratio.  Its value is the sum over the input frames to the left (n, n-1,
// ...) and to the right (n+1, n+2, ...) of that time of  x * (win[offset + i*step] + eta * delta[offset + i*step]):
// half a Kaiser-windowed sinc tabulated at 512 entries per zero crossing (built on the host from Config-free
// constants, matchering_b200/resample.py), read `step` entries apart with linear interpolation between
// entries; all index arithmetic in float64 in resampy's own order of operations, so the same table entries are
// picked.  One thread per output frame, both channels; about 2 * 64 / min(1, ratio) taps each.  The table
// (32769 entries, two doubles each, 512 KB) lives in L2; neighbouring threads read neighbouring inputs and, for
// ratios near 1, neighbouring table entries.
#include "kernels.cuh"

namespace mgb {

namespace {

__global__ void __launch_bounds__(256)
resample_kernel(const float2* __restrict__ x, long long n_orig, float2* __restrict__ y, long long n_out, double time_increment,
                double scale, int num_table, int index_step, const double2* __restrict__ win_delta, int nwin) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    const double time_register = (double)t * time_increment;
    const long long n = (long long)time_register;
    double acc_l = 0.0, acc_r = 0.0;
    // left wing: x[n - i]
    double frac = scale * (time_register - (double)n);
    double index_frac = frac * (double)num_table;
    int offset = (int)index_frac;
    double eta = index_frac - (double)offset;
    long long i_max = (nwin - offset) / index_step;
    if (n + 1 < i_max) i_max = n + 1;
    for (long long i = 0; i < i_max; ++i) {
        const double2 wd = win_delta[offset + (int)i * index_step];
        const double weight = wd.x + eta * wd.y;
        const float2 v = x[n - i];
        acc_l += weight * (double)v.x;
        acc_r += weight * (double)v.y;
    }
    // right wing: x[n + k + 1]
    frac = scale - frac;
    index_frac = frac * (double)num_table;
    offset = (int)index_frac;
    eta = index_frac - (double)offset;
    long long k_max = (nwin - offset) / index_step;
    if (n_orig - n - 1 < k_max) k_max = n_orig - n - 1;
    for (long long k = 0; k < k_max; ++k) {
        const double2 wd = win_delta[offset + (int)k * index_step];
        const double weight = wd.x + eta * wd.y;
        const float2 v = x[n + k + 1];
        acc_l += weight * (double)v.x;
        acc_r += weight * (double)v.y;
    }
    y[t] = make_float2((float)acc_l, (float)acc_r);
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" {

int64_t mgb_resample_frames(int64_t frames_in, int32_t rate_in, int32_t rate_out) {
    if (frames_in < 0 || rate_in <= 0 || rate_out <= 0) return -1;
    // resampy: int(n * sr_new / sr_orig) -- integer product, float division, truncation
    return (int64_t)((double)(frames_in * (int64_t)rate_out) / (double)rate_in);
}

int mgb_resample(const float* d_in_lr, int64_t frames_in, int32_t rate_in, float* d_out_lr, int64_t frames_out, int32_t rate_out,
                 const double* d_win_delta, int32_t nwin, int32_t num_table, void* stream) {
    MGB_REQUIRE(d_in_lr && d_out_lr && d_win_delta, MGB_ERR_INVALID, "resample: NULL argument");
    MGB_REQUIRE(rate_in > 0 && rate_out > 0 && frames_in > 0 && nwin > 1 && num_table > 0, MGB_ERR_INVALID, "resample: bad sizes");
    MGB_REQUIRE(frames_out == mgb_resample_frames(frames_in, rate_in, rate_out), MGB_ERR_INVALID,
                "resample: %lld output frames, the reference produces %lld", (long long)frames_out,
                (long long)mgb_resample_frames(frames_in, rate_in, rate_out));
    MGB_REQUIRE((reinterpret_cast<uintptr_t>(d_win_delta) & 15) == 0, MGB_ERR_INVALID, "resample: table is not 16-byte aligned");
    if (frames_out == 0) return MGB_OK;
    const double sample_ratio = (double)rate_out / (double)rate_in;
    const double scale = sample_ratio < 1.0 ? sample_ratio : 1.0;
    const int index_step = (int)(scale * (double)num_table);
    MGB_REQUIRE(index_step >= 1, MGB_ERR_UNSUPPORTED, "resample: ratio %g is below the table's resolution", sample_ratio);
    const unsigned blocks = (unsigned)((frames_out + 255) / 256);
    return launch("resample_kernel", resample_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const float2*)d_in_lr,
                  (long long)frames_in, (float2*)d_out_lr, (long long)frames_out, 1.0 / sample_ratio, scale, (int)num_table,
                  index_step, (const double2*)d_win_delta, (int)nwin);
}

}  // extern "C"
