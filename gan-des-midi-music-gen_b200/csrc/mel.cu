// GAN-DES mel front end on the device (/root/reference/GAN_DES/util.py:37-61: torchaudio MelSpectrogram(n_fft = 2048, hop, n_mels, f_min, f_max)
// followed by AmplitudeToDB(top_db)); three steps:
//   1. stft_power_kernel   framing (centre = True, reflect padding), periodic Hann window, 2048-point FFT, |X|^2 -> power rows [b * T][1025]
//                          One CTA transforms TWO frames with one complex FFT (frame a in the real part, frame b in the imaginary part,
//                          split by conjugate symmetry), radix-2 Stockham autosort in shared memory, twiddles from a sincospi table (fp32).
//   2. mmg_gemm_tc (tf32)  mel[b][m][t] = sum_f power[b*T + t][f] fb[f][m]: the filter bank as the N operand, stored straight to (B, n_mels, T)
//   3. power_to_db_kernel  10 log10(max(x, 1e-10)), floor at (max over the spectrogram) - top_db          (torchaudio.functional.amplitude_to_DB)
#include "common.cuh"

namespace {

constexpr int NFFT = 2048, NBINS = NFFT / 2 + 1, LOG2N = 11;

__global__ void __launch_bounds__(256) stft_power_kernel(const float* __restrict__ wave, long long L, long long wave_pitch, int hop, int T, long long n_frames,
                                                         float* __restrict__ power, int pitch) {
    __shared__ float2 buf[2][NFFT];
    __shared__ float2 tw[NFFT / 2];
    const int tid = threadIdx.x;
    const long long fa = 2LL * blockIdx.x, fb = fa + 1;
    const bool has_b = fb < n_frames;
    const long long ba = fa / T, bb = has_b ? fb / T : 0;
    const long long oa = (fa - ba * T) * hop - NFFT / 2, ob = has_b ? (fb - bb * T) * hop - NFFT / 2 : 0;
    const float* wa = wave + ba * wave_pitch;
    const float* wb = wave + bb * wave_pitch;
    for (int n = tid; n < NFFT; n += 256) {
        const float w = 0.5f - 0.5f * cospif((float)n * (1.f / 1024.f));          // torch.hann_window(2048, periodic=True)
        long long ia = oa + n, ib = ob + n;
        if (ia < 0) ia = -ia;
        if (ia >= L) ia = 2 * (L - 1) - ia;                                       // pad_mode = "reflect"
        if (ib < 0) ib = -ib;
        if (ib >= L) ib = 2 * (L - 1) - ib;
        buf[0][n] = make_float2(wa[ia] * w, has_b ? wb[ib] * w : 0.f);
    }
    for (int k = tid; k < NFFT / 2; k += 256) {
        float s, c;
        sincospif(-(float)k * (1.f / 1024.f), &s, &c);                            // e^{-2 pi i k / 2048}
        tw[k] = make_float2(c, s);
    }
    __syncthreads();
    int cur = 0;
#pragma unroll 1
    for (int st = 0; st < LOG2N; ++st) {
        const int Ns = 1 << st;
        const float2* in = buf[cur];
        float2* out = buf[cur ^ 1];
#pragma unroll
        for (int r = 0; r < NFFT / 2 / 256; ++r) {
            const int j = tid + r * 256;
            const int k = j & (Ns - 1);
            const float2 w = tw[k << (LOG2N - 1 - st)];
            const float2 u0 = in[j], x1 = in[j + NFFT / 2];
            const float2 u1 = make_float2(x1.x * w.x - x1.y * w.y, x1.x * w.y + x1.y * w.x);
            const int j0 = (j << 1) - k;
            out[j0] = make_float2(u0.x + u1.x, u0.y + u1.y);
            out[j0 + Ns] = make_float2(u0.x - u1.x, u0.y - u1.y);
        }
        cur ^= 1;
        __syncthreads();
    }
    const float2* Z = buf[cur];
    float* pa = power + fa * pitch;
    float* pb = power + fb * pitch;
    for (int k = tid; k < NBINS; k += 256) {
        const float2 zk = Z[k], zn = Z[(NFFT - k) & (NFFT - 1)];
        const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);          // X_a = (Z[k] + conj Z[N-k]) / 2
        const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);         // X_b = (Z[k] - conj Z[N-k]) / (2i)
        pa[k] = ar * ar + ai * ai;
        if (has_b) pb[k] = br * br + bi * bi;
    }
    for (int k = NBINS + tid; k < pitch; k += 256) {
        pa[k] = 0.f;
        if (has_b) pb[k] = 0.f;
    }
}

// one CTA per spectrogram of n values: out = max(10 log10(max(x, 1e-10)), max_db - top_db)   (top_db < 0: no floor)
__global__ void __launch_bounds__(1024) power_to_db_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, float top_db) {
    __shared__ float red[32];
    const float* xs = x + (long long)blockIdx.x * n;
    float* os = out + (long long)blockIdx.x * n;
    float mx = -INFINITY;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, xs[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (threadIdx.x == 0) red[0] = mx;
    }
    __syncthreads();
    const float floor_db = top_db >= 0.f ? 10.f * log10f(fmaxf(red[0], 1e-10f)) - top_db : -INFINITY;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) os[i] = fmaxf(10.f * log10f(fmaxf(xs[i], 1e-10f)), floor_db);
}

}  // namespace

extern "C" int mmg_stft_power_f32(const float* wave, int B, long long L, long long wave_pitch, int n_fft, int hop, float* power, int pitch, void* stream) {
    MMG_REQUIRE(wave && power && B > 0 && hop > 0, MMG_EINVAL, "stft_power: bad argument");
    MMG_REQUIRE(n_fft == NFFT, MMG_EUNSUPPORTED, "stft_power: only n_fft = 2048 is built (GAN_DES/util.py:37)");
    MMG_REQUIRE(L > NFFT / 2, MMG_EINVAL, "stft_power: reflect padding needs more than n_fft / 2 samples");       // torch.stft raises for shorter inputs too
    MMG_REQUIRE(pitch >= NBINS && wave_pitch >= L, MMG_EINVAL, "stft_power: pitch smaller than the row");
    const int T = (int)(1 + L / hop);
    const long long n_frames = (long long)B * T;
    stft_power_kernel<<<(unsigned)((n_frames + 1) / 2), 256, 0, (cudaStream_t)stream>>>(wave, L, wave_pitch, hop, T, n_frames, power, pitch);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

extern "C" int mmg_power_to_db_f32(const float* x, float* out, int n_spectrograms, long long n, float top_db, void* stream) {
    MMG_REQUIRE(x && out && n_spectrograms > 0 && n > 0, MMG_EINVAL, "power_to_db: bad argument");
    power_to_db_kernel<<<n_spectrograms, 1024, 0, (cudaStream_t)stream>>>(x, out, n, top_db);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}
