// Piano-roll rasteriser: note-event stream -> two 128 x W grids (velocity-at-onset, duration-fill).
//
// Replaces the CPython loop of the reference's generate_piano_roll
// (/root/reference/MMGAN_MIDI_DES/datasets.py:27-54) for a whole batch of songs.
// Integer-only apart from the float64 running time sum, which is kept SEQUENTIAL per song
// (one thread walks a song; a parallel scan would not be bit-exact) and rounded half-to-even
// exactly like Python's round().
//
// Two kernels:
//   raster_steps_kernel  one thread per song: t += dt; step = rn(t); finds the cut-off index
//                        (first message with step >= S, or first note_on with step >= W) and
//                        writes u16 steps for the messages before it.
//   raster_fill_kernel   one CTA per song: stable counting sort of the note messages by pitch
//                        (so each of the 128 rows sees its own events in message order), one
//                        thread per pitch row replays them into a shared-memory tile, then the
//                        whole CTA streams the tile out coalesced in the requested dtype.
// HBM-bound by design: every message is read once (12 B) (+6 B scratch round trip), every output
// cell is written exactly once.
#include "common.cuh"

namespace {

constexpr int FILL_THREADS = 512;
constexpr int FILL_WARPS = FILL_THREADS / 32;
constexpr int STEP_CHUNK = 8;

__global__ void __launch_bounds__(32) raster_steps_kernel(const double* __restrict__ dt, const uint32_t* __restrict__ meta,
                                                           const int64_t* __restrict__ offsets, int64_t n_songs, int S, int W,
                                                           uint16_t* __restrict__ steps, int32_t* __restrict__ cut,
                                                           int32_t* __restrict__ status) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_songs; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = offsets[s];
        const int64_t n = offsets[s + 1] - a;
        double t = 0.0;
        int64_t stop = n;
        int st = 0;
        for (int64_t i0 = 0; i0 < n && stop == n; i0 += STEP_CHUNK) {
            double d[STEP_CHUNK];
            uint32_t m[STEP_CHUNK];
#pragma unroll
            for (int j = 0; j < STEP_CHUNK; ++j) {      // independent loads first, dependent adds after
                const bool ok = i0 + j < n;
                d[j] = ok ? dt[a + i0 + j] : 0.0;
                m[j] = ok ? meta[a + i0 + j] : 0u;
            }
#pragma unroll
            for (int j = 0; j < STEP_CHUNK; ++j) {
                if (i0 + j < n && stop == n) {
                    t = __dadd_rn(t, d[j]);                          // datasets.py:35
                    const long long step = __double2ll_rn(t);        // :36 round-half-even
                    const uint32_t kind = m[j] & 0xFFu, pitch = (m[j] >> 8) & 0xFFu;
                    bool halt = step >= S;                            // :37-38 (any message kind)
                    if (step < 0) { halt = true; st |= 1; }           // dt < 0 is outside the contract
                    if ((kind == 1u || kind == 2u) && pitch >= 128u) { halt = true; st |= 2; }
                    if (kind == 1u && step >= W) halt = true;         // IndexError -> bare except (:41,:46)
                    if (halt) stop = i0 + j;
                    else steps[a + i0 + j] = (uint16_t)step;
                }
            }
        }
        cut[s] = (int32_t)stop;
        if (status) status[s] = st;
    }
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(unsigned v);
template <> __device__ __forceinline__ float to_out<float>(unsigned v) { return (float)v; }
template <> __device__ __forceinline__ uint8_t to_out<uint8_t>(unsigned v) { return (uint8_t)(v > 255u ? 255u : v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(unsigned v) { return __float2bfloat16((float)v); }

// packed sorted entry: step[0:16) | velocity[16:24) | kind[24:32)
template <typename OutT>
__global__ void __launch_bounds__(FILL_THREADS) raster_fill_kernel(const uint32_t* __restrict__ meta, const int64_t* __restrict__ offsets,
                                                                    const uint16_t* __restrict__ steps, const int32_t* __restrict__ cut,
                                                                    uint32_t* __restrict__ sorted, int W, int lo, int hi, int Wt,
                                                                    OutT* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem);               // [FILL_WARPS][128] -> cursors
    uint32_t* seg_begin = hist + FILL_WARPS * 128;                    // [129]
    uint16_t* dur_t = reinterpret_cast<uint16_t*>(seg_begin + 132);   // [128][Wt]
    uint8_t* roll_t = reinterpret_cast<uint8_t*>(dur_t + 128 * Wt);   // [128][Wt]

    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a = offsets[s];
    const int n = cut[s];
    const uint32_t* m = meta + a;
    const uint16_t* st = steps + a;
    uint32_t* srt = sorted + a;
    const int Wo = hi - lo;

    // ---- 1. per-warp pitch histograms over contiguous message slices
    for (int i = tid; i < FILL_WARPS * 128; i += FILL_THREADS) hist[i] = 0;
    __syncthreads();
    const int slice = ((n + FILL_WARPS - 1) / FILL_WARPS + 31) & ~31;
    const int w_lo = min(n, warp * slice), w_hi = min(n, w_lo + slice);
    for (int i = w_lo + lane; i < w_hi; i += 32) {
        const uint32_t v = m[i], kind = v & 0xFFu;
        if (kind == 1u || kind == 2u) atomicAdd(&hist[warp * 128 + ((v >> 8) & 0x7Fu)], 1u);
    }
    __syncthreads();
    // ---- 2. exclusive scan: pitch-major, warp-minor  (thread p owns pitch p)
    if (tid < 128) {
        uint32_t tot = 0;
        for (int w = 0; w < FILL_WARPS; ++w) { const uint32_t c = hist[w * 128 + tid]; hist[w * 128 + tid] = tot; tot += c; }
        seg_begin[tid + 1] = tot;       // per-pitch totals, scanned below
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        seg_begin[0] = 0;
        for (int p = 0; p < 128; ++p) { run += seg_begin[p + 1]; seg_begin[p + 1] = run; }
    }
    __syncthreads();
    // ---- 3. stable scatter: each warp walks its slice in message order, 32 at a time
    for (int i0 = w_lo; i0 < w_hi; i0 += 32) {
        const int i = i0 + lane;
        uint32_t v = 0, kind = 0;
        if (i < w_hi) { v = m[i]; kind = v & 0xFFu; }
        const bool note = (kind == 1u || kind == 2u);
        const uint32_t pitch = (v >> 8) & 0x7Fu;
        const uint32_t key = note ? pitch : 256u + lane;              // non-notes form singleton groups
        const uint32_t grp = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (note) base = hist[warp * 128 + pitch];
        __syncwarp();
        if (note && rank == 0) hist[warp * 128 + pitch] = base + __popc(grp);
        __syncwarp();
        if (note) srt[seg_begin[pitch] + base + rank] = (uint32_t)st[i] | (((v >> 16) & 0xFFu) << 16) | (kind << 24);
    }
    __syncthreads();

    // ---- 4. per column window: replay rows into the tile, stream it out
    OutT* out_roll = out + (size_t)s * 2 * 128 * Wo;
    OutT* out_dur = out_roll + (size_t)128 * Wo;
    for (int c0 = 0; c0 < W; c0 += Wt) {
        const int c1 = min(W, c0 + Wt), wt = c1 - c0;
        {   // zero the tile (dur_t and roll_t are contiguous: 3*128*Wt bytes, Wt % 16 == 0)
            uint4* z = reinterpret_cast<uint4*>(dur_t);
            const int nz = (3 * 128 * Wt) / 16;
            for (int i = tid; i < nz; i += FILL_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        if (tid < 128) {
            const int p = tid;
            const uint32_t b = seg_begin[p], e = seg_begin[p + 1];
            int on = 0;                                               // note_on_time[p] starts at 0 (:33)
            for (uint32_t k = b; k < e; ++k) {
                const uint32_t ent = srt[k];
                const int step = (int)(ent & 0xFFFFu);
                if ((ent >> 24) == 1u) {                              // note_on (:39-42)
                    if (step >= c0 && step < c1) roll_t[p * Wt + step - c0] = (uint8_t)((ent >> 16) & 0xFFu);
                    on = step;
                } else {                                              // note_off (:43-45): dur[p, on:step] = step-on
                    const int f0 = max(on, c0), f1 = min(min(step, W), c1);
                    const uint16_t val = (uint16_t)(step - on);
                    for (int c = f0; c < f1; ++c) dur_t[p * Wt + c - c0] = val;
                }
            }
        }
        __syncthreads();
        // coalesced write-out of columns [max(c0,lo), min(c1,hi)) (the :49-54 re-slice)
        const int x0 = max(c0, lo), x1 = min(c1, hi);
        if (x1 > x0) {
            const int span = x1 - x0;
            for (int i = tid; i < 128 * span; i += FILL_THREADS) {
                const int p = i / span, c = x0 + (i - p * span);
                out_roll[(size_t)p * Wo + (c - lo)] = to_out<OutT>(roll_t[p * Wt + c - c0]);
                out_dur[(size_t)p * Wo + (c - lo)] = to_out<OutT>(dur_t[p * Wt + c - c0]);
            }
        }
        __syncthreads();
        (void)wt;
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// python slice clip of [a:b) on a length-n axis
inline void clip_slice(long a, long b, long n, long* lo, long* hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    *lo = a; *hi = b > a ? b : a;
}

}  // namespace

extern "C" {

// width of the arrays the reference returns after its re-slice (datasets.py:49-54)
int mmg_raster_out_width(int start, int end) {
    long W = (long)end - start, lo, hi;
    if (W < 0) return -1;
    if (end < 128) clip_slice(start, end, W, &lo, &hi); else clip_slice(0, end, W, &lo, &hi);
    return (int)(hi - lo);
}

size_t mmg_raster_workspace_bytes(int64_t n_songs, int64_t total_events) {
    return align_up((size_t)total_events * 2, 256) + align_up((size_t)total_events * 4, 256) + align_up((size_t)n_songs * 4, 256);
}

// out: (n_songs, 2, 128, Wout) of out_dtype (0 = float32, 1 = bfloat16, 2 = uint8 saturating)
int mmg_raster_piano_roll(const double* dt, const uint32_t* meta, const int64_t* offsets, int64_t n_songs, int64_t total_events,
                          int sequence_length, int start, int end, int out_dtype, void* out, int32_t* status,
                          void* workspace, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(n_songs >= 0 && total_events >= 0, MMG_EINVAL, "raster: negative sizes");
    const long W = (long)end - start;
    MMG_REQUIRE(W >= 0, MMG_EINVAL, "raster: end-start must be >= 0");
    const int S = sequence_length < 0 ? end + 20 : sequence_length;     // datasets.py:14-15
    MMG_REQUIRE(S <= 65535 && W <= 65535, MMG_EUNSUPPORTED, "raster: sequence_length and width must be <= 65535");
    MMG_REQUIRE(out_dtype >= 0 && out_dtype <= 2, MMG_EINVAL, "raster: bad out_dtype %d", out_dtype);
    if (n_songs == 0) return MMG_OK;
    MMG_REQUIRE(offsets && out && workspace, MMG_EINVAL, "raster: null pointer");
    MMG_REQUIRE(total_events == 0 || (dt && meta), MMG_EINVAL, "raster: null event arrays");
    MMG_REQUIRE(ws_bytes >= mmg_raster_workspace_bytes(n_songs, total_events), MMG_EWORKSPACE, "raster: workspace too small");
    MMG_REQUIRE(n_songs <= 0x7fffffff, MMG_EUNSUPPORTED, "raster: too many songs");
    long lo, hi;
    if (end < 128) clip_slice(start, end, W, &lo, &hi); else clip_slice(0, end, W, &lo, &hi);
    if (hi - lo == 0) return MMG_OK;                                    // nothing to write

    unsigned char* ws = (unsigned char*)workspace;
    uint16_t* steps = (uint16_t*)ws;
    uint32_t* sorted = (uint32_t*)(ws + align_up((size_t)total_events * 2, 256));
    int32_t* cut = (int32_t*)((unsigned char*)sorted + align_up((size_t)total_events * 4, 256));

    const int g1 = (int)((n_songs + 31) / 32);
    raster_steps_kernel<<<g1, 32, 0, stream>>>(dt, meta, offsets, n_songs, S, (int)W, steps, cut, status);
    MMG_LAUNCH_CHECK();

    int Wt = (int)((W + 15) / 16 * 16);
    if (Wt > 512) Wt = 512;
    if (Wt < 16) Wt = 16;
    const size_t smem = (size_t)(FILL_WARPS * 128 + 132) * 4 + (size_t)3 * 128 * Wt;
#define MMG_RASTER_LAUNCH(T)                                                                                              \
    do {                                                                                                                  \
        MMG_CUDA(cudaFuncSetAttribute(raster_fill_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        raster_fill_kernel<T><<<(int)n_songs, FILL_THREADS, smem, stream>>>(meta, offsets, steps, cut, sorted, (int)W,    \
                                                                            (int)lo, (int)hi, Wt, (T*)out);               \
    } while (0)
    if (out_dtype == 0) MMG_RASTER_LAUNCH(float);
    else if (out_dtype == 1) MMG_RASTER_LAUNCH(__nv_bfloat16);
    else MMG_RASTER_LAUNCH(uint8_t);
#undef MMG_RASTER_LAUNCH
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
