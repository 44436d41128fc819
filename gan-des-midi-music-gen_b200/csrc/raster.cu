// Piano-roll rasteriser: note-event stream -> two 128 x W grids (velocity-at-onset, duration-fill).
//
// Replaces the CPython loop of the reference's generate_piano_roll
// (/root/reference/MMGAN_MIDI_DES/datasets.py:27-54) for a whole batch of songs.
// Integer-only apart from the float64 running time sum, which stays SEQUENTIAL per song (a parallel
// scan would not be bit-exact) and is rounded half-to-even exactly like Python's round().
//
// One fused kernel, one WARP per song (all songs in flight at once; nothing but the output leaves
// the SM):
//   1. the warp zero-fills its song's output planes (coalesced 16-byte stores);
//   2. per chunk of 256 messages: 32 lanes stage dt into shared memory (coalesced, next chunk
//      prefetched into registers), lane 0 runs the dependent t += dt chain in place, then all lanes
//      round to steps and evaluate the cut-off rule (first message with step >= S, first note_on with
//      step >= W) with ballots;
//   3. per 32 messages, in message order: lanes resolve same-pitch dependencies with match_any /
//      ballot (which note_on arms a note_off, which note_on is the last writer of a cell) and scatter
//      velocities / duration fills straight into the output (each note touches a few cells).
//      __syncwarp() between sub-steps gives the last-writer-wins order of the reference loop.
// HBM traffic = 12 B per message read once + every output cell written once (+ the touched cells).
#include "common.cuh"

namespace {

constexpr int RW = 4;     // warps (= songs) per CTA
constexpr int CH = 256;   // messages per chain chunk
constexpr int CJ = CH / 32;

template <typename OutT>
__device__ __forceinline__ OutT to_out(unsigned v);
template <> __device__ __forceinline__ float to_out<float>(unsigned v) { return (float)v; }
template <> __device__ __forceinline__ uint8_t to_out<uint8_t>(unsigned v) { return (uint8_t)(v > 255u ? 255u : v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(unsigned v) { return __float2bfloat16((float)v); }

template <typename OutT>
__global__ void __launch_bounds__(RW * 32) raster_fused_kernel(const double* __restrict__ dt, const uint32_t* __restrict__ meta,
                                                                const int64_t* __restrict__ offsets, int64_t n_songs, int S, int W,
                                                                int lo, int hi, OutT* __restrict__ out, int32_t* __restrict__ status) {
    __shared__ double tbuf[RW][CH];
    __shared__ int on_time[RW][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t song = (int64_t)blockIdx.x * RW + warp;
    if (song >= n_songs) return;                       // warp-uniform; no block-level barrier is used below
    const unsigned lt_mask = (1u << lane) - 1u, gt_mask = ~lt_mask & ~(1u << lane);
    const int64_t a0 = offsets[song];
    const int64_t n = offsets[song + 1] - a0;
    const int Wo = hi - lo;
    OutT* __restrict__ oroll = out + (size_t)song * 2 * 128 * Wo;
    OutT* __restrict__ odur = oroll + (size_t)128 * Wo;
    double* tb = tbuf[warp];
    int* on = on_time[warp];

    {   // 1. zero-fill (2*128*Wo*sizeof(OutT) bytes, a multiple of 16; the region is 16-byte aligned)
        uint4* z = reinterpret_cast<uint4*>(oroll);
        const int cnt = (int)((size_t)2 * 128 * Wo * sizeof(OutT) / 16);
        for (int i = lane; i < cnt; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        for (int p = lane; p < 128; p += 32) on[p] = 0;            // note_on_time = zeros(128) (:33)
    }
    __syncwarp();

    double d[CJ];
    uint32_t m[CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int64_t i = lane + 32 * j;
        d[j] = i < n ? dt[a0 + i] : 0.0;
        m[j] = i < n ? meta[a0 + i] : 0u;
    }
    double t = 0.0;                                                 // my_time (:32), carried by lane 0
    int st = 0;
    for (int64_t i0 = 0; i0 < n; i0 += CH) {
        uint32_t cm[CJ];
#pragma unroll
        for (int j = 0; j < CJ; ++j) { tb[lane + 32 * j] = d[j]; cm[j] = m[j]; }
#pragma unroll
        for (int j = 0; j < CJ; ++j) {                              // prefetch the next chunk behind the chain
            const int64_t i = i0 + CH + lane + 32 * j;
            d[j] = i < n ? dt[a0 + i] : 0.0;
            m[j] = i < n ? meta[a0 + i] : 0u;
        }
        __syncwarp();
        const int cnt = (int)((n - i0) < CH ? (n - i0) : CH);
        if (lane == 0) {                                            // 2. sequential float64 running sum (:35)
            const int lim8 = (cnt + 7) & ~7;                        // padding holds 0.0: t + 0.0 == t
            for (int k = 0; k < lim8; k += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = tb[k + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) { t = __dadd_rn(t, v[u]); tb[k + u] = t; }
            }
        }
        __syncwarp();
        int sj[CJ];
        int first_halt = CH;
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int e = lane + 32 * j;
            const long long step = __double2ll_rn(tb[e]);           // :36 round-half-even
            const uint32_t kind = cm[j] & 0xFFu, pitch = (cm[j] >> 8) & 0xFFu;
            const bool note = kind == 1u || kind == 2u;
            bool halt = step >= S;                                  // :37-38, every message kind
            halt |= step < 0;                                       // dt < 0: outside the contract (flagged)
            halt |= note && pitch >= 128u;                          // IndexError in the reference
            halt |= kind == 1u && step >= W;                        // IndexError -> bare except (:41,:46)
            const unsigned hm = __ballot_sync(0xffffffffu, e < cnt && halt);
            if (hm && first_halt == CH) {
                const int src = __ffs(hm) - 1;
                first_halt = 32 * j + src;
                const int bits = (step < 0 ? 1 : 0) | ((note && pitch >= 128u) ? 2 : 0);
                st = __shfl_sync(0xffffffffu, bits, src);
            }
            sj[j] = (int)(step < 0 ? 0 : (step > 0x7fffffff ? 0x7fffffff : step));
        }
        const int lim = cnt < first_halt ? cnt : first_halt;
        __syncwarp();
        // 3. replay, 32 messages at a time, in message order
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            if (32 * j >= lim) break;
            const bool valid = lane + 32 * j < lim;
            const uint32_t v = cm[j];
            const uint32_t kind = valid ? (v & 0xFFu) : 0u;
            const int p = (int)((v >> 8) & 0x7Fu);
            const int s = sj[j];
            const bool is_on = kind == 1u, is_off = kind == 2u;
            const unsigned pg = __match_any_sync(0xffffffffu, (is_on || is_off) ? (unsigned)p : 128u + lane);
            const unsigned onm = __ballot_sync(0xffffffffu, is_on), offm = __ballot_sync(0xffffffffu, is_off);
            // the note_on that arms this message: latest earlier note_on of the same pitch in this group, else carried state
            const unsigned lower_on = pg & onm & lt_mask;
            const int s_prev = __shfl_sync(0xffffffffu, s, lower_on ? 31 - __clz(lower_on) : lane);
            const int a = lower_on ? s_prev : on[p];
            // a note_on is overwritten if the next note_on of its pitch lands on the same step (steps never decrease)
            const unsigned higher_on = pg & onm & gt_mask;
            const int s_next = __shfl_sync(0xffffffffu, s, higher_on ? __ffs(higher_on) - 1 : lane);
            __syncwarp();
            if (is_on) {                                            // :39-42
                if (!(higher_on && s_next == s) && s >= lo && s < hi) oroll[(size_t)p * Wo + (s - lo)] = to_out<OutT>((v >> 16) & 0xFFu);
                if (!higher_on) on[p] = s;
            }
            // :43-45  durations[p, a:s] = s - a.  One note_off at a time, in message order (later fills overwrite
            // earlier ones); the whole warp writes each range, so the stores are coalesced along the row.
            unsigned rem = offm;
            while (rem) {
                const int src = __ffs(rem) - 1;
                rem &= rem - 1;
                const int pp = __shfl_sync(0xffffffffu, p, src), aa = __shfl_sync(0xffffffffu, a, src), ss = __shfl_sync(0xffffffffu, s, src);
                const int c0 = aa > lo ? aa : lo;
                int c1 = ss < W ? ss : W;
                c1 = c1 < hi ? c1 : hi;
                const OutT val = to_out<OutT>((unsigned)(ss - aa));
                OutT* row = odur + (size_t)pp * Wo - lo;
                for (int c = c0 + lane; c < c1; c += 32) row[c] = val;
                __syncwarp();
            }
        }
        if (first_halt < CH) break;
    }
    if (status && lane == 0) status[song] = st;
}

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

// the fused kernel keeps everything on chip: no scratch is needed (kept in the ABI for layout changes)
size_t mmg_raster_workspace_bytes(int64_t n_songs, int64_t total_events) {
    (void)n_songs; (void)total_events;
    return 0;
}

// out: (n_songs, 2, 128, Wout) of out_dtype (0 = float32, 1 = bfloat16, 2 = uint8 saturating); fully written.
int mmg_raster_piano_roll(const double* dt, const uint32_t* meta, const int64_t* offsets, int64_t n_songs, int64_t total_events,
                          int sequence_length, int start, int end, int out_dtype, void* out, int32_t* status,
                          void* workspace, size_t ws_bytes, void* stream_) {
    (void)workspace; (void)ws_bytes;
    cudaStream_t stream = (cudaStream_t)stream_;
    MMG_REQUIRE(n_songs >= 0 && total_events >= 0, MMG_EINVAL, "raster: negative sizes");
    const long W = (long)end - start;
    MMG_REQUIRE(W >= 0, MMG_EINVAL, "raster: end-start must be >= 0");
    const long S = sequence_length < 0 ? (long)end + 20 : sequence_length;     // datasets.py:14-15
    MMG_REQUIRE(S <= 0x3fffffff && W <= 0x3fffffff, MMG_EUNSUPPORTED, "raster: sequence_length / width too large");
    MMG_REQUIRE(out_dtype >= 0 && out_dtype <= 2, MMG_EINVAL, "raster: bad out_dtype %d", out_dtype);
    if (n_songs == 0) return MMG_OK;
    {
        long lo0, hi0;
        if (end < 128) clip_slice(start, end, W, &lo0, &hi0); else clip_slice(0, end, W, &lo0, &hi0);
        if (hi0 - lo0 == 0) return MMG_OK;                              // zero-width output: nothing to write
    }
    MMG_REQUIRE(offsets && out, MMG_EINVAL, "raster: null pointer");
    MMG_REQUIRE(total_events == 0 || (dt && meta), MMG_EINVAL, "raster: null event arrays");
    MMG_REQUIRE(((uintptr_t)out & 15) == 0, MMG_EINVAL, "raster: out must be 16-byte aligned");
    long lo, hi;
    if (end < 128) clip_slice(start, end, W, &lo, &hi); else clip_slice(0, end, W, &lo, &hi);
    if (hi - lo == 0) return MMG_OK;                                    // nothing to write
    const long long blocks = (n_songs + RW - 1) / RW;
    MMG_REQUIRE(blocks <= 0x7fffffff, MMG_EUNSUPPORTED, "raster: too many songs");
    if (out_dtype == 0)
        raster_fused_kernel<float><<<(int)blocks, RW * 32, 0, stream>>>(dt, meta, offsets, n_songs, (int)S, (int)W, (int)lo, (int)hi, (float*)out, status);
    else if (out_dtype == 1)
        raster_fused_kernel<__nv_bfloat16><<<(int)blocks, RW * 32, 0, stream>>>(dt, meta, offsets, n_songs, (int)S, (int)W, (int)lo, (int)hi, (__nv_bfloat16*)out, status);
    else
        raster_fused_kernel<uint8_t><<<(int)blocks, RW * 32, 0, stream>>>(dt, meta, offsets, n_songs, (int)S, (int)W, (int)lo, (int)hi, (uint8_t*)out, status);
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
