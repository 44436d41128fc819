// Piano-roll rasteriser: note-event stream -> two 128 x W grids (velocity-at-onset, duration-fill).
//
// Replaces the CPython loop of the reference's generate_piano_roll
// (/root/reference/MMGAN_MIDI_DES/datasets.py:27-54) for a whole batch of songs.
// Integer-only apart from the float64 running time sum, which stays SEQUENTIAL per song (a parallel
// scan would not be bit-exact) and is rounded half-to-even exactly like Python's round().
//
// Two paths, both bit-exact.
// STREAM (default; no workspace): raster_stream_kernel, one 64-thread CTA per song -- a chain warp runs the t += dt chain while a
//   replay warp, one chunk behind, rounds / cuts off / scatters into a just-in-time zero-filled output (see the kernel's header).
// SORT (a workspace is supplied and S <= 65535):
//   K1 raster_steps_kernel  one WARP per song: 32 lanes stage dt through shared memory, lane 0 runs the dependent
//      t += dt chain, all lanes round to steps, evaluate the cut-off rule with ballots and COMPACT the surviving
//      note_on / note_off messages into 4-byte records (step | off<<16 | pitch<<17 | vel<<24) in the workspace.
//   K2 raster_rows_kernel   one CTA per song: zero-fills the song's planes, stable-counting-sorts the notes by pitch
//      (per-warp segment histograms, no atomics), then one warp per pitch replays the pitch's list 32 notes at a time
//      with a closed form of the sequential rules, writing every touched cell exactly once (see the kernel's header).
// HBM traffic = 12 B per message read once + every output cell written once (SORT: + 4 B per surviving note written and
// read back through the workspace).
#include "common.cuh"

namespace {

template <typename OutT>
__device__ __forceinline__ OutT to_out(unsigned v);
template <> __device__ __forceinline__ float to_out<float>(unsigned v) { return (float)v; }
template <> __device__ __forceinline__ uint8_t to_out<uint8_t>(unsigned v) { return (uint8_t)(v > 255u ? 255u : v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(unsigned v) { return __float2bfloat16((float)v); }

// ------------------------------------------------------------------------------------------------
// STREAM path: one CTA per song (one chain warp + SK_R replay warps, each replay warp owning the pitches p % SK_R == w), warp-specialised.
//   warp 0 (chain)   stages 512 dt values into one half of a double buffer (next chunk prefetched into registers) and lane 0 runs the
//                    dependent t += dt chain in place -- nothing else sits on the chain's critical path;
//   warps 1.. (replay) work one chunk behind; each rounds its share of the chunk, then all see the whole chunk: rounds the prefix sums to steps, evaluates the cut-off rule with ballots, and BINS the
//                    chunk's notes by pitch in shared memory, 32 messages at a time in message order (match_any ranks; 8 slots per
//                    pitch).  When a bin is full, and at the end of the song, the bins are DRAINED lane-per-pitch: every lane replays
//                    its pitch's notes with the reference's own two rules (rows are independent, so per-pitch message order is all
//                    that last-writer-wins needs), scattering velocities and duration fills straight into the output.
//                    The output is zero-filled JUST IN TIME, one 128-byte column block of all 256 rows at a time, right before the
//                    first note whose step reaches the block: steps never decrease, so the lines a drain touches were written a few
//                    microseconds earlier and are still in L2 -- DRAM sees every output line once, as a full line.
//   One __syncthreads() per chunk hands chunk k to the replay warp and buffer (k+1)&1 back to the chain warp.
// The whole batch is in flight at once (15 KB of shared memory per song); the kernel's floor is the fp64 add latency x messages per song.
// ------------------------------------------------------------------------------------------------
constexpr int SK_CH_LONG = 512;       // messages per chunk; short songs (a few hundred messages: the training loop's simulated songs) use 128 so that the
constexpr int SK_CH_SHORT = 128;      // chain of chunk k + 1 still overlaps the replay of chunk k
constexpr int SK_CAP = 8;             // bin capacity per pitch (a full bin triggers a drain of the warp's bins)
constexpr int SK_R = 2;               // replay warps per song; warp w owns the pitches p with p % SK_R == w (rows are independent)

constexpr int SK_TILE_BYTES = 12800;  // songs whose two planes fit this many bytes (the training loop's 2 x 128 x 50 uint8 rolls) are built in SHARED
                                      // memory and leave the SM as coalesced 16-byte stores: no scattered global stores at all

template <typename OutT, int R, int SK_CH, bool SMEM_OUT>
__global__ void __launch_bounds__(32 * (1 + R), R == 1 ? 12 : 9)
raster_stream_kernel(const double* __restrict__ dt, const uint32_t* __restrict__ meta, const int64_t* __restrict__ offsets, int S, int W, int lo,
                     int hi, OutT* __restrict__ out, int32_t* __restrict__ status) {
    constexpr int SK_CJ = SK_CH / 32;
    static_assert(SK_CJ % R == 0 && 128 % (32 * R) == 0 && SK_CH % 32 == 0, "replay warps must divide the chunk groups and the pitch range");
    __shared__ __align__(16) double tbuf[2][SK_CH];
    __shared__ int bin_s[SK_CAP][128];              // per-pitch bins: steps ...
    __shared__ uint16_t bin_v[SK_CAP][128];         // ... and off | velocity << 1, in message order
    __shared__ int bcnt[128];
    __shared__ int on[128];
    __shared__ int fh[R], fst[R];                   // per replay warp: first halting message of its share of the chunk, its status bits
    __shared__ int halt_flag;
    __shared__ uint4 otile[SMEM_OUT ? SK_TILE_BYTES / 16 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t song = blockIdx.x;
    const int64_t a0 = offsets[song];
    const int64_t n = offsets[song + 1] - a0;
    const int64_t nchunks = (n + SK_CH - 1) / SK_CH;
    if (threadIdx.x == 0) halt_flag = 0;

    if (warp == 0) {
        // ---------------------------------------------------------------- chain warp
        double d[SK_CJ];
#pragma unroll
        for (int j = 0; j < SK_CJ; ++j) { const int64_t i = lane + 32 * j; d[j] = i < n ? dt[a0 + i] : 0.0; }
        double t = 0.0;                                             // my_time (datasets.py:32), carried by lane 0
        for (int64_t k = 0; k <= nchunks; ++k) {
            if (k < nchunks) {
                double* tb = tbuf[k & 1];
#pragma unroll
                for (int j = 0; j < SK_CJ; ++j) tb[lane + 32 * j] = d[j];
#pragma unroll
                for (int j = 0; j < SK_CJ; ++j) {                   // prefetch the next chunk behind the chain
                    const int64_t i = (k + 1) * SK_CH + lane + 32 * j;
                    d[j] = i < n ? dt[a0 + i] : 0.0;
                }
                __syncwarp();
                if (lane == 0) {                                    // sequential float64 running sum (datasets.py:35); padding holds 0.0: t + 0.0 == t
                    const int64_t rem = n - k * SK_CH;
                    const int lim8 = rem >= SK_CH ? SK_CH : (((int)rem + 7) & ~7);
                    double2* tb2 = reinterpret_cast<double2*>(tb);
                    double2 v[4], w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = tb2[u];
                    for (int q = 0; q < lim8; q += 8) {
                        if (q + 8 < lim8) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) w[u] = tb2[(q >> 1) + 4 + u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            t = __dadd_rn(t, v[u].x); v[u].x = t;
                            t = __dadd_rn(t, v[u].y); v[u].y = t;
                            tb2[(q >> 1) + u] = v[u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = w[u];
                    }
                }
            }
            __syncthreads();
            if (halt_flag) break;
        }
        return;
    }

    // -------------------------------------------------------------------- replay warps
    constexpr int GJ = SK_CJ / R;                                   // groups of 32 messages this warp rounds per chunk
    constexpr int NP = 128 / R;                                     // pitches this warp owns
    const int rw = warp - 1;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int Wo = hi - lo;
    OutT* gout = out + (size_t)song * 2 * 128 * Wo;
    OutT* oroll = SMEM_OUT ? reinterpret_cast<OutT*>(otile) : gout;         // SMEM_OUT: the song's planes are built in shared memory
    OutT* odur = oroll + (size_t)128 * Wo;
    // zero fill: just in time by 128-byte column blocks when a row is a whole number of 16-byte words, else the whole song up front
    const int row_bytes = Wo * (int)sizeof(OutT);
    const bool jit = !SMEM_OUT && (row_bytes & 15) == 0;
    const int rowq = row_bytes >> 4;                                // 16-byte words per row
    const int nblk = jit ? (rowq + 7) >> 3 : 0;
    constexpr int CPB = 128 / (int)sizeof(OutT);                    // output columns per fill block
    int filled = 0;                                                 // fill blocks done (of this warp's rows)
    auto fill_block = [&](int b) {                                  // this warp's 2 * NP rows (both planes of its pitches), 8 words each
        uint4* z = reinterpret_cast<uint4*>(oroll);
#pragma unroll 4
        for (int idx = lane; idx < 2 * NP * 8; idx += 32) {
            const int q = b * 8 + (idx & 7), rr = idx >> 3;
            const int row = (rr / NP) * 128 + rw + R * (rr % NP);
            if (q < rowq) z[(size_t)row * rowq + q] = make_uint4(0, 0, 0, 0);
        }
    };
    if (!jit) {
        uint4* z = reinterpret_cast<uint4*>(oroll);
        const int cnt = (int)((size_t)2 * 128 * Wo * sizeof(OutT) / 16);        // the song's planes are a multiple of 16 bytes, 16-byte aligned
        for (int i = lane + 32 * rw; i < cnt; i += 32 * R) z[i] = make_uint4(0, 0, 0, 0);
        if (R > 1) asm volatile("bar.sync 1, %0;" ::"r"(32 * R) : "memory");   // rows are shared between the replay warps in this mode
    }
    for (int i = lane; i < NP; i += 32) { on[rw + R * i] = 0; bcnt[rw + R * i] = 0; }  // note_on_time = zeros(128) (:33); empty bins
    __syncwarp();
    // Drain: lane-per-pitch replay of the binned notes, literally the reference loop (:39-45) restricted to one pitch row -- rows are
    // independent, and a pitch's notes sit in its bin in message order, so last-writer-wins is preserved.
    auto drain = [&]() {
#pragma unroll 1
        for (int q = 0; q < NP / 32; ++q) {
            const int p = rw + R * (lane + 32 * q);
            const int c = bcnt[p];
            if (c == 0) continue;
            int on_t = on[p];
            OutT* rrow = oroll + (size_t)p * Wo - lo;
            OutT* drow = odur + (size_t)p * Wo - lo;
#pragma unroll 1
            for (int i = 0; i < c; ++i) {
                const int s = bin_s[i][p];
                const unsigned ov = bin_v[i][p];
                if (!(ov & 1u)) {                                   // note_on (:39-42)
                    if (s >= lo && s < hi) rrow[s] = to_out<OutT>(ov >> 1);
                    on_t = s;
                } else {                                            // note_off (:43-45): durations[p, on:s] = s - on
                    int c1 = s < W ? s : W;
                    c1 = c1 < hi ? c1 : hi;
                    const OutT val = to_out<OutT>((unsigned)(s - on_t));
                    for (int cc = on_t > lo ? on_t : lo; cc < c1; ++cc) drow[cc] = val;
                }
            }
            on[p] = on_t;
            bcnt[p] = 0;
        }
        __syncwarp();
    };
    uint32_t m[GJ];                                                 // meta of this warp's share of the next chunk
#pragma unroll
    for (int j = 0; j < GJ; ++j) { const int64_t i = lane + 32 * (rw * GJ + j); m[j] = i < n ? meta[a0 + i] : 0u; }
    int st = 0;
    for (int64_t k = 0; k <= nchunks; ++k) {
        if (k >= 1) {
            const int64_t i0 = (k - 1) * SK_CH;
            double* tb = tbuf[(k - 1) & 1];
            int2* sm = reinterpret_cast<int2*>(tb);                 // each slot is rewritten in place as (step, meta) once it is rounded
            const int cnt = (int)((n - i0) < SK_CH ? (n - i0) : SK_CH);
            int first_halt = SK_CH, st_new = 0;
#pragma unroll
            for (int j = 0; j < GJ; ++j) {
                const int e = lane + 32 * (rw * GJ + j);
                const long long step = __double2ll_rn(tb[e]);       // :36 round-half-even
                const uint32_t kind = m[j] & 0xFFu, pitch = (m[j] >> 8) & 0xFFu;
                const bool note = kind == 1u || kind == 2u;
                bool halt = step >= S;                              // :37-38, every message kind
                halt |= step < 0;                                   // dt < 0: outside the contract (flagged)
                halt |= note && pitch >= 128u;                      // IndexError in the reference
                halt |= kind == 1u && step >= W;                    // IndexError -> bare except (:41,:46)
                const unsigned hm = __ballot_sync(0xffffffffu, e < cnt && halt);
                if (hm && first_halt == SK_CH) {
                    const int src = __ffs(hm) - 1;
                    first_halt = 32 * (rw * GJ + j) + src;
                    const int bits = (step < 0 ? 1 : 0) | ((note && pitch >= 128u) ? 2 : 0);
                    st_new = __shfl_sync(0xffffffffu, bits, src);
                }
                sm[e] = make_int2((int)(step < 0 ? 0 : (step > 0x7fffffff ? 0x7fffffff : step)), (int)m[j]);
            }
#pragma unroll
            for (int j = 0; j < GJ; ++j) {                          // prefetch this warp's share of the next chunk's meta
                const int64_t i = i0 + SK_CH + lane + 32 * (rw * GJ + j);
                m[j] = i < n ? meta[a0 + i] : 0u;
            }
            if (R > 1) {                                            // the earliest halt over all shares decides
                if (lane == 0) { fh[rw] = first_halt; fst[rw] = st_new; }
                asm volatile("bar.sync 1, %0;" ::"r"(32 * R) : "memory");
                first_halt = SK_CH;
#pragma unroll
                for (int w = R - 1; w >= 0; --w)
                    if (fh[w] < SK_CH) { first_halt = fh[w]; st_new = fst[w]; }
            }
            if (first_halt < SK_CH) st = st_new;
            const int lim = cnt < first_halt ? cnt : first_halt;
            // bin this warp's pitches, 32 messages at a time, in message order (rolled loop: the body must stay in the instruction cache)
            // (the match_any of group g + 1 is issued before group g is processed: its ~200-cycle latency hides behind the binning)
            auto load_group = [&](int e0, uint32_t& v, int& s, bool& note, unsigned& grp) {
                const int2 rec = sm[e0 + lane];
                v = (uint32_t)rec.y;
                s = rec.x;
                const uint32_t kind = (e0 + lane < lim) ? (v & 0xFFu) : 0u;
                const int p = (int)((v >> 8) & 0x7Fu);
                note = (kind == 1u || kind == 2u) && (R == 1 || (p % R) == rw);
                grp = __match_any_sync(0xffffffffu, note ? (unsigned)p : 128u + lane);
            };
            uint32_t v_n = 0; int s_n = 0; bool note_n = false; unsigned grp_n = 0;
            if (lim > 0) load_group(0, v_n, s_n, note_n, grp_n);
#pragma unroll 1
            for (int e0 = 0; e0 < lim; e0 += 32) {
                const uint32_t v = v_n;
                const int s = s_n;
                const bool note = note_n;
                const unsigned grp = grp_n;
                if (e0 + 32 < lim) load_group(e0 + 32, v_n, s_n, note_n, grp_n);
                const uint32_t kind = v & 0xFFu;
                const int p = (int)((v >> 8) & 0x7Fu);
                unsigned pend = __ballot_sync(0xffffffffu, note);
                if (!pend) continue;
                if (jit) {                                          // zero the column blocks this group reaches
                    int need = __reduce_max_sync(0xffffffffu, note ? s : 0);
                    need = (need < hi ? need : hi - 1) - lo;
                    if (need >= filled * CPB) {
                        while (filled < nblk && filled * CPB <= need) fill_block(filled++);
                        __syncwarp();
                    }
                }
                bool mine = note;
                while (true) {
                    const unsigned g = grp & pend;                  // still-unbinned notes of my pitch
                    const int base = bcnt[p];
                    const int r = __popc(g & lt_mask);
                    const bool fits = mine && base + r < SK_CAP;
                    if (fits) {
                        bin_s[base + r][p] = s;
                        bin_v[base + r][p] = (uint16_t)((kind == 2u ? 1u : 0u) | (((v >> 16) & 0xFFu) << 1));
                    }
                    __syncwarp();
                    if (mine && r == 0) { const int c = __popc(g); bcnt[p] = base + (c < SK_CAP - base ? c : SK_CAP - base); }
                    __syncwarp();
                    mine = mine && !fits;
                    pend = __ballot_sync(0xffffffffu, mine);
                    if (!pend) break;
                    drain();                                        // a bin is full: replay everything binned so far, then go on
                }
            }
            if (first_halt < SK_CH && rw == 0 && lane == 0) halt_flag = 1;
        }
        __syncthreads();
        if (halt_flag) break;
    }
    __syncwarp();
    drain();
    while (filled < nblk) fill_block(filled++);                     // columns no note of this warp's pitches reached
    if (SMEM_OUT) {                                                 // the finished planes leave the SM once, as coalesced 16-byte stores
        if (R > 1) asm volatile("bar.sync 1, %0;" ::"r"(32 * R) : "memory");
        const int cnt = (int)((size_t)2 * 128 * Wo * sizeof(OutT) / 16);
        uint4* g = reinterpret_cast<uint4*>(gout);
        for (int i = lane + 32 * rw; i < cnt; i += 32 * R) g[i] = otile[i];
    }
    if (status && rw == 0 && lane == 0) status[song] = st;
}

// ------------------------------------------------------------------------------------------------
// fast path, K1: time chain + cut-off + note compaction (one warp per song)
// ------------------------------------------------------------------------------------------------
constexpr int K1_WARPS = 4;
constexpr int K1_CH = 512;            // messages per chunk
constexpr int K1_CJ = K1_CH / 32;
constexpr int K1_TB = K1_CH + K1_CH / 16;      // prefix-sum buffer of a warp; the parallel mode pads every 16 doubles by one (conflict-free transposition)

// The time steps of datasets.py:35-36 are round_half_even(s_i) with s_i the SEQUENTIAL float64 running sum.  Running that chain on one lane
// costs >= one dependent DADD per message (round 1: 261 us for 19 M messages, the GPU 7x under-subscribed).  SPECULATE AND VERIFY instead:
//   * a parallel prefix sum A_i of the same dt (any order of float64 additions of non-negative terms: |A_i - T_i| <= gamma_h T_i with T_i the
//     exact sum and h the height of the addition tree, here h <= 64), while the sequential sum obeys |s_i - T_i| <= gamma_i T_i
//     (gamma_k = k u / (1 - k u), u = 2^-53; Higham, Accuracy and Stability of Numerical Algorithms, section 4.2);
//   * hence |s_i - A_i| <= (gamma_i + gamma_64) T_i < delta_i := 2 (i + 64) 2^-53 A_i, and round_half_even(s_i) == round(A_i) whenever no
//     k + 1/2 lies within delta_i of A_i (rounding only changes at half-integers);
//   * a song with ANY message that fails this test (or a negative / non-finite dt, where the bound does not hold) is redone from its first
//     message by the exact sequential chain.  Bit-exactness is therefore preserved by construction; for continuous dt the fallback rate is
//     ~ n * 2 delta ~ 3e-5 per 15 000-message song, while streams made of exact half-integers (the rounding tests) always take the chain.
// PAR = parallel mode.  Returns false when the song has to be redone sequentially.
template <bool PAR>
__device__ __forceinline__ bool steps_song(const double* __restrict__ dt, const uint32_t* __restrict__ meta, int64_t a0, int64_t n, int S, int W,
                                           double* tb, uint32_t* __restrict__ nout, int lane, int& count_out, unsigned& bigvel_out, int& st_out) {
    const unsigned lt_mask = (1u << lane) - 1u;
    auto TI = [](int e) { return PAR ? e + (e >> 4) : e; };
    double d[K1_CJ];
    uint32_t m[K1_CJ];
#pragma unroll
    for (int j = 0; j < K1_CJ; ++j) {
        const int64_t i = lane + 32 * j;
        d[j] = i < n ? dt[a0 + i] : 0.0;
        m[j] = i < n ? meta[a0 + i] : 0u;
    }
    double t = 0.0;                                                // my_time (datasets.py:32): lane 0's chain, or the chunk carry of the scan
    int st = 0, count = 0;
    unsigned bigvel = 0;                                           // any kept note with velocity > 127 (does not fit K2's packed shared-memory record)
    for (int64_t i0 = 0; i0 < n; i0 += K1_CH) {
        uint32_t cm[K1_CJ];
        bool bad = false;                                          // PAR: a term the error bound does not cover
#pragma unroll
        for (int j = 0; j < K1_CJ; ++j) {
            tb[TI(lane + 32 * j)] = d[j];
            cm[j] = m[j];
            if (PAR) bad |= !(d[j] >= 0.0 && d[j] <= 1.7e308);
        }
#pragma unroll
        for (int j = 0; j < K1_CJ; ++j) {                          // prefetch the next chunk behind the sums
            const int64_t i = i0 + K1_CH + lane + 32 * j;
            d[j] = i < n ? dt[a0 + i] : 0.0;
            m[j] = i < n ? meta[a0 + i] : 0u;
        }
        __syncwarp();
        const int cnt = (int)((n - i0) < K1_CH ? (n - i0) : K1_CH);
        if (PAR) {
            if (__any_sync(0xffffffffu, bad)) return false;
            // lane l owns messages 16 l .. 16 l + 15 of the chunk: local prefix, warp scan of the lane totals, chunk carry (padding holds 0.0)
            double p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = tb[17 * lane + j];
#pragma unroll
            for (int j = 1; j < 16; ++j) p[j] = __dadd_rn(p[j - 1], p[j]);
            double inc = p[15];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc = __dadd_rn(inc, v);
            }
            double excl = __shfl_up_sync(0xffffffffu, inc, 1);    // sum of the lanes below
            if (lane == 0) excl = 0.0;
            const double off = __dadd_rn(t, excl);
#pragma unroll
            for (int j = 0; j < 16; ++j) tb[17 * lane + j] = __dadd_rn(off, p[j]);
            t = __shfl_sync(0xffffffffu, __dadd_rn(off, p[15]), 31);
        } else if (lane == 0) {                                    // sequential float64 running sum (datasets.py:35)
            // 8 elements per trip: 4 LDS.128 for the next trip are in flight behind the 8 dependent adds, 4 STS.128 write the prefix sums back
            const int lim8 = (cnt + 7) & ~7;                       // padding holds 0.0: t + 0.0 == t
            double2* tb2 = reinterpret_cast<double2*>(tb);
            double2 v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = tb2[u];
            for (int k = 0; k < lim8; k += 8) {
                if (k + 8 < lim8) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) w[u] = tb2[(k >> 1) + 4 + u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    t = __dadd_rn(t, v[u].x); v[u].x = t;
                    t = __dadd_rn(t, v[u].y); v[u].y = t;
                    tb2[(k >> 1) + u] = v[u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = w[u];
            }
        }
        __syncwarp();
        int first_halt = K1_CH;
        uint32_t rec[K1_CJ];
        unsigned keep = 0;                                         // bit j: message lane+32j is a note inside the contract
        bool unsure = false;
#pragma unroll
        for (int j = 0; j < K1_CJ; ++j) {
            const int e = lane + 32 * j;
            const double tv = tb[TI(e)];
            const long long step = __double2ll_rn(tv);             // datasets.py:36 round-half-even
            if (PAR && e < cnt) {                                  // is round(s_i) certain to equal round(A_i)?
                const double dist = fabs((tv - floor(tv)) - 0.5);
                const double delta = (double)(i0 + e + 64) * tv * 2.220446049250313e-16;      // 2 (i + 64) 2^-53 A_i
                unsure |= !(dist > delta);
            }
            const uint32_t kind = cm[j] & 0xFFu, pitch = (cm[j] >> 8) & 0xFFu;
            const bool note = kind == 1u || kind == 2u;
            bool halt = step >= S;                                 // :37-38, every message kind
            halt |= step < 0;                                      // dt < 0: outside the contract (flagged)
            halt |= note && pitch >= 128u;                         // IndexError in the reference
            halt |= kind == 1u && step >= W;                       // IndexError -> bare except (:41,:46)
            const unsigned hm = __ballot_sync(0xffffffffu, e < cnt && halt);
            if (hm && first_halt == K1_CH) {
                const int src = __ffs(hm) - 1;
                first_halt = 32 * j + src;
                const int bits = (step < 0 ? 1 : 0) | ((note && pitch >= 128u) ? 2 : 0);
                st = __shfl_sync(0xffffffffu, bits, src);
            }
            rec[j] = ((uint32_t)step & 0xFFFFu) | ((kind == 2u ? 1u : 0u) << 16) | ((pitch & 0x7Fu) << 17) | (((cm[j] >> 16) & 0xFFu) << 24);
            if (note && e < cnt) keep |= 1u << j;
        }
        if (PAR && __any_sync(0xffffffffu, unsure)) return false;
        const int lim = cnt < first_halt ? cnt : first_halt;
#pragma unroll
        for (int j = 0; j < K1_CJ; ++j) {                          // ordered compaction, 32 messages per ballot
            const bool k = ((keep >> j) & 1u) && (lane + 32 * j < lim);
            const unsigned km = __ballot_sync(0xffffffffu, k);
            if (k) nout[count + __popc(km & lt_mask)] = rec[j];
            bigvel |= __ballot_sync(0xffffffffu, k && (rec[j] >> 31));
            count += __popc(km);
        }
        if (first_halt < K1_CH) break;
        __syncwarp();
    }
    count_out = count; bigvel_out = bigvel; st_out = st;
    return true;
}

// mode: 0 = speculate and verify (sequential chain only for the songs that need it), 1 = sequential chain for every song (round 1's kernel)
__global__ void __launch_bounds__(K1_WARPS * 32, 3) raster_steps_kernel(const double* __restrict__ dt, const uint32_t* __restrict__ meta,
                                                                      const int64_t* __restrict__ offsets, int64_t n_songs, int S, int W,
                                                                      uint32_t* __restrict__ notes, int32_t* __restrict__ note_count,
                                                                      int32_t* __restrict__ status, int mode, unsigned long long* __restrict__ n_fallback) {
    __shared__ __align__(16) double tbuf[K1_WARPS][K1_TB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t song = (int64_t)blockIdx.x * K1_WARPS + warp;
    if (song >= n_songs) return;
    const int64_t a0 = offsets[song];
    const int64_t n = offsets[song + 1] - a0;
    int count = 0, st = 0;
    unsigned bigvel = 0;
    bool done = false;
    if (mode == 0) {
        done = steps_song<true>(dt, meta, a0, n, S, W, tbuf[warp], notes + a0, lane, count, bigvel, st);
        if (!done && lane == 0 && n_fallback) atomicAdd(n_fallback, 1ull);
        __syncwarp();
    }
    if (!done) steps_song<false>(dt, meta, a0, n, S, W, tbuf[warp], notes + a0, lane, count, bigvel, st);
    if (lane == 0) {
        note_count[song] = count | (bigvel ? 0x40000000 : 0);
        if (status) status[song] = st;
    }
}

// ------------------------------------------------------------------------------------------------
// fast path, K2: stable counting sort of the song's notes by pitch, then a write-once replay (one CTA per song)
//
// For one pitch the notes arrive in message order with non-decreasing steps, so the sequential rules of
// datasets.py:39-45 have a closed form that needs only each note's neighbours in its pitch list:
//   note_on  i : roll[p, s_i] = vel_i unless the NEXT note_on of the pitch has the same step (last writer wins);
//   note_off j : a_j = step of the latest note_on before it (0 if none); it fills [a_j, s_j) with s_j - a_j, and is
//                overwritten from a_{j'} onwards by the next note_off j' (a is non-decreasing, s is non-decreasing),
//                so it OWNS exactly the columns [a_j, min(a_{j'}, s_j, W)).
// Every cell is therefore written at most once after the CTA's zero fill (which stays in L2 for the few microseconds in
// between: HBM sees each output line once).
// ------------------------------------------------------------------------------------------------
constexpr int K2_THREADS = 128;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int K2_CAP = 6656;                  // notes whose sorted copy fits shared memory (3 bytes each: 19.5 KB -> 9+ CTAs per SM, every song of a
                                              // MAESTRO-scale batch resident at once); longer songs sort through the global workspace instead

constexpr int K2_ROW_CAP = 304;               // output widths up to this many columns are built row by row in shared memory (u8 velocity + u16 duration per
                                              // cell and warp) and leave the SM as coalesced full-row stores: no zero-fill pass, no scattered global stores

template <typename OutT>
__global__ void __launch_bounds__(K2_THREADS) raster_rows_kernel(const uint32_t* __restrict__ notes, uint32_t* __restrict__ sorted,
                                                                  const int32_t* __restrict__ note_count, const int64_t* __restrict__ offsets,
                                                                  int W, int lo, int hi, OutT* __restrict__ out) {
    __shared__ __align__(16) unsigned char scratch[K2_WARPS * K2_ROW_CAP * 3];      // phases 1-3: hist; phase 4: the warps' row buffers
    static_assert(sizeof(int) * K2_WARPS * 128 <= sizeof(scratch), "hist must fit the scratch area");
    int (*hist)[128] = reinterpret_cast<int (*)[128]>(scratch);       // per-warp-segment pitch histogram, then running scatter base
    __shared__ int pstart[129];
    __shared__ int wsum[4];
    __shared__ uint16_t s_step[K2_CAP];          // sorted notes, shared-memory form: step | (off | velocity << 1)
    __shared__ uint8_t s_ov[K2_CAP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u, gt_mask = ~lt_mask & ~(1u << lane);
    const int64_t song = blockIdx.x;
    const uint32_t* nin = notes + offsets[song];
    uint32_t* srt = sorted + offsets[song];
    const int nc = note_count[song];
    const int n = nc & 0x3fffffff;
    const bool in_smem = n <= K2_CAP && !(nc & 0x40000000);
    auto put = [&](int pos, uint32_t r) {
        if (in_smem) { s_step[pos] = (uint16_t)r; s_ov[pos] = (uint8_t)(((r >> 16) & 1u) | ((r >> 24) << 1)); }
        else srt[pos] = r;
    };
    auto get = [&](int pos) -> uint32_t {
        if (in_smem) { const uint32_t ov = s_ov[pos]; return (uint32_t)s_step[pos] | ((ov & 1u) << 16) | ((ov >> 1) << 24); }
        return srt[pos];
    };
    const int Wo = hi - lo;
    OutT* __restrict__ oroll = out + (size_t)song * 2 * 128 * Wo;
    OutT* __restrict__ odur = oroll + (size_t)128 * Wo;
    // Zero fill.  When 8 rows are a multiple of 16 bytes the fill is done LATER, 8 pitch rows at a time by the warp that replays those
    // pitches right afterwards, so the few cells the notes overwrite are still in L2 (a whole-song fill up front is evicted before the
    // replay gets to it: every touched line then costs a DRAM read-modify-write).
    const bool rowbuf = Wo <= K2_ROW_CAP && sizeof(OutT) == 4;        // rows built in shared memory (measured: +5 % for float32 rows, -3 % for uint8 rows)
    const bool late_fill = !rowbuf && ((size_t)8 * Wo * sizeof(OutT)) % 16 == 0;
    if (!late_fill && !rowbuf) {   // 2*128*Wo*sizeof(OutT) bytes, a multiple of 16; 16-byte aligned
        uint4* z = reinterpret_cast<uint4*>(oroll);
        const int cnt = (int)((size_t)2 * 128 * Wo * sizeof(OutT) / 16);
        for (int i = tid; i < cnt; i += K2_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < K2_WARPS * 128; i += K2_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    // 1. histogram of this warp's contiguous segment (multiple of 32 notes); 4 groups of loads in flight
    const int seg = ((n + K2_WARPS * 32 - 1) / (K2_WARPS * 32)) * 32;
    const int s_lo = warp * seg, s_hi = min(n, s_lo + seg);
    for (int i0 = s_lo; i0 < s_hi; i0 += 128) {
        uint32_t rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u + lane; rr[u] = i < s_hi ? nin[i] : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + 32 * u >= s_hi) break;
            const bool valid = i0 + 32 * u + lane < s_hi;
            const int p = valid ? (int)((rr[u] >> 17) & 0x7Fu) : 128 + lane;
            const unsigned grp = __match_any_sync(0xffffffffu, p);
            if (valid && !(grp & lt_mask)) hist[warp][p] += __popc(grp);
            __syncwarp();
        }
    }
    __syncthreads();
    // 2. pitch offsets (exclusive scan over 128 pitches) and per-segment bases
    {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < K2_WARPS; ++w) tot += hist[w][tid];
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int base = inc - tot;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        pstart[tid] = base;
        if (tid == 127) pstart[128] = base + tot;
#pragma unroll
        for (int w = 0; w < K2_WARPS; ++w) { const int c = hist[w][tid]; hist[w][tid] = base; base += c; }
    }
    __syncthreads();
    // 3. stable scatter (each warp in message order over its own segment)
    for (int i0 = s_lo; i0 < s_hi; i0 += 128) {
        uint32_t rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u + lane; rr[u] = i < s_hi ? nin[i] : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + 32 * u >= s_hi) break;
            const bool valid = i0 + 32 * u + lane < s_hi;
            const uint32_t r = rr[u];
            const int p = valid ? (int)((r >> 17) & 0x7Fu) : 128 + lane;
            const unsigned grp = __match_any_sync(0xffffffffu, p);
            if (valid) put(hist[warp][p] + __popc(grp & lt_mask), r);
            __syncwarp();
            if (valid && !(grp & lt_mask)) hist[warp][p] += __popc(grp);
            __syncwarp();
        }
    }
    __syncthreads();            // zero fill and the sorted notes are visible to the whole CTA; hist is dead from here on
    uint16_t* row_d = reinterpret_cast<uint16_t*>(scratch + warp * (K2_ROW_CAP * 3));      // this warp's duration / velocity row (output column c - lo)
    uint8_t* row_r = scratch + warp * (K2_ROW_CAP * 3) + K2_ROW_CAP * 2;
    if (rowbuf) {
        for (int c = lane; c < K2_ROW_CAP; c += 32) { row_d[c] = 0; row_r[c] = 0; }
        __syncwarp();
    }
    // 4. per-pitch replay, one warp per pitch, 32 notes at a time with the unresolved last on / off carried forward
    for (int pi = 0; pi < 128 / K2_WARPS; ++pi) {
        const int p = ((pi >> 3) * K2_WARPS + warp) * 8 + (pi & 7);     // blocks of 8 consecutive pitches, dealt round-robin to the warps
        if (late_fill && (pi & 7) == 0) {
            const int cnt = (int)((size_t)8 * Wo * sizeof(OutT) / 16);
            uint4* z0 = reinterpret_cast<uint4*>(oroll + (size_t)p * Wo);
            uint4* z1 = reinterpret_cast<uint4*>(odur + (size_t)p * Wo);
            for (int i = lane; i < cnt; i += 32) { z0[i] = make_uint4(0, 0, 0, 0); z1[i] = make_uint4(0, 0, 0, 0); }
            __syncwarp();
        }
        const int b0 = pstart[p], b1 = pstart[p + 1];
        if (b0 == b1) {
            if (rowbuf) {                                   // a pitch without notes: its two rows are zeros
                OutT* gr = oroll + (size_t)p * Wo;
                OutT* gd = odur + (size_t)p * Wo;
                for (int c = lane; c < Wo; c += 32) { gr[c] = to_out<OutT>(0u); gd[c] = to_out<OutT>(0u); }
            }
            continue;
        }
        OutT* rrow = oroll + (size_t)p * Wo - lo;
        OutT* drow = odur + (size_t)p * Wo - lo;
        auto put_r = [&](int c, unsigned v) { if (rowbuf) row_r[c - lo] = (uint8_t)v; else rrow[c] = to_out<OutT>(v); };
        auto put_d = [&](int c, unsigned v) { if (rowbuf) row_d[c - lo] = (uint16_t)v; else drow[c] = to_out<OutT>(v); };
        int on_carry = 0;                                   // note_on_time[p] (:33)
        int pend_on_s = -1, pend_on_v = 0;                  // last note_on seen, not yet known to be the last writer of its cell
        int pend_a = 0, pend_s = -1;                        // last note_off seen, its right neighbour still unknown
        for (int i0 = b0; i0 < b1; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < b1;
            const uint32_t r = valid ? get(i) : 0u;
            const int s = (int)(r & 0xFFFFu);
            const bool is_off = valid && ((r >> 16) & 1u), is_on = valid && !((r >> 16) & 1u);
            const unsigned onm = __ballot_sync(0xffffffffu, is_on), offm = __ballot_sync(0xffffffffu, is_off);
            const unsigned lower_on = onm & lt_mask;
            const int s_prev_on = __shfl_sync(0xffffffffu, s, lower_on ? 31 - __clz(lower_on) : 0);
            const int a = lower_on ? s_prev_on : on_carry;                       // meaningful for note_offs
            const unsigned higher_on = onm & gt_mask, higher_off = offm & gt_mask;
            const int s_next_on = __shfl_sync(0xffffffffu, s, higher_on ? __ffs(higher_on) - 1 : 0);
            const int a_next_off = __shfl_sync(0xffffffffu, a, higher_off ? __ffs(higher_off) - 1 : 0);
            const int s_first_on = __shfl_sync(0xffffffffu, s, onm ? __ffs(onm) - 1 : 0);
            const int a_first_off = __shfl_sync(0xffffffffu, a, offm ? __ffs(offm) - 1 : 0);
            // resolve what the previous chunk left pending
            if (onm && pend_on_s >= 0) {
                if (lane == 0 && pend_on_s != s_first_on && pend_on_s >= lo && pend_on_s < hi) put_r(pend_on_s, (unsigned)pend_on_v);
                pend_on_s = -1;
            }
            if (offm && pend_s >= 0) {
                int c1 = pend_s < W ? pend_s : W;
                c1 = c1 < a_first_off ? c1 : a_first_off;
                c1 = c1 < hi ? c1 : hi;
                const unsigned v = (unsigned)(pend_s - pend_a);
                for (int c = (pend_a > lo ? pend_a : lo) + lane; c < c1; c += 32) put_d(c, v);
                pend_s = -1;
            }
            // this chunk's notes
            if (is_on && higher_on && s_next_on != s && s >= lo && s < hi) put_r(s, r >> 24);
            if (is_off && higher_off) {
                int c1 = s < W ? s : W;
                c1 = c1 < a_next_off ? c1 : a_next_off;
                c1 = c1 < hi ? c1 : hi;
                const unsigned v = (unsigned)(s - a);
                int c = a > lo ? a : lo;
                if (rowbuf) {
                    // this lane's cells [c, c1) of the shared-memory row: head to 8-byte alignment, 4 cells per store, tail (the per-lane spans
                    // differ, so the loop runs as long as the longest one: 4x fewer trips)
                    int j = c - lo;
                    const int j1 = c1 - lo;
                    const uint16_t v16 = (uint16_t)v;
                    const uint32_t v32 = (uint32_t)v16 * 0x10001u;
                    if ((j & 1) && j < j1) row_d[j++] = v16;
                    if ((j & 2) && j + 2 <= j1) { *reinterpret_cast<uint32_t*>(row_d + j) = v32; j += 2; }
                    for (; j + 4 <= j1; j += 4) *reinterpret_cast<uint2*>(row_d + j) = make_uint2(v32, v32);
                    if (j + 2 <= j1) { *reinterpret_cast<uint32_t*>(row_d + j) = v32; j += 2; }
                    if (j < j1) row_d[j] = v16;
                } else {
                    for (; c < c1; ++c) put_d(c, v);
                }
            }
            if (onm) {
                const int last = 31 - __clz(onm);
                pend_on_s = __shfl_sync(0xffffffffu, s, last);
                pend_on_v = __shfl_sync(0xffffffffu, (int)(r >> 24), last);
                on_carry = pend_on_s;
            }
            if (offm) {
                const int last = 31 - __clz(offm);
                pend_a = __shfl_sync(0xffffffffu, a, last);
                pend_s = __shfl_sync(0xffffffffu, s, last);
            }
        }
        if (pend_on_s >= 0 && lane == 0 && pend_on_s >= lo && pend_on_s < hi) put_r(pend_on_s, (unsigned)pend_on_v);
        if (pend_s >= 0) {
            int c1 = pend_s < W ? pend_s : W;
            c1 = c1 < hi ? c1 : hi;
            const unsigned v = (unsigned)(pend_s - pend_a);
            for (int c = (pend_a > lo ? pend_a : lo) + lane; c < c1; c += 32) put_d(c, v);
        }
        if (rowbuf) {                                       // the finished rows leave the SM once, coalesced; the buffers go back to zero
            __syncwarp();
            OutT* gr = oroll + (size_t)p * Wo;
            OutT* gd = odur + (size_t)p * Wo;
            for (int c = lane; c < Wo; c += 32) {
                const unsigned vr = row_r[c], vd = row_d[c];
                row_r[c] = 0; row_d[c] = 0;
                gr[c] = to_out<OutT>(vr);
                gd[c] = to_out<OutT>(vd);
            }
            __syncwarp();
        }
    }
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

// fast path scratch: two 4-byte records per message (compacted notes, then sorted by pitch) + one int32 note count per song.
// A smaller (or NULL) workspace selects the generic single-kernel path, which needs none.
size_t mmg_raster_workspace_bytes(int64_t n_songs, int64_t total_events) {
    if (n_songs < 0 || total_events < 0) return 0;
    return (size_t)total_events * 8 + (((size_t)n_songs * 4 + 7) & ~(size_t)7) + 32;
}

static int g_raster_mode = 0;
// 0 (default): time steps by speculate-and-verify (parallel prefix sum, exact sequential chain only for the songs that need it);
// 1: the sequential chain for every song.  Both are bit-exact; process-wide; for A/B measurements and tests.
int mmg_raster_set_mode(int mode) { g_raster_mode = mode; return MMG_OK; }

// out: (n_songs, 2, 128, Wout) of out_dtype (0 = float32, 1 = bfloat16, 2 = uint8 saturating); fully written.
int mmg_raster_piano_roll(const double* dt, const uint32_t* meta, const int64_t* offsets, int64_t n_songs, int64_t total_events,
                          int sequence_length, int start, int end, int out_dtype, void* out, int32_t* status,
                          void* workspace, size_t ws_bytes, void* stream_) {
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
    MMG_REQUIRE(n_songs <= 0x7fffffff, MMG_EUNSUPPORTED, "raster: too many songs");
    if (S <= 65535 && workspace && ws_bytes >= mmg_raster_workspace_bytes(n_songs, total_events) && ((uintptr_t)workspace & 7) == 0) {
        const size_t rec_bytes = ((size_t)total_events * 4 + 15) & ~(size_t)15;
        uint32_t* notes = (uint32_t*)workspace;
        uint32_t* sorted = (uint32_t*)((unsigned char*)workspace + rec_bytes);
        int32_t* counts = (int32_t*)((unsigned char*)workspace + 2 * rec_bytes);
        // the last 8 bytes of the workspace: number of songs that fell back to the sequential chain (cumulative; the caller may zero it)
        unsigned long long* n_fb = (unsigned long long*)((unsigned char*)workspace + mmg_raster_workspace_bytes(n_songs, total_events) - 8);
        raster_steps_kernel<<<(int)((n_songs + K1_WARPS - 1) / K1_WARPS), K1_WARPS * 32, 0, stream>>>(dt, meta, offsets, n_songs, (int)S, (int)W, notes,
                                                                                                   counts, status, g_raster_mode, n_fb);
        MMG_LAUNCH_CHECK();
        if (out_dtype == 0)
            raster_rows_kernel<float><<<(int)n_songs, K2_THREADS, 0, stream>>>(notes, sorted, counts, offsets, (int)W, (int)lo, (int)hi, (float*)out);
        else if (out_dtype == 1)
            raster_rows_kernel<__nv_bfloat16><<<(int)n_songs, K2_THREADS, 0, stream>>>(notes, sorted, counts, offsets, (int)W, (int)lo, (int)hi, (__nv_bfloat16*)out);
        else
            raster_rows_kernel<uint8_t><<<(int)n_songs, K2_THREADS, 0, stream>>>(notes, sorted, counts, offsets, (int)W, (int)lo, (int)hi, (uint8_t*)out);
        MMG_LAUNCH_CHECK();
        return MMG_OK;
    }
    const bool short_songs = total_events <= 1024 * n_songs;           // average song length decides the chunk size
    const size_t esz = out_dtype == 0 ? 4 : out_dtype == 1 ? 2 : 1;
    const bool tile = short_songs && (size_t)2 * 128 * (hi - lo) * esz <= (size_t)SK_TILE_BYTES;
#define MMG_STREAM(T, CH, SM) raster_stream_kernel<T, SK_R, CH, SM><<<(int)n_songs, 32 * (1 + SK_R), 0, stream>>>(dt, meta, offsets, (int)S, (int)W, (int)lo, (int)hi, (T*)out, status)
#define MMG_STREAM_T(T) do { if (tile) MMG_STREAM(T, SK_CH_SHORT, true); else if (short_songs) MMG_STREAM(T, SK_CH_SHORT, false); else MMG_STREAM(T, SK_CH_LONG, false); } while (0)
    if (out_dtype == 0) MMG_STREAM_T(float);
    else if (out_dtype == 1) MMG_STREAM_T(__nv_bfloat16);
    else MMG_STREAM_T(uint8_t);
#undef MMG_STREAM_T
#undef MMG_STREAM
    MMG_LAUNCH_CHECK();
    return MMG_OK;
}

}  // extern "C"
