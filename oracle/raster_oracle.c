/* ORACLE (test infrastructure, not product code) -- piano-roll rasteriser in plain C.
 *
 * Scalar CPU restatement of /root/reference/MMGAN_MIDI_DES/datasets.py:27-54 on the
 * post-mido event stream (same contract as oracle/raster_oracle.py, which is pinned
 * against the unmodified reference through tests/golden/raster_*.npz; this file is
 * pinned against raster_oracle.py by tests/test_oracle_raster.py).
 * Used only by tests/, smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/libraster_oracle.so oracle/raster_oracle.c -lm
 * (this image's gcc has no libgomp; callers thread over song ranges from Python -- ctypes drops the GIL)
 * (no -ffast-math: the float64 running sum and rint() must stay IEEE, round-half-even)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* python slice clip of [a:b) on length n */
static inline void clip_slice(long a, long b, long n, long* lo, long* hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    *lo = a; *hi = b;
}

/* one song -> roll/dur planes (128 x Wout) float32, row-major. returns #note_on messages applied */
static long raster_one(const double* dt, const uint32_t* meta, long n, long S, long start, long end,
                       float* roll_out, float* dur_out, double* roll, double* dur) {
    const long W = end - start;
    long on_time[128];
    long notes = 0;
    memset(on_time, 0, sizeof(on_time));
    memset(roll, 0, sizeof(double) * 128 * (size_t)W);
    memset(dur, 0, sizeof(double) * 128 * (size_t)W);
    double t = 0.0;                                  /* datasets.py:32 */
    for (long i = 0; i < n; ++i) {
        t += dt[i];                                  /* :35 sequential f64 sum */
        long s = (long)rint(t);                      /* :36 round-half-even */
        if (s >= S) break;                           /* :37-38 */
        unsigned kind = meta[i] & 0xFF, p = (meta[i] >> 8) & 0xFF, v = (meta[i] >> 16) & 0xFF;
        if (kind == 1) {                             /* :39-42 */
            long c = s;
            if (c < 0) c += W;
            if (c < 0 || c >= W || p >= 128) break;  /* IndexError -> bare except (:46) */
            roll[p * W + c] = (double)v;
            on_time[p] = s;
            ++notes;
        } else if (kind == 2) {                      /* :43-45 */
            if (p >= 128) break;
            long a = on_time[p], lo, hi;
            clip_slice(a, s, W, &lo, &hi);
            for (long c = lo; c < hi; ++c) dur[p * W + c] = (double)(s - a);
        }
    }
    /* :49-54 re-slice */
    long lo, hi;
    if (end < 128) clip_slice(start, end, W, &lo, &hi); else clip_slice(0, end, W, &lo, &hi);
    long Wo = hi > lo ? hi - lo : 0;
    for (long p = 0; p < 128; ++p)
        for (long c = 0; c < Wo; ++c) {
            roll_out[p * Wo + c] = (float)roll[p * W + lo + c];
            dur_out[p * Wo + c] = (float)dur[p * W + lo + c];
        }
    return notes;
}

long mmg_oracle_out_width(long start, long end) {
    long W = end - start, lo, hi;
    if (W < 0) return -1;
    if (end < 128) clip_slice(start, end, W, &lo, &hi); else clip_slice(0, end, W, &lo, &hi);
    return hi > lo ? hi - lo : 0;
}

/* out: (n_songs, 2, 128, Wout) float32 for songs [song_lo, song_hi). returns note_on messages applied, or -1 */
long mmg_oracle_raster_batch(const double* dt, const uint32_t* meta, const int64_t* offsets, long song_lo, long song_hi,
                             long S, long start, long end, float* out) {
    const long W = end - start;
    if (W < 0) return -1;
    const long Wo = mmg_oracle_out_width(start, end);
    long total = 0;
    double* roll = (double*)malloc(sizeof(double) * 128 * (size_t)(W > 0 ? W : 1));
    double* dur = (double*)malloc(sizeof(double) * 128 * (size_t)(W > 0 ? W : 1));
    for (long s = song_lo; s < song_hi; ++s) {
        float* o = out + (size_t)s * 2 * 128 * Wo;
        total += raster_one(dt + offsets[s], meta + offsets[s], (long)(offsets[s + 1] - offsets[s]), S, start, end,
                            o, o + 128 * Wo, roll, dur);
    }
    free(roll);
    free(dur);
    return total;
}
