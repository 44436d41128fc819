// Standard MIDI File reader on the host (no device code): the input side of the rasteriser.  The reference reads its songs with
// `for msg in mido.MidiFile(path)` (MMGAN_MIDI_DES/datasets.py:18,34; mido 1.3.2 is a dependency that is not in the reference tree): all tracks
// merged by absolute tick (stable: a lower track index wins a tie), every end_of_track meta dropped and ONE re-appended at the last tick, and the
// message times converted to seconds with the RUNNING tempo -- delta_seconds = delta_ticks * (tempo * 1e-6 / ticks_per_beat), the tempo switching
// after each set_tempo message (mido.midifiles.midifiles: merge_tracks, tick2second, MidiFile.__iter__).  This file restates that published
// algorithm in C++ so that MAESTRO-scale files (10^4 .. 10^5 messages) reach the device rasteriser without a Python loop per message; the
// Python restatement it replaced lives on as the checker (oracle/smf_oracle.py) and the two agree bit for bit on the reference's 30 shipped .mid
// files and on generated files (tests/test_smf_native.py).
#include "common.cuh"
#include "../../include/mmgan_b200.h"

#include <algorithm>
#include <string.h>
#include <vector>

namespace {

struct SmfEvent {
    int64_t tick;
    uint32_t meta;       // kind | pitch << 8 | velocity << 16   (kind 0 other / 1 note_on / 2 note_off: the rasteriser's record)
    int32_t tempo;       // microseconds per beat of a set_tempo meta, else -1
    bool eot;            // end_of_track meta
};

// variable-length quantity; false when the buffer ends inside it
inline bool vlq(const unsigned char* b, size_t n, size_t& i, uint32_t& v) {
    v = 0;
    for (int k = 0; k < 5; ++k) {
        if (i >= n) return false;
        const unsigned char c = b[i++];
        v = (v << 7) | (c & 0x7Fu);
        if (!(c & 0x80u)) return true;
    }
    return false;
}

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint32_t be16(const unsigned char* p) { return ((uint32_t)p[0] << 8) | p[1]; }

// one MTrk body -> events with absolute ticks; nullptr on success, else what was wrong
const char* parse_track(const unsigned char* b, size_t n, std::vector<SmfEvent>& out) {
    size_t i = 0;
    int64_t t = 0;
    unsigned status = 0;
    while (i < n) {
        uint32_t d;
        if (!vlq(b, n, i, d)) return "truncated delta time";
        t += d;
        if (i >= n) return "track ends after a delta time";
        const unsigned c = b[i];
        if (c == 0xFF) {                                    // meta event: FF type len data
            if (i + 1 >= n) return "truncated meta event";
            const unsigned typ = b[i + 1];
            size_t j = i + 2;
            uint32_t ln;
            if (!vlq(b, n, j, ln) || j + ln > n) return "truncated meta event";
            int32_t tempo = -1;
            if (typ == 0x51 && ln == 3) tempo = (int32_t)(((uint32_t)b[j] << 16) | ((uint32_t)b[j + 1] << 8) | b[j + 2]);
            out.push_back({t, 0u, tempo, typ == 0x2F});
            i = j + ln;
        } else if (c == 0xF0 || c == 0xF7) {                // sysex
            size_t j = i + 1;
            uint32_t ln;
            if (!vlq(b, n, j, ln) || j + ln > n) return "truncated sysex event";
            out.push_back({t, 0u, -1, false});
            i = j + ln;
        } else {
            if (c & 0x80u) { status = c; ++i; }             // else: running status
            const unsigned hi = status & 0xF0u;
            int nbytes = (hi == 0xC0u || hi == 0xD0u) ? 1 : 2;
            if (status >= 0xF0u) nbytes = status == 0xF1u ? 1 : status == 0xF2u ? 2 : status == 0xF3u ? 1 : 0;      // system common / realtime
            if (i + (size_t)nbytes > n) return "truncated channel message";
            const unsigned d1 = nbytes >= 1 ? b[i] : 0u, d2 = nbytes >= 2 ? b[i + 1] : 0u;
            i += (size_t)nbytes;
            const uint32_t kind = hi == 0x90u ? 1u : hi == 0x80u ? 2u : 0u;          // a note_on with velocity 0 stays a note_on (mido does not convert it)
            out.push_back({t, kind ? (kind | (d1 << 8) | (d2 << 16)) : 0u, -1, false});
        }
    }
    return nullptr;
}

}  // namespace

extern "C" {

int64_t mmg_smf_max_messages(size_t len) { return (int64_t)(len / 2 + 2); }     // a message takes at least 2 bytes; + the re-appended end_of_track

int mmg_smf_parse(const unsigned char* data, size_t len, double* dt, uint32_t* meta, int64_t* abs_tick, int64_t capacity, int64_t* n_messages,
                  int* ticks_per_beat, int64_t* tempo_tick, int32_t* tempo_us, int64_t tempo_capacity, int64_t* n_tempo) {
    MMG_REQUIRE(data && dt && meta && n_messages && ticks_per_beat && capacity > 0, MMG_EINVAL, "smf_parse: bad arguments");
    MMG_REQUIRE(len >= 14 && memcmp(data, "MThd", 4) == 0, MMG_EINVAL, "not a Standard MIDI File");
    const uint32_t hlen = be32(data + 4), fmt = be16(data + 8), ntrk = be16(data + 10), div = be16(data + 12);
    MMG_REQUIRE(!(div & 0x8000u), MMG_EUNSUPPORTED, "SMPTE time division is not supported");
    MMG_REQUIRE(div > 0, MMG_EINVAL, "ticks per beat is zero");
    MMG_REQUIRE(fmt != 2, MMG_EUNSUPPORTED, "can't merge tracks in type 2 (asynchronous) file");
    std::vector<SmfEvent> ev;
    ev.reserve(len / 3);
    size_t pos = 8 + (size_t)hlen;
    for (uint32_t tr = 0; tr < ntrk; ++tr) {
        for (;;) {                                          // chunks that are not tracks are skipped
            MMG_REQUIRE(pos + 8 <= len, MMG_EINVAL, "track %u of %u is missing", tr + 1, ntrk);
            if (memcmp(data + pos, "MTrk", 4) == 0) break;
            pos += 8 + (size_t)be32(data + pos + 4);
        }
        const size_t ln = be32(data + pos + 4);
        MMG_REQUIRE(pos + 8 + ln <= len, MMG_EINVAL, "track %u is truncated", tr + 1);
        const char* err = parse_track(data + pos + 8, ln, ev);
        MMG_REQUIRE(err == nullptr, MMG_EINVAL, "track %u: %s", tr + 1, err);
        pos += 8 + ln;
    }
    std::stable_sort(ev.begin(), ev.end(), [](const SmfEvent& a, const SmfEvent& b) { return a.tick < b.tick; });      // mido.merge_tracks
    int64_t end_tick = 0;
    for (const SmfEvent& e : ev) end_tick = e.tick > end_tick ? e.tick : end_tick;
    int64_t n = 0, nt = 0, prev = 0;
    double tempo = 500000.0;
    auto emit = [&](int64_t tick, uint32_t m, int32_t new_tempo) -> bool {
        if (n >= capacity) return false;
        const int64_t dticks = tick - prev;
        // mido.tick2second with the tempo in force BEFORE this message; the same three IEEE operations as the Python expression
        // int(dticks) * (tempo * 1e-6 / div)
        dt[n] = dticks > 0 ? (double)dticks * ((tempo * 1e-6) / (double)div) : 0.0;
        meta[n] = m;
        if (abs_tick) abs_tick[n] = tick;
        ++n;
        prev = tick;
        if (new_tempo >= 0) {
            tempo = (double)new_tempo;
            if (tempo_tick && tempo_us && nt < tempo_capacity) { tempo_tick[nt] = tick; tempo_us[nt] = new_tempo; }
            ++nt;
        }
        return true;
    };
    for (const SmfEvent& e : ev)
        if (!e.eot) MMG_REQUIRE(emit(e.tick, e.meta, e.tempo), MMG_EWORKSPACE, "smf_parse: more than %lld messages (see mmg_smf_max_messages)", (long long)capacity);
    MMG_REQUIRE(emit(end_tick, 0u, -1), MMG_EWORKSPACE, "smf_parse: more than %lld messages (see mmg_smf_max_messages)", (long long)capacity);
    *n_messages = n;
    *ticks_per_beat = (int)div;
    if (n_tempo) *n_tempo = nt;
    MMG_REQUIRE(!(tempo_tick && tempo_us) || nt <= tempo_capacity, MMG_EWORKSPACE, "smf_parse: %lld tempo changes, room for %lld", (long long)nt, (long long)tempo_capacity);
    return MMG_OK;
}

// Quarter-note beat grid along the tempo map, from tick 0 up to last_tick (the host-side stand-in for pretty_midi.get_beats that datasets.py:57
// pads / truncates to beats_length; parity with pretty_midi is unpinned): beats[j] = seconds at tick j * ticks_per_beat, integrated
// piecewise over the set_tempo messages (tempo_tick / tempo_us in stream order; 500000 us per beat before the first).  n = last_tick /
// ticks_per_beat + 1 values.
int mmg_smf_beat_grid(const int64_t* tempo_tick, const int32_t* tempo_us, int64_t n_tempo, int ticks_per_beat, int64_t last_tick, double* beats,
                      int64_t capacity, int64_t* n_beats) {
    MMG_REQUIRE(ticks_per_beat > 0 && last_tick >= 0 && n_beats && (n_tempo == 0 || (tempo_tick && tempo_us)), MMG_EINVAL, "smf_beat_grid: bad arguments");
    const int64_t need = last_tick / ticks_per_beat + 1;
    *n_beats = need;
    MMG_REQUIRE(beats && capacity >= need, MMG_EWORKSPACE, "smf_beat_grid: %lld beats, room for %lld", (long long)need, (long long)capacity);
    const double div = (double)ticks_per_beat;
    double t_sec = 0.0, tempo = 500000.0;
    int64_t tick = 0, k = 0, n = 0;                          // k = tempo changes consumed
    while (tick <= last_tick) {
        beats[n++] = t_sec;
        const int64_t nxt = tick + ticks_per_beat;
        while (k < n_tempo && tempo_tick[k] < nxt) {         // integrate across the tempo changes inside this beat
            const int64_t c_tick = tempo_tick[k] > tick ? tempo_tick[k] : tick;
            const double seg = (double)(c_tick - tick) * ((tempo * 1e-6) / div);
            t_sec = t_sec + seg;
            tick = c_tick;
            tempo = (double)tempo_us[k];
            ++k;
        }
        const double seg = (double)(nxt - tick) * ((tempo * 1e-6) / div);
        t_sec = t_sec + seg;
        tick = nxt;
    }
    return MMG_OK;
}

}  // extern "C"
