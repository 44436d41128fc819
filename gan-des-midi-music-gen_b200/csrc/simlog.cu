// Simulator log -> note-event stream on the host (no device code): the caller of the rasteriser inside the training loop.
// For every generated sample the reference writes the DES log to ./logs/simulation.log, turns it into a mido track with MidiGenerator
// (MMGAN_MIDI_DES/sim_log_to_midi.py:13-218), saves a .mid, re-reads it through mido and rasterises in Python (:238-277 -> datasets.py:13-70) --
// twice per sample and iteration.  This file is that conversion for a whole BATCH of songs, natively and on all host cores: log text in, the
// post-mido message streams (delta seconds as float64 + packed kind / pitch / velocity records) out, ready for one H2D copy and one
// mmg_raster_piano_roll launch.  Every quirk that shapes the output is kept (absolute times written as deltas, the while-iterating removal
// in save_midi, the index-based clean-up, the count-%-100 save rule, Python's floor modulo, float32 products of a float32 gen2 row);
// tests/test_simlog.py and tests/test_live_reference.py hold it to the vectors and live runs of the unmodified reference, bit for bit.
#include "common.cuh"
#include "../../include/mmgan_b200.h"

#include <atomic>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

enum MsgType : int { SET_TEMPO = 0, TIME_SIG, KEY_SIG, PROGRAM, NOTE_ON, NOTE_OFF, END_OF_TRACK };
struct Msg {
    int type;
    long long time, a, b;
    bool operator==(const Msg& o) const { return type == o.type && time == o.time && a == o.a && b == o.b; }
};

struct PyError {};                                           // any exception inside the reference's try block

struct Tok {                                                 // a matched group of the log line (points into the log text)
    const char* p;
    int n;
    std::string str() const { return std::string(p, (size_t)n); }
};

inline long long pymod(long long x, long long m) {           // Python's %: the result has the sign of the divisor; ZeroDivisionError
    if (m == 0) throw PyError();
    long long r = x % m;
    if (r != 0 && ((r < 0) != (m < 0))) r += m;
    return r;
}

// int(g * k) with g a float32 (numpy scalar: the product stays float32) or a float64 value; truncation towards zero; non-finite -> exception
template <typename T>
inline long long trunc_mul(T g, long long k) {
    const T p = g * (T)k;
    if (!(p == p) || p > (T)9.0e18 || p < (T)-9.0e18) throw PyError();
    return (long long)p;
}

struct Song {
    const long long* instruments; int n_instr;
    const long long* note_levels; int n_notes;
    long long skip[3], base, tempo, var;
    std::vector<Msg> track;
    bool saved = false;
    long long previous_time = 0, current_instrument = 0;
    // the reference keys its dictionaries by the server id AS WRITTEN ('3', '03' and '3.0' are three keys).  Canonical ids below 64 -- every id a
    // real log holds -- live in arrays; any other spelling goes to the maps
    struct Future { long long time, velocity, service; bool set; };
    static constexpr int FAST = 64;
    long long ql_fast[FAST];
    bool ql_set[FAST];
    Future fut_fast[FAST];
    std::unordered_map<std::string, long long> queue_lengths;
    std::unordered_map<std::string, Future> future_events;

    template <typename T>
    void init(const T* g) {                                  // MidiGenerator.__init__, sim_log_to_midi.py:21-46
        for (int i = 0; i < 3; ++i) { const long long s = trunc_mul<T>(g[i], 10); skip[i] = s > 2 ? s : 2; }
        base = trunc_mul<T>(g[3], 90);
        if (base < 50) base = 80;
        tempo = trunc_mul<T>(g[4], 1000000);
        if (tempo > 16777215) tempo = 16777215;
        if (tempo == 0) tempo = 500000;
        var = trunc_mul<T>(g[5], 63);
        if (var == 0) var = 30;
        (void)pymod(trunc_mul<T>(g[5], 11), 11);             // the key signature's index (no effect on the stream)
        track.reserve(512);
        track = {{SET_TEMPO, 0, tempo, 0}, {TIME_SIG, 0, 0, 0}, {KEY_SIG, 0, 0, 0}, {PROGRAM, 0, 0, 0}};      // generate_midi, :72-96
        for (int i = 0; i < FAST; ++i) { ql_fast[i] = 0; ql_set[i] = false; fut_fast[i] = {0, 0, 0, false}; }
    }

    // index of a server id as the reference's dictionaries know it: keys are str(i), so only the canonical decimal spelling matches
    static long long canonical(const Tok& s) {               // value of a canonical decimal spelling (no dot, no leading zero), else -1
        if (s.n <= 0 || s.n > 9 || (s.n > 1 && s.p[0] == '0')) return -1;
        long long v = 0;
        for (int i = 0; i < s.n; ++i) {
            if (s.p[i] < '0' || s.p[i] > '9') return -1;
            v = v * 10 + (s.p[i] - '0');
        }
        return v;
    }

    void process_line(const Tok& a1, const Tok& a2, const Tok& a3, bool arrival) {      // :99-180
        // int(float(array1)); the regex admits digits and one dot only.  Up to 15 characters the nearest double cannot reach the next integer, so
        // the truncation is the integer part as written; longer spellings go through strtod (correctly rounded, like float())
        double tf;
        if (a1.n <= 15) {
            long long ip = 0;
            for (int i = 0; i < a1.n && a1.p[i] != '.'; ++i) ip = ip * 10 + (a1.p[i] - '0');
            tf = (double)ip;
        } else tf = strtod(a1.str().c_str(), nullptr);
        if (!(tf < 200.0) || track.size() >= 500) return;    // max(0, .) of a non-negative number; `midi_time < 200 and len(track) < 500`
        long long midi_time = (long long)tf;
        if (previous_time > midi_time) midi_time = previous_time;
        long long cust = 0;                                  // int(array2): int('1.5') raises; (19+ digits: outside this port)
        if (a2.n > 18) throw PyError();
        for (int i = 0; i < a2.n; ++i) {
            if (a2.p[i] == '.') throw PyError();
            cust = cust * 10 + (a2.p[i] - '0');
        }
        const bool hit = pymod(cust, skip[0]) == 0 || pymod(cust, skip[1]) == 0 || pymod(cust, skip[2]) == 0;
        if (!hit) return;
        const long long id = canonical(a3);
        const bool fast = id >= 0 && id < FAST;
        if (arrival) {
            long long ql;
            if (fast) { ql = ++ql_fast[id]; ql_set[id] = true; }
            else ql = ++queue_lengths[a3.str()];
            if (ql >= 127 && ql < 254) { ql = 254 - ql; ql = ql < 0 ? 0 : ql; ql = ql > 127 ? 127 : ql; }
            else if (ql >= 254) { ql = pymod(ql, 127); ql = ql > 127 ? 127 : ql; }
            const long long max_c = base + var;
            long long cid = base - var + cust;
            if (cid > max_c) cid = max_c - pymod(cid, max_c);
            const Future ev = {midi_time, pymod(cid, 126), ql, true};
            if (fast) fut_fast[id] = ev;
            else future_events[a3.str()] = ev;
            const long long on_time = previous_time > ev.time ? previous_time : ev.time;
            previous_time = on_time;
            if (id < 0 || id >= n_instr) throw PyError();     // KeyError: self.instruments[array3]
            if (current_instrument != instruments[id]) {
                current_instrument = instruments[id];
                track.push_back({PROGRAM, on_time, instruments[id], 0});
            }
            if (id >= n_notes) throw PyError();               // KeyError: self.note_offsets[array3]
            track.push_back({NOTE_ON, on_time, note_levels[id], ev.velocity});
        } else {
            const Future* ev = nullptr;
            if (fast) { if (fut_fast[id].set) ev = &fut_fast[id]; }
            else {
                auto it = future_events.find(a3.str());
                if (it != future_events.end()) ev = &it->second;
            }
            if (ev) {
                const long long t = midi_time + (ev->service > 0 ? ev->service : 0);    // ev.time + (midi_time - ev.time) + max(0, service_time)
                const long long off_time = previous_time > t ? previous_time : t;
                previous_time = off_time;
                if (id < 0 || id >= n_instr) throw PyError();
                if (current_instrument != instruments[id]) {
                    current_instrument = instruments[id];
                    track.push_back({PROGRAM, off_time, instruments[id], 0});
                }
                if (id >= n_notes) throw PyError();
                track.push_back({NOTE_OFF, off_time, note_levels[id], ev->velocity});
            }
            if (fast) {
                if (ql_set[id]) ql_fast[id] -= 1;
                else { ql_fast[id] = 0; ql_set[id] = true; }
            } else {
                auto q = queue_lengths.find(a3.str());
                if (q != queue_lengths.end()) q->second -= 1;
                else queue_lengths[a3.str()] = 0;
            }
        }
    }

    void save() {                                            // save_midi + clean_midi_file, :182-218
        for (size_t i = 0; i < track.size(); ++i) {          // `for msg in track: if msg.time > 200: track.remove(msg)`: the list shrinks under the
            const Msg m = track[i];                          // iterator, so the element after a removed one is never looked at; remove() takes
            if (m.time > 200) {                              // the FIRST message equal by value
                for (size_t j = 0; j < track.size(); ++j)
                    if (track[j] == m) { track.erase(track.begin() + (long)j); break; }
            }
        }
        track.push_back({END_OF_TRACK, 0, 0, 0});
        std::unordered_map<long long, long long> on_time;
        std::vector<char> drop(track.size(), 0);
        for (size_t j = 0; j < track.size(); ++j) {
            const Msg& m = track[j];
            if (m.type == NOTE_ON) {
                auto it = on_time.find(m.a);
                if (it != on_time.end() && it->second > 0) drop[j] = 1;
                else on_time[m.a] = m.time;
            } else if (m.type == NOTE_OFF) {
                auto it = on_time.find(m.a);
                if (it == on_time.end() || it->second == 0) drop[j] = 1;
                else it->second = 0;
            }
            if (m.time > 200) drop[j] = 1;
        }
        size_t w = 0;
        for (size_t j = 0; j < track.size(); ++j)
            if (!drop[j]) track[w++] = track[j];
        track.resize(w);
        saved = true;
    }

    // what `for msg in mid` yields (mido merge_tracks + fix_end_of_track + tick2second, 480 ticks per beat): <= track.size() + 1 messages
    long long emit(double* dt, uint32_t* meta, long long capacity) const {
        long long n = 0;
        double tempo_now = 500000.0;
        long long now = 0, prev = 0, accum = 0;
        auto put = [&](long long ticks, long long m) {
            if (n >= capacity || m < 0 || m > 0xFFFFFFFFll) throw PyError();
            dt[n] = ticks > 0 ? (double)ticks * ((tempo_now * 1e-6) / 480.0) : 0.0;
            meta[n] = (uint32_t)m;
            ++n;
        };
        if (saved) {
            for (const Msg& m : track) {                     // message times are absolute simulation times but are played back as deltas (:149-170);
                now += m.time;                               // they are never negative, so the merged order is the track order
                const long long delta = now - prev;
                prev = now;
                if (m.type == END_OF_TRACK) { accum += delta; continue; }
                const long long ticks = delta + accum;
                accum = 0;
                const long long kind = m.type == NOTE_ON ? 1 : m.type == NOTE_OFF ? 2 : 0;
                put(ticks, kind ? (kind | (m.a << 8) | (m.b << 16)) : 0);
                if (m.type == SET_TEMPO) tempo_now = (double)m.a;
            }
        }
        put(accum, 0);                                       // the one end_of_track mido appends
        return n;
    }
};

// `INFO:root:<num> - <num> - <num> - (arrival|departure)` matched at the start of the line like re.match; <num> = [0-9]*\.[0-9]+|[0-9]+ .
// What follows a number is the literal " - ", so the alternation has one way to succeed: the longest digits[.digits] token.
inline bool take_number(const char*& p, const char* end, Tok& out) {
    const char* s = p;
    while (p < end && *p >= '0' && *p <= '9') ++p;
    if (p + 1 < end && *p == '.' && p[1] >= '0' && p[1] <= '9') {
        ++p;
        while (p < end && *p >= '0' && *p <= '9') ++p;
    } else if (p == s) return false;
    out = {s, (int)(p - s)};
    return true;
}
inline bool take_lit(const char*& p, const char* end, const char* lit) {
    const size_t n = strlen(lit);
    if ((size_t)(end - p) < n || memcmp(p, lit, n) != 0) return false;
    p += n;
    return true;
}

template <typename T>
long long convert_song(const char* log, size_t len, const long long* instruments, int n_instr, const long long* note_levels, int n_notes, const T* g,
                       int generate, double* dt, uint32_t* meta, long long capacity) {
    Song s;
    s.instruments = instruments; s.n_instr = n_instr; s.note_levels = note_levels; s.n_notes = n_notes;
    s.init<T>(g);
    long long count = 0;
    const long long cap = 5000;
    const char* p = log;
    const char* end = log + len;
    Tok a1, a2, a3;
    while (p < end) {                                        // process_adjsim_log, :254-266
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = eol ? eol : end;
        if (++count > cap) break;
        const char* q = p;
        bool arrival = false, ok = take_lit(q, le, "INFO:root:") && take_number(q, le, a1) && take_lit(q, le, " - ") && take_number(q, le, a2) &&
                                   take_lit(q, le, " - ") && take_number(q, le, a3) && take_lit(q, le, " - ");
        if (ok) {
            if (take_lit(q, le, "arrival")) arrival = true;
            else ok = take_lit(q, le, "departure");
        }
        if (ok) s.process_line(a1, a2, a3, arrival);
        p = eol ? eol + 1 : end;
    }
    if ((count % 100 == 0 && !generate) || generate) s.save();       // :268-272: otherwise the track never reaches the file
    return s.emit(dt, meta, capacity);
}

long long convert_any(const char* log, size_t len, const long long* instruments, int n_instr, const long long* note_levels, int n_notes, const void* g,
                      int g_is_f32, int generate, double* dt, uint32_t* meta, long long capacity) {
    try {
        return g_is_f32 ? convert_song<float>(log, len, instruments, n_instr, note_levels, n_notes, (const float*)g, generate, dt, meta, capacity)
                        : convert_song<double>(log, len, instruments, n_instr, note_levels, n_notes, (const double*)g, generate, dt, meta, capacity);
    } catch (const PyError&) {
        return -1;
    }
}

}  // namespace

extern "C" {

int mmg_simlog_max_messages(void) { return 512; }            // the track is capped at 500 messages (+ one program change / note pair, + end_of_track)

int mmg_simlog_to_events(const char* log, size_t log_len, const int64_t* instruments, int n_instruments, const int64_t* note_levels, int n_note_levels,
                         const void* gen2, int gen2_len, int gen2_is_f32, int generate, double* dt, uint32_t* meta, int64_t capacity, int64_t* n_messages) {
    MMG_REQUIRE((log || log_len == 0) && instruments && note_levels && gen2 && dt && meta && n_messages && n_instruments >= 0 && n_note_levels >= 0,
                MMG_EINVAL, "simlog_to_events: bad arguments");
    MMG_REQUIRE(gen2_len >= 6, MMG_EINVAL, "index 5 is out of bounds for axis 0 with size %d", gen2_len);
    MMG_REQUIRE(capacity >= mmg_simlog_max_messages(), MMG_EWORKSPACE, "simlog_to_events: room for %d messages is required", mmg_simlog_max_messages());
    const long long n = convert_any(log, log_len, (const long long*)instruments, n_instruments, (const long long*)note_levels, n_note_levels, gen2, gen2_is_f32,
                                    generate, dt, meta, capacity);
    MMG_REQUIRE(n >= 0, MMG_EINVAL, "Error in processing log file");
    *n_messages = n;
    return MMG_OK;
}

// A batch of songs on n_threads host threads (0 = all cores).  Song s: log text logs[log_offsets[s] .. log_offsets[s+1]), row s of instruments
// (n_songs x n_instruments), note_levels (n_songs x n_note_levels) and gen2 (n_songs x gen2_len).  dt / meta need n_songs x 512 entries; the streams
// come back packed at the front, song s at [offsets[s], offsets[s+1]).  A song the reference would refuse ("Error in processing log file") fails the call.
int mmg_simlog_batch_to_events(const char* logs, const int64_t* log_offsets, int64_t n_songs, const int64_t* instruments, int n_instruments,
                               const int64_t* note_levels, int n_note_levels, const void* gen2, int gen2_len, int gen2_is_f32, int generate, double* dt,
                               uint32_t* meta, int64_t* offsets, int n_threads) {
    MMG_REQUIRE(n_songs >= 0 && offsets && (n_songs == 0 || (logs && log_offsets && instruments && note_levels && gen2 && dt && meta)), MMG_EINVAL,
                "simlog_batch_to_events: bad arguments");
    MMG_REQUIRE(gen2_len >= 6, MMG_EINVAL, "index 5 is out of bounds for axis 0 with size %d", gen2_len);
    const int slot = mmg_simlog_max_messages();
    std::vector<long long> counts((size_t)n_songs, 0);
    std::atomic<long long> next(0), failed(-1);
    auto work = [&]() {
        for (;;) {
            const long long s = next.fetch_add(1);
            if (s >= n_songs) return;
            const size_t gstride = (size_t)gen2_len * (gen2_is_f32 ? 4 : 8);
            counts[(size_t)s] = convert_any(logs + log_offsets[s], (size_t)(log_offsets[s + 1] - log_offsets[s]), (const long long*)instruments + s * n_instruments,
                                            n_instruments, (const long long*)note_levels + s * n_note_levels, n_note_levels, (const char*)gen2 + (size_t)s * gstride,
                                            gen2_is_f32, generate, dt + s * slot, meta + s * slot, slot);
            if (counts[(size_t)s] < 0) { long long none = -1; failed.compare_exchange_strong(none, s); }
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((long long)nt > n_songs) nt = (int)(n_songs > 0 ? n_songs : 1);
    std::vector<std::thread> pool;
    for (int i = 1; i < nt; ++i) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    MMG_REQUIRE(failed.load() < 0, MMG_EINVAL, "Error in processing log file (song %lld)", failed.load());
    int64_t o = 0;
    for (int64_t s = 0; s < n_songs; ++s) {                  // compaction to the front (the destination never overtakes the source)
        offsets[s] = o;
        if (o != s * slot) {
            memmove(dt + o, dt + s * slot, (size_t)counts[(size_t)s] * sizeof(double));
            memmove(meta + o, meta + s * slot, (size_t)counts[(size_t)s] * sizeof(uint32_t));
        }
        o += counts[(size_t)s];
    }
    offsets[n_songs] = o;
    return MMG_OK;
}

}  // extern "C"
