"""Drop-in mirror of the reference's ``MMGAN_MIDI_DES/sim_log_to_midi.py``
(/root/reference/MMGAN_MIDI_DES/sim_log_to_midi.py:13-277): the step between the host DES and the rasteriser.

The reference turns the simulator's log lines into a mido track, saves it, and hands the ``mido.MidiFile`` to
``generate_piano_roll`` which plays it back through mido (ticks -> seconds) and rasterises in a CPython loop.  Here the same
``MidiGenerator`` state machine produces the POST-MIDO MESSAGE STREAM directly -- an :class:`~.datasets.EventStream` (delta seconds as
float64, packed kind / pitch / velocity) -- with no mido objects and no file round trip, and the rasterisation runs on the device
(``mmg_raster_piano_roll``).  ``process_adjsim_log`` keeps the reference's signature (it still reads ``./logs/simulation.log`` by
default) and accepts the log lines in memory as well; ``sim_logs_to_event_batch`` packs a whole batch of simulated songs into the pinned
``(dt, meta, offsets)`` triple that ``trainer.HostBatchPipeline`` copies and rasterises on the GPU.

Every quirk that shapes the output is kept (file:line refer to the reference):
  * message ``time`` fields hold ABSOLUTE simulation times but are written -- and played back -- as delta ticks (:149-154,163-170);
  * ``save_midi`` removes late messages while iterating over the list (every removal skips the next element, :184-186) and
    ``clean_midi_file`` drops re-struck notes / orphan note_offs by LIST INDEX (:201-218);
  * ``process_adjsim_log`` only attaches the track to the file when the number of log lines is a multiple of 100 or ``generate`` is set
    (:268-272) -- otherwise the roll is empty;
  * ``instruments=None`` / ``note_levels=None`` fail with TypeError because the ``range`` parameter shadows the builtin (:14,52,62).
The mido-side arithmetic restated here (track merge, ``tick * (tempo * 1e-6 / ticks_per_beat)``, tempo switching after the set_tempo
message, 480 ticks per beat) is mido 1.3.2's published behaviour; parity at that boundary is unpinned.
"""
import re
import struct

import numpy as np

from . import datasets as ds

KEYS = ['C', 'C#', 'D', 'E', 'F', 'F#', 'G', 'G#m', 'A', 'A#m', 'B']
TICKS_PER_BEAT = 480              # mido.MidiFile() default
DEFAULT_TEMPO = 500000

# message = [type, time, a, b]: note_on / note_off: a = note, b = velocity; program_change: a = program; set_tempo: a = tempo
_T, _TIME, _A, _B = 0, 1, 2, 3


class MidiGenerator:
    """sim_log_to_midi.py:13-218 on plain lists instead of mido objects."""

    def __init__(self, n, baseline=80, range=30, instruments=None, note_levels=None, gen2_output=None):
        self.n = n
        self.baseline = baseline
        self.range = range
        self.track = []
        self.tracks = []                                   # mid.tracks (:19,195)
        self.gen2_output = gen2_output
        g = gen2_output
        self.skip_1 = max(2, int(g[0] * 10))               # :23-31 (the == 0 branches are dead: max(2, .) >= 2)
        self.skip_2 = max(2, int(g[1] * 10))
        self.skip_3 = max(2, int(g[2] * 10))
        self.base = int(g[3] * 90)                         # :32-34
        if self.base < 50:
            self.base = 80
        self.tempo = min(int(g[4] * 1000000), 16777215)    # :35-37
        if self.tempo == 0:
            self.tempo = 500000
        self.var = int(g[5] * int(126 / 2))                # :39-41
        if self.var == 0:
            self.var = 30
        self.key_signature = KEYS[int(g[5] * 11) % 11]     # :44-46
        if note_levels is None or instruments is None:     # :52,62: `range` is the int parameter there
            raise TypeError("'int' object is not callable")
        self.note_offsets = {str(i): int(v) for i, v in enumerate(note_levels)}     # :48-50
        self.queue_lengths = {}
        self.instruments = {str(i): int(v) for i, v in enumerate(instruments)}      # :57-60
        self.future_events = {}
        self.generate_midi()
        self.previous_time = 0
        self.current_instrument = 0

    def generate_midi(self):                               # :72-96
        self.track.append(["set_tempo", 0, self.tempo, 0])
        self.track.append(["time_signature", 0, 0, 0])
        self.track.append(["key_signature", 0, 0, 0])
        self.track.append(["program_change", 0, 0, 0])

    def process_line(self, processed_line):                # :99-180
        array1, array2, array3, array4 = processed_line
        midi_time = max(0, int(float(array1)))
        if not (midi_time < 200 and len(self.track) < 500):
            return
        if self.previous_time > midi_time:
            midi_time = self.previous_time
        cust = int(array2)
        hit = cust % self.skip_1 == 0 or cust % self.skip_2 == 0 or cust % self.skip_3 == 0
        if array4 == 'arrival' and hit:
            self.queue_lengths[array3] = self.queue_lengths.get(array3, 0) + 1
            queue_length = self.queue_lengths[array3]
            if 127 <= queue_length < 2 * 127:
                queue_length = min(127, max(0, 2 * 127 - queue_length))
            elif queue_length >= 2 * 127:
                queue_length = min(127, max(0, queue_length % 127))
            max_customer_id = self.base + self.var
            customer_id = self.base - self.var + cust
            if customer_id > max_customer_id:
                customer_id = max_customer_id - (customer_id % max_customer_id)
            ev = self.future_events[array3] = {'time': int(midi_time), 'velocity': int(customer_id) % 126, 'service_time': int(queue_length)}
            on_time = int(max(self.previous_time, ev['time']))
            self.previous_time = on_time
            if self.current_instrument != self.instruments[array3]:
                self.current_instrument = self.instruments[array3]
                self.track.append(["program_change", on_time, self.instruments[array3], 0])
            self.track.append(["note_on", on_time, int(self.note_offsets[array3]), ev['velocity']])
        elif array4 == 'departure' and hit:
            if array3 in self.future_events:
                ev = self.future_events[array3]
                off_time = int(max(self.previous_time, int(ev['time'] + (midi_time - ev['time']) + max(0, ev['service_time']))))
                self.previous_time = off_time
                if self.current_instrument != self.instruments[array3]:
                    self.current_instrument = self.instruments[array3]
                    self.track.append(["program_change", off_time, self.instruments[array3], 0])
                self.track.append(["note_off", off_time, int(self.note_offsets[array3]), ev['velocity']])
            if array3 in self.queue_lengths:
                self.queue_lengths[array3] -= 1
            else:
                self.queue_lengths[array3] = 0
        elif array4 == 'processing' and hit:
            self.future_events[array3]['service_time'] += midi_time

    def save_midi(self, filename=None):                    # :182-199
        i = 0
        while i < len(self.track):                         # `for msg in track: if msg.time > 200: track.remove(msg)`: a removal shifts the
            msg = self.track[i]                            # list under the iterator, so the element after a removed one is never looked at
            if msg[_TIME] > 200:
                self.track.remove(msg)                     # first equal message, like list.remove on mido messages (equality by value)
            i += 1
        self.track.append(["end_of_track", 0, 0, 0])
        self.clean_midi_file()
        self.tracks.append(self.track)
        if filename is not None:
            write_smf(filename, self.tracks, TICKS_PER_BEAT)

    def clean_midi_file(self):                             # :201-218
        note_on_times, drop = {}, []
        for j, msg in enumerate(self.track):
            if msg[_T] == 'note_on':
                if note_on_times.get(msg[_A], 0) > 0:
                    drop.append(j)
                else:
                    note_on_times[msg[_A]] = msg[_TIME]
            elif msg[_T] == 'note_off':
                if note_on_times.get(msg[_A], 0) == 0:
                    drop.append(j)
                else:
                    note_on_times[msg[_A]] = 0
            if msg[_TIME] > 200 and j not in drop:
                drop.append(j)
        for j in sorted(drop, reverse=True):
            self.track.pop(j)

    # ---- what `for msg in self.mid` yields (mido.MidiFile.__iter__), as an EventStream
    def event_stream(self):
        msgs = []
        for tr in self.tracks:                             # merge_tracks: absolute ticks, stable sort
            now = 0
            for m in tr:
                now += m[_TIME]
                msgs.append((now, m))
        msgs.sort(key=lambda x: x[0])
        dt, meta = [], []
        tempo, now, accum = DEFAULT_TEMPO, 0, 0
        for t_abs, m in msgs:
            delta = t_abs - now
            now = t_abs
            if m[_T] == "end_of_track":                    # fix_end_of_track: removed, its delta goes to the next message
                accum += delta
                continue
            ticks = delta + accum
            accum = 0
            dt.append(ticks * (tempo * 1e-6 / TICKS_PER_BEAT) if ticks > 0 else 0.0)
            kind = ds.KIND_ON if m[_T] == "note_on" else ds.KIND_OFF if m[_T] == "note_off" else ds.KIND_OTHER
            meta.append(kind | (m[_A] << 8) | (m[_B] << 16) if kind else 0)
            if m[_T] == "set_tempo":
                tempo = m[_A]
        dt.append(accum * (tempo * 1e-6 / TICKS_PER_BEAT) if accum > 0 else 0.0)     # the one end_of_track mido appends
        meta.append(0)
        return ds.EventStream(np.array(dt, dtype=np.float64), np.array(meta, dtype=np.uint32))


class LogLineProcessor:                                    # :225-234
    def __init__(self, regex_format):
        self.regex_format = re.compile(regex_format)

    def process_line(self, line):
        match = self.regex_format.match(line)
        if match:
            return match.group(1), match.group(2), match.group(3), match.group(4)
        return None


LOG_REGEX = r"INFO:root:([0-9]*\.[0-9]+|[0-9]+) - ([0-9]*\.[0-9]+|[0-9]+) - ([0-9]*\.[0-9]+|[0-9]+) - (arrival|departure)"


def sim_log_to_event_stream(log_lines, instruments, note_levels, gen2_output, generate=False, midi_path=None):
    """The host half of ``process_adjsim_log`` (:238-275): log lines -> MidiGenerator -> the message stream ``generate_piano_roll`` would
    iterate over.  Returns (EventStream, MidiGenerator)."""
    log_processor = LogLineProcessor(LOG_REGEX)
    count, cap = 0, 5000
    gen = MidiGenerator(n=cap, baseline=70, range=50, instruments=instruments, note_levels=note_levels, gen2_output=gen2_output)
    try:
        for line in log_lines:
            count += 1
            if count > cap:
                break
            processed = log_processor.process_line(line)
            if processed:
                gen.process_line(processed)
    except Exception:
        raise ValueError("Error in processing log file")
    if (count % 100 == 0 and not generate) or generate:    # :268-272: otherwise the track never reaches the file
        gen.save_midi(midi_path)
    return gen.event_stream(), gen


def process_adjsim_log(n=5000, baseline=70, range=50, instruments=np.arange(0, 16), note_levels=None, gen2_output=None, count=0, start=0, end=30,
                       generate=False, log_lines=None, log_path="./logs/simulation.log", midi_path=None):
    """sim_log_to_midi.py:238-277 -> ``(piano_roll, durations, beats)``.  ``log_lines``: the simulator's log in memory (skips the file IPC);
    ``midi_path``: also write the .mid the reference writes (``./adj_sim_outputs/midi/simulation.mid`` / ``generation.mid``)."""
    if note_levels is None:
        note_levels = np.random.randint(0, 127, 16)        # the reference draws its default once, at import (:238)
    if gen2_output is None:
        raise TypeError("'NoneType' object is not subscriptable")      # :22: MidiGenerator indexes gen2_output before the None check at :248
    if log_lines is None:
        try:
            with open(log_path, 'r') as f:
                log_lines = f.readlines()
        except Exception:
            raise ValueError("Error in processing log file")
    if midi_path is None and instruments is not None and isinstance(gen2_output, np.ndarray) and gen2_output.ndim == 1 and len(gen2_output) >= 6:
        # nothing to write to disk: the native conversion (csrc/simlog.cu), same stream bit for bit
        stream = sim_log_to_event_stream_native(log_lines, instruments, note_levels, gen2_output, generate)
    else:
        stream, _ = sim_log_to_event_stream(log_lines, instruments, note_levels, gen2_output, generate, midi_path)
    return ds.generate_piano_roll(stream, start=start, end=end)


def _log_text(lines):
    """the log lines as one '\\n'-separated byte string (what ``readlines`` split); a str / bytes object is taken as the file's text"""
    if isinstance(lines, bytes):
        return lines
    if isinstance(lines, str):
        return lines.encode()
    return "".join(l if l.endswith("\n") else l + "\n" for l in lines).encode()


def _gen2_rows(g):
    g = np.asarray(g)
    return (np.ascontiguousarray(g, dtype=np.float32), 1) if g.dtype == np.float32 else (np.ascontiguousarray(g, dtype=np.float64), 0)


def sim_log_to_event_stream_native(log_lines, instruments, note_levels, gen2_output, generate=False):
    """``sim_log_to_event_stream(...)[0]`` through the C-ABI library (``mmg_simlog_to_events``, csrc/simlog.cu): the same message stream, bit for
    bit, without the Python state machine (no ``MidiGenerator`` object comes back and no .mid is written)."""
    import ctypes
    lib = ds.N.lib()
    text = _log_text(log_lines)
    ins, nl = np.ascontiguousarray(np.asarray(instruments), dtype=np.int64), np.ascontiguousarray(np.asarray(note_levels), dtype=np.int64)
    g, is_f32 = _gen2_rows(gen2_output)
    cap = lib.mmg_simlog_max_messages()
    dt, meta, n = np.empty(cap, dtype=np.float64), np.empty(cap, dtype=np.uint32), ctypes.c_int64(0)
    rc = lib.mmg_simlog_to_events(text, len(text), ins.ctypes.data, len(ins), nl.ctypes.data, len(nl), g.ctypes.data, len(g), is_f32, int(bool(generate)),
                                  dt.ctypes.data, meta.ctypes.data, cap, ctypes.byref(n))
    if rc != 0:
        raise ValueError(lib.mmg_last_error().decode())
    return ds.EventStream(dt[:n.value].copy(), meta[:n.value].copy())


def sim_logs_to_event_batch(logs, instruments, note_levels, gen2_outputs, generate=False, threads=0):
    """A batch of simulated songs -> pinned ``(dt float64, meta int32, offsets int64)`` host tensors, the ``fake_*_events`` entry of a
    ``trainer.HostBatchPipeline`` batch.  ``logs[i]`` are song i's log lines (or its log text); ``instruments[i]`` / ``note_levels[i]`` /
    ``gen2_outputs[i]`` the per-song arguments ``matrix_to_midi`` passes (matrix_sim_process.py:171).  The conversion runs natively on ``threads``
    host threads (0 = all cores): ``mmg_simlog_batch_to_events``; ``sim_log_to_event_stream`` is the same thing per song in Python."""
    import torch
    lib = ds.N.lib()
    texts = [_log_text(lg) for lg in logs]
    S = len(texts)
    log_off = np.zeros(S + 1, dtype=np.int64)
    np.cumsum([len(t) for t in texts], out=log_off[1:])
    blob = b"".join(texts)
    ins = np.ascontiguousarray(np.asarray(instruments), dtype=np.int64).reshape(S, -1) if S else np.zeros((0, 1), dtype=np.int64)
    nl = np.ascontiguousarray(np.asarray(note_levels), dtype=np.int64).reshape(S, -1) if S else np.zeros((0, 1), dtype=np.int64)
    g, is_f32 = _gen2_rows(gen2_outputs)
    g = g.reshape(S, -1) if S else np.zeros((0, 6), dtype=g.dtype)
    slot = lib.mmg_simlog_max_messages()
    dt, meta, off = np.empty(S * slot, dtype=np.float64), np.empty(S * slot, dtype=np.uint32), np.zeros(S + 1, dtype=np.int64)
    rc = lib.mmg_simlog_batch_to_events(blob, log_off.ctypes.data, S, ins.ctypes.data, ins.shape[1], nl.ctypes.data, nl.shape[1], g.ctypes.data, g.shape[1],
                                        is_f32, int(bool(generate)), dt.ctypes.data, meta.ctypes.data, off.ctypes.data, int(threads))
    if rc != 0:
        raise ValueError(lib.mmg_last_error().decode())
    total = int(off[-1])
    arrays = (dt[:total].copy(), meta[:total].view(np.int32).copy(), off)
    return tuple(torch.from_numpy(a).pin_memory() if torch.cuda.is_available() else torch.from_numpy(a) for a in arrays)


# ----------------------------------------------------------------------------------------------
# minimal Standard MIDI File writer (what mid.save() leaves on disk for the host consumers; datasets.read_smf reads it back)
# ----------------------------------------------------------------------------------------------
def _vlq(v):
    v = int(v)
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def write_smf(path, tracks, ticks_per_beat=TICKS_PER_BEAT):
    chunks = []
    for tr in tracks:
        body = bytearray()
        for typ, t, a, b in tr:
            body += _vlq(max(0, t))
            if typ == "note_on":
                body += bytes([0x90, a & 0x7F, b & 0x7F])
            elif typ == "note_off":
                body += bytes([0x80, a & 0x7F, b & 0x7F])
            elif typ == "program_change":
                body += bytes([0xC0, a & 0x7F])
            elif typ == "set_tempo":
                body += b"\xFF\x51\x03" + int(a).to_bytes(3, "big")
            elif typ == "time_signature":
                body += b"\xFF\x58\x04\x04\x02\x18\x08"
            elif typ == "key_signature":
                body += b"\xFF\x59\x02\x00\x00"
            elif typ == "end_of_track":
                body += b"\xFF\x2F\x00"
        chunks.append(b"MTrk" + struct.pack(">I", len(body)) + bytes(body))
    with open(path, "wb") as f:
        f.write(b"MThd" + struct.pack(">IHHH", 6, 1, len(tracks), ticks_per_beat) + b"".join(chunks))
