"""ORACLE (test infrastructure, not product code) -- piano-roll rasteriser.

CPU restatement of the raster core of the reference's ``generate_piano_roll``
(/root/reference/MMGAN_MIDI_DES/datasets.py:13-70), on the *post-mido* event
stream: per message a delta time in seconds (float64), a kind, a pitch and a
velocity.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import this module.

Pinned against the unmodified reference function (run through the mido shim of
``oracle/_refimport.py``) by ``oracle/make_golden.py`` -> ``tests/golden/raster_*.npz``
(SURVEY.md Appendix A vectors K1-K10 plus random streams).

Event encoding shared with the CUDA path (include/mmgan_b200.h):
    kind: 0 = any other message, 1 = note_on, 2 = note_off
    meta (uint32) = kind | pitch << 8 | velocity << 16
"""
import numpy as np

KIND_OTHER, KIND_ON, KIND_OFF = 0, 1, 2


def pack_meta(kind, pitch, vel):
    kind = np.asarray(kind, dtype=np.uint32)
    pitch = np.asarray(pitch, dtype=np.uint32)
    vel = np.asarray(vel, dtype=np.uint32)
    return (kind | (pitch << 8) | (vel << 16)).astype(np.uint32)


def unpack_meta(meta):
    meta = np.asarray(meta, dtype=np.uint32)
    return (meta & 0xFF).astype(np.uint8), ((meta >> 8) & 0xFF).astype(np.uint8), ((meta >> 16) & 0xFF).astype(np.uint8)


def time_steps(dt):
    """datasets.py:32-36 -- ``my_time += msg.time; int(round(my_time))``.

    Sequential float64 running sum (np.cumsum on float64 is a sequential loop,
    same rounding as the Python ``+=``), then round-half-even (Python ``round``
    on a float == np.rint)."""
    t = np.cumsum(np.asarray(dt, dtype=np.float64))
    return np.rint(t).astype(np.int64)


def raster_events(dt, kind, pitch, vel, sequence_length=100, start=0, end=50):
    """Returns (piano_roll, durations) float64, exactly as datasets.py:27-54.

    Loop termination (datasets.py:37-38, 41, 46): stop at the first message of
    ANY kind whose step >= sequence_length, or at the first note_on whose step
    >= W = end-start (IndexError swallowed by the bare ``except``).  A note_off
    with W <= step < sequence_length does not stop; its slice store clips."""
    if sequence_length is None:                      # datasets.py:14-15
        sequence_length = end + 20
    W = end - start
    if W < 0:
        raise ValueError("end-start must be >= 0")
    roll = np.zeros((128, W))                        # :28
    dur = np.zeros((128, W))                         # :29
    on_time = np.zeros(128, dtype=np.int64)          # :33
    steps = time_steps(dt)
    for i in range(len(steps)):
        s = int(steps[i])
        if s >= sequence_length:                     # :37-38
            break
        k = int(kind[i])
        if k == KIND_ON:                             # :39-42
            if s >= W or s < -W:                     # numpy IndexError -> except -> loop ends
                break
            roll[pitch[i], s] = vel[i]
            on_time[pitch[i]] = s
        elif k == KIND_OFF:                          # :43-45
            a = int(on_time[pitch[i]])
            dur[pitch[i], a:s] = s - a               # python slice semantics (clips, may be empty)
    if end < 128:                                    # :49-54  (len(piano_roll) == 128 rows)
        roll = roll[:, start:end]
        dur = dur[:, start:end]
    else:
        roll = roll[:, :end]
        dur = dur[:, :end]
    return roll, dur


def pad_beats(beats, beats_length=50):
    """datasets.py:57-65 -- zero-pad or truncate the beat vector (host side)."""
    beats = np.asarray(beats, dtype=np.float64)
    if len(beats) < beats_length:
        beats = np.pad(beats, (0, beats_length - len(beats)))
    elif len(beats) > beats_length:
        beats = beats[:beats_length]
    return beats


def out_width(sequence_length, start, end):
    """Width of the returned arrays after the :49-54 re-slice."""
    W = end - start
    if end < 128:
        return len(range(W)[start:end])
    return len(range(W)[:end])


def synth_songs(n_songs, n_events, T, seed=0, p_on=0.2, p_off=0.2):
    """synthetic MAESTRO-scale event streams.

    dt ~ Exp(mean=T/E) f64; kind in {on p_on, off p_off, other}; pitch
    randint(21,109); velocity randint(0,128) (0 included)."""
    rng = np.random.default_rng(seed)
    E = n_songs * n_events
    dt = rng.exponential(T / n_events, size=E)
    u = rng.random(E)
    kind = np.where(u < p_on, KIND_ON, np.where(u < p_on + p_off, KIND_OFF, KIND_OTHER)).astype(np.uint8)
    pitch = rng.integers(21, 109, size=E).astype(np.uint8)
    vel = rng.integers(0, 128, size=E).astype(np.uint8)
    offsets = (np.arange(n_songs + 1, dtype=np.int64) * n_events)
    return dt, pack_meta(kind, pitch, vel), offsets


def raster_batch(dt, meta, offsets, sequence_length, start, end):
    """Batch driver over ragged songs -> (S,2,128,Wout) float32 (what the Dataset hands to torch)."""
    kind, pitch, vel = unpack_meta(meta)
    S = len(offsets) - 1
    Wo = out_width(sequence_length, start, end)
    out = np.zeros((S, 2, 128, Wo), dtype=np.float32)
    for s in range(S):
        a, b = int(offsets[s]), int(offsets[s + 1])
        r, d = raster_events(dt[a:b], kind[a:b], pitch[a:b], vel[a:b], sequence_length, start, end)
        out[s, 0], out[s, 1] = r, d
    return out


# ----------------------------------------------------------------------------------------------
# plain-C twin (oracle/raster_oracle.c) -- same contract, used where the Python loop is too slow
# ----------------------------------------------------------------------------------------------
def _load_c():
    import ctypes, os, subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "libraster_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", here, "libraster_oracle.so"])
    lib = ctypes.CDLL(so)
    lib.mmg_oracle_raster_batch.restype = ctypes.c_long
    lib.mmg_oracle_raster_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long,
                                            ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
    lib.mmg_oracle_out_width.restype = ctypes.c_long
    lib.mmg_oracle_out_width.argtypes = [ctypes.c_long, ctypes.c_long]
    return lib


_C = None


def raster_batch_c(dt, meta, offsets, sequence_length, start, end, n_threads=1):
    """C twin of raster_batch. Returns (out (S,2,128,Wo) f32, n_note_on_applied)."""
    global _C
    if _C is None:
        _C = _load_c()
    if sequence_length is None:
        sequence_length = end + 20
    dt = np.ascontiguousarray(dt, dtype=np.float64)
    meta = np.ascontiguousarray(meta, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    S = len(offsets) - 1
    Wo = _C.mmg_oracle_out_width(start, end)
    if Wo < 0:
        raise ValueError("end-start must be >= 0")
    out = np.zeros((S, 2, 128, Wo), dtype=np.float32)

    def run(lo, hi):
        return _C.mmg_oracle_raster_batch(dt.ctypes.data, meta.ctypes.data, offsets.ctypes.data, lo, hi,
                                          sequence_length, start, end, out.ctypes.data)
    if n_threads <= 1 or S < 2:
        return out, run(0, S)
    from concurrent.futures import ThreadPoolExecutor
    bounds = np.linspace(0, S, n_threads + 1).astype(int)
    with ThreadPoolExecutor(n_threads) as ex:
        notes = sum(ex.map(lambda ab: run(int(ab[0]), int(ab[1])), zip(bounds[:-1], bounds[1:])))
    return out, notes


def preprocess_reference(streams, sample_size=300, sequence_length=50, beats_length=50):
    """ORACLE restatement of the MAESTRO pickling loop, /root/reference/MMGAN_MIDI_DES/data_viewing_and_processing.ipynb cells 10-11, on
    post-mido streams given as (dt, meta, beats) tuples: plain Python message loop (cell 10, which returns ``total_time`` = the step of the
    last message visited), then the slicing of cell 11 (50-step slices, slice 0 skipped, wrong-shaped slices dropped).
    Returns a list of (roll (128,L) f32, dur (128,L) f32, beats (beats_length,) f32) numpy triples."""
    out = []
    for dt, meta, beats in streams:
        kind, pitch, vel = unpack_meta(meta)
        roll = np.zeros((128, sample_size)); dur = np.zeros((128, sample_size))
        on = np.zeros(128)
        t, total = 0.0, 0
        for i in range(len(dt)):
            t += float(dt[i])
            step = int(round(t))
            total = step
            if step >= sample_size:
                break
            if kind[i] == KIND_ON:
                roll[pitch[i], step] = vel[i]
                on[pitch[i]] = step
            elif kind[i] == KIND_OFF:
                a = int(round(on[pitch[i]]))
                dur[pitch[i], a:step] = step - a
        b = np.asarray(beats, dtype=np.float64)
        b = np.pad(b, (0, beats_length - len(b))) if len(b) < beats_length else b[:beats_length]
        for j in range(int(np.floor(total / sequence_length))):
            rs, ds = roll[:, j * sequence_length:(j + 1) * sequence_length], dur[:, j * sequence_length:(j + 1) * sequence_length]
            if rs.shape[1] == sequence_length and ds.shape[1] == sequence_length and j != 0:
                out.append((rs.astype(np.float32), ds.astype(np.float32), b.astype(np.float32)))
    return out
