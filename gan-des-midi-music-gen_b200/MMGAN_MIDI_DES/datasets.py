"""Drop-in mirror of the reference's ``MMGAN_MIDI_DES/datasets.py``
(/root/reference/MMGAN_MIDI_DES/datasets.py:13-123): ``generate_piano_roll`` and the three
``MaestroDataset*`` classes, with the note-event -> 128-pitch-grid rasterisation done by the CUDA
kernel ``mmg_raster_piano_roll`` (csrc/raster.cu) instead of the CPython loop (:27-54).

The third-party pieces stay on the host: SMF parsing + tick->second (mido 1.3.2 in the reference;
here a small built-in reader that follows mido's merge/tempo rules, or any iterable of mido-like
message objects) and the beat grid (pretty_midi 0.2.10 ``get_beats`` in the reference; here
``EventStream.beats``).  Parity at those two boundaries is unpinned: the raster input
is defined as the post-mido event stream.

There is no CPU fallback: rasterisation needs a CUDA device and the built library.
"""
import ctypes
import glob
import os
import pickle
import struct

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import _native as N

KIND_OTHER, KIND_ON, KIND_OFF = 0, 1, 2
_KIND_OF_TYPE = {"note_on": KIND_ON, "note_off": KIND_OFF}
_TORCH_OUT = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2}


class EventStream:
    """One song as the post-mido message stream: ``dt`` seconds (float64) and packed ``meta``
    (kind | pitch << 8 | velocity << 16), plus the beat times the reference takes from pretty_midi."""

    def __init__(self, dt, meta, beats=(), filename=None):
        self.dt = np.ascontiguousarray(dt, dtype=np.float64)
        self.meta = np.ascontiguousarray(meta, dtype=np.uint32)
        if self.dt.shape != self.meta.shape or self.dt.ndim != 1:
            raise ValueError("dt and meta must be 1-D arrays of equal length")
        self.beats = np.asarray(beats, dtype=np.float64)
        self.filename = filename

    def __len__(self):
        return len(self.dt)

    @classmethod
    def from_arrays(cls, dt, kind, pitch, velocity, beats=(), filename=None):
        meta = (np.asarray(kind, dtype=np.uint32) | (np.asarray(pitch, dtype=np.uint32) << 8) | (np.asarray(velocity, dtype=np.uint32) << 16))
        return cls(dt, meta, beats, filename)

    @classmethod
    def from_messages(cls, messages, beats=(), filename=None):
        """From any iterable of mido-like messages (``.type``, ``.time`` seconds, ``.note``, ``.velocity``)."""
        dt, meta = [], []
        for m in messages:
            k = _KIND_OF_TYPE.get(m.type, KIND_OTHER)
            dt.append(float(m.time))
            meta.append(k | (int(m.note) << 8) | (int(m.velocity) << 16) if k else 0)
        return cls(np.array(dt, dtype=np.float64), np.array(meta, dtype=np.uint32), beats, filename)

    @classmethod
    def from_midi_file(cls, path):
        return read_smf(path)


# ----------------------------------------------------------------------------------------------
# Standard MIDI File reader: native (csrc/smf.cu, mmg_smf_parse); follows mido 1.3.2: merge_tracks, tick2second, running tempo
# ----------------------------------------------------------------------------------------------
def read_smf(path):
    """Parses a type-0/1 Standard MIDI File into an EventStream the way ``for msg in mido.MidiFile``
    yields it (datasets.py:18,34): tracks merged by absolute tick (stable), end_of_track metas
    dropped and one re-appended, delta seconds = ticks * (tempo * 1e-6 / ticks_per_beat) with the
    tempo switching after each set_tempo message.  The parsing is ``mmg_smf_parse`` of the C-ABI library (host code);
    ``oracle/smf_oracle.py`` is its Python checker."""
    with open(path, "rb") as f:
        raw = f.read()
    return parse_smf_bytes(raw, filename=path)


def read_smf_many(paths, workers=None):
    """``[read_smf(p) for p in paths]`` on a thread pool (the file read and the native parse both run without the interpreter lock): the
    MAESTRO corpus is 1276 files (notebook cell 10).  Results are in input order."""
    paths = list(paths)
    workers = min(32, os.cpu_count() or 1) if workers is None else int(workers)
    if len(paths) <= 1 or workers <= 1:
        return [read_smf(p) for p in paths]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(read_smf, paths))


def parse_smf_bytes(raw, filename=None):
    """``read_smf`` for a file image already in memory (bytes)."""
    name = filename if filename is not None else "<bytes>"
    if raw[:4] != b"MThd" or len(raw) < 14:
        raise ValueError(f"{name}: not a Standard MIDI File")
    fmt, _, div = struct.unpack(">HHH", raw[8:14])
    if div & 0x8000:
        raise ValueError("SMPTE time division is not supported")
    if fmt == 2:
        raise TypeError("can't merge tracks in type 2 (asynchronous) file")
    lib = N.lib()
    cap = int(lib.mmg_smf_max_messages(len(raw)))
    dt, meta, ticks = np.empty(cap, dtype=np.float64), np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=np.int64)
    t_tick, t_us = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int32)
    n, tpb, nt = ctypes.c_int64(0), ctypes.c_int(0), ctypes.c_int64(0)
    buf = (ctypes.c_ubyte * len(raw)).from_buffer_copy(raw)
    rc = lib.mmg_smf_parse(buf, len(raw), dt.ctypes.data, meta.ctypes.data, ticks.ctypes.data, cap, ctypes.byref(n), ctypes.byref(tpb),
                           t_tick.ctypes.data, t_us.ctypes.data, cap, ctypes.byref(nt))
    if rc != 0:
        raise ValueError(f"{name}: {lib.mmg_last_error().decode()}")
    n, nt = n.value, nt.value
    dt, meta, ticks = dt[:n].copy(), meta[:n].copy(), ticks[:n]
    # beat grid: quarter-note beats along the tempo map up to the last note event (host-side estimate of
    # pretty_midi.get_beats; parity at this boundary is unpinned)
    note_ticks = ticks[(meta & 0xFF) != 0]
    last = int(note_ticks.max()) if len(note_ticks) else 0
    beats = np.empty(last // tpb.value + 1, dtype=np.float64)
    nb = ctypes.c_int64(0)
    N.check(lib.mmg_smf_beat_grid(t_tick.ctypes.data, t_us.ctypes.data, nt, tpb.value, last, beats.ctypes.data, len(beats), ctypes.byref(nb)), "mmg_smf_beat_grid")
    return EventStream(dt, meta, beats, filename=filename)


# ----------------------------------------------------------------------------------------------
# rasterisation (device)
# ----------------------------------------------------------------------------------------------
def out_width(start, end):
    """Width of what the reference returns after its re-slice (datasets.py:49-54)."""
    return N.lib().mmg_raster_out_width(int(start), int(end))


# default kernel path of rasterize_events (csrc/raster.cu): "stream", "sort", or "auto" = stream for batches of short songs (the training loop's
# call shape: a few hundred messages per simulated song, where it is 1.7x faster), sort for long ones (MAESTRO windows: on par / 10 % faster)
RASTER_PATH = os.environ.get("MMG_RASTER_PATH", "auto")
RASTER_AUTO_MAX_EVENTS_PER_SONG = 4096


def rasterize_events(dt, meta, offsets, sequence_length=100, start=0, end=50, out_dtype=torch.float32, status=False, out=None, workspace=None,
                     path=None):
    """Device-resident batch rasterisation: ``dt`` (E,) float64, ``meta`` (E,) int32-typed packed u32,
    ``offsets`` (S+1,) int64, all CUDA tensors -> (S, 2, 128, Wout) tensor of ``out_dtype``.
    ``out`` (a contiguous (S,2,128,Wout) CUDA tensor, its dtype wins) and ``workspace`` (uint8, at least
    ``raster_workspace_bytes(S, E)``) let a caller that rasterises every step reuse its buffers (no allocation on the stream).
    ``path``: "stream" (warp-specialised single kernel, no workspace) or "sort" (chain + compaction kernel, sort-by-pitch write-once kernel);
    both are bit-exact, None = ``RASTER_PATH`` ("auto": by messages per song)."""
    path = RASTER_PATH if path is None else path
    if path == "auto":
        path = "stream" if dt.numel() <= RASTER_AUTO_MAX_EVENTS_PER_SONG * max(1, offsets.numel() - 1) else "sort"
    if path not in ("stream", "sort"):
        raise ValueError("path must be 'stream', 'sort' or 'auto'")
    N.require_cuda(dt, meta, offsets)
    if end - start < 0:
        raise ValueError("end-start must be >= 0")
    S, E = offsets.numel() - 1, dt.numel()
    Wo = out_width(start, end)
    if out is None:
        out = torch.empty(S, 2, 128, Wo, device=dt.device, dtype=out_dtype)      # the kernel writes every cell
    else:
        N.require_cuda(out)
        if tuple(out.shape) != (S, 2, 128, Wo) or not out.is_contiguous() or out.dtype not in _TORCH_OUT:
            raise ValueError(f"out must be a contiguous {(S, 2, 128, Wo)} float32 / bfloat16 / uint8 tensor")
        out_dtype = out.dtype
    st = torch.zeros(S, device=dt.device, dtype=torch.int32) if status else None
    ws_bytes = raster_workspace_bytes(S, E) if path == "sort" else 0
    if ws_bytes == 0:
        ws = None
    elif workspace is None:
        ws = torch.empty(ws_bytes, device=dt.device, dtype=torch.uint8) if ws_bytes else None
    else:
        N.require_cuda(workspace)
        if workspace.dtype != torch.uint8 or workspace.numel() < ws_bytes:
            raise ValueError(f"workspace must be a uint8 tensor of at least {ws_bytes} bytes")
        ws = workspace
    N.call("mmg_raster_piano_roll", N.ptr(dt), N.ptr(meta), N.ptr(offsets), S, E, -1 if sequence_length is None else int(sequence_length),
           int(start), int(end), _TORCH_OUT[out_dtype], N.ptr(out), N.ptr(st), N.ptr(ws), ws_bytes, N.stream())
    return (out, st) if status else out


def raster_workspace_bytes(n_songs, n_events):
    """Scratch bytes the fast path of ``mmg_raster_piano_roll`` wants for a batch of this size."""
    return int(N.lib().mmg_raster_workspace_bytes(int(n_songs), int(n_events)))


def pack_streams(streams, device="cuda"):
    """Host EventStreams -> (dt, meta, offsets) CUDA tensors (one pinned staging copy each)."""
    lens = np.array([len(s) for s in streams], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    dt = np.concatenate([s.dt for s in streams]) if len(streams) else np.zeros(0)
    meta = np.concatenate([s.meta for s in streams]) if len(streams) else np.zeros(0, dtype=np.uint32)
    to = lambda a: torch.from_numpy(a).pin_memory().to(device, non_blocking=True)
    return to(dt.astype(np.float64)), to(meta.astype(np.uint32).view(np.int32)), to(offsets)


def rasterize_batch(streams, sequence_length=100, start=0, end=50, device="cuda", out_dtype=torch.float32):
    """Batch of EventStreams -> (S, 2, 128, Wout) device tensor (plane 0 roll, plane 1 durations)."""
    dt, meta, offsets = pack_streams(streams, device)
    return rasterize_events(dt, meta, offsets, sequence_length, start, end, out_dtype)


def _as_stream(midi_input):
    if isinstance(midi_input, EventStream):
        return midi_input
    if isinstance(midi_input, str):
        return read_smf(midi_input)
    try:                                                  # a real mido.MidiFile, when mido is installed
        import mido
        if isinstance(midi_input, mido.MidiFile):
            s = EventStream.from_messages(midi_input, filename=midi_input.filename)
            try:
                import pretty_midi
                s.beats = np.asarray(pretty_midi.PrettyMIDI(midi_input.filename).get_beats(), dtype=np.float64)
            except ImportError:
                pass
            return s
    except ImportError:
        pass
    raise ValueError("midi_input must be a file path or a mido.MidiFile object")


def generate_piano_roll(midi_input, sequence_length=100, beats_length=50, start=0, end=50):
    """datasets.py:13-70 -- returns ``(piano_roll, durations, beats)`` as float64 numpy arrays of shape
    (128, Wout), (128, Wout), (beats_length,).  ``midi_input``: a path, a ``mido.MidiFile`` (when mido is
    installed) or an :class:`EventStream`."""
    s = _as_stream(midi_input)
    out = rasterize_batch([s], sequence_length, start, end)[0].cpu().numpy().astype(np.float64)
    beats = np.asarray(s.beats, dtype=np.float64)
    if len(beats) < beats_length:                         # :60-62
        beats = np.pad(beats, (0, beats_length - len(beats)))
    elif len(beats) > beats_length:                       # :63-65
        beats = beats[:beats_length]
    return out[0], out[1], beats


# ----------------------------------------------------------------------------------------------
# MAESTRO preprocessing (data_viewing_and_processing.ipynb cells 10-11 -> data/preprocessed_data_50.pkl)
# ----------------------------------------------------------------------------------------------
def total_time_steps(stream, sample_size):
    """``total_time`` of the notebook's generate_piano_roll (cell 10): the time step of the last message the loop visited, i.e. of the
    first message whose step reaches ``sample_size`` (the loop assigns it before it breaks) or of the last message.  The float64 running
    sum is sequential (np.cumsum) and the rounding is half-to-even (np.rint), like ``int(round(my_time))``."""
    if len(stream) == 0:
        return 0
    steps = np.rint(np.cumsum(stream.dt)).astype(np.int64)
    over = np.nonzero(steps >= sample_size)[0]
    return int(steps[over[0]] if len(over) else steps[-1])


def preprocess_maestro(inputs, sample_size=300, sequence_length=50, beats_length=50, device="cuda"):
    """The MAESTRO pickling pipeline of the reference (notebook cells 10-11) with ONE batched rasterisation on the device: every file
    (path or :class:`EventStream`) is rasterised over a ``sample_size``-step window, cut into ``sequence_length``-step slices, slice 0 is
    skipped, and each kept slice becomes a ``(piano_roll (128,L), durations (128,L), beats (beats_length,))`` triple of float32 CPU
    tensors -- the list ``MaestroDatasetPickle`` unpickles (datasets.py:73-87)."""
    inputs = list(inputs)
    parsed = iter(read_smf_many([x for x in inputs if isinstance(x, str)]))              # all paths at once, on a thread pool
    streams = [next(parsed) if isinstance(x, str) else _as_stream(x) for x in inputs]
    if not streams:
        return []
    rolls = rasterize_batch(streams, sample_size, 0, sample_size, device=device).cpu()       # (S, 2, 128, sample_size), bit-exact
    out = []
    for s, r in zip(streams, rolls):
        beats = np.asarray(s.beats, dtype=np.float64)
        beats = np.pad(beats, (0, beats_length - len(beats))) if len(beats) < beats_length else beats[:beats_length]
        beats_t = torch.from_numpy(beats).float()
        n = int(np.floor(total_time_steps(s, sample_size) / sequence_length))
        for i in range(1, n):                                # slice 0 is skipped (cell 11: `and i != 0`)
            a = i * sequence_length
            if a + sequence_length > sample_size:            # the reference's shape check drops slices past the window
                break
            out.append((r[0, :, a:a + sequence_length].clone(), r[1, :, a:a + sequence_length].clone(), beats_t))
    return out


def save_preprocessed(triples, path):
    """Writes the list the way cell 11 does (pickle), readable by ``MaestroDatasetPickle``."""
    with open(path, "wb") as f:
        pickle.dump(triples, f)


# ----------------------------------------------------------------------------------------------
# Dataset classes (datasets.py:73-123)
# ----------------------------------------------------------------------------------------------
def _data_path(*parts):
    """The reference hard-codes Windows separators ('data\\\\name'); accept either layout."""
    win = "\\".join(parts)
    return win if os.path.exists(win) else os.path.join(*parts)


class MaestroDatasetPickle(Dataset):
    """Pickled list of (piano_roll (128,W), durations (128,W), beats (50,)) float tensors (datasets.py:73-87)."""

    def __init__(self, pickle_file_name, sequence_length=100, beats_length=50, device="cpu"):
        self.device = device
        with open(_data_path("data", pickle_file_name), "rb") as f:
            self.data = pickle.load(f)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        piano_roll, durations, beats = self.data[idx]
        return piano_roll.to(self.device), durations.to(self.device), beats.to(self.device)


class MaestroDatasetTorch(Dataset):
    """One ``torch.save``d triple per item under data/tensors (datasets.py:90-100)."""

    def __init__(self, root_dir, sequence_length=100, beats_length=50, device="cpu"):
        self.data_dir = root_dir
        self.device = device
        self.file_list = sorted(glob.glob(_data_path("data", "tensors", "*.pt")))

    def __len__(self):
        return len(self.file_list)

    def __getitem__(self, idx):
        return torch.load(self.file_list[idx])


class MaestroDatasetMidi(Dataset):
    """Rasterises one MAESTRO .midi per item (datasets.py:103-123).  Like the reference it globs
    data/maestro-v3.0.0 unless ``root_dir`` is an explicit list of paths / EventStreams."""

    def __init__(self, root_dir, sequence_length=100, beats_length=50, device="cpu"):
        self.root_dir = root_dir
        self.sequence_length = sequence_length
        self.beats_length = beats_length
        self.device = device
        if isinstance(root_dir, (list, tuple)):
            self.file_list = list(root_dir)
        else:
            self.file_list = sorted(glob.glob(_data_path("data", "maestro-v3.0.0", "**", "*.midi"), recursive=True))

    def __len__(self):
        return len(self.file_list)

    def __getitem__(self, idx):
        piano_roll, durations, beats = generate_piano_roll(self.file_list[idx], self.sequence_length, self.beats_length)
        f = lambda a: torch.from_numpy(a).float().to(self.device)
        return f(piano_roll), f(durations), f(beats)
