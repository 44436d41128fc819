"""Import harness for the UNMODIFIED reference (test infrastructure only).

This file is part of the oracle tooling: it is used by ``oracle/make_golden.py``
(run in the build container, where ``/root/reference`` exists) to execute the
reference's own code and freeze golden vectors under ``tests/golden/``.
Nothing on the product path may import it; the GPU box has no /root/reference.

Recipe = SURVEY.md Appendix B: ``sys.modules`` stubs for the third-party
packages the reference imports but this image lacks, a functional ``mido`` /
``pretty_midi`` shim so that ``MMGAN_MIDI_DES/datasets.py:13-70`` runs on
pre-parsed event lists, and cwd/sys.path = the flat script directory.
"""
import os
import sys
import types
import importlib

REF_ROOT = os.environ.get("MMG_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    """A module whose every attribute is another stub and which is callable."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Stub(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __iter__(self):
        return iter(())


class ShimMessage:
    """What ``for msg in mido.MidiFile`` yields: .type .time(seconds) .note .velocity"""
    __slots__ = ("type", "time", "note", "velocity")

    def __init__(self, type, time, note=0, velocity=0):
        self.type, self.time, self.note, self.velocity = type, time, note, velocity


class ShimBuildMessage:
    """Stands in for mido.Message / mido.MetaMessage on the BUILD side (sim_log_to_midi.py:19-20,87-96,153-172): keyword attributes,
    equality by value (mido compares ``vars``; ``list.remove`` in save_midi relies on it) and ``copy(time=...)``."""

    def __init__(self, type, **kw):
        self.type = type
        self.time = kw.pop("time", 0)
        self.__dict__.update(kw)

    def copy(self, **kw):
        m = ShimBuildMessage.__new__(ShimBuildMessage)
        m.__dict__.update(self.__dict__)
        m.__dict__.update(kw)
        return m

    def __eq__(self, other):
        return isinstance(other, ShimBuildMessage) and vars(self) == vars(other)

    __hash__ = None


class ShimMidiTrack(list):
    pass


def _merge_tracks(tracks):
    """mido 1.3.2 ``merge_tracks``: absolute ticks, stable sort by time, back to deltas, every end_of_track removed (its delta carried
    to the next message) and one appended at the end (``fix_end_of_track``)."""
    msgs = []
    for tr in tracks:
        now = 0
        for m in tr:
            now += m.time
            msgs.append(m.copy(time=now))
    msgs.sort(key=lambda m: m.time)
    out, now, accum = [], 0, 0
    for m in msgs:
        delta = m.time - now
        now = m.time
        if m.type == "end_of_track":
            accum += delta
        else:
            out.append(m.copy(time=delta + accum))
            accum = 0
    out.append(ShimBuildMessage("end_of_track", time=accum))
    return out


class ShimMidiFile:
    """Stands in for mido.MidiFile.  Read side: an iterable of ShimMessage with .filename (pre-parsed events).  Build side
    (``mido.MidiFile()`` with no arguments): ``tracks`` / ``ticks_per_beat`` / ``save`` and iteration with mido 1.3.2's published
    semantics -- merged tracks, delta ticks -> seconds with the running tempo (``tick * (tempo * 1e-6 / ticks_per_beat)``), the tempo
    switching after the set_tempo message is yielded."""
    beats = ()

    def __init__(self, events=None, filename=None, beats=()):
        self.events = None if events is None else list(events)
        self.filename = filename if filename is not None else self
        self.beats = beats
        self.tracks = []
        self.ticks_per_beat = 480
        self.type = 1

    def save(self, filename=None):
        self.saved_as = filename

    def __iter__(self):
        if self.events is not None:
            return iter(self.events)
        return self._play()

    def _play(self):
        tempo = 500000
        for m in _merge_tracks(self.tracks):
            delta = m.time * (tempo * 1e-6 / self.ticks_per_beat) if m.time > 0 else 0
            yield m.copy(time=delta)
            if m.type == "set_tempo":
                tempo = m.tempo


class ShimPrettyMIDI:
    def __init__(self, f):
        self._beats = getattr(f, "beats", ())

    def get_beats(self):
        import numpy as np
        return np.asarray(self._beats, dtype=np.float64)


_STUBS = ["torchviz", "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
          "matplotlib.patches", "matplotlib.lines", "librosa", "librosa.display", "midi2audio",
          "IPython", "IPython.display", "seaborn"]


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    mido = _Stub("mido")
    mido.MidiFile = ShimMidiFile
    mido.Message = ShimBuildMessage
    mido.MetaMessage = ShimBuildMessage
    mido.MidiTrack = ShimMidiTrack
    sys.modules["mido"] = mido
    pm = _Stub("pretty_midi")
    pm.PrettyMIDI = ShimPrettyMIDI
    sys.modules["pretty_midi"] = pm


def import_mmgan():
    """Returns (network_tests, datasets) modules of MMGAN_MIDI_DES, unmodified."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    _install_stubs()
    d = os.path.join(REF_ROOT, "MMGAN_MIDI_DES")
    if d not in sys.path:
        sys.path.insert(0, d)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ds = importlib.import_module("datasets")
            nt = importlib.import_module("network_tests")
    finally:
        os.chdir(cwd)
    return nt, ds


def import_simlog():
    """Returns the unmodified MMGAN_MIDI_DES/sim_log_to_midi.py module (MidiGenerator, LogLineProcessor, process_adjsim_log)."""
    import_mmgan()
    d = os.path.join(REF_ROOT, "MMGAN_MIDI_DES")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        m = importlib.import_module("sim_log_to_midi")
    finally:
        os.chdir(cwd)
    return m


def import_gandes():
    """Returns the GAN_DES/SIMNN.py module (run in a SEPARATE process from import_mmgan)."""
    import torch.utils.data.dataset as tds
    if not hasattr(tds, "T_co"):
        tds.T_co = tds._T_co
    _install_stubs()
    d = os.path.join(REF_ROOT, "GAN_DES")
    if d not in sys.path:
        sys.path.insert(0, d)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        import warnings, io, contextlib
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            m = importlib.import_module("SIMNN")
    finally:
        os.chdir(cwd)
    return m


def import_gandes_util():
    """Returns the UNMODIFIED GAN_DES/util.py (mel-spectrogram helpers): librosa is stubbed (only the two librosa-based helpers need it), the
    REAL torchaudio of this image computes ``get_melspectrogram_db_tensor`` (util.py:37-61).  Run in a separate process from the other imports."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    sys.modules.setdefault("librosa", _Stub("librosa"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_gandes_util", os.path.join(REF_ROOT, "GAN_DES", "util.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
