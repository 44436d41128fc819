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


class ShimMidiFile:
    """Stands in for mido.MidiFile: an iterable of ShimMessage with .filename."""
    beats = ()

    def __init__(self, events=None, filename=None, beats=()):
        self.events = list(events or [])
        self.filename = filename if filename is not None else self
        self.beats = beats

    def __iter__(self):
        return iter(self.events)


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
