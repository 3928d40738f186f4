"""Import the *real* reference functions from ``/root/reference`` (build
container only -- the directory does not exist on the GPU box).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Used by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and by the ``not gpu`` tests to pin the numpy restatement in
``oracle/ice_oracle.py`` against the reference itself when it is present.

``hicexplorer/hicCorrectMatrix.py:1-25`` imports a handful of third-party
modules that are absent here (``past``, ``hicmatrix``, ``krbalancing``,
``matplotlib``, ``cooler``, ``unidecode``); they are replaced by empty stub
modules because none of them is touched by ``MAD`` / ``filter_by_zscore`` /
``iterativeCorrection``.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hicexplorer", "iterativeCorrection.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns (iterativeCorrection, MAD, filter_by_zscore, parse_arguments)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    try:
        import past.builtins  # noqa: F401
    except ImportError:
        _stub("past")
        _stub("past.builtins", zip=zip, map=map)
    for name in ("unidecode",):
        try:
            __import__(name)
        except ImportError:
            _stub(name, unidecode=lambda s: s)
    try:
        import cooler  # noqa: F401
    except ImportError:
        fileops = _stub("cooler.fileops", is_cooler=lambda p: str(p).endswith(".cool"))
        _stub("cooler", fileops=fileops)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib", use=lambda *a, **k: None, rcParams={})
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.gridspec = _stub("matplotlib.gridspec")
        mpl.ticker = _stub("matplotlib.ticker", MultipleLocator=object, FormatStrFormatter=object)
    try:
        import hicmatrix  # noqa: F401
    except ImportError:
        hm = _stub("hicmatrix")
        hm.HiCMatrix = _stub("hicmatrix.HiCMatrix")
    try:
        import krbalancing  # noqa: F401
    except ImportError:
        _stub("krbalancing", __all__=[])
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # hicexplorer/__init__.py:21-22 calls multiprocessing.set_start_method('fork') at import time; neutralise
    # that side effect (it raises if a start method is already set and must not leak into the test process)
    import multiprocessing
    saved = multiprocessing.set_start_method
    multiprocessing.set_start_method = lambda *a, **k: None
    try:
        from hicexplorer.iterativeCorrection import iterativeCorrection
        from hicexplorer.hicCorrectMatrix import MAD, filter_by_zscore, parse_arguments
    finally:
        multiprocessing.set_start_method = saved
    return iterativeCorrection, MAD, filter_by_zscore, parse_arguments
