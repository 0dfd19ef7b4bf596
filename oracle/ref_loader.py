"""Load the *real* reference hot-path module by file path.  TEST INFRASTRUCTURE.

``import pyremap`` needs xarray/netCDF4/pyproj, which this image lacks, but
``pyremap/remapper/remap_numpy.py`` itself only needs numpy, scipy and the name
``xarray``.  We execute that file, unmodified and in place (never copied), with
whatever module ``sys.modules['xarray']`` holds -- the real one if installed,
else the small stand-in ``tests/minixarray.py``.

``/root/reference`` exists only in the authoring container: callers must check
:func:`available` and skip otherwise (the GPU box has no reference tree; there
the committed ``tests/golden`` fixtures made by this loader are used instead).
"""

from __future__ import annotations

import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get('PYREMAP_REFERENCE_ROOT', '/root/reference')
_HOT_PATH = os.path.join(REFERENCE_ROOT, 'pyremap', 'remapper', 'remap_numpy.py')
_mod = None


def available():
    return os.path.isfile(_HOT_PATH)


def load(xarray_module=None):
    """Return the reference ``remap_numpy`` module (executed from its own file)."""
    global _mod
    if _mod is not None and xarray_module is None:
        return _mod
    if not available():
        raise FileNotFoundError(_HOT_PATH)
    if xarray_module is not None:
        sys.modules['xarray'] = xarray_module
    elif 'xarray' not in sys.modules:
        try:
            import xarray  # noqa: F401
        except ImportError:
            here = os.path.dirname(os.path.abspath(__file__))
            tests_dir = os.path.join(os.path.dirname(here), 'tests')
            if tests_dir not in sys.path:
                sys.path.insert(0, tests_dir)
            import minixarray
            sys.modules['xarray'] = minixarray
    spec = importlib.util.spec_from_file_location('_pyremap_ref_remap_numpy',
                                                  _HOT_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mod = mod
    return mod


class _Var:
    """Stands in for an xarray variable: only ``.values`` is read."""

    def __init__(self, values):
        self.values = values


class FakeRemapper:
    """The three attributes ``_remap_numpy_array`` touches (remap_numpy.py:250-270)."""

    def __init__(self, matrix, frac_b, dst_grid_dims):
        self._matrix = matrix
        self._ds_map = {'dst_grid_dims': _Var(dst_grid_dims),
                        'frac_b': _Var(frac_b)}


def reference_remap_array(matrix, frac_b, dst_grid_dims, in_field, remap_axes,
                          renormalization_threshold):
    """Run the reference's own ``_remap_numpy_array`` (remap_numpy.py:223)."""
    ref = load()
    return ref._remap_numpy_array(FakeRemapper(matrix, frac_b, dst_grid_dims),
                                  in_field, list(remap_axes),
                                  renormalization_threshold)
