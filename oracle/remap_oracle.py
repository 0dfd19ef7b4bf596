"""CPU restatement of pyremap's in-memory weight application.  TEST INFRASTRUCTURE ONLY.

This file is the *checker* (and the timed CPU baseline of ``bench.py``); the
product (``pyremap_b200``) never imports it.  See ``oracle/__init__.py``.

Reference being restated (all paths under ``/root/reference``):

* ``pyremap/remapper/remap_numpy.py:72-139``  ``_load_mapping``  -> :func:`build_matrix`
* ``pyremap/remapper/remap_numpy.py:223-297`` ``_remap_numpy_array`` -> :func:`remap_array`
  (closed form) and :func:`remap_array_stepwise` (same NumPy pass structure, used
  only so that the CPU baseline does the same amount of work as the reference)
* the arithmetic itself lives in a third-party dependency that is NOT vendored in
  the reference tree: **scipy** (``pyproject.toml:31``, unpinned; this image has
  scipy 1.18.1).  ``csr_matrix.dot`` -> ``_sparsetools.csr_matvecs`` computes, for
  every row ``i`` in order and every stored entry ``jj`` of that row in stored
  (column-sorted) order, ``Y[i, :] += data[jj] * X[indices[jj], :]`` with a
  separate multiply and add (no FMA) starting from ``+0.0``.  That published
  algorithm is restated in :func:`spmm_ordered` (NumPy) and in
  ``oracle/csr_oracle.c`` (plain C); both are bit-equal to scipy's own kernel
  (``tests/test_oracle.py``).

Parity: PINNED against the reference module executed in the authoring container
(``tests/golden/*.npz`` made by ``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix

CANONICAL_NAN = np.float64(np.nan)


# --------------------------------------------------------------------------
# a3: _load_mapping  (remap_numpy.py:134-137)
# --------------------------------------------------------------------------
def build_matrix(S, row, col, n_b, n_a):
    """Destination-by-source weight matrix from the map file's 1-based triplets.

    Follows ``remap_numpy.py:134-137``: shift ``row``/``col`` to 0-based and let
    scipy build a CSR matrix (column-sorted rows, duplicate entries summed).
    """
    r0 = np.asarray(row) - 1
    c0 = np.asarray(col) - 1
    return csr_matrix((np.asarray(S), (r0, c0)), shape=(int(n_b), int(n_a)))


# --------------------------------------------------------------------------
# a6: scipy csr_matvecs, restated
# --------------------------------------------------------------------------
def spmm_ordered(indptr, indices, data, X):
    """``Y = A @ X`` in exactly scipy's order of operations, without scipy.

    Vectorised over rows: pass ``j`` adds the ``j``-th stored entry of every row
    that has one.  Each row therefore still sees its entries in stored order,
    one rounded multiply followed by one rounded add per entry (NumPy ufuncs do
    not fuse), starting from +0.0 -- the published ``csr_matvecs`` recurrence.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    data = np.asarray(data, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    squeeze = X.ndim == 1
    if squeeze:
        X = X[:, None]
    n_row = indptr.size - 1
    Y = np.zeros((n_row, X.shape[1]), dtype=np.float64)
    counts = np.diff(indptr)
    max_count = int(counts.max()) if n_row else 0
    for j in range(max_count):
        rows = np.nonzero(counts > j)[0]
        jj = indptr[rows] + j
        Y[rows] = Y[rows] + data[jj][:, None] * X[indices[jj]]
    return Y[:, 0] if squeeze else Y


def spmm_rowloop(indptr, indices, data, X):
    """The same recurrence as a literal double loop (tiny cases only)."""
    X = np.asarray(X, dtype=np.float64)
    n_row = len(indptr) - 1
    Y = np.zeros((n_row, X.shape[1]), dtype=np.float64)
    for i in range(n_row):
        for jj in range(indptr[i], indptr[i + 1]):
            Y[i, :] = Y[i, :] + data[jj] * X[indices[jj], :]
    return Y


# --------------------------------------------------------------------------
# a5: _remap_numpy_array  (remap_numpy.py:223-297)
# --------------------------------------------------------------------------
def _flatten(in_field, remap_axes):
    """remap_numpy.py:236-256: put the remap axes first and flatten both groups."""
    remap_axes = [int(a) for a in remap_axes]
    extra_axes = [a for a in range(in_field.ndim) if a not in remap_axes]
    extra_shape = [in_field.shape[a] for a in extra_axes]
    n_src = int(np.prod([in_field.shape[a] for a in remap_axes]))
    k = int(np.prod(extra_shape)) if extra_axes else 1
    flat = in_field.transpose(remap_axes + extra_axes).reshape((n_src, k))
    return flat, extra_shape


def _unflatten(flat, dst_grid_dims_file_order, extra_shape, remap_axes):
    """remap_numpy.py:280-295: dst dims go where the first source dim was."""
    dst_dims = [int(d) for d in np.asarray(dst_grid_dims_file_order)[::-1]]
    n_dst = len(dst_dims)
    full = np.reshape(flat, dst_dims + list(extra_shape))
    first = int(min(remap_axes))
    tail = list(range(n_dst, n_dst + len(extra_shape)))
    order = tail[:first] + list(range(n_dst)) + tail[first:]
    return np.transpose(full, order)


def remap_flat(matrix, frac_b, X, valid, threshold):
    """Closed form of ``remap_numpy.py:258-278`` on a flat ``[n_a, K]`` field.

    ``valid`` is ``None`` (unmasked branch: divide by ``frac_b`` where it is
    > 0) or a boolean ``[n_a, K]`` array (masked branch: data under the mask
    contributes ``S * 0.0``, the denominator is ``S @ valid`` and an element is
    kept iff that denominator is strictly greater than ``threshold``).

    Returns ``(values, keep)``; ``values`` is meaningful where ``keep``.
    """
    if valid is not None:
        weight = np.asarray(valid, dtype=np.float64)
        # the reference multiplies a float mask into a MaskedArray; NumPy leaves
        # the *mask operand's* value (0.0) under the mask (remap_numpy.py:264)
        x0 = np.where(valid, np.asarray(X, dtype=np.float64), 0.0)
        num = matrix.dot(x0)
        den = matrix.dot(weight)
        keep = den > threshold
    else:
        num = matrix.dot(np.asarray(X))
        num = np.asarray(num, dtype=np.float64)
        den = np.repeat(np.asarray(frac_b, dtype=np.float64)[:, None],
                        num.shape[1], axis=1)
        keep = den > 0.0
    values = num.copy()
    np.divide(num, den, out=values, where=keep)
    return values, keep


def remap_array(matrix, frac_b, dst_grid_dims, in_field, remap_axes,
                renormalization_threshold):
    """Restatement of ``_remap_numpy_array`` (remap_numpy.py:223-297).

    ``dst_grid_dims`` is in the map file's (Fortran) order, as stored in
    ``ds_map['dst_grid_dims']``.  Returns a ``numpy.ma.MaskedArray`` like the
    reference does.
    """
    is_ma = isinstance(in_field, np.ma.MaskedArray)
    flat, extra_shape = _flatten(in_field, remap_axes)
    if is_ma and renormalization_threshold is not None:
        valid = np.logical_not(np.ma.getmaskarray(flat))
        values, keep = remap_flat(matrix, frac_b, np.ma.getdata(flat), valid,
                                  renormalization_threshold)
    else:
        values, keep = remap_flat(matrix, frac_b, np.ma.getdata(flat), None,
                                  None)
    out = np.ma.masked_array(values, mask=np.logical_not(keep))
    return _unflatten(out, dst_grid_dims, extra_shape, remap_axes)


def nanfilled(masked):
    """What xarray stores for a MaskedArray (``remap_numpy.py:209-218``)."""
    return np.where(np.ma.getmaskarray(masked), np.nan, np.ma.getdata(masked))


def remap_array_stepwise(matrix, frac_b, dst_grid_dims, in_field, remap_axes,
                         renormalization_threshold):
    """Same result as :func:`remap_array`, but with the reference's pass
    structure (one full-array NumPy pass per line of ``remap_numpy.py:256-278``)
    so that timing it measures the work the reference does: the permuting copy,
    the float mask, the masked multiply, two sparse products, the comparison,
    the boolean-indexed divide and the final masked-array construction.
    Used only as the ``cpu_baseline`` / ``--impl reference`` arm of bench.py.
    """
    flat, extra_shape = _flatten(in_field, remap_axes)
    k = flat.shape[1]
    renorm = (isinstance(flat, np.ma.MaskedArray)
              and renormalization_threshold is not None)
    if renorm:
        live = np.array(~flat.mask, float)              # :263
        num = matrix.dot(live * flat)                   # :264
        den = matrix.dot(live)                          # :265
        keep = den > renormalization_threshold          # :266
    else:
        num = matrix.dot(flat)                          # :268
        fb = np.asarray(frac_b)
        den = fb.reshape((fb.size, 1)).repeat(k, axis=1)  # :270-273
        keep = den > 0.0                                # :274
    num[keep] /= den[keep]                              # :277
    out = np.ma.masked_array(num, mask=~keep)           # :278
    return _unflatten(out, dst_grid_dims, extra_shape, remap_axes)
