"""Array-level weight application on the GPU: the body of the reference's
``_remap_numpy_array`` (``/root/reference/pyremap/remapper/remap_numpy.py:223-297``)
expressed as one fused kernel launch through the C ABI.

Layout.  The reference permutes the field so the source (remap) axes come first,
flattens to ``[nSrc, K]`` (a full copy, ``:256``), multiplies, and permutes back
(``:280-295``).  When the remap axes are adjacent and in order -- every field
pyremap's own callers produce -- the field already *is* ``[B, nSrc, L]`` in memory
(``B`` = leading extra dims, ``L`` = trailing extra dims), and the wanted output
``[B, nDst, L]`` is exactly what a batched launch writes: no copy on either side.
Only ``L == 1`` with ``B > 1`` (source dims last, e.g. ``(time, lat, lon)``) is
transposed on the device so that K is contiguous for the gather.

PyTorch is used as the device-buffer carrier and for H2D/D2H copies only.
"""

from __future__ import annotations

import numpy as np

from . import _cabi
from ._cabi import (F32, F64, KERNEL_AUTO, KERNEL_WROW_F32, MODE_FRACB, MODE_MASKED, MODE_RAW,
                    B200RemapError)

_MAX_BATCH = 65535


def _torch():
    try:
        import torch
    except ImportError as exc:  # pragma: no cover
        raise B200RemapError(-3, 'PyTorch is required as the device-tensor '
                             'carrier') from exc
    return torch


def require_cuda(device=None):
    """Resolve the CUDA device to run on; fail loudly when there is none."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise B200RemapError(
            -3, 'no CUDA device is available; pyremap_b200 runs on NVIDIA B200 '
            '(sm_100a) only and has no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    if device.type != 'cuda':
        raise B200RemapError(-1, f'device must be a CUDA device, got {device}')
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    return device


class Layout:
    """How a field of ``shape`` with ``remap_axes`` maps onto ``[B, nSrc, L]``."""

    def __init__(self, shape, remap_axes, dst_dims):
        shape = tuple(int(s) for s in shape)
        remap_axes = [int(a) for a in remap_axes]
        if len(set(remap_axes)) != len(remap_axes) or not remap_axes:
            raise ValueError(f'invalid remap_axes {remap_axes}')
        for a in remap_axes:
            if a < 0 or a >= len(shape):
                raise ValueError(f'remap axis {a} out of range for {len(shape)}-d field')
        self.shape = shape
        self.remap_axes = remap_axes
        self.extra_axes = [a for a in range(len(shape)) if a not in remap_axes]
        self.extra_shape = [shape[a] for a in self.extra_axes]
        self.dst_dims = [int(d) for d in dst_dims]
        self.n_src = int(np.prod([shape[a] for a in remap_axes], dtype=np.int64))
        self.n_dst = int(np.prod(self.dst_dims, dtype=np.int64))
        self.K = int(np.prod(self.extra_shape, dtype=np.int64)) if self.extra_axes else 1
        first = min(remap_axes)
        self.first = first
        self.adjacent = remap_axes == list(range(first, first + len(remap_axes)))
        if self.adjacent:
            self.B = int(np.prod(shape[:first], dtype=np.int64)) if first else 1
            self.L = int(np.prod(shape[first + len(remap_axes):], dtype=np.int64))
            self.out_shape = shape[:first] + tuple(self.dst_dims) + shape[first + len(remap_axes):]
        else:
            self.B, self.L = 1, self.K
            lead = [shape[a] for a in self.extra_axes[:first]]
            tail = [shape[a] for a in self.extra_axes[first:]]
            self.out_shape = tuple(lead) + tuple(self.dst_dims) + tuple(tail)

    def unpermute_axes(self):
        """Axes order that moves ``dst_dims + extra`` back (remap_numpy.py:288-295)."""
        n_dst = len(self.dst_dims)
        tail = list(range(n_dst, n_dst + len(self.extra_shape)))
        return tail[:self.first] + list(range(n_dst)) + tail[self.first:]


def _as_device_tensor(array, device, torch):
    """numpy / torch (any device) -> contiguous f32/f64 CUDA tensor on ``device``."""
    if isinstance(array, torch.Tensor):
        t = array
        if t.dtype not in (torch.float64, torch.float32):
            t = t.to(torch.float64)
        return t.to(device, non_blocking=True)
    a = np.asarray(array)
    if a.dtype not in (np.float64, np.float32):
        # ints / float16 / bool: the reference upcasts through the float64 weights
        a = a.astype(np.float64)
    if not a.dtype.isnative:
        a = a.astype(a.dtype.newbyteorder('='))
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:
        a = a.copy()
    t = torch.from_numpy(a)
    return t.to(device, non_blocking=True)


def _wants_f32(out_dtype):
    if out_dtype is None:
        return False
    dt = np.dtype(out_dtype) if not str(out_dtype).startswith('torch.') else \
        np.dtype(str(out_dtype).split('.')[1])
    if dt == np.float64:
        return False
    if dt == np.float32:
        return True
    raise ValueError(f'out_dtype must be float64 or float32, got {out_dtype}')


def _dtype_code(t, torch):
    return F64 if t.dtype == torch.float64 else F32


def device_any_nan(x, stream=None):
    """True iff the CUDA tensor ``x`` contains a NaN (kernel K4, early exit)."""
    torch = _torch()
    x = x if x.is_contiguous() else x.contiguous()
    with torch.cuda.device(x.device):
        flag = torch.empty(1, dtype=torch.int32, device=x.device)
        st = (stream or torch.cuda.current_stream(x.device)).cuda_stream
        _cabi.any_nan(x.data_ptr(), _dtype_code(x, torch), x.numel(),
                      flag.data_ptr(), st)
        return bool(flag.item())


def _transpose2d(t, nbatch, rows, cols, torch):
    """[nbatch, rows, cols] -> [nbatch, cols, rows] through kernel K5."""
    out = torch.empty((nbatch, cols, rows), dtype=t.dtype, device=t.device)
    _cabi.transpose(t.data_ptr(), out.data_ptr(), t.element_size(), nbatch, rows,
                    cols, torch.cuda.current_stream(t.device).cuda_stream)
    return out


def _permuted(t, order, torch):
    """``t.permute(order)`` materialised C-contiguous by the library's own kernel
    (``b200remap_permute``) -- uint8, float32 or float64 CUDA tensors, any strides."""
    shape = [t.shape[a] for a in order]
    strides = [t.stride(a) for a in order]
    out = torch.empty(shape, dtype=t.dtype, device=t.device)
    if out.numel():
        _cabi.permute(t.data_ptr(), out.data_ptr(), t.element_size(), shape, strides,
                      torch.cuda.current_stream(t.device).cuda_stream)
    return out


def apply_weights(matrix, dst_dims, field, remap_axes, threshold=None, *,
                  valid=None, mode='auto', device=None, want_keep=False,
                  return_torch=False, kernel=KERNEL_AUTO, out_dtype=None, out=None):
    """Remap ``field`` and return the NaN-filled float64 result.

    Parameters
    ----------
    matrix : pyremap_b200.mapfile.WeightMatrix
    dst_dims : destination grid dims in C order (``dst_grid_dims[::-1]``)
    field : numpy array or torch tensor (host or CUDA), any numeric dtype
    remap_axes : positions of the source dims in ``field``
    threshold : renormalisation threshold or None
    valid : optional boolean array shaped like ``field`` (True = use the value);
        selects the masked branch with an explicit mask (a ``MaskedArray``'s
        ``~mask``).  Without it the masked branch derives validity from
        ``!isnan``, as ``_remap_data_array`` does (remap_numpy.py:202-204).
    mode : 'auto' | 'raw' | 'fracb' | 'masked'.  'auto' reproduces the
        reference's branch selection: masked iff a threshold is given and the
        field has a mask (explicit, or any NaN anywhere), else ``frac_b``.
    want_keep : also return the boolean keep mask (``~`` of the reference's
        output mask).
    out : optional preallocated HOST result (numpy array or CPU tensor, C-contiguous, shape of
        the result, dtype ``out_dtype``) for host inputs: pinned memory receives the
        device->host copies directly, other memory is filled through a pinned staging ring.
    out_dtype : ``None`` / ``float64`` (the reference's result type) or ``float32``: every
        element is then the float64 result rounded to nearest float32 (= the reference's
        result ``.astype(float32)``), written by the kernel itself -- half the output traffic.

    Returns ``out`` or ``(out, keep)``; numpy arrays unless ``return_torch``.
    """
    torch = _torch()
    device = require_cuda(device if device is not None else (
        field.device if isinstance(field, torch.Tensor) and field.is_cuda else None))
    y_f32 = _wants_f32(out_dtype)
    lay = Layout(field.shape, remap_axes, dst_dims)
    if lay.n_src != matrix.shape[1]:
        raise ValueError(f'field has {lay.n_src} source cells but the map has '
                         f'{matrix.shape[1]}')
    if lay.n_dst != matrix.shape[0]:
        raise ValueError(f'destination dims {lay.dst_dims} do not match the map '
                         f'({matrix.shape[0]} rows)')
    if (valid is None and not want_keep and not return_torch and lay.adjacent
            and not (lay.L == 1 and lay.B > 1) and lay.L > 0 and lay.n_dst > 0):
        host = _host_view(field, torch)
        if host is not None:
            return _apply_host_streamed(matrix, lay, host, threshold, mode, device, kernel, torch,
                                        y_f32, out)
    if out is not None:
        raise ValueError('out= is only supported for contiguous host arrays remapped along '
                         'adjacent axes (the streamed path)')
    csr = matrix.on_device(device.index)

    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        x = _as_device_tensor(field, device, torch)
        v = None
        if valid is not None:
            vv = valid if isinstance(valid, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(np.asarray(valid, dtype=np.uint8)))
            v = vv.to(device=device, dtype=torch.uint8, non_blocking=True)
            if tuple(v.shape) != tuple(x.shape):
                raise ValueError('valid must have the same shape as field')

        # ---- branch selection (remap_numpy.py:202-204, 258-261) ----
        if mode == 'auto':
            if threshold is None:
                mode_code = MODE_FRACB
            elif v is not None:
                mode_code = MODE_MASKED
            else:
                mode_code = MODE_MASKED if device_any_nan(x, stream) else MODE_FRACB
        else:
            mode_code = {'raw': MODE_RAW, 'fracb': MODE_FRACB,
                         'masked': MODE_MASKED}[mode]
        if mode_code == MODE_MASKED and threshold is None:
            raise ValueError('the masked branch needs a renormalization threshold')
        if mode_code != MODE_MASKED:
            v = None
        if mode_code == MODE_FRACB and not csr.has_frac_b:
            raise ValueError('the map has no frac_b; cannot take the unmasked branch')
        thr = float(threshold) if threshold is not None else 0.0

        # ---- bring the field to [B, nSrc, L] with L contiguous ----
        transposed = False
        if lay.adjacent:
            B, L = lay.B, lay.L
            x3 = x.contiguous().view(B, lay.n_src, L)
            v3 = None if v is None else v.contiguous().view(B, lay.n_src, L)
            if L == 1 and B > 1:
                # source dims last: make the batch the contiguous K axis -- padded to a multiple
                # of 4 (zero columns, dropped again on the way back) so that the gathers run in
                # 256-bit lanes whatever the number of time slices (K = 365: 697 -> 392 us)
                Kp = lay.B if v3 is not None else (lay.B + 3) // 4 * 4
                if Kp == lay.B:
                    x3 = _transpose2d(x3, 1, B, lay.n_src, torch).view(1, lay.n_src, B)
                else:
                    xt = torch.zeros((1, lay.n_src, Kp), dtype=x3.dtype, device=device)
                    _cabi.transpose(x3.data_ptr(), xt.data_ptr(), x3.element_size(), 1, B, lay.n_src,
                                    stream.cuda_stream, ld_out=Kp)
                    x3 = xt
                if v3 is not None:
                    v3 = _permuted(v3.view(B, lay.n_src), [1, 0], torch).view(1, lay.n_src, B)
                transposed, B, L = True, 1, Kp
        else:
            order = lay.remap_axes + lay.extra_axes
            x3 = _permuted(x, order, torch).view(1, lay.n_src, lay.K)
            v3 = None if v is None else _permuted(v, order, torch).view(1, lay.n_src, lay.K)
            B, L = 1, lay.K

        y3 = torch.empty((B, lay.n_dst, L), dtype=torch.float32 if y_f32 else torch.float64,
                         device=device)
        k3 = torch.empty((B, lay.n_dst, L), dtype=torch.uint8,
                         device=device) if want_keep else None
        if L > 0 and lay.n_dst > 0:
            for b0 in range(0, B, _MAX_BATCH):
                nb = min(_MAX_BATCH, B - b0)
                csr.spmm(x3[b0].data_ptr(), _dtype_code(x3, torch), L, L, nb,
                         lay.n_src * L, y3[b0].data_ptr(), L, lay.n_dst * L,
                         mode_code, thr,
                         valid_ptr=None if v3 is None else v3[b0].data_ptr(),
                         keep_ptr=None if k3 is None else k3[b0].data_ptr(),
                         kernel=kernel, stream=stream.cuda_stream, y_f32=y_f32)

        # ---- back to the caller's layout (remap_numpy.py:280-295) ----
        def restore(t3):
            if lay.adjacent:
                if transposed:
                    Kp = t3.shape[-1]            # the padded batch axis (>= lay.B)
                    if t3.dtype == torch.uint8:
                        t3 = _permuted(t3.view(lay.n_dst, Kp)[:, :lay.B], [1, 0], torch)
                    else:
                        back = torch.empty((lay.B, lay.n_dst), dtype=t3.dtype, device=t3.device)
                        _cabi.transpose(t3.data_ptr(), back.data_ptr(), t3.element_size(), 1,
                                        lay.n_dst, lay.B, stream.cuda_stream, ld_in=Kp)
                        t3 = back
                return t3.reshape(lay.out_shape)
            full = t3.reshape(lay.dst_dims + lay.extra_shape)
            return _permuted(full, lay.unpermute_axes(), torch)

        out = restore(y3)
        keep = restore(k3).to(torch.bool) if want_keep else None
        if not return_torch:
            out = out.cpu().numpy()
            keep = None if keep is None else keep.cpu().numpy()
    return (out, keep) if want_keep else out


# --------------------------------------------------------------------------
# host arrays: streamed H2D -> kernel -> D2H
# --------------------------------------------------------------------------
_STREAMS = {}


class _Trace:
    """Phase timer of the streamed path, printed when ``B200REMAP_TRACE=1``."""

    def __init__(self):
        import os
        import time
        self.on = os.environ.get('B200REMAP_TRACE') == '1'
        self.clock = time.perf_counter
        self.t = self.clock()
        self.marks = []

    def mark(self, name):
        if self.on:
            now = self.clock()
            self.marks.append((name, (now - self.t) * 1e3))
            self.t = now

    def report(self, what):
        if self.on:
            import sys
            print('[b200remap] ' + what + ': ' +
                  ', '.join(f'{n} {ms:.2f} ms' for n, ms in self.marks), file=sys.stderr)


def _side_streams(device, torch):
    key = (device.index,)
    if key not in _STREAMS:
        _STREAMS[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return _STREAMS[key]


_PINNED = {}
_STREAM_LOCKS = {}


def _stream_lock(device):
    """The streamed path of a device (side streams, pinned staging rings) serves one call at a
    time; concurrent callers queue here."""
    import threading
    return _STREAM_LOCKS.setdefault(device.index, threading.Lock())


def _pinned_ring(torch, tag, count, nbytes):
    """``count`` persistent pinned byte buffers of at least ``nbytes`` (per process, grown on
    demand, never returned): page-locking fresh memory runs at ~2 GB/s, far below PCIe, so the
    staging blocks of the streamed path are allocated once and reused by every call."""
    ring = _PINNED.get(tag)
    if ring is None or len(ring) < count or ring[0].numel() < nbytes:
        ring = [torch.empty(max(int(nbytes), 1), dtype=torch.uint8, pin_memory=True)
                for _ in range(count)]
        _PINNED[tag] = ring
    return ring


class _Buf:
    """One recycled result buffer: a CPU uint8 tensor, page-locked once it has been released."""

    def __init__(self, tensor):
        self.tensor, self.pinned = tensor, False


class _ResultPool:
    """Recycled result buffers of the streamed path.

    A fresh pageable result costs a CPU copy out of the pinned D2H ring plus the first-touch
    page faults of the new array (C3: ~8 ms per 193 MB slice); page-locking a fresh block inside
    a call would cost ~0.5 ms per MB.  So results are numpy arrays backed by buffers of this
    pool.  The first result of a size is ordinary pageable memory (filled through the ring, as
    before).  When the caller drops it -- the array and every view of it -- a helper thread
    page-locks the buffer (``cudaHostRegister``) and parks it; the next result of that size is
    then written by the device->host copies directly, with no CPU copy and no page faults.
    Callers that keep every result alive simply stay on the pageable path.  The pool holds at
    most ``B200REMAP_RESULT_POOL_GB`` (default 8) of buffers; beyond that results are plain arrays.
    """

    def __init__(self):
        import os
        import threading
        self.free = []
        self.bytes = 0
        self.lock = threading.Lock()
        self.budget = int(float(os.environ.get('B200REMAP_RESULT_POOL_GB', '8')) * (1 << 30))
        self.worker = None

    def take(self, nbytes):
        """A parked page-locked buffer that fits ``nbytes`` without wasting half of it."""
        with self.lock:
            best = None
            for i, buf in enumerate(self.free):
                n = buf.tensor.numel()
                if buf.pinned and nbytes <= n <= 2 * nbytes and (best is None or n < self.free[best].tensor.numel()):
                    best = i
            return self.free.pop(best) if best is not None else None

    def fresh(self, nbytes, torch):
        """A new pageable buffer the pool will recycle, or None when the budget is used up."""
        with self.lock:
            while self.bytes + nbytes > self.budget and self.free:
                self._drop(self.free.pop(0), torch)
            if self.bytes + nbytes > self.budget:
                return None
            self.bytes += nbytes
        return _Buf(torch.empty(max(int(nbytes), 1), dtype=torch.uint8))

    def _drop(self, buf, torch):
        if buf.pinned:
            torch.cuda.cudart().cudaHostUnregister(buf.tensor.data_ptr())
        self.bytes -= buf.tensor.numel()

    def release(self, buf, device_index):
        """Finalizer of a handed-out array: page-lock (once) and park the buffer, off-thread."""
        try:
            if self.worker is None:
                from concurrent.futures import ThreadPoolExecutor
                self.worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix='b200remap-pin')
            self.worker.submit(self._park, buf, device_index)
        except RuntimeError:            # interpreter shutting down
            pass

    def _park(self, buf, device_index):
        torch = _torch()
        if not buf.pinned:
            try:
                torch.cuda.set_device(device_index)
                err = torch.cuda.cudart().cudaHostRegister(buf.tensor.data_ptr(),
                                                           buf.tensor.numel(), 1)   # portable
                buf.pinned = int(err) == 0
            except Exception:           # noqa: BLE001 - stay pageable, but do not keep it
                buf.pinned = False
        with self.lock:
            if buf.pinned:
                self.free.append(buf)
            else:
                self.bytes -= buf.tensor.numel()

    def wait_idle(self):
        """Block until every released buffer has been parked (tests, benchmarks)."""
        if self.worker is not None:
            self.worker.submit(lambda: None).result()


_RESULTS = _ResultPool()


def _new_result(shape, np_dtype, torch, device_index):
    """(host tensor of ``shape``, numpy array to return, direct) -- ``direct``: the tensor is
    page-locked, device->host copies may target it."""
    import weakref
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(np_dtype).itemsize
    buf = None
    if nbytes >= (1 << 20):
        buf = _RESULTS.take(nbytes) or _RESULTS.fresh(nbytes, torch)
    if buf is None:
        arr = np.empty(shape, dtype=np_dtype)
        return torch.from_numpy(arr), arr, False
    t_dtype = torch.float32 if np.dtype(np_dtype) == np.float32 else torch.float64
    t = buf.tensor[:nbytes].view(t_dtype).view(shape)
    arr = t.numpy()              # its base is the tensor, so every view of arr keeps arr alive
    fin = weakref.finalize(arr, _RESULTS.release, buf, device_index)
    fin.atexit = False
    return t, arr, buf.pinned


_POOL = []


def _copy_pool():
    """One helper thread that moves finished slices from the pinned ring into the result while
    the calling thread packs / enqueues the next ones."""
    if not _POOL:
        from concurrent.futures import ThreadPoolExecutor
        _POOL.append(ThreadPoolExecutor(max_workers=1, thread_name_prefix='b200remap-copy'))
    return _POOL[0]


def _chunks(nbytes, piece=1 << 20):
    off = np.arange(0, max(int(nbytes), 1), piece, dtype=np.int64)
    return off, np.minimum(piece, nbytes - off).astype(np.int64)


def _host_out(out, shape, y_dtype, torch):
    """Validate a caller-provided result buffer; returns it as a CPU tensor of ``shape``."""
    t = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
    if t.is_cuda or t.dtype != y_dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f'out must be a C-contiguous host array of shape {tuple(shape)} and '
                         f'dtype {y_dtype}')
    return t


def _host_view(field, torch):
    """A C-contiguous float32/float64 CPU tensor sharing ``field``'s memory, or None."""
    if isinstance(field, torch.Tensor):
        if field.is_cuda or field.dtype not in (torch.float64, torch.float32):
            return None
        return field if field.is_contiguous() else None
    a = field
    if isinstance(a, np.ma.MaskedArray) or not isinstance(a, np.ndarray):
        return None
    if a.dtype not in (np.float64, np.float32) or not a.dtype.isnative:
        return None
    if not a.flags.c_contiguous or not a.flags.writeable or a.size == 0:
        return None
    return torch.from_numpy(a)


class _Job:
    """One variable of a streamed call: host field ``[B, nSrc, L]`` in, host result
    ``[B, nDst, L]`` out, with its own branch (``mode_code``) and element types."""

    def __init__(self, lay, host, mode_code, thr, y_f32, out_t, direct=None):
        self.lay, self.host, self.mode_code, self.thr = lay, host, mode_code, thr
        self.y_f32, self.out_t = y_f32, out_t
        # direct: device->host copies may land in out_t itself (page-locked memory)
        self.direct = out_t.is_pinned() if direct is None else direct


def _host_mode(host, threshold, mode):
    """Branch selection over the whole variable on the host (remap_numpy.py:202-204,258-261):
    a native early-exit scan, because only the source rows the map touches travel."""
    if mode == 'auto':
        if threshold is None:
            mode_code = MODE_FRACB
        else:
            mode_code = MODE_MASKED if _cabi.host_any_nan(host.numpy()) else MODE_FRACB
    else:
        mode_code = {'raw': MODE_RAW, 'fracb': MODE_FRACB, 'masked': MODE_MASKED}[mode]
    if mode_code == MODE_MASKED and threshold is None:
        raise ValueError('the masked branch needs a renormalization threshold')
    return mode_code


def _apply_host_streamed(matrix, lay, host, threshold, mode, device, kernel, torch, y_f32=False,
                         out=None):
    """Host field in, host result out: a one-variable call of :func:`_stream_jobs`."""
    mode_code = _host_mode(host, threshold, mode)
    thr = float(threshold) if threshold is not None else 0.0
    y_dtype = torch.float32 if y_f32 else torch.float64
    out_shape = (lay.B, lay.n_dst, lay.L)
    arr, direct = None, None
    if out is None:
        out_t, arr, direct = _new_result(out_shape, np.float32 if y_f32 else np.float64, torch,
                                         device.index)
    else:
        out_t = _host_out(out, lay.out_shape, y_dtype, torch).view(out_shape)
    job = _Job(lay, host.view(lay.B, lay.n_src, lay.L), mode_code, thr, y_f32, out_t, direct)
    _stream_jobs(matrix, [job], device, kernel, torch)
    return arr.reshape(lay.out_shape) if out is None else out


#: rows up to this many bytes are packed side by side with other variables' (K-concatenation)
_THIN_ROW_BYTES = 256


def apply_weights_many(matrix, dst_dims, fields, threshold=None, *, device=None,
                       kernel=KERNEL_AUTO):
    """Remap several host variables through ONE streamed pipeline.

    ``fields``: list of ``(array, remap_axes)``.  The reference remaps the variables of a
    Dataset one by one (``ds.map(_remap_data_array)``, remap_numpy.py:42-55) and each pays the
    whole chain ``.values`` -> NumPy passes -> result.  Here every variable is one *job* of a
    single H2D / kernel / D2H pipeline: the weights are uploaded once, the side streams, pinned
    staging rings and device buffers are shared, and the exposed first H2D / last D2H of a
    variable overlap with its neighbours' work.  Branch selection stays per variable, exactly
    as in the reference.  Variables the streamed path does not cover (non-contiguous, remap
    axes not adjacent, integer dtype, ``(time, lat, lon)`` layouts) go through
    :func:`apply_weights` one by one.  Returns the list of NaN-filled float64 results.
    """
    torch = _torch()
    device = require_cuda(device)
    results = [None] * len(fields)
    jobs, where = [], []
    thin = {}            # (branch, dtype) -> [(index, layout, host view)] of K-concatenable variables
    for i, (field, remap_axes) in enumerate(fields):
        lay = Layout(field.shape, remap_axes, dst_dims)
        if lay.n_src != matrix.shape[1]:
            raise ValueError(f'field has {lay.n_src} source cells but the map has '
                             f'{matrix.shape[1]}')
        if lay.n_dst != matrix.shape[0]:
            raise ValueError(f'destination dims {lay.dst_dims} do not match the map '
                             f'({matrix.shape[0]} rows)')
        host = _host_view(field, torch) if (
            lay.adjacent and not (lay.L == 1 and lay.B > 1) and lay.L > 0 and lay.n_dst > 0) else None
        if host is None:
            results[i] = apply_weights(matrix, dst_dims, field, remap_axes, threshold,
                                       device=device, kernel=kernel)
            continue
        mode_code = _host_mode(host, threshold, 'auto')
        if lay.B == 1 and lay.L * host.element_size() <= _THIN_ROW_BYTES:
            thin.setdefault((mode_code, host.dtype), []).append((i, lay, host))
            continue
        out_t, arr, direct = _new_result((lay.B, lay.n_dst, lay.L), np.float64, torch, device.index)
        jobs.append(_Job(lay, host.view(lay.B, lay.n_src, lay.L), mode_code,
                         float(threshold) if threshold is not None else 0.0, False, out_t, direct))
        where.append((i, arr, None))
    # K-concatenation: variables with short rows -- 2-D fields such as
    # (Time=1, nCells), a few levels -- that take the same branch are packed side by side into
    # ONE [nSrc, sum(L)] job: one launch whose gathers read sum(L) contiguous elements per
    # source row instead of one launch per variable reading 8 bytes per row.  Every column is
    # computed exactly as it would be alone (the recurrence is per element).
    for (mode_code, _), members in thin.items():
        widths = [lay.L for _, lay, _ in members]
        if len(members) == 1:
            i, lay, host = members[0]
            packed = host.view(1, lay.n_src, lay.L)
        else:
            packed = torch.cat([host.view(lay.n_src, lay.L) for _, lay, host in members],
                               dim=1).view(1, members[0][1].n_src, sum(widths))
        lay_p = Layout((members[0][1].n_src, sum(widths)), [0], dst_dims)
        out_t, arr, direct = _new_result((1, lay_p.n_dst, lay_p.L), np.float64, torch, device.index)
        jobs.append(_Job(lay_p, packed, mode_code,
                         float(threshold) if threshold is not None else 0.0, False, out_t, direct))
        where.append((None, arr, members))
    if jobs:
        _stream_jobs(matrix, jobs, device, kernel, torch)
        for (i, arr, members), job in zip(where, jobs):
            if members is None:
                results[i] = arr.reshape(job.lay.out_shape)
            elif len(members) == 1:
                results[members[0][0]] = arr.reshape(members[0][1].out_shape)
            else:
                wide = arr.reshape(job.lay.n_dst, job.lay.L)
                at = 0
                for j, lay, _ in members:
                    results[j] = np.ascontiguousarray(wide[:, at:at + lay.L]).reshape(lay.out_shape)
                    at += lay.L
    return results


def _stream_jobs(matrix, jobs, device, kernel, torch):
    """Every leading-axis slice of every job through one 3-stage pipeline, one fused launch each.

    * only the source rows the map touches are copied to the GPU (regional maps: a few
      contiguous runs, :meth:`WeightMatrix.cover_exact`);
    * per slice: H2D on a copy stream, kernel on the current stream, D2H into pinned memory on
      a second copy stream -- the three overlap across slices AND across variables (PCIe is
      full duplex) with double-buffered device tensors;
    * results: a caller-provided pinned ``out`` receives the D2H copies directly; otherwise they
      land in a persistent pinned ring and CPU threads move each slice into the (pageable)
      result while the next slices are in flight -- page-locking a fresh result block per call
      would cost ~0.5 ms per MB.
    """
    trace = _Trace()
    cov = matrix.cover_exact()
    csr = matrix.on_device(device.index) if cov is None else \
        matrix.on_device_cover(device.index, exact=True)
    n_x = matrix.shape[1] if cov is None else cov['n_cover']
    n_dst = matrix.shape[0]
    for job in jobs:
        if job.mode_code == MODE_FRACB and not csr.has_frac_b:
            raise ValueError('the map has no frac_b; cannot take the unmasked branch')
    import os
    want = os.environ.get('B200REMAP_H2D', 'auto')       # 'dma' | 'gather' | 'auto' (experiments)
    # CPU threads for packing / copying out: the cores this process may use, shared fairly with
    # the other ranks of the node (torchrun exports LOCAL_WORLD_SIZE)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    ranks_here = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    threads = max(1, min(16, cores // ranks_here))
    total = sum(j.lay.B for j in jobs)
    nbuf = min(2, total)
    x_bytes = max(n_x * j.lay.L * j.host.element_size() for j in jobs)
    y_bytes = max(n_dst * j.lay.L * j.out_t.element_size() for j in jobs)
    rows_dev = None

    with _stream_lock(device), torch.cuda.device(device):
        compute = torch.cuda.current_stream(device)
        s_in, s_out = _side_streams(device, torch)
        trace.mark('weights on device')
        xd = [torch.empty(x_bytes, dtype=torch.uint8, device=device) for _ in range(nbuf)]
        yd = [torch.empty(y_bytes, dtype=torch.uint8, device=device) for _ in range(nbuf)]
        any_pageable = any(not j.host.is_pinned() for j in jobs)
        any_indirect = any(not j.direct for j in jobs)
        stage = _pinned_ring(torch, ('in', device.index), nbuf, x_bytes) if any_pageable else None
        ring_out = _pinned_ring(torch, ('out', device.index), nbuf, y_bytes) if any_indirect else None
        trace.mark('buffers')
        pending = [None] * nbuf     # copy-out job of the slice parked in ring_out[i]
        stage_free = [None] * nbuf  # the DMA that last read stage[i] has finished
        x_free = [None] * nbuf      # kernel that last read xd[i] has finished
        y_free = [None] * nbuf      # D2H that last read yd[i] has finished

        def copy_out(i, dst_ptr, nbytes, ev):   # runs on the helper thread (both calls drop the GIL)
            ev.synchronize()
            c_off, c_len = _chunks(nbytes)
            _cabi.host_pack_runs(ring_out[i].data_ptr(), dst_ptr, c_off, c_off, c_len, threads)

        def drain(i):
            if pending[i] is not None:
                pending[i].result()
                pending[i] = None

        s_in.wait_stream(compute)
        k = 0
        try:
            for job in jobs:
                lay, src = job.lay, job.host
                L, esz = lay.L, src.element_size()
                row_bytes = L * esz
                code = _dtype_code(src, torch)
                pinned = src.is_pinned()
                direct = job.direct
                slice_out = n_dst * L * job.out_t.element_size()
                # how this variable's touched rows travel
                pack = dma = None
                gather = False
                if cov is None:
                    runs = (np.zeros(1, np.int64), np.zeros(1, np.int64),
                            np.array([lay.n_src * row_bytes], np.int64))
                else:
                    runs = (np.ascontiguousarray(cov['run_start'] * row_bytes),
                            np.ascontiguousarray(cov['run_pos'] * row_bytes),
                            np.ascontiguousarray(cov['run_len'] * row_bytes))
                if not pinned:
                    # pageable input: the copy engines cannot read it, and a plain cudaMemcpy stages
                    # it at ~10 GB/s.  CPU threads pack the touched runs into a pinned staging block
                    # (~25 GB/s) that crosses PCIe as one DMA, overlapped with the next slice's packing
                    pack = runs
                else:
                    # pinned input.  Few long contiguous runs: one batched DMA submission per slice
                    # (copy engines run at full rate beside the D2H of results; SM loads from host
                    # memory do not).  Many short runs: the GPU gathers the rows itself.
                    n_runs = int(runs[0].size)
                    zero_copy = cov is not None and row_bytes % 16 == 0 and src.data_ptr() % 16 == 0
                    if want == 'gather' and zero_copy or (
                            want == 'auto' and zero_copy and not (
                                n_runs <= 16384 and n_x * row_bytes >= n_runs * 32768)):
                        gather = True
                        key = ('rows_dev', device.index)
                        if key not in cov:
                            cov[key] = torch.from_numpy(cov['rows']).to(device)
                        rows_dev = cov[key]
                    else:
                        dma = runs
                for b in range(lay.B):
                    i = k % nbuf
                    k += 1
                    x_view = xd[i][:n_x * row_bytes]
                    y_view = yd[i][:slice_out]
                    if pack is not None:
                        if stage_free[i] is not None:
                            stage_free[i].synchronize()      # the DMA of slice k-2 has read stage[i]
                        _cabi.host_pack_runs(src[b].data_ptr(), stage[i].data_ptr(), pack[0], pack[1],
                                             pack[2], threads)
                    if x_free[i] is not None:
                        s_in.wait_event(x_free[i])
                    with torch.cuda.stream(s_in):
                        if pack is not None:
                            x_view.copy_(stage[i][:n_x * row_bytes], non_blocking=True)
                            stage_free[i] = torch.cuda.Event()
                            stage_free[i].record(s_in)
                        elif gather:
                            _cabi.gather_rows(src[b].data_ptr(), x_view.data_ptr(), rows_dev.data_ptr(),
                                              n_x, row_bytes, row_bytes, s_in.cuda_stream)
                        else:
                            _cabi.copy_runs(src[b].data_ptr(), x_view.data_ptr(), dma[0], dma[1],
                                            dma[2], s_in.cuda_stream)
                        ready = torch.cuda.Event()
                        ready.record(s_in)
                    compute.wait_event(ready)
                    if y_free[i] is not None:
                        compute.wait_event(y_free[i])
                    csr.spmm(x_view.data_ptr(), code, L, L, 1, 0, y_view.data_ptr(), L, 0,
                             job.mode_code, job.thr, kernel=kernel, stream=compute.cuda_stream,
                             y_f32=job.y_f32)
                    done = torch.cuda.Event()
                    done.record(compute)
                    x_free[i] = done
                    s_out.wait_event(done)
                    if ring_out is not None:
                        drain(i)             # the slice parked in ring_out[i] has reached its result
                    dst_host = job.out_t[b]
                    with torch.cuda.stream(s_out):
                        if direct:
                            dst_host.view(-1).view(torch.uint8).copy_(y_view, non_blocking=True)
                        else:
                            ring_out[i][:slice_out].copy_(y_view, non_blocking=True)
                        fin = torch.cuda.Event()
                        fin.record(s_out)
                    y_free[i] = fin
                    if not direct:
                        pending[i] = _copy_pool().submit(copy_out, i, dst_host.data_ptr(), slice_out, fin)
            trace.mark('enqueue')
        finally:
            # whatever happened above, nothing may still read the shared pinned rings or write a
            # result when the lock is released: drain the helper thread and the three streams
            for i in range(nbuf):
                try:
                    drain(i)
                except Exception:      # noqa: BLE001 - the original exception matters more
                    pass
            for t in xd + yd:            # the side streams still use these buffers
                t.record_stream(s_in)
                t.record_stream(s_out)
            s_in.synchronize()
            compute.synchronize()
            s_out.synchronize()
        trace.mark('drain')
    trace.report(f'streamed remap: {len(jobs)} variable(s), {total} slice(s), rows_copied={n_x}')
