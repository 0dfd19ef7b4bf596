"""ctypes binding of ``libb200remap.so`` (the C ABI declared in ``include/b200remap.h``).

There is deliberately no fallback: if the shared object cannot be built or
loaded, or no sm_100 device is present, every compute entry raises
:class:`B200RemapError`.  PyTorch tensors are used only as device-buffer carriers
(``data_ptr()``) and for the current CUDA stream.
"""

from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

MODE_RAW, MODE_FRACB, MODE_MASKED = 0, 1, 2
KERNEL_AUTO, KERNEL_LANES_K, KERNEL_WROW, KERNEL_SELL, KERNEL_WROW_F32 = 0, 1, 7, 8, 9
KERNEL_NAMES = {1: 'lanes_k_kernel', 7: 'wrow_kernel', 8: 'sell_kernel', 9: 'wrow_kernel'}
F64, F32 = 0, 1

#: every symbol ``include/b200remap.h`` declares
EXPORTED_SYMBOLS = (
    'b200remap_abi_version', 'b200remap_last_error', 'b200remap_device_count',
    'b200remap_device_arch', 'b200remap_csr_create', 'b200remap_csr_destroy',
    'b200remap_csr_info', 'b200remap_spmm', 'b200remap_any_nan',
    'b200remap_transpose', 'b200remap_set_tunable', 'b200remap_debug_divide',
    'b200remap_host_any_nan', 'b200remap_gather_rows', 'b200remap_copy_runs',
    'b200remap_spmm_f32out', 'b200remap_coo_to_csr', 'b200remap_host_pack_runs',
    'b200remap_auto_kernel', 'b200remap_debug_divide_masked', 'b200remap_permute',
    'b200remap_transpose_ld',
)


class B200RemapError(RuntimeError):
    """A call into libb200remap failed (``code`` < 0: library, > 0: cudaError_t)."""

    def __init__(self, code, message):
        super().__init__(f'libb200remap error {code}: {message}')
        self.code = code


_lib = None
_lock = threading.Lock()


def library_path():
    return _build.LIB_PATH


def load_library():
    """Load (building first if the sources are newer) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get('B200REMAP_LIB') or _build.LIB_PATH   # override: experiment builds
        if path == _build.LIB_PATH and _build.is_stale():
            try:
                _build.build_library()
            except Exception as exc:
                if not os.path.exists(path):
                    raise B200RemapError(
                        -3, f'{path} is missing and could not be built ({exc}); '
                        'pyremap_b200 has no CPU fallback') from exc
                import warnings
                warnings.warn(
                    f'{path} is older than its sources and could not be rebuilt ({exc}); '
                    'loading the stale library -- kernel fixes in the sources are NOT in effect',
                    RuntimeWarning, stacklevel=2)
        try:
            lib = ctypes.CDLL(path)
        except OSError as exc:
            raise B200RemapError(
                -3, f'cannot load {path}: {exc}; pyremap_b200 has no CPU '
                'fallback') from exc
        i32, i64, dbl, vp = (ctypes.c_int, ctypes.c_int64, ctypes.c_double,
                             ctypes.c_void_p)
        lib.b200remap_abi_version.argtypes = []
        lib.b200remap_abi_version.restype = i32
        lib.b200remap_last_error.argtypes = []
        lib.b200remap_last_error.restype = ctypes.c_char_p
        lib.b200remap_device_count.argtypes = [ctypes.POINTER(i32)]
        lib.b200remap_device_arch.argtypes = [i32, ctypes.POINTER(i32)]
        lib.b200remap_csr_create.argtypes = [i32, i64, i64, i64, vp, vp, vp, vp,
                                             i32, ctypes.POINTER(vp)]
        lib.b200remap_csr_destroy.argtypes = [vp]
        lib.b200remap_csr_destroy.restype = None
        lib.b200remap_csr_info.argtypes = [vp, ctypes.POINTER(i64)]
        lib.b200remap_auto_kernel.argtypes = [vp, i32, i64]
        lib.b200remap_auto_kernel.restype = i32
        lib.b200remap_spmm.argtypes = [vp, vp, i32, i64, i64, i64, i64, vp, vp,
                                       i64, i64, vp, i32, dbl, i32, vp]
        lib.b200remap_spmm_f32out.argtypes = lib.b200remap_spmm.argtypes
        lib.b200remap_coo_to_csr.argtypes = [i32, i64, i64, i64, vp, vp, vp, i32, vp, vp, vp,
                                             ctypes.POINTER(i64), vp]
        lib.b200remap_host_pack_runs.argtypes = [vp, vp, vp, vp, vp, i64, i32]
        lib.b200remap_any_nan.argtypes = [vp, i32, i64, vp, vp]
        lib.b200remap_transpose.argtypes = [vp, vp, i32, i64, i64, i64, vp]
        lib.b200remap_transpose_ld.argtypes = [vp, vp, i32, i64, i64, i64, i64, i64, vp]
        lib.b200remap_transpose_ld.restype = i32
        lib.b200remap_permute.argtypes = [vp, vp, i32, i32, ctypes.POINTER(i64),
                                          ctypes.POINTER(i64), vp]
        lib.b200remap_permute.restype = i32
        lib.b200remap_set_tunable.argtypes = [i32, i32]
        lib.b200remap_debug_divide.argtypes = [vp, vp, vp, i64, vp]
        lib.b200remap_debug_divide_masked.argtypes = [vp, vp, vp, i64, vp]
        lib.b200remap_debug_divide_masked.restype = i32
        lib.b200remap_host_any_nan.argtypes = [vp, i32, i64, i32, ctypes.POINTER(i32)]
        lib.b200remap_gather_rows.argtypes = [vp, vp, vp, i64, i64, i64, vp]
        lib.b200remap_copy_runs.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
        for name in ('b200remap_device_count', 'b200remap_device_arch',
                     'b200remap_csr_create', 'b200remap_csr_info',
                     'b200remap_spmm', 'b200remap_any_nan',
                     'b200remap_transpose', 'b200remap_set_tunable',
                     'b200remap_debug_divide', 'b200remap_host_any_nan',
                     'b200remap_gather_rows', 'b200remap_copy_runs',
                     'b200remap_spmm_f32out', 'b200remap_coo_to_csr',
                     'b200remap_host_pack_runs'):
            getattr(lib, name).restype = i32
        if lib.b200remap_abi_version() != 1:
            raise B200RemapError(-2, 'ABI version mismatch')
        # experiments: B200REMAP_TUNABLES="13=2,7=20" presets b200remap_set_tunable knobs
        for item in filter(None, os.environ.get('B200REMAP_TUNABLES', '').split(',')):
            which, _, value = item.partition('=')
            lib.b200remap_set_tunable(int(which), int(value))
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        msg = load_library().b200remap_last_error()
        raise B200RemapError(code, msg.decode('utf-8', 'replace') if msg else '')


def device_count():
    n = ctypes.c_int(0)
    code = load_library().b200remap_device_count(ctypes.byref(n))
    return n.value if code == 0 else 0


def set_tunable(which, value):
    check(load_library().b200remap_set_tunable(int(which), int(value)))


def _np_ptr(arr):
    return None if arr is None else ctypes.c_void_p(arr.ctypes.data)


class DeviceCSR:
    """Owner of one ``b200remap_csr`` handle (device copy of the weight matrix)."""

    def __init__(self, indptr, indices, data, frac_b, n_col, device=0):
        import numpy as np
        lib = load_library()
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.float64)
        if frac_b is not None:
            frac_b = np.ascontiguousarray(frac_b, dtype=np.float64)
            if frac_b.size != indptr.size - 1:
                raise ValueError('frac_b must have one entry per destination row')
        handle = ctypes.c_void_p()
        check(lib.b200remap_csr_create(
            int(device), indptr.size - 1, int(n_col), data.size, _np_ptr(indptr),
            _np_ptr(indices), _np_ptr(data), _np_ptr(frac_b), 0,
            ctypes.byref(handle)))
        self._handle = handle
        self._lib = lib
        info = (ctypes.c_int64 * 8)()
        check(lib.b200remap_csr_info(handle, info))
        (self.n_row, self.n_col, self.nnz, self.n_touched, self.max_row_nnz,
         self.n_empty_rows, self.device, has_frac) = [int(v) for v in info]
        self.has_frac_b = bool(has_frac)

    def auto_kernel(self, x_dtype=F64, K=80):
        """The kernel selector ``KERNEL_AUTO`` resolves to for this matrix and field row."""
        code = self._lib.b200remap_auto_kernel(self._handle, int(x_dtype), int(K))
        if code < 0:
            check(code)
        return code

    def close(self):
        if getattr(self, '_handle', None):
            self._lib.b200remap_csr_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def spmm(self, x_ptr, x_dtype, K, ldx, nbatch, x_batch_stride, y_ptr, ldy,
             y_batch_stride, mode, threshold=0.0, valid_ptr=None,
             keep_ptr=None, kernel=KERNEL_AUTO, stream=0, y_f32=False):
        """Raw pointer-level call of ``b200remap_spmm`` (asynchronous); ``y_f32``: the result
        buffer holds float32 (``b200remap_spmm_f32out``)."""
        if not self._handle:
            raise B200RemapError(-1, 'DeviceCSR is closed')
        fn = self._lib.b200remap_spmm_f32out if y_f32 else self._lib.b200remap_spmm
        check(fn(
            self._handle, ctypes.c_void_p(x_ptr), int(x_dtype), int(K), int(ldx),
            int(nbatch), int(x_batch_stride),
            ctypes.c_void_p(valid_ptr) if valid_ptr else None,
            ctypes.c_void_p(y_ptr), int(ldy), int(y_batch_stride),
            ctypes.c_void_p(keep_ptr) if keep_ptr else None, int(mode),
            float(threshold), int(kernel),
            ctypes.c_void_p(stream) if stream else None))


def any_nan(x_ptr, x_dtype, n, flag_ptr, stream=0):
    check(load_library().b200remap_any_nan(
        ctypes.c_void_p(x_ptr), int(x_dtype), int(n), ctypes.c_void_p(flag_ptr),
        ctypes.c_void_p(stream) if stream else None))


def transpose(in_ptr, out_ptr, elem_size, nbatch, rows, cols, stream=0, ld_in=None, ld_out=None):
    check(load_library().b200remap_transpose_ld(
        ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), int(elem_size),
        int(nbatch), int(rows), int(cols), int(cols if ld_in is None else ld_in),
        int(rows if ld_out is None else ld_out),
        ctypes.c_void_p(stream) if stream else None))


def permute(in_ptr, out_ptr, elem_size, shape, in_strides, stream=0):
    """``out`` (C-contiguous, ``shape``) <- ``in`` read with ``in_strides`` (elements)."""
    n = len(shape)
    arr = ctypes.c_int64 * n
    check(load_library().b200remap_permute(
        ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), int(elem_size), n,
        arr(*[int(v) for v in shape]), arr(*[int(v) for v in in_strides]),
        ctypes.c_void_p(stream) if stream else None))


def debug_divide(a_ptr, b_ptr, q_ptr, n, stream=0, masked=False):
    lib = load_library()
    fn = lib.b200remap_debug_divide_masked if masked else lib.b200remap_debug_divide
    check(fn(
        ctypes.c_void_p(a_ptr), ctypes.c_void_p(b_ptr), ctypes.c_void_p(q_ptr), int(n),
        ctypes.c_void_p(stream) if stream else None))


def host_any_nan(array, threads=None):
    """True iff the C-contiguous float32/float64 numpy array holds a NaN (native, early exit)."""
    import numpy as np
    a = array
    if a.dtype not in (np.float64, np.float32) or not a.flags.c_contiguous:
        raise ValueError('host_any_nan needs a C-contiguous float32/float64 array')
    if threads is None:
        threads = min(16, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity')
                      else (os.cpu_count() or 1))
    out = ctypes.c_int(0)
    check(load_library().b200remap_host_any_nan(
        ctypes.c_void_p(a.ctypes.data), F64 if a.dtype == np.float64 else F32, a.size,
        int(threads), ctypes.byref(out)))
    return bool(out.value)


def gather_rows(src_ptr, dst_ptr, rows_ptr, n_rows, row_bytes, src_row_bytes, stream=0):
    check(load_library().b200remap_gather_rows(
        ctypes.c_void_p(src_ptr), ctypes.c_void_p(dst_ptr), ctypes.c_void_p(rows_ptr),
        int(n_rows), int(row_bytes), int(src_row_bytes),
        ctypes.c_void_p(stream) if stream else None))


def copy_runs(src_ptr, dst_ptr, src_off, dst_off, nbytes, stream, use_batch=True):
    """Batched DMA of contiguous runs; ``src_off``/``dst_off``/``nbytes`` are int64 numpy arrays
    (bytes) that must stay alive until the call returns."""
    import numpy as np
    src_off = np.ascontiguousarray(src_off, dtype=np.int64)
    dst_off = np.ascontiguousarray(dst_off, dtype=np.int64)
    nbytes = np.ascontiguousarray(nbytes, dtype=np.int64)
    if not (src_off.size == dst_off.size == nbytes.size):
        raise ValueError('src_off, dst_off and nbytes must have the same length')
    check(load_library().b200remap_copy_runs(
        ctypes.c_void_p(src_ptr), ctypes.c_void_p(dst_ptr),
        ctypes.c_void_p(src_off.ctypes.data), ctypes.c_void_p(dst_off.ctypes.data),
        ctypes.c_void_p(nbytes.ctypes.data), int(src_off.size), 1 if use_batch else 0,
        ctypes.c_void_p(stream) if stream else None))



def coo_to_csr_device(device, n_row, n_col, row, col, S, indptr_ptr, indices_ptr, data_ptr, stream=0):
    """``b200remap_coo_to_csr`` on host triplets (int32 row/col 0-based, float64 S, numpy);
    the three output pointers are device buffers of n_row+1 / n_s / n_s elements.  Returns nnz."""
    nnz = ctypes.c_int64(0)
    check(load_library().b200remap_coo_to_csr(
        int(device), int(n_row), int(n_col), int(S.size),
        ctypes.c_void_p(row.ctypes.data), ctypes.c_void_p(col.ctypes.data),
        ctypes.c_void_p(S.ctypes.data), 0, ctypes.c_void_p(indptr_ptr),
        ctypes.c_void_p(indices_ptr), ctypes.c_void_p(data_ptr), ctypes.byref(nnz),
        ctypes.c_void_p(stream) if stream else None))
    return int(nnz.value)


def host_pack_runs(src_ptr, dst_ptr, src_off, dst_off, nbytes, threads=8):
    """CPU-thread packing of runs between two host buffers (offsets / sizes: int64 numpy, bytes);
    releases the GIL for the duration of the call."""
    check(load_library().b200remap_host_pack_runs(
        ctypes.c_void_p(src_ptr), ctypes.c_void_p(dst_ptr),
        ctypes.c_void_p(src_off.ctypes.data), ctypes.c_void_p(dst_off.ctypes.data),
        ctypes.c_void_p(nbytes.ctypes.data), int(src_off.size), int(threads)))
