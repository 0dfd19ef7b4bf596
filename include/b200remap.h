/* b200remap.h -- C ABI of libb200remap.so: pyremap's weight-application hot path
 * on NVIDIA B200 (sm_100a).
 *
 * pyremap (reference tree /root/reference, v2.4.0) has no FFI layer of its own:
 * the native code its hot path executes is scipy's
 *     csr_matvecs(n_row, n_col, n_vecs, Ap, Aj, Ax, Xx, Yx)
 * reached from `matrix.dot(...)` at pyremap/remapper/remap_numpy.py:264,265,268,
 * wrapped by ~10 NumPy passes (remap_numpy.py:256-278).  The entry points below are
 * what a ctypes binding inside pyremap/remapper/remap_numpy.py would bind to replace
 * exactly that (see INTEGRATION.md for the stub):
 *
 *   b200remap_csr_create   <- scipy.sparse.csr_matrix(...) result of _load_mapping,
 *                             remap_numpy.py:134-137 (+ frac_b read at :270)
 *   b200remap_spmm         <- matrix.dot(...) x1 or x2 plus remap_numpy.py:258-278
 *                             (mask product, threshold, divide, NaN fill), fused
 *   b200remap_any_nan      <- np.isnan(field) / np.count_nonzero(mask) > 0,
 *                             remap_numpy.py:202-204 (branch selection)
 *   b200remap_transpose, b200remap_transpose_ld, b200remap_permute
 *                          <- in_field.transpose(...).reshape(...) / np.transpose,
 *                             remap_numpy.py:256 and :295 (layout only)
 *   b200remap_coo_to_csr   <- csr_matrix((S, (row, col)), shape=(n_b, n_a)) itself,
 *                             remap_numpy.py:134-137, for large maps
 *   b200remap_host_any_nan, b200remap_copy_runs, b200remap_host_pack_runs,
 *   b200remap_gather_rows  <- `field = da.values` reaching the device, remap_numpy.py:201:
 *                             only the source rows the map touches travel
 *
 * Conventions
 *   - plain C: pointers and sizes only; no exceptions or aborts cross the boundary.
 *   - return value 0 = success; < 0 = library error (B200REMAP_E_*); > 0 = cudaError_t.
 *     b200remap_last_error() returns a thread-local description of the last failure.
 *   - X / Y / flags are caller-owned DEVICE pointers on the handle's device; launches
 *     are asynchronous on `cuda_stream` (a cudaStream_t / CUstream, NULL = default).
 *   - a handle is immutable after create: concurrent calls on different streams are safe.
 *   - there is no CPU fallback: without a usable CUDA device every compute entry fails.
 */
#ifndef B200REMAP_H_
#define B200REMAP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REMAP_ABI_VERSION 1

#if defined(__GNUC__)
#define B200REMAP_API __attribute__((visibility("default")))
#else
#define B200REMAP_API
#endif

/* library error codes (negative) */
#define B200REMAP_E_INVALID   (-1)  /* bad argument (null, negative size, misaligned, ...) */
#define B200REMAP_E_UNSUPPORTED (-2) /* valid request this build cannot serve */
#define B200REMAP_E_NODEVICE  (-3)  /* no CUDA device / wrong architecture */
#define B200REMAP_E_NOMEM     (-4)  /* host allocation failed */

/* element type of the field X (the output Y is always float64, like the reference,
 * whose S is float64: remap_numpy.py:136) */
#define B200REMAP_F64 0
#define B200REMAP_F32 1

/* what b200remap_spmm computes per output element (i = destination row, k = column) */
#define B200REMAP_MODE_RAW    0  /* Y = S@X                                   (csr_matvecs) */
#define B200REMAP_MODE_FRACB  1  /* unmasked branch, remap_numpy.py:268-278:
                                    Y = frac_b[i] > 0 ? (S@X)/frac_b[i] : NaN            */
#define B200REMAP_MODE_MASKED 2  /* masked branch, remap_numpy.py:263-266,277-278:
                                    v = valid(x); num = S@(v ? x : 0.0); den = S@(v ? 1.0 : 0.0)
                                    Y = den > threshold ? num/den : NaN
                                    valid(x) = explicit byte mask if given, else !isnan(x) */

/* kernel selection (0 = let the library choose: rows of at most 8 entries on average -> WROW,
 * or LANES_K for thin fields (<= 32 bytes per row) and small launches (< 48 MB of gathers);
 * longer rows -> SELL;
 * b200remap_auto_kernel reports the choice) */
#define B200REMAP_KERNEL_AUTO     0
#define B200REMAP_KERNEL_LANES_K  1  /* lanes across K on the plain CSR, 4-deep gather loop   */
/* 2..6: selectors of experiments (shared-memory staged kernels for long rows, a non-persistent
 * binned kernel, TMA / cp.async staged pipelines, persistent binned CTAs); they lost to
 * LANES_K / WROW on B200 and were removed -- E_INVALID now */
#define B200REMAP_KERNEL_SELL     8  /* lanes across K on a sliced-ELL (SELL-32) copy of the
                                        entries: a warp that walks 32 rows reads entry j of all of
                                        them as one line; built for maps with > 8 entries per row */
#define B200REMAP_KERNEL_WROW_F32 9  /* WROW with float32 products and sums, never chosen by AUTO:
                                        float32 fields and float32 results only
                                        (b200remap_spmm_f32out, no explicit mask).  The masked
                                        denominator -- all the keep decision depends on -- stays
                                        the exact float64 recurrence: NaN / mask placement is
                                        bit-identical to the reference, values agree within 1e-6
                                        relative (the north star's bar for float32 fields)      */
#define B200REMAP_KERNEL_WROW     7  /* warp tiles of the binned view, claimed dynamically in item
                                        order by persistent warps; no CTA barrier              */

typedef struct b200remap_csr b200remap_csr;

B200REMAP_API int b200remap_abi_version(void);
B200REMAP_API const char *b200remap_last_error(void);

/* number of visible CUDA devices and compute capability (major*10+minor) of `device` */
B200REMAP_API int b200remap_device_count(int *count);
B200REMAP_API int b200remap_device_arch(int device, int *sm);

/* Build the library-owned device copy of a CSR weight matrix in scipy's canonical form
 * (rows = destination cells, column indices sorted within a row, no duplicates), plus
 * the optional destination fraction `frac_b` [n_row] (required for MODE_FRACB).
 * Input pointers are HOST pointers unless ptrs_are_device != 0. */
B200REMAP_API int b200remap_csr_create(int device, int64_t n_row, int64_t n_col, int64_t nnz,
                         const int32_t *indptr, const int32_t *indices,
                         const double *data, const double *frac_b,
                         int ptrs_are_device, b200remap_csr **out);
B200REMAP_API void b200remap_csr_destroy(b200remap_csr *csr);

/* the selector B200REMAP_KERNEL_AUTO resolves to for this matrix (>= 1), or a negative error */
B200REMAP_API int b200remap_auto_kernel(const b200remap_csr *csr, int x_dtype, int64_t K);

/* info[0..7] = n_row, n_col, nnz, n_touched (distinct source rows referenced),
 *              max nnz per row, number of empty rows, device, has_frac_b */
B200REMAP_API int b200remap_csr_info(const b200remap_csr *csr, int64_t info[8]);

/* Y[b] = remap(S, X[b]) for b in [0, nbatch):
 *   X[b] = X + b*x_batch_stride, row-major [n_col, K] with leading dimension ldx (elements)
 *   Y[b] = Y + b*y_batch_stride, row-major [n_row, K] with leading dimension ldy, float64
 *   valid    (nullable) explicit validity bytes laid out exactly like X (MODE_MASKED only)
 *   keep_out (nullable) receives the keep flag of every output element, laid out like Y
 * Per row the stored entries are accumulated in stored order with a separately rounded
 * multiply and add starting from +0.0, i.e. bit-for-bit scipy's csr_matvecs. */
B200REMAP_API int b200remap_spmm(const b200remap_csr *csr, const void *X, int x_dtype, int64_t K,
                   int64_t ldx, int64_t nbatch, int64_t x_batch_stride,
                   const uint8_t *valid, double *Y, int64_t ldy,
                   int64_t y_batch_stride, uint8_t *keep_out, int mode,
                   double threshold, int kernel, void *cuda_stream);

/* The same product with a float32 result: every element is the float64 value b200remap_spmm
 * writes, rounded to nearest float32 (so it equals numpy's `.astype(float32)` of the reference's
 * result bit for bit; NaN placement unchanged).  Halves the output traffic for float32 workflows
 * (SURVEY §8f rank 3).  Y strides are in float32 elements. */
B200REMAP_API int b200remap_spmm_f32out(const b200remap_csr *csr, const void *X, int x_dtype,
                   int64_t K, int64_t ldx, int64_t nbatch, int64_t x_batch_stride,
                   const uint8_t *valid, float *Y, int64_t ldy, int64_t y_batch_stride,
                   uint8_t *keep_out, int mode, double threshold, int kernel,
                   void *cuda_stream);

/* *flag_dev (device int32) := 1 if any of the n elements of X is NaN, else 0.
 * Blocks stop reading as soon as a NaN has been seen anywhere. */
B200REMAP_API int b200remap_any_nan(const void *X, int x_dtype, int64_t n, int32_t *flag_dev,
                      void *cuda_stream);

/* The same test on a HOST buffer with `threads` CPU threads and early exit (*out = 0/1): used
 * when the field lives in host memory and only the source rows the map touches are copied
 * to the GPU -- the branch test of remap_numpy.py:202-204 is over the whole variable. */
B200REMAP_API int b200remap_host_any_nan(const void *X, int x_dtype, int64_t n, int threads,
                           int *out);

/* out[b][c][r] = in[b][r][c]  (batched 2-D transpose; rows x cols -> cols x rows),
 * used for field-major layouts either side of the product.  elem_size is 4 or 8. */
B200REMAP_API int b200remap_transpose(const void *in, void *out, int elem_size, int64_t nbatch,
                        int64_t rows, int64_t cols, void *cuda_stream);
/* the same with leading dimensions: in[b][r][c] at in + (b * rows + r) * ld_in + c, out[b][c][r] at
 * out + (b * cols + c) * ld_out + r (ld_in >= cols, ld_out >= rows).  Lets the batch axis of a
 * source-dims-last field, (time, lat, lon), become a K axis padded to a multiple of 4 -- the
 * 256-bit lanes of the gather (K = 365: 697 -> 392 us on the C2 map) -- without an extra copy. */
B200REMAP_API int b200remap_transpose_ld(const void *in, void *out, int elem_size, int64_t nbatch,
                                         int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out,
                                         void *cuda_stream);

/* out (C-contiguous, `shape`) [i0]...[i_{n-1}] = in[sum_d i_d * in_strides[d]]: a general axis
 * permutation on the device for the layouts the native batched launch does not cover -- remap
 * axes that are not adjacent (reference remap_numpy.py:236-256 going in, :280-295 coming out).
 * elem_size 1, 4 or 8 bytes; 1 <= ndim <= 8; strides in elements. */
B200REMAP_API int b200remap_permute(const void *in, void *out, int elem_size, int ndim,
                                    const int64_t *shape, const int64_t *in_strides,
                                    void *cuda_stream);

/* dst[i, :] = src[rows_dev[i], :] for i < n_rows, rows of row_bytes (multiple of 16) bytes;
 * src rows are src_row_bytes apart.  `src` is any device-accessible pointer -- in particular
 * pinned (mapped) host memory, which makes this the host->device transfer of exactly the
 * source rows the map touches (replaces the full-field copy of `da.values`,
 * remap_numpy.py:201, for regional maps). */
B200REMAP_API int b200remap_gather_rows(const void *src, void *dst, const int32_t *rows_dev,
                          int64_t n_rows, int64_t row_bytes, int64_t src_row_bytes,
                          void *cuda_stream);

/* n_runs independent copies dst + dst_off[i] <- src + src_off[i] of bytes[i] bytes (offsets and
 * sizes are host arrays, in bytes) enqueued on `cuda_stream` as ONE batched DMA submission
 * (cudaMemcpyBatchAsync; a loop of cudaMemcpyAsync when use_batch == 0 or the batch API refuses).
 * With `src` in pinned host memory this moves the contiguous runs of touched source rows
 * through the copy engines, which -- unlike SM loads from host memory -- run at full PCIe rate
 * while a device->host copy of results is in flight (replaces the full-field copy of
 * `da.values`, remap_numpy.py:201).  `cuda_stream` must be a real (non-legacy) stream. */
B200REMAP_API int b200remap_copy_runs(const void *src, void *dst, const int64_t *src_off,
                        const int64_t *dst_off, const int64_t *bytes, int64_t n_runs,
                        int use_batch, void *cuda_stream);

/* The same runs copied by CPU threads between two HOST buffers: packs the touched source-row
 * runs of a slice that lives in pageable memory into one pinned staging block, which then
 * crosses PCIe as a single DMA (pageable memory cannot be read by the copy engines directly;
 * a plain cudaMemcpy of the bridged runs moved 417 MB per C3 slice at ~10 GB/s). */
B200REMAP_API int b200remap_host_pack_runs(const void *src, void *dst, const int64_t *src_off,
                             const int64_t *dst_off, const int64_t *bytes, int64_t n_runs,
                             int threads);

/* Map loader on the GPU: 0-based COO triplets -> the canonical CSR that
 * `csr_matrix((S, (row, col)), shape=(n_row, n_col))` builds in _load_mapping
 * (remap_numpy.py:134-137): rows in order, columns sorted within a row, duplicates of one
 * (row, col) summed left to right in file order.  row/col/S are host pointers unless
 * ptrs_are_device; indptr_dev[n_row+1], indices_dev[n_s], data_dev[n_s] are caller-owned DEVICE
 * buffers (the first *nnz_out entries of indices/data are valid).  Synchronises the stream. */
B200REMAP_API int b200remap_coo_to_csr(int device, int64_t n_row, int64_t n_col, int64_t n_s,
                         const int32_t *row, const int32_t *col, const double *S,
                         int ptrs_are_device, int32_t *indptr_dev, int32_t *indices_dev,
                         double *data_dev, int64_t *nnz_out, void *cuda_stream);

/* diagnostic: q[i] = a[i] / b[i] (device pointers) through the library's shared-reciprocal
 * division, which must equal IEEE-754 division bit for bit (pinned by the test-suite) */
B200REMAP_API int b200remap_debug_divide(const double *a, const double *b, double *q, int64_t n,
                           void *cuda_stream);
/* the same through the branch-free division of the masked epilogue (what MODE_MASKED runs for
 * num / den); every denominator is treated as kept, so q[i] must equal a[i] / b[i] for finite
 * positive b[i] */
B200REMAP_API int b200remap_debug_divide_masked(const double *a, const double *b, double *q,
                                  int64_t n, void *cuda_stream);

/* tuning knobs for experiments (process-wide; 0 restores the default):
 *   0: LANES_K: target threads per CTA (32..384, default 160); WROW: 3..6 = 4..32 lanes per row
 *   1: host_pack_runs: 1 = non-temporal stores instead of memcpy
 *   2: 1 = b200remap_csr_create builds the sliced-ELL view for every map (default: long rows only)
 *   3: cap on the vector width (1, 2, 4)
 *   4: binning segment length in units of 8 rows (read by b200remap_csr_create; default 256)
 *   7: WROW: resident warps per SM (default: occupancy limit, 24)
 *   8: WROW: 1 = static round-robin schedule instead of dynamic in-order claiming
 *  12: WROW: slices per sweep of the tiles inside a launch (default 4)
 *  14: WROW: warps per CTA (1, 2, 4; default 4)      15: WROW: shared-memory carve-out in percent */
B200REMAP_API int b200remap_set_tunable(int which, int value);

#ifdef __cplusplus
}
#endif
#endif /* B200REMAP_H_ */
