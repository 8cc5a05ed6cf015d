/*
 * b200dsp.h -- C ABI of the B200-native FIR / SOS-IIR / integer multirate engine.
 *
 * This is the drop-in boundary for the hot path of mwickert/scikit-dsp-comm
 * (SURVEY.md section 8b).  The reference has no FFI of its own -- its boundary is the
 * Python surface in src/sk_dsp_comm/multirate_helper.py:85-192 and
 * src/sk_dsp_comm/sigsys.py:3031-3083, whose arithmetic is delegated to
 * scipy.signal.lfilter / sosfilt.  Each entry point below names the reference call it
 * replaces.  INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, opaque plan handles.  No C++/torch types.
 *   - every `x`, `y`, `hist`, `zi`, `zf`, `ws` pointer is a DEVICE pointer on the current
 *     CUDA device; `*_host` pointers are host pointers read synchronously during the call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work
 *     is enqueued asynchronously on it; nothing is allocated or synchronised inside the
 *     filter calls: plans own their coefficient buffers and every kernel table (tensor-core tap
 *     matrices, scan matrices), all built in the *_plan_create call, and are read-only afterwards
 *     -- one plan may be used from several threads and inside CUDA-graph capture.  (The only host
 *     state a filter call may add is a mutex-guarded cache of scan-matrix powers the first time an
 *     SOS plan sees a new call length on the float64 / complex path; no device allocation.)
 *   - return value: 0 = OK, negative = error (B200DSP_E_*); b200dsp_last_error() gives the
 *     thread-local message.  No exception ever crosses this boundary.
 *   - sample counts are int64_t (a 2^31-sample stream does not fit int32).
 *   - sm_100a only.  There is no CPU fallback and no other GPU back-end.
 */
#ifndef B200DSP_H
#define B200DSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DSP_VERSION 200 /* 0.2.0 */

/* sample dtypes (complex = interleaved re,im; filter coefficients are always real) */
enum {
    B200DSP_F32 = 0,  /* float            */
    B200DSP_F64 = 1,  /* double           */
    B200DSP_C64 = 2,  /* float2  (re,im)  */
    B200DSP_C128 = 3  /* double2 (re,im)  */
};

enum {
    B200DSP_OK = 0,
    B200DSP_E_BADARG = -1,      /* NULL pointer, non-positive factor, ...             */
    B200DSP_E_DTYPE = -2,       /* dtype code not in the enum above                   */
    B200DSP_E_UNSUPPORTED = -3, /* e.g. filter too long for on-chip staging           */
    B200DSP_E_CUDA = -4,        /* a CUDA runtime call or launch failed               */
    B200DSP_E_WORKSPACE = -5    /* workspace missing or too small                     */
};

int b200dsp_version(void);
const char *b200dsp_last_error(void);
/* SM count / compute capability of the current device (grid sizing, diagnostics). */
int b200dsp_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------ FIR ----------------
 * A FIR plan holds the taps of one multirate_FIR object (multirate_helper.py:95-101,
 * `self.b`) on the device in the layouts the kernels want.  taps_host: ntaps doubles. */
typedef struct b200dsp_fir_plan b200dsp_fir_plan;
int b200dsp_fir_plan_create(const double *taps_host, int32_t ntaps, b200dsp_fir_plan **plan);
void b200dsp_fir_plan_destroy(b200dsp_fir_plan *plan);
int32_t b200dsp_fir_plan_ntaps(const b200dsp_fir_plan *plan);

/* y[i] = sum_k b[k] * xe[i-k], i in [0,n); xe = [hist ; x].
 * Replaces multirate_FIR.filter -> signal.lfilter(self.b,[1],x)  (multirate_helper.py:104-109).
 * hist: the ntaps-1 samples preceding x (overlap-save halo of a sharded stream),
 *       or NULL for the reference's zero initial state.  lfilter's zi/zf state is expressed through
 *       hist as well: zf = this call on ntaps-1 zero samples with hist = the block's tail.
 * Also replaces sigsys.os_filter / oa_filter (sigsys.py:482-598), which compute the same output
 * through FFT frames -- as this entry does itself for float32 / complex64 filters of 257..2049 taps
 * (overlap-save, 4096-point frames, csrc/fir_fft.cu). */
int b200dsp_fir_filter(const b200dsp_fir_plan *plan, int dtype, const void *x, const void *hist,
                       void *y, int64_t n, void *stream);

/* The same for `rows` independent streams of n samples each (row r at x + r*x_row_stride samples, output at
 * y + r*y_row_stride), zero initial state per row: scipy.signal.lfilter filters the LAST axis of an N-D array,
 * which is what multirate_FIR.filter hands it (multirate_helper.py:108).  One C call; the rows are queued back to
 * back on `stream`. */
int b200dsp_fir_filter_batch(const b200dsp_fir_plan *plan, int dtype, const void *x, void *y, int64_t rows,
                             int64_t n, int64_t x_row_stride, int64_t y_row_stride, void *stream);

/* y[L*m+r] = L * sum_q b[L*q+r] * xe[m-q], m in [0,n): n*L outputs.
 * Replaces multirate_FIR.up -> lfilter(b,[1], L*upsample(x,L))  (multirate_helper.py:112-118)
 * without materialising the zero-stuffed stream.  hist: ceil((ntaps-1)/L) preceding input
 * samples or NULL.  Use b200dsp_fir_up_hist_len() for the exact count.
 * The same entry serves the reference's pulse-shaping call sites lfilter(b,1,upsample(x,ns))
 * (digitalcom.py:488,666,1047,1676,1821; sigsys.py:2149,2201) with a plan holding b/ns. */
int b200dsp_fir_up(const b200dsp_fir_plan *plan, int dtype, const void *x, const void *hist,
                   void *y, int64_t n, int32_t L, void *stream);
int32_t b200dsp_fir_up_hist_len(const b200dsp_fir_plan *plan, int32_t L);

/* y[m] = sum_k b[k] * xe[M*m-k], m in [0, floor(n/M)).
 * Replaces multirate_FIR.dn -> downsample(lfilter(b,[1],x), M)  (multirate_helper.py:121-127); only the kept
 * outputs are stored (and, on the polyphase kernels, computed).  hist: ntaps-1 preceding samples or NULL. */
int b200dsp_fir_dn(const b200dsp_fir_plan *plan, int dtype, const void *x, const void *hist,
                   void *y, int64_t n, int32_t M, void *stream);

/* ------------------------------------------------------------------ SOS IIR ------------
 * A SOS plan holds the biquad cascade of one multirate_IIR object (multirate_helper.py:159-166,
 * `self.sos`): sos_host = nsec rows [b0 b1 b2 a0 a1 a2] (row-major doubles, a0 must be 1 --
 * scipy's _validate_sos rule) plus the host-precomputed chunk-transition matrices that the
 * parallel-prefix kernels use. */
typedef struct b200dsp_sos_plan b200dsp_sos_plan;
int b200dsp_sos_plan_create(const double *sos_host, int32_t nsec, b200dsp_sos_plan **plan);
void b200dsp_sos_plan_destroy(b200dsp_sos_plan *plan);
int32_t b200dsp_sos_plan_nsec(const b200dsp_sos_plan *plan);

/* Bytes of device scratch b200dsp_sos_filter needs for an n-sample call of this dtype / L / M. */
size_t b200dsp_sos_workspace_bytes(const b200dsp_sos_plan *plan, int dtype, int64_t n, int32_t L, int32_t M);

/* Biquad cascade in direct-form-II-transposed arithmetic, run as a parallel-prefix
 * recurrence (chunked zero-state pass -> scan of chunk carries -> corrected pass).
 *   L == 1, M == 1 : y = sosfilt(sos, x)                         multirate_helper.py:169-174
 *   L  > 1         : y = sosfilt(sos, L*upsample(x,L)), n*L outputs          (:177-183)
 *   M  > 1         : y = downsample(sosfilt(sos,x), M), floor(n/M) outputs   (:186-192)
 *                    (long calls stage the full-rate stream once in the workspace; an IIR runs at the
 *                    full rate in any case)
 * zi: initial state, 2*nsec values per channel laid out [section][2][channel] in the
 *     dtype's real scalar type (scipy's zi layout), or NULL = zeros (the reference never
 *     passes zi).  zf: final state written in the same layout, or NULL.
 * ws / ws_bytes: scratch of at least b200dsp_sos_workspace_bytes(plan, dtype, n, L, M). */
int b200dsp_sos_filter(const b200dsp_sos_plan *plan, int dtype, const void *x, void *y, int64_t n,
                       int32_t L, int32_t M, const void *zi, void *zf, void *ws, size_t ws_bytes,
                       void *stream);

/* `rows` independent streams (last axis of an N-D array, as sosfilt filters it), L = M = 1, zero initial state per
 * row; ws as for a single n-sample call (the rows share it). */
int b200dsp_sos_filter_batch(const b200dsp_sos_plan *plan, int dtype, const void *x, void *y, int64_t rows, int64_t n,
                             int64_t x_row_stride, int64_t y_row_stride, void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------ rate change --------
 * y[i*L] = x[i], zeros elsewhere; n*L outputs.  Replaces sigsys.upsample (sigsys.py:3031-3053).
 * Pure index map: bit exact.  (The reference widens to >= float64; that dtype policy lives
 * in the Python shim, not here.) */
int b200dsp_upsample(int dtype, const void *x, void *y, int64_t n, int32_t L, void *stream);

/* y[m] = x[m*M+p], m in [0, floor(n/M)).  Replaces sigsys.downsample (sigsys.py:3056-3083).
 * Pure index map: bit exact. */
int b200dsp_downsample(int dtype, const void *x, void *y, int64_t n, int32_t M, int32_t p,
                       void *stream);

/* y = a + j*b, n samples.  dtype_in = dtype of a and b: real (F32/F64: y is the complex stream (a, b)) or complex
 * (C64/C128: y = (a.re - b.im, a.im + b.re)).  Second half of a complex-tap FIR: scipy.signal.lfilter accepts complex
 * b (multirate_helper.py:108 passes self.b through); by linearity lfilter(br + j*bi, 1, x) = lfilter(br,1,x) +
 * j*lfilter(bi,1,x), i.e. two real-tap plans and this combination. */
int b200dsp_combine_complex(int dtype_in, const void *a, const void *b, void *y, int64_t n, void *stream);

/* ------------------------------------------------------------------ diagnostics --------
 * Number of kernel launches issued through this library by the calling thread since the
 * last reset (bench.py's "gpu_launches" claim is read from here, not estimated). */
int64_t b200dsp_launch_count(void);
void b200dsp_launch_count_reset(void);

/* Select an implementation variant for the FIR filter kernels (tuning / A-B measurements):
 * 0 = default heuristic.  Other values are documented in DESIGN.md. */
void b200dsp_set_fir_variant(int variant);

/* Same for the SOS cascade: 0 = default heuristic, 1 = force the three-kernel scan path,
 * 2 = force the single-pass tensor-core kernel (float32) whenever the plan has its tables. */
void b200dsp_set_sos_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* B200DSP_H */
