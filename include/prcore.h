/*
 * prcore.h -- C ABI of libprcore.so, the B200 (sm_100a) passive-radar DSP core.
 *
 * This is the drop-in boundary for the hot path of Max-Manning/passiveRadar.
 * The reference has no FFI of its own (it is pure Python), so each entry point
 * below names the *Python function* whose arithmetic it replaces; the Python
 * wrappers in passiveradar_b200/ keep those functions' signatures and bind these
 * symbols through ctypes (see INTEGRATION.md for the stub a maintainer of the
 * reference would add).
 *
 * Conventions
 *   - complex64 data are interleaved (re, im) float pairs: `prc_c64`.
 *   - every function returns PRC_OK (0) or a negative PRC_E_* code; the message
 *     for the calling thread is available from prc_last_error().
 *   - `mem_kind` says where the data pointers live:
 *         PRC_MEM_HOST   (0)  host memory; the library copies H2D / D2H itself
 *         PRC_MEM_DEVICE (1)  device memory on `device`; no copies are made
 *   - `stream` is a cudaStream_t (as void*).  NULL: the library uses a private
 *     stream owned by the calling thread and the call is synchronous.  Non-NULL:
 *     all work is enqueued on that stream; unless PRC_FLAG_ASYNC is set the call
 *     waits for it to finish before returning.
 *   - the library is re-entrant: dask's threaded scheduler calls the Python
 *     wrappers concurrently (main.py:169-194); ctypes drops the GIL, every
 *     (thread | stream) owns its workspace, errors are thread-local.
 *   - inputs are never modified; outputs are caller-allocated.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with PRC_E_CUDA.
 */
#ifndef PRCORE_H
#define PRCORE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct prc_c64 { float re, im; } prc_c64;

#define PRC_VERSION 100          /* 0.1.0 */

/* status codes */
#define PRC_OK            0
#define PRC_E_INVALID    -1      /* bad argument (message says which) */
#define PRC_E_CUDA       -2      /* CUDA runtime error / no device */
#define PRC_E_SINGULAR   -3      /* LS normal equations not positive definite */
#define PRC_E_NOMEM      -4

/* mem_kind */
#define PRC_MEM_HOST      0
#define PRC_MEM_DEVICE    1

/* flags */
#define PRC_FLAG_ASYNC        1u   /* do not synchronise `stream` before returning */
#define PRC_FLAG_WINDOW_F32   2u   /* `window` points at float, not double */

/* ---- lifecycle / diagnostics ------------------------------------------------ */
int         prc_version(void);
const char* prc_last_error(void);                /* thread-local, never NULL */
int         prc_device_count(int* count);
int         prc_init(int device);                /* optional: create the context early */
void        prc_shutdown(void);                  /* free every workspace and private stream */
int         prc_sync(int device, void* stream);  /* cudaStreamSynchronize on `stream` (or the thread's) */
uint64_t    prc_launch_count(void);              /* kernels launched by this library so far */

/* per-kernel device timing (CUDA events on the launching stream around every launch).
 * enable -> run -> collect (synchronises) -> read(idx) for idx < prc_profile_kernels() */
int prc_profile_enable(int on);
int prc_profile_collect(void);
int prc_profile_kernels(void);
int prc_profile_read(int idx, const char** name, double* total_ms, uint64_t* launches);
int prc_profile_reset(void);

/* pinned host memory helpers (so numpy buffers can take the async copy path) */
int prc_host_alloc(void** ptr, uint64_t bytes);
int prc_host_free(void* ptr);
int prc_host_register(void* ptr, uint64_t bytes);
int prc_host_unregister(void* ptr);

/* ---- cross-ambiguity function ---------------------------------------------------
 * Replaces fast_xambg(), reference passiveRadar/range_doppler_processing.py:12-90,
 * for inputs already padded to their final length `n` (the zero-padding of :52-55,
 * named windows of :57-58 and the (F, R+1, 1) reshape stay in the Python wrapper).
 *
 *   out[f*(range_bins+1) + k],  f = 0..freq_bins-1 (fftshift-ed Doppler),
 *                               k = 0..range_bins   (delay range_bins - k)
 *
 * window : n doubles (or floats with PRC_FLAG_WINDOW_F32) or NULL      (:83-84)
 * dtaps  : decimator taps (doubles) or NULL for the reference's default boxcar of
 *          int(n/freq_bins)+1 ones (:69-72); ndtaps = their count         (:73-78)
 */
int prc_xambg_c64(const prc_c64* ref, const prc_c64* srv, int64_t n,
                  int range_bins, int freq_bins,
                  const void* window, const double* dtaps, int64_t ndtaps,
                  prc_c64* out, int mem_kind, int device, void* stream, unsigned flags);

/* ---- block least-squares clutter filter ---------------------------------------------
 * Replaces LS_Filter(), reference passiveRadar/clutter_removal.py:6-56.
 *   out  : n cleaned surveillance samples
 *   taps : filter_len + peek taps, or NULL                                    (:53-56)
 * The Gram matrix of :39 is Hermitian Toeplitz (circular autocorrelation of ref);
 * the library computes its first column and the right-hand side of :45 as lag
 * correlations, solves (T + reg I) w = rhs in float64 and applies the circular FIR
 * of :51.
 */
int prc_ls_filter_c64(const prc_c64* ref, const prc_c64* srv, int64_t n,
                      int filter_len, int peek, float reg,
                      prc_c64* out, prc_c64* taps,
                      int mem_kind, int device, void* stream, unsigned flags);

/* ---- Toeplitz least-squares clutter filter (what main.py actually calls) ---------------------------
 * prc_ls_toeplitz_c64 replaces LS_Filter_Toeplitz(), reference passiveRadar/clutter_removal.py:109-160:
 * reference rolled by -peek (:139), LINEAR (zero-padded) auto/cross correlations (:142-147), Levinson
 * solve without regularisation (:150, float64 here as in SciPy), linear convolution + subtraction
 * (:153-155).  prc_ls_multiple_c64 replaces LS_Filter_Multiple() (:162-187): the filter is applied once
 * per Doppler bin on the running residual, the reference being frequency-shifted by bin Hz first
 * (signal_utils.py:24-27, float32 phase ramp as in the reference).  doppler_bins == NULL means [0].
 * Results are complex64; the Python wrappers widen them to complex128 like the reference returns.
 */
int prc_ls_toeplitz_c64(const prc_c64* ref, const prc_c64* srv, int64_t n, int filter_len, int peek,
                        prc_c64* out, prc_c64* taps, int mem_kind, int device, void* stream, unsigned flags);
int prc_ls_multiple_c64(const prc_c64* ref, const prc_c64* srv, int64_t n, int filter_len, int peek,
                        double sample_rate, const double* doppler_bins, int nbins,
                        prc_c64* out, prc_c64* taps_last,
                        int mem_kind, int device, void* stream, unsigned flags);

/* ---- NLMS / block-NLMS clutter filter ---------------------------------------------------
 * Replaces NLMS_filter(), reference passiveRadar/clutter_removal.py:189-249
 * (block_len == 1) and provides block_NLMS (block_len > 1; not in the reference,
 * defined in DESIGN.md).
 *   init_taps : filter_len + peek initial taps or NULL (zeros)              (:218-225)
 *   out       : n samples; out[0:filter_len] and out[n-peek:] are zero        (:231,:244)
 *   taps_out  : final taps or NULL                                            (:246-249)
 */
int prc_nlms_c64(const prc_c64* ref, const prc_c64* srv, int64_t n,
                 int filter_len, int peek, float mu, int block_len,
                 const prc_c64* init_taps, prc_c64* out, prc_c64* taps_out,
                 int mem_kind, int device, void* stream, unsigned flags);

/* ---- one CPI frame: LS_Filter -> fast_xambg with filterLen = range_bins-like chaining ----
 * The composition main.py performs per chunk (main.py:169-194 with LS_Filter in place
 * of LS_Filter_Multiple): cleaned = LS_Filter(ref, srv, filter_len, reg, peek);
 * map = fast_xambg(ref, cleaned, range_bins, freq_bins, n, window).
 * The cleaned surveillance channel stays in device memory.
 *   taps_out, cleaned_out : optional (NULL to skip)
 */
int prc_frame_c64(const prc_c64* ref, const prc_c64* srv, int64_t n,
                  int filter_len, int peek, float reg,
                  int range_bins, int freq_bins, const void* window,
                  prc_c64* out_map, prc_c64* taps_out, prc_c64* cleaned_out,
                  int mem_kind, int device, void* stream, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* PRCORE_H */
