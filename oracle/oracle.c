/*
 * oracle.c -- CPU restatement of the reference's FIR / SOS-IIR / integer-rate-change
 * arithmetic.  TEST INFRASTRUCTURE ONLY: nothing in the product path
 * (scikit-dsp-comm_b200/) may link, import or execute this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * What it restates (reference = mwickert/scikit-dsp-comm @ /root/reference):
 *   - multirate_FIR.filter/.up/.dn   src/sk_dsp_comm/multirate_helper.py:104-127
 *       -> scipy.signal.lfilter(b,[1],x), FIR branch == np.convolve(b,x)[:len(x)]
 *          (scipy 1.18.1 scipy/signal/_signaltools.py lfilter, len(a)==1 branch)
 *   - multirate_IIR.filter/.up/.dn   src/sk_dsp_comm/multirate_helper.py:169-192
 *       -> scipy.signal.sosfilt(sos,x): direct-form-II-transposed biquad cascade,
 *          sample-outer / section-inner loop, zero initial state
 *          (scipy _sosfilt.pyx; the loop body is quoted in SURVEY.md section 3.3)
 *   - sigsys.upsample / downsample   src/sk_dsp_comm/sigsys.py:3031-3083
 *
 * All arithmetic is IEEE double, sequential accumulation in tap order.  The FIR
 * loops are OpenMP-parallel over output samples (each output is an independent
 * dot product, so threading does not change any result bit); the SOS loop is
 * inherently sequential in n and stays single-threaded like the reference.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* y[n] = sum_{k<K} b[k] * xe[n-k],  xe = [hist (K-1 samples) ; x], hist==NULL -> zeros.
 * nch = 1 (real) or 2 (interleaved complex with real taps).
 * multirate_helper.py:108 -> lfilter FIR branch. */
void oracle_fir_f64(const double *b, int32_t K, const double *x, const double *hist,
                    int64_t n, int32_t nch, double *y)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int c = 0; c < nch; ++c) {
            double acc = 0.0;
            for (int32_t k = 0; k < K; ++k) {
                int64_t j = i - k;
                double v;
                if (j >= 0) v = x[j * nch + c];
                else if (hist) v = hist[((int64_t)(K - 1) + j) * nch + c];
                else v = 0.0;
                acc += b[k] * v;
            }
            y[i * nch + c] = acc;
        }
    }
}

/* .up(x, L): y = lfilter(b,[1], L*upsample(x,L))  (multirate_helper.py:116-117),
 * evaluated through the polyphase identity y[L*m+r] = L * sum_q b[L*q+r] * x[m-q]
 * (identical because the skipped products are exact zeros).
 * Output length L*n. */
void oracle_fir_up_f64(const double *b, int32_t K, const double *x, int64_t n,
                       int32_t nch, int32_t L, double *y)
{
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; ++m) {
        for (int32_t r = 0; r < L; ++r) {
            for (int c = 0; c < nch; ++c) {
                double acc = 0.0;
                for (int32_t k = r; k < K; k += L) {
                    int64_t j = m - (k - r) / L;
                    if (j < 0) break;
                    acc += b[k] * ((double)L * x[j * nch + c]);
                }
                y[(m * L + r) * nch + c] = acc;
            }
        }
    }
}

/* .dn(x, M): y = downsample(lfilter(b,[1],x), M)  (multirate_helper.py:125-126),
 * y[m] = sum_k b[k] x[M*m-k], m < floor(n/M). */
void oracle_fir_dn_f64(const double *b, int32_t K, const double *x, int64_t n,
                       int32_t nch, int32_t M, double *y)
{
    int64_t nout = n / M;
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < nout; ++m) {
        for (int c = 0; c < nch; ++c) {
            double acc = 0.0;
            for (int32_t k = 0; k < K; ++k) {
                int64_t j = m * M - k;
                if (j < 0) break;
                acc += b[k] * x[j * nch + c];
            }
            y[m * nch + c] = acc;
        }
    }
}

/* scipy.signal.sosfilt restated:
 *   for n: x_cur = x[n]
 *     for s: x_new = b0*x_cur + z[s][0]
 *            z[s][0] = b1*x_cur - a1*x_new + z[s][1]
 *            z[s][1] = b2*x_cur - a2*x_new
 *            x_cur = x_new
 *     y[n] = x_cur
 * sos rows are [b0 b1 b2 a0(=1) a1 a2]; zi/zf are nsec*2*nch doubles laid out
 * [s][0..1][c] (may be NULL: zero initial state / final state discarded).
 * Called from multirate_helper.py:173,182,190. */
void oracle_sosfilt_f64(const double *sos, int32_t nsec, const double *x, int64_t n,
                        int32_t nch, const double *zi, double *zf, double *y)
{
    double z[64][2][2];
    if (nsec > 64) return;
    for (int s = 0; s < nsec; ++s)
        for (int q = 0; q < 2; ++q)
            for (int c = 0; c < nch; ++c)
                z[s][q][c] = zi ? zi[(s * 2 + q) * nch + c] : 0.0;
    for (int64_t i = 0; i < n; ++i) {
        for (int c = 0; c < nch; ++c) {
            double x_cur = x[i * nch + c];
            for (int s = 0; s < nsec; ++s) {
                const double *q = sos + 6 * s;
                double x_new = q[0] * x_cur + z[s][0][c];
                z[s][0][c] = q[1] * x_cur - q[4] * x_new + z[s][1][c];
                z[s][1][c] = q[2] * x_cur - q[5] * x_new;
                x_cur = x_new;
            }
            y[i * nch + c] = x_cur;
        }
    }
    if (zf)
        for (int s = 0; s < nsec; ++s)
            for (int q = 0; q < 2; ++q)
                for (int c = 0; c < nch; ++c)
                    zf[(s * 2 + q) * nch + c] = z[s][q][c];
}

/* sigsys.upsample (sigsys.py:3050-3052): y[n*L] = x[n], zeros elsewhere.  esz = element bytes. */
void oracle_upsample_bytes(const void *x, int64_t n, int32_t L, int32_t esz, void *y)
{
    memset(y, 0, (size_t)n * L * esz);
    for (int64_t i = 0; i < n; ++i)
        memcpy((char *)y + (size_t)i * L * esz, (const char *)x + (size_t)i * esz, esz);
}

/* sigsys.downsample (sigsys.py:3080-3082): y[m] = x[m*M+p], m < floor(n/M). */
void oracle_downsample_bytes(const void *x, int64_t n, int32_t M, int32_t p, int32_t esz, void *y)
{
    int64_t nout = n / M;
    for (int64_t m = 0; m < nout; ++m)
        memcpy((char *)y + (size_t)m * esz, (const char *)x + (size_t)(m * M + p) * esz, esz);
}

/* scipy.signal.lfilter with a denominator: direct form II transposed,
 *   y[n] = b0 x[n] + z0;  z_i = b_{i+1} x[n] - a_{i+1} y[n] + z_{i+1}   (a normalised by a0),
 * zero initial state.  Called by rate_change.up/.dn (multirate_helper.py:73-74,81-82) and
 * sigsys.interp24/deci24 (sigsys.py:2971-3027).  b and a are padded to the same length nc. */
void oracle_lfilter_ba_f64(const double *b, const double *a, int32_t nc, const double *x, int64_t n, double *y)
{
    double z[64];
    if (nc > 64 || nc < 1) return;
    for (int i = 0; i < nc; ++i) z[i] = 0.0;
    const double a0 = a[0];
    for (int64_t i = 0; i < n; ++i) {
        const double xi = x[i];
        /* same operation order as scipy's _linear_filter inner loop */
        const double yi = z[0] + (b[0] / a0) * xi;
        for (int k = 1; k < nc; ++k) {
            if (k + 1 < nc) z[k - 1] = (z[k] + (b[k] / a0) * xi) - (a[k] / a0) * yi;
            else z[k - 1] = (b[k] / a0) * xi - (a[k] / a0) * yi;
        }
        y[i] = yi;
    }
}
