// fir_direct.cu -- tiled direct-form / polyphase FIR on CUDA cores (sm_100a).
//
// One kernel template covers the three FIR entry points of the reference's multirate_FIR
// (src/sk_dsp_comm/multirate_helper.py:104-127):
//   filter : y[i]      = sum_k b[k]      x[i-k]                       (L = M = 1)
//   up(L)  : y[L m + r] = sum_q L b[Lq+r] x[m-q]        r in [0,L)     (polyphase interpolation)
//   dn(M)  : y[m]      = sum_c sum_q b[Mq+c] x[M(m-q)-c]  c in [0,M)   (polyphase decimation)
// In all three the inner work is "sum_q h[q] * xin[m-q]" over a unit-stride phase stream, so a
// block stages, per tile of TILE = NT*R consecutive m positions,
//   * the tap phases h_ph[q] (zero padded to KQ, a multiple of R) in shared memory,
//   * the M input phase streams xin_c[m] = x[M m - c] (one stream for filter / up), each with a
//     KQ-sample left apron, de-interleaved while they are loaded with coalesced reads,
// and every thread produces R consecutive outputs with a register sliding window: per tap
// it issues R (complex: 2R) FMAs against ONE new shared-memory sample, so the FMA pipe -- not
// the LSU / shared-memory port -- is the limiter (DESIGN.md "FIR kernel" has the arithmetic).
// Shared-memory rows are padded by one sample every R samples so that the R-strided
// per-lane window accesses are bank-conflict free.  Results leave through a shared-memory
// transpose so global stores are fully coalesced.
#include "common.cuh"
#include <vector>

namespace b200dsp {

template <typename S, typename C>
struct FirArgs {
    const S *x;
    const S *hist;
    S *y;
    const C *taps;     // ntaps raw taps (device)
    int64_t n_in;      // input samples
    int64_t n_m;       // m-domain length: filter n, up n, dn floor(n/M)
    int32_t ntaps;
    int32_t kq;        // taps per phase, padded to a multiple of R
    int32_t hist_len;  // samples available in hist (logically x[-hist_len..-1])
    int32_t L, M;      // at most one of them > 1
    int32_t lg;        // output phases staged per store group (1 <= lg <= L)
};

template <int R>
__device__ __forceinline__ int pad_idx(int i) { return i + i / R; }

template <typename C> struct __align__(16) TapVec { C v[16 / sizeof(C)]; };

// PK: complex64 only -- one packed FFMA2 (fma.rn.f32x2, new on sm_100) per complex MAC.  The
// (re,im) accumulator and sample pairs are 64-bit register operands, so every FFMA2 reads two
// full-width operands (acc pair, sample pair; the (t,t) tap pair sits in the operand-reuse
// cache) -- no even/odd register-bank conflicts, half the issue slots of two scalar FFMAs.
template <typename S, bool PK> struct MacOp {
    template <typename C>
    static __device__ __forceinline__ void run(S &acc, C t, const S &x) { tap_fma(acc, t, x); }
};
template <> struct MacOp<float2, true> {
    static __device__ __forceinline__ void run(float2 &acc, float t, const float2 &x) {
        acc = __ffma2_rn(make_float2(t, t), x, acc);
    }
};

template <typename S, typename C, int R, int NT, bool PK = false>
__global__ void __launch_bounds__(NT) fir_poly_kernel(const FirArgs<S, C> a)
{
    constexpr int TILE = NT * R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int kq = a.kq;
    const int P = a.L * a.M;                        // number of tap phases
    const int len_ph = TILE + kq;                   // samples per staged input phase
    const int PS = pad_idx<R>(len_ph) + 1;          // padded phase stride
    C *hs = reinterpret_cast<C *>(smem_raw);
    size_t taps_bytes = ((size_t)P * kq * sizeof(C) + 15) & ~(size_t)15;
    S *xs = reinterpret_cast<S *>(smem_raw + taps_bytes);
    S *st = (a.L == 1) ? xs : xs + (size_t)a.M * PS;   // staging aliases xs when there is one group

    const int64_t m_t = (int64_t)blockIdx.x * TILE;

    // ---- stage tap phases: h[ph][q] = L * b[P q + ph] (zero padded) ----
    {
        const C gain = (C)a.L;
        for (int v = tid; v < P * kq; v += NT) {
            int ph = v / kq, q = v - ph * kq;
            int64_t k = (int64_t)q * P + ph;
            hs[v] = (k < a.ntaps) ? gain * a.taps[k] : (C)0;
        }
    }
    // ---- stage input phases (coalesced over the contiguous global range) ----
    {
        const int M = a.M;
        const int total = M * len_ph;
        const int64_t g_base = (int64_t)M * (m_t - kq) - (M - 1);
        for (int v = tid; v < total; v += NT) {
            int i = (M == 1) ? v : v / M;
            int c = (M == 1) ? 0 : (M - 1 - (v - i * M));
            int64_t g = g_base + v;
            S val = zero_of(S());
            if (g >= 0) {
                if (g < a.n_in) val = a.x[g];
            } else if (a.hist != nullptr) {
                int64_t hidx = (int64_t)a.hist_len + g;
                if (hidx >= 0) val = a.hist[hidx];
            }
            xs[(size_t)c * PS + pad_idx<R>(i)] = val;
        }
    }
    __syncthreads();

    const int nkb = kq / R;
    // first window group of this thread: padded index of sample (kq + tid*R)
    const int grp0 = (nkb + tid) * (R + 1);

    for (int rg = 0; rg < a.L; rg += a.lg) {
        const int lgc = min(a.lg, a.L - rg);
        for (int rr = 0; rr < lgc; ++rr) {
            S acc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = zero_of(S());

            for (int c = 0; c < a.M; ++c) {
                const int ph = (a.L > 1) ? (rg + rr) : c;
                const C *h = hs + ph * kq;
                const S *xc = xs + (size_t)c * PS;
                S w[R];
                // slots 1..R-1 <- x[m0+1 .. m0+R-1]
#pragma unroll
                for (int s = 1; s < R; ++s) w[s] = xc[grp0 + s];
                const S *gp = xc + grp0;     // group holding x[m0 - kb*R]
                for (int kb = 0; kb < nkb; ++kb) {
                    constexpr int TV = 16 / (int)sizeof(C);      // taps per 16-byte broadcast load
#pragma unroll
                    for (int k4 = 0; k4 < R; k4 += TV) {
                        const TapVec<C> tv = *reinterpret_cast<const TapVec<C> *>(h + kb * R + k4);
#pragma unroll
                        for (int u = 0; u < TV; ++u) {
                            const int kk = k4 + u;
                            // new sample x[m0 - (kb*R+kk)] -> slot (R-kk)%R
                            w[(R - kk) % R] = (kk == 0) ? gp[0] : gp[-1 - kk];
                            const C t = tv.v[u];
#pragma unroll
                            for (int j = 0; j < R; ++j) MacOp<S, PK>::run(acc[j], t, w[(j - kk + R) % R]);
                        }
                    }
                    gp -= (R + 1);
                }
            }
            if (a.L == 1) __syncthreads();      // xs is dead: staging may overwrite it
#pragma unroll
            for (int j = 0; j < R; ++j) st[pad_idx<R>((tid * R + j) * lgc + rr)] = acc[j];
        }
        __syncthreads();
        // ---- coalesced store of this phase group ----
        const int total_o = TILE * lgc;
        for (int u = tid; u < total_o; u += NT) {
            int m = (lgc == 1) ? u : u / lgc;
            int r2 = (lgc == 1) ? 0 : (u - m * lgc);
            int64_t mg = m_t + m;
            if (mg < a.n_m) a.y[mg * a.L + rg + r2] = st[pad_idx<R>(u)];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Short-phase interpolation: up(L) with few taps per phase (pulse shaping: 97-tap SRC at 8
// samples/symbol is 13 taps per phase).  fir_poly_kernel pads every phase to a multiple of its R
// and stages L*TILE outputs in shared memory, both of which hurt when Q = ceil(K/L) is small; here
// a thread owns ONE input position m and produces all L outputs y[L m .. L m + L-1] in chunks of
// RP phases held in registers: per tap index q one shared-memory sample (lane-consecutive, conflict
// free) and one broadcast vector of RP taps feed RP FMAs.  Each thread's outputs are contiguous, so
// they leave as 16-byte vector stores straight from registers (a warp covers 32*L consecutive
// elements); the kernel is used when L*sizeof(sample) is a multiple of 16 so that every m starts
// on a 16-byte boundary.  Accumulation order (q ascending, gain folded into the taps) is the same
// as fir_poly_kernel's.
template <typename S> struct VecStore;
template <> struct VecStore<float> {
    static constexpr int EPV = 4;
    static __device__ __forceinline__ void st(float *p, const float *a) {
        *reinterpret_cast<float4 *>(p) = make_float4(a[0], a[1], a[2], a[3]);
    }
};
template <> struct VecStore<float2> {
    static constexpr int EPV = 2;
    static __device__ __forceinline__ void st(float2 *p, const float2 *a) {
        *reinterpret_cast<float4 *>(p) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
    }
};
template <> struct VecStore<double> {
    static constexpr int EPV = 2;
    static __device__ __forceinline__ void st(double *p, const double *a) {
        *reinterpret_cast<double2 *>(p) = make_double2(a[0], a[1]);
    }
};
template <> struct VecStore<double2> {
    static constexpr int EPV = 1;
    static __device__ __forceinline__ void st(double2 *p, const double2 *a) { *p = a[0]; }
};

template <typename S, typename C, int RP, int NT, bool PK>
__global__ void __launch_bounds__(NT) fir_up_short_kernel(const FirArgs<S, C> a)
{
    constexpr int TV = 16 / (int)sizeof(C);          // taps per 16-byte broadcast load
    constexpr int EPV = VecStore<S>::EPV;            // output elements per 16-byte store
    static_assert(RP % TV == 0 && RP % EPV == 0, "phase chunk must be whole vectors");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int Q = a.kq;                              // taps per phase (not padded)
    const int L = a.L;
    const int LP = ((L + RP - 1) / RP) * RP;         // phases padded to whole chunks
    C *hs = reinterpret_cast<C *>(smem_raw);         // hs[q*LP + r] = L * b[L q + r]
    const size_t taps_bytes = ((size_t)Q * LP * sizeof(C) + 15) & ~(size_t)15;
    S *xs = reinterpret_cast<S *>(smem_raw + taps_bytes);   // xs[i] = x[m_t - (Q-1) + i]
    const int64_t m_t = (int64_t)blockIdx.x * NT;
    {
        const C gain = (C)L;
        for (int v = tid; v < Q * LP; v += NT) {
            const int q = v / LP, r = v - q * LP;
            const int64_t k = (int64_t)q * L + r;
            hs[v] = (r < L && k < a.ntaps) ? gain * a.taps[k] : (C)0;
        }
    }
    for (int v = tid; v < NT + Q - 1; v += NT) {
        const int64_t g = m_t - (Q - 1) + v;
        S val = zero_of(S());
        if (g >= 0) {
            if (g < a.n_in) val = a.x[g];
        } else if (a.hist != nullptr) {
            const int64_t hidx = (int64_t)a.hist_len + g;
            if (hidx >= 0) val = a.hist[hidx];
        }
        xs[v] = val;
    }
    __syncthreads();
    const int64_t m = m_t + tid;
    if (m >= a.n_m) return;
    S *yo = a.y + m * L;
    const S *xp = xs + tid + (Q - 1);                // x[m]; x[m-q] = xp[-q]
    for (int r0 = 0; r0 < L; r0 += RP) {
        S acc[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) acc[r] = zero_of(S());
        const C *hp = hs + r0;
        for (int q = 0; q < Q; ++q) {
            const S xv = xp[-q];
#pragma unroll
            for (int r4 = 0; r4 < RP; r4 += TV) {
                const TapVec<C> tv = *reinterpret_cast<const TapVec<C> *>(hp + r4);
#pragma unroll
                for (int u = 0; u < TV; ++u) MacOp<S, PK>::run(acc[r4 + u], tv.v[u], xv);
            }
            hp += LP;
        }
#pragma unroll
        for (int r = 0; r < RP; r += EPV)
            if (r0 + r + EPV <= L) VecStore<S>::st(yo + r0 + r, acc + r);     // L % EPV == 0: no ragged tail
    }
}

// ------------------------------------------------------------------------------------------
struct b200dsp_fir_plan_impl {
    int32_t ntaps;
    float *taps_f32;
    double *taps_f64;
    double *taps_host;   // host copy
    void *tc2_amat;      // taps-stationary tensor-core path: 128 x (320 | 320 with b_lo rows zero) fp16 for TMEM (or NULL)
    int32_t tc2_sb_exp;
    // float32 tensor-core paths (fir_tc_real.cu): tap matrices per (mode, factor), built on first use
    void *tcr_mat[4][5];      // pointers into tcr_pool
    int32_t tcr_sb[4][5];
    int8_t tcr_state[4][5];   // 1 ready, otherwise this (mode, factor) does not fit the tensor-core kernel
    void *tcr_pool;           // one allocation for all of them (plan_create)
    void *fft_tables;         // fir_fft.cu: digit-reversed spectrum of the taps + twiddles (2 <= ntaps <= 2049), or NULL
    int32_t sm_count;
};

// fir_tc_real.cu
int tcr_build(const double *taps, int ntaps, int mode, int P, unsigned char *out, int *sb_exp);
int tcr_matrix_bytes(int mode, int P);
int launch_fir_tc_real(int mode, int P, const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                       const void *amat_dev, int sb_exp, int ntaps, int sm_count, cudaStream_t stream);
// fir_fft.cu
int fft_table_floats();
int fft_build_tables(const double *taps, int ntaps, float *out);
int launch_fir_fft(bool real, const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                   const void *tables_dev, int ntaps, int sm_count, cudaStream_t stream);
// fir_tc2.cu
int tc2_build_tap_matrix(const double *taps, int ntaps, unsigned char *out, int *sb_exp);
int tc2_matrix_bytes();
int launch_fir_tc2(const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                   const void *amat_dev, int sb_exp, int ntaps, int tile_rows, int sm_count, cudaStream_t stream,
                   int32_t M = 1, int32_t L = 1);

static thread_local int g_fir_variant = 0;

template <typename S, int R, int NT, bool PK = false>
static int launch_fir(const b200dsp_fir_plan_impl *p, const void *x, const void *hist, void *y,
                      int64_t n, int64_t n_m, int32_t L, int32_t M, int32_t hist_len,
                      cudaStream_t stream)
{
    using C = typename Sample<S>::C;
    constexpr int TILE = NT * R;
    FirArgs<S, C> a;
    a.x = static_cast<const S *>(x);
    a.hist = static_cast<const S *>(hist);
    a.y = static_cast<S *>(y);
    a.taps = sizeof(C) == 4 ? reinterpret_cast<const C *>(p->taps_f32)
                            : reinterpret_cast<const C *>(p->taps_f64);
    a.n_in = n;
    a.n_m = n_m;
    a.ntaps = p->ntaps;
    const int P = L * M;
    int taps_per_phase = (p->ntaps + P - 1) / P;
    a.kq = ((taps_per_phase + R - 1) / R) * R;
    a.hist_len = hist_len;
    a.L = L;
    a.M = M;
    const int len_ph = TILE + a.kq;
    const size_t PS = (size_t)(len_ph + len_ph / R) + 1;
    const size_t taps_bytes = ((size_t)P * a.kq * sizeof(C) + 15) & ~(size_t)15;
    const size_t xs_bytes = (size_t)M * PS * sizeof(S);
    size_t smem;
    if (L == 1) {
        a.lg = 1;
        size_t st_bytes = (size_t)(TILE + TILE / R + 1) * sizeof(S);
        smem = taps_bytes + (xs_bytes > st_bytes ? xs_bytes : st_bytes);
    } else {
        if (taps_bytes + xs_bytes + (size_t)(TILE + TILE / R + 1) * sizeof(S) > kMaxSmemPerBlock)
            return B200DSP_E_UNSUPPORTED;
        size_t room = kMaxSmemPerBlock - taps_bytes - xs_bytes;
        // keep the staging area modest so several blocks stay resident per SM
        size_t budget = room < 96 * 1024 ? room : 96 * 1024;
        int lg = (int)(budget / ((size_t)(TILE + TILE / R + 1) * sizeof(S)));
        if (lg < 1) lg = 1;
        if (lg > L) lg = L;
        a.lg = lg;
        size_t st_elems = (size_t)TILE * lg;
        smem = taps_bytes + xs_bytes + (st_elems + st_elems / R + 1) * sizeof(S);
    }
    if (smem > kMaxSmemPerBlock) return B200DSP_E_UNSUPPORTED;
    auto kern = fir_poly_kernel<S, C, R, NT, PK>;
    B200_CHECK_CUDA(allow_smem(kern, smem));
    int64_t tiles = (n_m + TILE - 1) / TILE;
    if (tiles > 2147483647LL) {
        set_error("fir: too many tiles (%lld)", (long long)tiles);
        return B200DSP_E_UNSUPPORTED;
    }
    kern<<<(unsigned)tiles, NT, smem, stream>>>(a);
    B200_CHECK_LAUNCH("fir_poly_kernel");
    return B200DSP_OK;
}

// ---- last resort: no staging limits ---------------------------------------------------------------------
// One thread per output, taps and samples straight from global memory (L1 / L2 keep them).  Serves the shapes
// whose staging does not fit fir_poly_kernel's shared-memory tile: decimation by more than ~50 (the reference
// accepts any M, multirate_helper.py:121-127) or filters of many thousand taps.  Same accumulation order as the
// oracle's direct sum (k ascending), float64 accumulators for the float64 dtypes, float32 otherwise.
template <typename S, typename C>
__global__ void __launch_bounds__(256) fir_generic_kernel(const FirArgs<S, C> a, int64_t n_out)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    // output o: up -> (m, r) = (o / L, o % L), taps r, r+L, ... against x[m], x[m-1], ...; filter / dn: all taps
    // against x[M o - k]
    const int64_t m = (a.L > 1) ? o / a.L : o;
    const int r = (a.L > 1) ? (int)(o - m * a.L) : 0;
    const int64_t pos0 = (a.L > 1) ? m : o * (int64_t)a.M;
    const C gain = (a.L > 1) ? (C)a.L : (C)1;
    S acc = zero_of(S());
    for (int t = r, q = 0; t < a.ntaps; t += a.L, ++q) {
        const int64_t pos = pos0 - q;
        S v;
        if (pos >= 0) {
            if (pos >= a.n_in) continue;
            v = a.x[pos];
        } else {
            const int64_t h = (int64_t)a.hist_len + pos;
            if (a.hist == nullptr || h < 0) break;              // zero initial state from here on
            v = a.hist[h];
        }
        tap_fma(acc, a.taps[t] * gain, v);
    }
    a.y[o] = acc;
}

template <typename S>
static int launch_fir_generic(const b200dsp_fir_plan_impl *p, const void *x, const void *hist, void *y,
                              int64_t n, int64_t n_m, int32_t L, int32_t M, int32_t hist_len, cudaStream_t stream)
{
    using C = typename Sample<S>::C;
    FirArgs<S, C> a;
    a.x = static_cast<const S *>(x);
    a.hist = static_cast<const S *>(hist);
    a.y = static_cast<S *>(y);
    a.taps = sizeof(C) == 4 ? reinterpret_cast<const C *>(p->taps_f32) : reinterpret_cast<const C *>(p->taps_f64);
    a.n_in = n;
    a.n_m = n_m;
    a.ntaps = p->ntaps;
    a.kq = 0;
    a.hist_len = hist_len;
    a.L = L;
    a.M = M;
    a.lg = 1;
    const int64_t n_out = n_m * L;
    const int64_t blocks = (n_out + 255) / 256;
    if (blocks > 2147483647LL) {
        set_error("fir: too many blocks (%lld)", (long long)blocks);
        return B200DSP_E_UNSUPPORTED;
    }
    fir_generic_kernel<S, C><<<(unsigned)blocks, 256, 0, stream>>>(a, n_out);
    B200_CHECK_LAUNCH("fir_generic_kernel");
    return B200DSP_OK;
}

// Try progressively smaller tiles until the staging fits in shared memory.
template <typename S, int R, bool PK = false>
static int launch_fir_fit(const b200dsp_fir_plan_impl *p, const void *x, const void *hist, void *y,
                          int64_t n, int64_t n_m, int32_t L, int32_t M, int32_t hist_len,
                          cudaStream_t stream, int nt_first)
{
    int rc = B200DSP_E_UNSUPPORTED;
    if (nt_first >= 256) {
        rc = launch_fir<S, R, 256, PK>(p, x, hist, y, n, n_m, L, M, hist_len, stream);
        if (rc != B200DSP_E_UNSUPPORTED) return rc;
    }
    if (nt_first >= 128) {
        rc = launch_fir<S, R, 128, PK>(p, x, hist, y, n, n_m, L, M, hist_len, stream);
        if (rc != B200DSP_E_UNSUPPORTED) return rc;
    }
    rc = launch_fir<S, R, 64, PK>(p, x, hist, y, n, n_m, L, M, hist_len, stream);
    if (rc != B200DSP_E_UNSUPPORTED) return rc;
    rc = launch_fir<S, R, 32, PK>(p, x, hist, y, n, n_m, L, M, hist_len, stream);
    if (rc == B200DSP_E_UNSUPPORTED)      // staging does not fit even the smallest tile: unstaged kernel
        rc = launch_fir_generic<S>(p, x, hist, y, n, n_m, L, M, hist_len, stream);
    return rc;
}

// up(L) through fir_up_short_kernel when the filter has few taps per phase and every output row starts on a
// 16-byte boundary; B200DSP_E_UNSUPPORTED = not eligible (caller falls back to fir_poly_kernel).
template <typename S, int RP, bool PK = false>
static int launch_fir_up_short(const b200dsp_fir_plan_impl *p, const void *x, const void *hist, void *y,
                               int64_t n, int32_t L, int32_t hist_len, cudaStream_t stream)
{
    using C = typename Sample<S>::C;
    constexpr int NT = 256;
    constexpr int QMAX = 32;
    const int Q = (p->ntaps + L - 1) / L;
    if (L < 2 || Q > QMAX || ((size_t)L * sizeof(S)) % 16 != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0)
        return B200DSP_E_UNSUPPORTED;
    const int LP = ((L + RP - 1) / RP) * RP;
    const size_t taps_bytes = ((size_t)Q * LP * sizeof(C) + 15) & ~(size_t)15;
    const size_t smem = taps_bytes + (size_t)(NT + Q - 1) * sizeof(S);
    if (smem > 96 * 1024) return B200DSP_E_UNSUPPORTED;
    const int64_t tiles = (n + NT - 1) / NT;
    if (tiles > 2147483647LL) return B200DSP_E_UNSUPPORTED;
    FirArgs<S, C> a;
    a.x = static_cast<const S *>(x);
    a.hist = static_cast<const S *>(hist);
    a.y = static_cast<S *>(y);
    a.taps = sizeof(C) == 4 ? reinterpret_cast<const C *>(p->taps_f32)
                            : reinterpret_cast<const C *>(p->taps_f64);
    a.n_in = n;
    a.n_m = n;
    a.ntaps = p->ntaps;
    a.kq = Q;
    a.hist_len = hist_len;
    a.L = L;
    a.M = 1;
    a.lg = 1;
    auto kern = fir_up_short_kernel<S, C, RP, NT, PK>;
    B200_CHECK_CUDA(allow_smem(kern, smem));
    kern<<<(unsigned)tiles, NT, smem, stream>>>(a);
    B200_CHECK_LAUNCH("fir_up_short_kernel");
    return B200DSP_OK;
}

// The TMEM tap matrices of the float32 tensor-core modes (filter; up / dn by 2..4) are all built in
// b200dsp_fir_plan_create into one device allocation: filter calls allocate nothing, synchronise nothing and
// never write to the plan (include/b200dsp.h contract; plans can be shared between threads and captured in graphs).
static bool tcr_ready(const b200dsp_fir_plan_impl *p, int mode, int P) { return p->tcr_state[mode][P] == 1; }

static int fir_dispatch(const b200dsp_fir_plan_impl *p, int dtype, const void *x, const void *hist,
                        void *y, int64_t n, int64_t n_m, int32_t L, int32_t M, int32_t hist_len,
                        cudaStream_t s)
{
    const int v = g_fir_variant;
    switch (dtype) {
    case B200DSP_F32: {
        // float32 streams on the tensor cores (fir_tc_real.cu): filter (<= 256 taps), up/dn by 2..4 with at
        // most 64 taps per phase; 16-byte aligned streams only.  v == 9 forces the CUDA-core kernel.
        const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
        const int mode = (L > 1) ? 2 : (M > 1 ? 3 : 1);
        const int P = (L > 1) ? L : (M > 1 ? M : 1);
        if (v == 0 && aligned && P <= 4 && n_m >= 16384 &&
            tcr_ready(p, mode, P))
            return launch_fir_tc_real(mode, P, x, hist, y, n, hist_len, p->tcr_mat[mode][P],
                                      p->tcr_sb[mode][P], p->ntaps, p->sm_count, s);
        }
        // Longer filters (257 .. 2049 taps): overlap-save FFT, two real frames per complex transform (fir_fft.cu)
        if (L == 1 && M == 1 && p->fft_tables != nullptr && (v == 16 || (v == 0 && p->ntaps > 256 && n >= 32768)))
            return launch_fir_fft(true, x, hist, y, n, hist_len, p->fft_tables, p->ntaps, p->sm_count, s);
        // dn(M) the phase-stream kernel above did not take (M > 4 -- the reference's default is 12 -- or more than 65
        // taps per phase): the tensor-core FILTER kernel with decimating stores.  All outputs are computed, but the
        // stream is read once at that kernel's rate (3x the polyphase kernel at M = 12).
        if (L == 1 && M > 1 && v == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && n >= 32768 && tcr_ready(p, 1, 1))
            return launch_fir_tc_real(1, M, x, hist, y, n, hist_len, p->tcr_mat[1][1], p->tcr_sb[1][1], p->ntaps,
                                      p->sm_count, s);
        if (L > 1 && v != 8 && v != 9) {       // few taps per phase: short-phase kernel (v == 8 / 9: fir_poly_kernel)
            const int rc = launch_fir_up_short<float, 8>(p, x, hist, y, n, L, hist_len, s);
            if (rc != B200DSP_E_UNSUPPORTED) return rc;
        }
        if (v == 1) return launch_fir_fit<float, 16>(p, x, hist, y, n, n_m, L, M, hist_len, s, 256);
        return launch_fir_fit<float, 32>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
    case B200DSP_C64:
        // default for long complex64 streams: block-Toeplitz GEMM on the tensor cores (fir_tc.cu).
        // v == 10 forces it for any length, v == 9 forces the CUDA-core kernel.
        // Long complex64 streams: block-Toeplitz GEMM on the tensor cores (fir_tc2.cu, taps in TMEM).
        // Needs 16-byte aligned streams (bulk TMA in, vector stores out); otherwise the CUDA-core kernel.
        //   v == 0 auto (96-row tiles, x_lo pass against zeroed b_lo rows: tile code 97) | 12 / 15 / 14 / 13 force the
        //   four-product tensor-core kernel with 64 / 80 / 96 / 128-row
        //   tiles | 9 force CUDA cores | 1..5 CUDA-core shapes
        if (L == 1 && M == 1) {
            const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
            if (p->tc2_amat != nullptr && (v == 12 || v == 13 || v == 14 || v == 15 || (v == 0 && n >= 32768 && aligned)))
                return launch_fir_tc2(x, hist, y, n, hist_len, p->tc2_amat, p->tc2_sb_exp, p->ntaps,
                                      v == 13 ? 128 : (v == 12 ? 64 : (v == 15 ? 80 : (v == 14 ? 96 : 97))), p->sm_count, s);
        }
        // dn(M) on long streams: the same kernel with a decimating epilogue -- every output is computed (an M-phase
        // tensor-core formulation would do 1/M of the MACs) but the stream is read once at the filter kernel's rate,
        // 4-5x the CUDA-core polyphase kernel at M = 4 .. 12.  v == 9 forces the CUDA-core kernel.
        if (L == 1 && M > 1 && p->tc2_amat != nullptr && v == 0 && n >= 32768 && n / M >= 1 &&
            (reinterpret_cast<uintptr_t>(x) & 15) == 0)
            return launch_fir_tc2(x, hist, y, n, hist_len, p->tc2_amat, p->tc2_sb_exp, p->ntaps, 96, p->sm_count, s, M);
        // Longer filters (257 .. 2049 taps): overlap-save with the in-shared-memory 4096-point FFT (fir_fft.cu);
        // v == 16 forces it for any filter it can take
        if (L == 1 && M == 1 && p->fft_tables != nullptr && (v == 16 || (v == 0 && p->ntaps > 256 && n >= 32768)))
            return launch_fir_fft(false, x, hist, y, n, hist_len, p->fft_tables, p->ntaps, p->sm_count, s);
        // up(L) with more than 32 taps per phase (the short-phase kernel's limit; e.g. 256 taps, L = 2 .. 7): the
        // tensor-core filter kernel on the zero-stuffed stream.  Only the input samples travel (one small bulk copy
        // per tile), the converter warps stuff the zeros; L-1 of L MACs multiply zeros, and it is still 2.4x the
        // CUDA-core polyphase kernel at L = 4 because the stream side runs at the filter kernel's rate.
        if (L > 1 && M == 1 && v == 0 && p->tc2_amat != nullptr && (p->ntaps + L - 1) / L > 32 && n * L >= 32768 &&
            (reinterpret_cast<uintptr_t>(x) & 15) == 0)
            return launch_fir_tc2(x, hist, y, n, hist_len, p->tc2_amat, p->tc2_sb_exp, p->ntaps, 96, p->sm_count, s, 1, L);
        if (L > 1 && v != 8 && v != 9) {
            const int rc = launch_fir_up_short<float2, 8, true>(p, x, hist, y, n, L, hist_len, s);
            if (rc != B200DSP_E_UNSUPPORTED) return rc;
        }
        if (v == 1) return launch_fir_fit<float2, 16>(p, x, hist, y, n, n_m, L, M, hist_len, s, 256);
        if (v == 2) return launch_fir_fit<float2, 16>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
        if (v == 3) return launch_fir_fit<float2, 32>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
        if (v == 4) return launch_fir_fit<float2, 16, true>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
        if (v == 5) return launch_fir_fit<float2, 16, true>(p, x, hist, y, n, n_m, L, M, hist_len, s, 256);
        // default: 32 outputs/thread, packed FFMA2 (fastest CUDA-core variant measured on B200)
        return launch_fir_fit<float2, 32, true>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
    case B200DSP_F64:
        if (L > 1 && v != 8 && v != 9) {
            const int rc = launch_fir_up_short<double, 4>(p, x, hist, y, n, L, hist_len, s);
            if (rc != B200DSP_E_UNSUPPORTED) return rc;
        }
        return launch_fir_fit<double, 16>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
    case B200DSP_C128:
        if (L > 1 && v != 8 && v != 9) {
            const int rc = launch_fir_up_short<double2, 4>(p, x, hist, y, n, L, hist_len, s);
            if (rc != B200DSP_E_UNSUPPORTED) return rc;
        }
        return launch_fir_fit<double2, 8>(p, x, hist, y, n, n_m, L, M, hist_len, s, 128);
    default:
        set_error("fir: bad dtype code %d", dtype);
        return B200DSP_E_DTYPE;
    }
}

}  // namespace b200dsp

using namespace b200dsp;

struct b200dsp_fir_plan : b200dsp_fir_plan_impl {};

extern "C" {

int b200dsp_fir_plan_create(const double *taps_host, int32_t ntaps, b200dsp_fir_plan **plan)
{
    if (!taps_host || !plan || ntaps < 1) {
        set_error("fir_plan_create: bad argument (ntaps=%d)", ntaps);
        return B200DSP_E_BADARG;
    }
    b200dsp_fir_plan *p = new b200dsp_fir_plan();
    p->ntaps = ntaps;
    p->taps_f32 = nullptr;
    p->taps_f64 = nullptr;
    p->tc2_amat = nullptr;
    p->tc2_sb_exp = 0;
    memset(p->tcr_mat, 0, sizeof(p->tcr_mat));
    memset(p->tcr_sb, 0, sizeof(p->tcr_sb));
    memset(p->tcr_state, 0, sizeof(p->tcr_state));
    p->taps_host = new double[ntaps];
    memcpy(p->taps_host, taps_host, sizeof(double) * ntaps);
    p->sm_count = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess)
            cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    float *tmp = new float[ntaps];
    for (int i = 0; i < ntaps; ++i) tmp[i] = (float)taps_host[i];
    cudaError_t e = cudaMalloc(&p->taps_f32, sizeof(float) * ntaps);
    if (e == cudaSuccess) e = cudaMalloc(&p->taps_f64, sizeof(double) * ntaps);
    if (e == cudaSuccess) e = cudaMemcpy(p->taps_f32, tmp, sizeof(float) * ntaps, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->taps_f64, taps_host, sizeof(double) * ntaps, cudaMemcpyHostToDevice);
    delete[] tmp;
    if (e == cudaSuccess && ntaps <= 256) {
        const int nb2 = tc2_matrix_bytes();
        unsigned char *h2 = new unsigned char[nb2];
        if (tc2_build_tap_matrix(taps_host, ntaps, h2, &p->tc2_sb_exp) == 0) {
            e = cudaMalloc(&p->tc2_amat, nb2);
            if (e == cudaSuccess) e = cudaMemcpy(p->tc2_amat, h2, nb2, cudaMemcpyHostToDevice);
        }
        delete[] h2;
    }
    // overlap-save FFT tables (complex64 and float32 streams, 2 .. 2049 taps)
    p->fft_tables = nullptr;
    if (e == cudaSuccess && ntaps >= 2 && ntaps <= 2049) {
        std::vector<float> tb((size_t)fft_table_floats());
        if (fft_build_tables(taps_host, ntaps, tb.data()) == 0) {
            e = cudaMalloc(&p->fft_tables, tb.size() * sizeof(float));
            if (e == cudaSuccess) e = cudaMemcpy(p->fft_tables, tb.data(), tb.size() * sizeof(float), cudaMemcpyHostToDevice);
        }
    }
    // float32 tensor-core tap matrices: mode 1 filter, 2 up(P), 3 dn(P), P = 2..4 -- every one that fits the filter
    p->tcr_pool = nullptr;
    if (e == cudaSuccess) {
        std::vector<unsigned char> pool;
        struct Slot { int mode, P; size_t off; };
        std::vector<Slot> slots;
        for (int mode = 1; mode <= 3; ++mode)
            for (int P = (mode == 1 ? 1 : 2); P <= (mode == 1 ? 1 : 4); ++P) {
                const int nb = tcr_matrix_bytes(mode, P);
                const size_t off = (pool.size() + 255) & ~(size_t)255;
                pool.resize(off + nb);
                int sb = 0;
                if (tcr_build(taps_host, ntaps, mode, P, pool.data() + off, &sb) == 0) {
                    slots.push_back({mode, P, off});
                    p->tcr_sb[mode][P] = sb;
                } else {
                    pool.resize(off);
                }
            }
        if (!slots.empty()) {
            e = cudaMalloc(&p->tcr_pool, pool.size());
            if (e == cudaSuccess) e = cudaMemcpy(p->tcr_pool, pool.data(), pool.size(), cudaMemcpyHostToDevice);
            if (e == cudaSuccess)
                for (const Slot &sl : slots) {
                    p->tcr_mat[sl.mode][sl.P] = static_cast<unsigned char *>(p->tcr_pool) + sl.off;
                    p->tcr_state[sl.mode][sl.P] = 1;
                }
        }
    }
    // cudaMemcpy from pageable host memory may return before the DMA to the device has completed; kernels
    // are launched on arbitrary (non-blocking) streams, so make the uploads visible to all of them now
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        set_error("fir_plan_create: %s", cudaGetErrorString(e));
        cudaFree(p->taps_f32);
        cudaFree(p->taps_f64);
        cudaFree(p->tc2_amat);
        cudaFree(p->tcr_pool);
        cudaFree(p->fft_tables);
        delete[] p->taps_host;
        delete p;
        return B200DSP_E_CUDA;
    }
    *plan = p;
    return B200DSP_OK;
}

void b200dsp_fir_plan_destroy(b200dsp_fir_plan *plan)
{
    if (!plan) return;
    cudaFree(plan->taps_f32);
    cudaFree(plan->taps_f64);
    cudaFree(plan->tc2_amat);
    cudaFree(plan->tcr_pool);
    cudaFree(plan->fft_tables);
    delete[] plan->taps_host;
    delete plan;
}

int32_t b200dsp_fir_plan_ntaps(const b200dsp_fir_plan *plan) { return plan ? plan->ntaps : 0; }

int b200dsp_fir_filter(const b200dsp_fir_plan *plan, int dtype, const void *x, const void *hist,
                       void *y, int64_t n, void *stream)
{
    if (!plan || n < 0 || (n > 0 && (!x || !y))) {
        set_error("fir_filter: bad argument");
        return B200DSP_E_BADARG;
    }
    if (n == 0) return B200DSP_OK;
    return fir_dispatch(plan, dtype, x, hist, y, n, n, 1, 1, plan->ntaps - 1, (cudaStream_t)stream);
}

int32_t b200dsp_fir_up_hist_len(const b200dsp_fir_plan *plan, int32_t L)
{
    if (!plan || L < 1) return 0;
    return (plan->ntaps - 1 + L - 1) / L;
}

int b200dsp_fir_up(const b200dsp_fir_plan *plan, int dtype, const void *x, const void *hist,
                   void *y, int64_t n, int32_t L, void *stream)
{
    if (!plan || n < 0 || L < 1 || (n > 0 && (!x || !y))) {
        set_error("fir_up: bad argument (L=%d)", L);
        return B200DSP_E_BADARG;
    }
    if (n == 0) return B200DSP_OK;
    return fir_dispatch(plan, dtype, x, hist, y, n, n, L, 1, b200dsp_fir_up_hist_len(plan, L),
                        (cudaStream_t)stream);
}

int b200dsp_fir_dn(const b200dsp_fir_plan *plan, int dtype, const void *x, const void *hist,
                   void *y, int64_t n, int32_t M, void *stream)
{
    if (!plan || n < 0 || M < 1 || (n > 0 && (!x || !y))) {
        set_error("fir_dn: bad argument (M=%d)", M);
        return B200DSP_E_BADARG;
    }
    int64_t n_m = n / M;
    if (n_m == 0) return B200DSP_OK;
    return fir_dispatch(plan, dtype, x, hist, y, n, n_m, 1, M, plan->ntaps - 1, (cudaStream_t)stream);
}

int b200dsp_fir_filter_batch(const b200dsp_fir_plan *plan, int dtype, const void *x, void *y, int64_t rows,
                             int64_t n, int64_t x_row_stride, int64_t y_row_stride, void *stream)
{
    if (!plan || rows < 0 || n < 0 || x_row_stride < n || y_row_stride < n || (rows > 0 && n > 0 && (!x || !y))) {
        set_error("fir_filter_batch: bad argument");
        return B200DSP_E_BADARG;
    }
    size_t es = 0;
    switch (dtype) {
    case B200DSP_F32: es = 4; break;
    case B200DSP_F64: es = 8; break;
    case B200DSP_C64: es = 8; break;
    case B200DSP_C128: es = 16; break;
    default: set_error("fir_filter_batch: bad dtype code %d", dtype); return B200DSP_E_DTYPE;
    }
    for (int64_t r = 0; r < rows && n > 0; ++r) {
        const int rc = fir_dispatch(plan, dtype, static_cast<const char *>(x) + (size_t)r * x_row_stride * es, nullptr,
                                    static_cast<char *>(y) + (size_t)r * y_row_stride * es, n, n, 1, 1,
                                    plan->ntaps - 1, (cudaStream_t)stream);
        if (rc != B200DSP_OK) return rc;
    }
    return B200DSP_OK;
}

void b200dsp_set_fir_variant(int variant) { g_fir_variant = variant; }

}  // extern "C"
