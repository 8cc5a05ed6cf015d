// resample.cu -- integer zero-stuffing / decimation index maps (sm_100a), bit exact.
// Replaces sigsys.upsample / sigsys.downsample (src/sk_dsp_comm/sigsys.py:3031-3083).
// Both are pure HBM streaming: grids are sized to a multiple of the SM count and every
// thread moves whole elements (4/8/16 bytes) with coalesced accesses on the dense side.
#include "common.cuh"

namespace b200dsp {

// y[i*L] = x[i]; other outputs zero.  Threads own 16-byte OUTPUT vectors (the L-times larger
// side) so stores are full-width and perfectly coalesced; one integer division per vector, the
// element index then advances incrementally.  Reads touch each x element once per vector that
// overlaps its row (L1/L2 hits).
template <typename E, typename IDX, bool GAIN = false>
__global__ void __launch_bounds__(256) upsample_vec_kernel(const E *__restrict__ x, E *__restrict__ y,
                                                          int64_t n_out, int32_t L, typename Sample<E>::C gain = 1)
{
    constexpr int VEC = 16 / (int)sizeof(E);
    struct __align__(16) Pack { E v[VEC]; };
    const int64_t n_vec = n_out / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t vi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vi < n_vec; vi += stride) {
        const IDX o = (IDX)(vi * VEC);
        IDX q = o / (IDX)L;
        int r = (int)(o - q * (IDX)L);
        Pack pk;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            pk.v[e] = (r == 0) ? (GAIN ? scale_of(x[q], gain) : x[q]) : zero_of(E());
            if (++r == L) { r = 0; ++q; }
        }
        reinterpret_cast<Pack *>(y)[vi] = pk;
    }
    // tail (n_out not a multiple of VEC)
    const int64_t t = n_vec * VEC + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_out) {
        int64_t q = t / L;
        y[t] = (q * L == t) ? (GAIN ? scale_of(x[q], gain) : x[q]) : zero_of(E());
    }
}

template <typename E>
__global__ void __launch_bounds__(256) upsample_kernel(const E *__restrict__ x, E *__restrict__ y,
                                                      int64_t n_out, int32_t L)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += stride) {
        int64_t q = o / L;
        E v = zero_of(E());
        if (q * L == o) v = x[q];
        y[o] = v;
    }
}

// y[m] = x[m*M + p]
template <typename E>
__global__ void __launch_bounds__(256) downsample_kernel(const E *__restrict__ x, E *__restrict__ y,
                                                        int64_t n_out, int32_t M, int32_t p)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += stride)
        y[o] = x[o * M + p];
}

static unsigned grid_for(int64_t n, int sm_count)
{
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count * 16;          // 16 resident 256-thread blocks per SM at most
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static int sm_count_cached()
{
    static thread_local int dev_cached = -1, sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != dev_cached) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dev_cached = dev;
    }
    return sms;
}

template <typename E>
static int up_launch(const void *x, void *y, int64_t n, int32_t L, cudaStream_t st)
{
    int64_t n_out = n * L;
    constexpr int VEC = 16 / (int)sizeof(E);
    if ((reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        unsigned g = grid_for((n_out + VEC - 1) / VEC, sm_count_cached());
        if (n_out < (int64_t)1 << 31)
            upsample_vec_kernel<E, uint32_t><<<g, 256, 0, st>>>((const E *)x, (E *)y, n_out, L);
        else
            upsample_vec_kernel<E, int64_t><<<g, 256, 0, st>>>((const E *)x, (E *)y, n_out, L);
    } else {
        upsample_kernel<E><<<grid_for(n_out, sm_count_cached()), 256, 0, st>>>((const E *)x, (E *)y, n_out, L);
    }
    B200_CHECK_LAUNCH("upsample_kernel");
    return B200DSP_OK;
}

template <typename E>
static int dn_launch(const void *x, void *y, int64_t n_out, int32_t M, int32_t p, cudaStream_t st)
{
    downsample_kernel<E><<<grid_for(n_out, sm_count_cached()), 256, 0, st>>>((const E *)x, (E *)y, n_out, M, p);
    B200_CHECK_LAUNCH("downsample_kernel");
    return B200DSP_OK;
}

// ---- internal helpers for the SOS cascade's rate-change staging (sos_scan.cu) ----
template <typename E>
static int up_gain_launch(const void *x, void *y, int64_t n, int32_t L, double gain, cudaStream_t st)
{
    using C = typename Sample<E>::C;
    const int64_t n_out = n * L;
    constexpr int VEC = 16 / (int)sizeof(E);
    unsigned g = grid_for((n_out + VEC - 1) / VEC, sm_count_cached());
    upsample_vec_kernel<E, int64_t, true><<<g, 256, 0, st>>>((const E *)x, (E *)y, n_out, L, (C)gain);
    B200_CHECK_LAUNCH("upsample_vec_kernel(gain)");
    return B200DSP_OK;
}

// y (16-byte aligned) = gain * zero-stuffed x
int resample_up_scaled(int dtype, const void *x, void *y, int64_t n, int32_t L, double gain, cudaStream_t st)
{
    switch (dtype) {
    case B200DSP_F32: return up_gain_launch<float>(x, y, n, L, gain, st);
    case B200DSP_F64: return up_gain_launch<double>(x, y, n, L, gain, st);
    case B200DSP_C64: return up_gain_launch<float2>(x, y, n, L, gain, st);
    case B200DSP_C128: return up_gain_launch<double2>(x, y, n, L, gain, st);
    }
    return B200DSP_E_DTYPE;
}

int resample_dn(int dtype, const void *x, void *y, int64_t n_out, int32_t M, cudaStream_t st)
{
    switch (dtype) {
    case B200DSP_F32: return dn_launch<float>(x, y, n_out, M, 0, st);
    case B200DSP_F64: return dn_launch<double>(x, y, n_out, M, 0, st);
    case B200DSP_C64: return dn_launch<float2>(x, y, n_out, M, 0, st);
    case B200DSP_C128: return dn_launch<double2>(x, y, n_out, M, 0, st);
    }
    return B200DSP_E_DTYPE;
}

}  // namespace b200dsp

using namespace b200dsp;

// y = a + j b : the two real-tap passes of a complex-tap FIR (lfilter accepts complex b; by linearity
// lfilter(br + j bi, 1, x) = lfilter(br, 1, x) + j lfilter(bi, 1, x)).  CIN: a, b complex (complex stream) or real.
template <typename R, typename C2, bool CIN>
__global__ void __launch_bounds__(256) combine_complex_kernel(const void *a, const void *b, C2 *__restrict__ y, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    C2 o;
    if (CIN) {
        const C2 u = static_cast<const C2 *>(a)[i], v = static_cast<const C2 *>(b)[i];
        o.x = u.x - v.y;
        o.y = u.y + v.x;
    } else {
        o.x = static_cast<const R *>(a)[i];
        o.y = static_cast<const R *>(b)[i];
    }
    y[i] = o;
}

extern "C" {

int b200dsp_combine_complex(int dtype_in, const void *a, const void *b, void *y, int64_t n, void *stream)
{
    if (n < 0 || (n > 0 && (!a || !b || !y))) {
        set_error("combine_complex: bad argument");
        return B200DSP_E_BADARG;
    }
    if (n == 0) return B200DSP_OK;
    const int64_t blocks = (n + 255) / 256;
    if (blocks > 2147483647LL) {
        set_error("combine_complex: too many blocks");
        return B200DSP_E_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype_in) {
    case B200DSP_F32: combine_complex_kernel<float, float2, false><<<(unsigned)blocks, 256, 0, st>>>(a, b, (float2 *)y, n); break;
    case B200DSP_F64: combine_complex_kernel<double, double2, false><<<(unsigned)blocks, 256, 0, st>>>(a, b, (double2 *)y, n); break;
    case B200DSP_C64: combine_complex_kernel<float, float2, true><<<(unsigned)blocks, 256, 0, st>>>(a, b, (float2 *)y, n); break;
    case B200DSP_C128: combine_complex_kernel<double, double2, true><<<(unsigned)blocks, 256, 0, st>>>(a, b, (double2 *)y, n); break;
    default:
        set_error("combine_complex: bad dtype code %d", dtype_in);
        return B200DSP_E_DTYPE;
    }
    B200_CHECK_LAUNCH("combine_complex_kernel");
    return B200DSP_OK;
}

int b200dsp_upsample(int dtype, const void *x, void *y, int64_t n, int32_t L, void *stream)
{
    if (n < 0 || L < 1 || (n > 0 && (!x || !y))) {
        set_error("upsample: bad argument (L=%d)", L);
        return B200DSP_E_BADARG;
    }
    if (n == 0) return B200DSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case B200DSP_F32: return up_launch<float>(x, y, n, L, st);
    case B200DSP_F64: return up_launch<double>(x, y, n, L, st);
    case B200DSP_C64: return up_launch<float2>(x, y, n, L, st);
    case B200DSP_C128: return up_launch<double2>(x, y, n, L, st);
    }
    set_error("upsample: bad dtype code %d", dtype);
    return B200DSP_E_DTYPE;
}

int b200dsp_downsample(int dtype, const void *x, void *y, int64_t n, int32_t M, int32_t p, void *stream)
{
    if (n < 0 || M < 1 || p < 0 || p >= M || (n > 0 && (!x || !y))) {
        set_error("downsample: bad argument (M=%d, p=%d)", M, p);
        return B200DSP_E_BADARG;
    }
    int64_t n_out = n / M;
    if (n_out == 0) return B200DSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case B200DSP_F32: return dn_launch<float>(x, y, n_out, M, p, st);
    case B200DSP_F64: return dn_launch<double>(x, y, n_out, M, p, st);
    case B200DSP_C64: return dn_launch<float2>(x, y, n_out, M, p, st);
    case B200DSP_C128: return dn_launch<double2>(x, y, n_out, M, p, st);
    }
    set_error("downsample: bad dtype code %d", dtype);
    return B200DSP_E_DTYPE;
}

}  // extern "C"
