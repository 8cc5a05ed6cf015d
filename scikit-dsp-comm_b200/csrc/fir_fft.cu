// fir_fft.cu -- complex64 / float32 FIR by overlap-save with a hand-written 4096-point FFT held in shared memory.
//
// For filters longer than the 256 taps the block-Toeplitz tensor-core kernel takes, a direct form costs O(K) per
// sample; overlap-save costs O(log N).  This is the engine behind the reference's FFT block filters
// sigsys.os_filter / oa_filter (src/sk_dsp_comm/sigsys.py:482-598), which compute the same causal FIR
// y = lfilter(h, 1, x) through N-point FFT frames, and behind multirate_FIR.filter
// (src/sk_dsp_comm/multirate_helper.py:104-109) for long complex64 filters.
//
// 256 threads per frame: 4096 input samples starting K-1 before the frame's first output, 4096-(K-1) valid outputs.
// 4096 = 16 x 16 x 16: three radix-16 passes, every thread computing one 16-point DFT in registers per pass.
// Forward = decimation in frequency (natural order in, digit-reversed out); the spectrum of the taps is stored in
// the SAME digit-reversed order (and pre-scaled by 1/4096), so the product needs no reordering; inverse =
// decimation in time (digit-reversed in, natural order out).  The last forward pass, the product and the first
// inverse pass work on the same 16 values of a thread, so a frame needs only four shared-memory exchanges, and two
// of those stay inside a half-warp (warp sync instead of a barrier).
// Shared-memory index p -> p + (p >> 4) (one pad word per 16) makes all three access patterns conflict free.
// Every table value a thread multiplies by -- spectrum, W_4096^(j t), W_256^(j (t & 15)) -- depends only on the
// thread index, so the tables are stored per thread, pair-interleaved for 128-bit loads, in shared memory (66 KB):
// read from a global table they are per-lane gathers, and the first version of this kernel was bound by exactly
// that (ncu: l1tex throughput 95 %).  A persistent CTA of 768 threads holds ONE copy of the tables and three
// frames in flight (three independent 256-thread groups on named barriers): 24 warps per SM, 168 KB.
// (Tried and dropped: 2 CTAs of 256 threads with the next frame prefetched in registers, 0.46 ms against 0.415 at
//  1024 taps; an L2 prefetch of the next frame, which only added spills.)
// float32 streams: the taps are real, so two consecutive real frames ride as the real and imaginary part of one
// complex transform.
// No cuFFT: the transform, the twiddle tables and the frame logic are all here.
#include "common.cuh"
#include <math.h>
#include <vector>

namespace b200dsp {
namespace fft {

constexpr int N = 4096;
constexpr int NT = 256;
constexpr int SM_FLOATS = N + N / 16;              // padded length of the frame buffer (float2 elements)
constexpr int T1_LEN = 4096, T2_LEN = 256;         // twiddle tables W_4096^(a b) [16][256] and W_256^(k n) [16][16]
constexpr int TABLE_LEN = N + T1_LEN + T2_LEN;     // float2 elements per plan: spectrum | T1 | T2
constexpr int GROUPS = 3;                          // frames in flight per CTA (256 threads each), sharing the tables
constexpr int SMEM_BYTES = (GROUPS * SM_FLOATS + TABLE_LEN) * 8;  // 3 x 34 KB frames + 66 KB tables = 168 KB
constexpr int CTAS_PER_SM = 1;

struct Args {
    const void *x;            // float2 (complex64) or float (REAL: two consecutive real frames ride as re / im)
    const void *hist;
    void *y;
    const float2 *H;          // TABLE_LEN values, each table [j >> 1][thread index][j & 1] for 128-bit loads:
                              // spectrum of the taps / 4096 (j = k2, index k0 * 16 + k1, k = k0 + 16 k1 + 256 k2),
                              // W_4096^(j b) (index b < 256), W_256^(j n) (index n < 16)
    int64_t frames;
    int64_t n;
    int32_t hist_len;
    int32_t ntaps;
    int32_t valid;            // outputs per frame = 4096 - (ntaps - 1)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// 4-point DFT in place: (a,b,c,d) -> (y0,y1,y2,y3), forward kernel exp(-2 pi i nk/4) (INV: conjugate)
template <bool INV>
__device__ __forceinline__ void fft4(float2 &a, float2 &b, float2 &c, float2 &d)
{
    const float2 s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = csub(b, d);
    // forward: -i * d1 = (d1.y, -d1.x); inverse: +i * d1 = (-d1.y, d1.x)
    const float2 jd = INV ? make_float2(-d1.y, d1.x) : make_float2(d1.y, -d1.x);
    a = cadd(s0, s1);
    c = csub(s0, s1);
    b = cadd(d0, jd);
    d = csub(d0, jd);
}

// 16-point DFT of v[0..15] in place; output X[k] ends up in v[4 (k & 3) + (k >> 2)]
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) fft4<INV>(v[j], v[j + 4], v[j + 8], v[j + 12]);
    // twiddles W16^(j q), q = row index after the first stage (v[j + 4 q])
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
    const float2 w1 = make_float2(C1, INV ? S1 : -S1), w2 = make_float2(R2, INV ? R2 : -R2), w3 = make_float2(S1, INV ? C1 : -C1);
    const float2 w6 = make_float2(-R2, INV ? R2 : -R2), w9 = make_float2(-C1, INV ? -S1 : S1);
    v[1 + 4] = cmul(v[1 + 4], w1);  v[2 + 4] = cmul(v[2 + 4], w2);  v[3 + 4] = cmul(v[3 + 4], w3);      // q = 1: j q = 1,2,3
    v[1 + 8] = cmul(v[1 + 8], w2);  v[2 + 8] = INV ? make_float2(-v[2 + 8].y, v[2 + 8].x) : make_float2(v[2 + 8].y, -v[2 + 8].x);  v[3 + 8] = cmul(v[3 + 8], w6);      // q = 2: 2,4,6
    v[1 + 12] = cmul(v[1 + 12], w3); v[2 + 12] = cmul(v[2 + 12], w6); v[3 + 12] = cmul(v[3 + 12], w9);  // q = 3: 3,6,9
#pragma unroll
    for (int q = 0; q < 4; ++q) fft4<INV>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// The transposed form: input x[j] taken from v[4 (j & 3) + (j >> 2)] (where fft16 leaves its outputs), output X[k]
// in v[k] -- so a forward transform, an element-wise product and an inverse transform chain in place.
template <bool INV>
__device__ __forceinline__ void fft16t(float2 (&v)[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) fft4<INV>(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    // element (column j, row q) now sits in v[4 j + q]; twiddle W16^(j q)
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
    const float2 w1 = make_float2(C1, INV ? S1 : -S1), w2 = make_float2(R2, INV ? R2 : -R2), w3 = make_float2(S1, INV ? C1 : -C1);
    const float2 w6 = make_float2(-R2, INV ? R2 : -R2), w9 = make_float2(-C1, INV ? -S1 : S1);
    v[4 + 1] = cmul(v[4 + 1], w1);  v[8 + 1] = cmul(v[8 + 1], w2);  v[12 + 1] = cmul(v[12 + 1], w3);
    v[4 + 2] = cmul(v[4 + 2], w2);
    v[8 + 2] = INV ? make_float2(-v[8 + 2].y, v[8 + 2].x) : make_float2(v[8 + 2].y, -v[8 + 2].x);
    v[12 + 2] = cmul(v[12 + 2], w6);
    v[4 + 3] = cmul(v[4 + 3], w3);  v[8 + 3] = cmul(v[8 + 3], w6);  v[12 + 3] = cmul(v[12 + 3], w9);
#pragma unroll
    for (int q = 0; q < 4; ++q) fft4<INV>(v[q], v[4 + q], v[8 + q], v[12 + q]);
}
// position of output k of fft16 inside v
__device__ __forceinline__ constexpr int o16(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ int pad(int p) { return p + (p >> 4); }

template <typename T>
__device__ __forceinline__ T load_sample(const Args &a, int64_t g)
{
    T z{};
    if (g >= 0) return (g < a.n) ? static_cast<const T *>(a.x)[g] : z;
    if (a.hist != nullptr) {
        const int64_t h = (int64_t)a.hist_len + g;
        if (h >= 0) return static_cast<const T *>(a.hist)[h];
    }
    return z;
}

// 16 table values of a thread, stored pair-interleaved ([j >> 1][thread][j & 1]): eight 128-bit loads
__device__ __forceinline__ void load_pairs(float2 (&w)[16], const float2 *base, int pair_stride)
{
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const float4 q = *reinterpret_cast<const float4 *>(base + (j >> 1) * pair_stride);
        w[j] = make_float2(q.x, q.y);
        w[j + 1] = make_float2(q.z, q.w);
    }
}

// The 16 samples a thread contributes to a frame.  REAL: the taps are real, so two real frames transform as one
// complex frame (re = frame 2f, im = frame 2f + 1) and come out of the inverse transform separated again.
template <bool REAL>
__device__ __forceinline__ void load_frame(const Args &a, int64_t frame, int t, float2 (&r)[16])
{
    if (!REAL) {
        const int64_t g0 = frame * a.valid - (a.ntaps - 1);        // first input sample of the frame
        if (g0 >= 0 && g0 + N <= a.n) {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = static_cast<const float2 *>(a.x)[g0 + t + 256 * j];
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = load_sample<float2>(a, g0 + t + 256 * j);
        }
    } else {
        const int64_t g0 = 2 * frame * a.valid - (a.ntaps - 1);
        if (g0 >= 0 && g0 + a.valid + N <= a.n) {
            const float *xa = static_cast<const float *>(a.x) + g0 + t;
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = make_float2(xa[256 * j], xa[256 * j + a.valid]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                r[j] = make_float2(load_sample<float>(a, g0 + t + 256 * j), load_sample<float>(a, g0 + a.valid + t + 256 * j));
        }
    }
}

// barrier over the 256 threads of one frame group (named barrier 1 + group)
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

template <bool REAL>
__global__ void __launch_bounds__(NT * GROUPS, CTAS_PER_SM) fir_fft_os_kernel(const Args a)
{
    extern __shared__ float2 smem_f2[];
    const int grp = threadIdx.x >> 8, t = threadIdx.x & 255;
    float2 *sm = smem_f2 + grp * SM_FLOATS;          // frame: (re, im) pairs, 64-bit accesses, conflict free per half-warp
    float2 *Hs = smem_f2 + GROUPS * SM_FLOATS;       // spectrum
    float2 *T1 = Hs + N;
    float2 *T2 = T1 + T1_LEN;
    for (int i = threadIdx.x; i < TABLE_LEN; i += NT * GROUPS) Hs[i] = a.H[i];
    float2 v[16];
    __syncthreads();
  for (int64_t frame = (int64_t)blockIdx.x * GROUPS + grp; frame < a.frames; frame += (int64_t)gridDim.x * GROUPS) {
    const int64_t out0 = (REAL ? 2 : 1) * frame * a.valid;         // first output of this frame (pair)
    // ---- forward pass 1 (over n2, stride 256): inputs straight from global memory (coalesced over t)
    load_frame<REAL>(a, frame, t, v);
    // (the tables are read into registers BEFORE each 16-point transform: the compiler cannot hoist a table load
    //  above the stores to the frame buffer by itself, and with 16 warps per SM an exposed LDS latency shows)
    float2 w[16];
    load_pairs(w, T1 + 2 * t, 512);                                // W_4096^(k0 t), k0 = 0..15
    fft16<false>(v);
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) sm[pad(k0 * 256 + t)] = (k0 == 0) ? v[o16(0)] : cmul(v[o16(k0)], w[k0]);
    group_sync(grp);
    // ---- forward pass 2 (over n1, stride 16) inside block k0
    {
        const int k0 = t >> 4, n0 = t & 15;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int p = pad(k0 * 256 + n1 * 16 + n0);
            v[n1] = sm[p];
        }
        load_pairs(w, T2 + 2 * n0, 32);                            // W_256^(k1 n0), k1 = 0..15
        fft16<false>(v);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {          // (a thread rewrites exactly the 16 positions it read: no barrier)
            const float2 r = (k1 == 0) ? v[o16(0)] : cmul(v[o16(k1)], w[k1]);
            const int p = pad(k0 * 256 + k1 * 16 + n0);
            sm[p] = r;
        }
    }
    __syncwarp();         // the 16 values thread (k0, k1) reads next were written by threads (k0, 0..15): same half-warp
    // ---- forward pass 3 (over n0), spectrum product, inverse pass 1 (over k2): all on the same 16 values
    {
#pragma unroll
        for (int n0 = 0; n0 < 16; ++n0) {
            const int p = pad(t * 16 + n0);
            v[n0] = sm[p];
        }
        load_pairs(w, Hs + 2 * t, 512);
        fft16<false>(v);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) v[o16(k2)] = cmul(v[o16(k2)], w[k2]);
        const int k1 = t & 15;
        load_pairs(w, T2 + 2 * k1, 32);                            // W_256^(k1 n0), n0 = 0..15, used conjugated
        fft16t<true>(v);                                           // in place: input where fft16 left it, output natural
#pragma unroll
        for (int n0 = 0; n0 < 16; ++n0) {
            const float2 r = (n0 == 0) ? v[0] : cmulc(v[n0], w[n0]);
            const int p = pad(t * 16 + n0);
            sm[p] = r;
        }
    }
    __syncwarp();         // same half-warp again
    // ---- inverse pass 2 (over k1) inside block k0
    {
        const int k0 = t >> 4, n0 = t & 15;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const int p = pad(k0 * 256 + k1 * 16 + n0);
            v[k1] = sm[p];
        }
        fft16<true>(v);
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) sm[pad(k0 * 256 + n1 * 16 + n0)] = v[o16(n1)];
    }
    group_sync(grp);
    // ---- inverse pass 3 (over k0, stride 256): natural order out; the first K-1 samples of the frame are aliased
    // (the twiddle conj(W_4096^(k0 (16 n1 + n0))) between inverse passes 2 and 3 is applied here, on the input side:
    //  with 16 n1 + n0 = t it is the same per-thread set W_4096^(k0 t) the first forward pass uses)
    load_pairs(w, T1 + 2 * t, 512);
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) {
        const float2 r = sm[pad(k0 * 256 + t)];
        v[k0] = (k0 == 0) ? r : cmulc(r, w[k0]);
    }
    fft16<true>(v);
    {
        // 32-bit bounds: q = position among this frame's outputs, lim = outputs the stream still has room for
        const int skip = a.ntaps - 1;
        const int64_t left = a.n - out0;
        const int lim = left > 2 * N ? 2 * N : (int)left;
        const int q0 = t - skip;
        if (!REAL) {
            float2 *yb = static_cast<float2 *>(a.y) + out0 + q0;
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int q = q0 + n2 * 256;
                if (q >= 0 && q < lim) yb[n2 * 256] = v[o16(n2)];
            }
        } else {
            float *yb = static_cast<float *>(a.y) + out0 + q0;
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int q = q0 + n2 * 256;
                if (q >= 0 && q < lim) yb[n2 * 256] = v[o16(n2)].x;
                if (q >= 0 && q + a.valid < lim) yb[n2 * 256 + a.valid] = v[o16(n2)].y;
            }
        }
    }
    // no barrier: forward pass 1 of the next frame writes exactly the positions this thread has just read
  }
}

}  // namespace fft

// ------------------------------------------------------------------------------------------ host
// Tables for one plan (fft::TABLE_LEN complex values): the spectrum of the taps / 4096 in the kernel's order, then
// the twiddles W_4096^(a b) as [16][256] and W_256^(k n) as [16][16].
int fft_table_floats() { return 2 * fft::TABLE_LEN; }

int fft_build_tables(const double *taps, int ntaps, float *out /* fft_table_floats() floats */)
{
    using namespace fft;
    if (ntaps < 2 || ntaps - 1 > N / 2) return -1;
    const double PI2 = 6.283185307179586476925286766559;
    // plain O(N K) DFT of the zero-padded taps in float64 (K <= 2049: 8 M complex MACs, once per plan)
    std::vector<double> c(N), s(N);
    for (int j = 0; j < N; ++j) {
        c[j] = cos(PI2 * j / N);
        s[j] = -sin(PI2 * j / N);
    }
    for (int k = 0; k < N; ++k) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n < ntaps; ++n) {
            const int idx = (int)(((long long)k * n) & (N - 1));
            re += taps[n] * c[idx];
            im += taps[n] * s[idx];
        }
        const int k0 = k & 15, k1 = (k >> 4) & 15, k2 = k >> 8;
        const int p = (k2 >> 1) * 512 + (k0 * 16 + k1) * 2 + (k2 & 1);
        out[2 * p] = (float)(re / N);
        out[2 * p + 1] = (float)(im / N);
    }
    for (int a = 0; a < 16; ++a)
        for (int b = 0; b < 256; ++b) {
            const int p = N + (a >> 1) * 512 + b * 2 + (a & 1);
            out[2 * p] = (float)c[(a * b) & (N - 1)];
            out[2 * p + 1] = (float)s[(a * b) & (N - 1)];
        }
    for (int k = 0; k < 16; ++k)
        for (int n = 0; n < 16; ++n) {
            const int p = N + T1_LEN + (k >> 1) * 32 + n * 2 + (k & 1);
            out[2 * p] = (float)c[(16 * k * n) & (N - 1)];
            out[2 * p + 1] = (float)s[(16 * k * n) & (N - 1)];
        }
    return 0;
}

int launch_fir_fft(bool real, const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                   const void *tables_dev, int ntaps, int sm_count, cudaStream_t stream)
{
    using namespace fft;
    Args a;
    a.x = x;
    a.hist = hist;
    a.y = y;
    a.H = static_cast<const float2 *>(tables_dev);
    a.n = n;
    a.hist_len = hist_len;
    a.ntaps = ntaps;
    a.valid = N - (ntaps - 1);
    const int64_t per_frame = (real ? 2 : 1) * (int64_t)a.valid;
    a.frames = (n + per_frame - 1) / per_frame;
    auto kern = real ? fir_fft_os_kernel<true> : fir_fft_os_kernel<false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);   // per device
    const int64_t want = (a.frames + GROUPS - 1) / GROUPS, resident = (int64_t)sm_count * CTAS_PER_SM;
    const unsigned grid = (unsigned)(want < resident ? want : resident);
    kern<<<grid, NT * GROUPS, SMEM_BYTES, stream>>>(a);
    B200_CHECK_LAUNCH("fir_fft_os_kernel");
    return B200DSP_OK;
}

}  // namespace b200dsp
