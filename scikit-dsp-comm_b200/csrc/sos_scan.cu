// sos_scan.cu -- second-order-section IIR cascade as a parallel-prefix recurrence (sm_100a).
//
// Replaces scipy.signal.sosfilt as called by the reference's multirate_IIR
// (src/sk_dsp_comm/multirate_helper.py:169-192).  sosfilt is a strictly sequential
// direct-form-II-transposed loop; here the same recurrence is evaluated in
// three launches per group of <= 8 sections.  The cascade is LTI (s[n+1] = A s[n] + B x[n]), so only
// state VECTORS are ever scanned, with constant host-computed matrices A^(256 * 2^j):
//
//   K1 sos_pass1   every thread streams its own chunk of 256 consecutive samples straight from global
//                  memory (128-bit loads; the 32 lanes of a warp walk 32 different 128-byte lines, L1 keeps
//                  the lines between a thread's consecutive loads -- no shared-memory staging, no
//                  transposition, DRAM sectors fully used) and runs the cascade from ZERO state: the
//                  2*nsec final state values are the "chunk carry".  A block-wide Kogge-Stone scan over
//                  the 256 carries of a tile (65536 samples) yields the tile carry.
//   K2 sos_tile_scan   one block: scans the tile carries with A^65536 (runs of tiles per thread +
//                  Kogge-Stone over the runs) -> true state at every tile start (fp64).
//   K3 sos_pass2   rescans the stored chunk carries with the tile's true start state folded in, then
//                  every thread re-runs the cascade over its chunk from its TRUE start state and streams
//                  the outputs back with 128-bit stores.
//
// Up-sampling (zero stuffing, x L gain) is fused into the chunk read and down-sampling into the
// write (element-wise path), so multirate_IIR.up/.dn never materialise the full-rate stream.
#include "common.cuh"
#include "sos_tc.cuh"
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

namespace b200dsp {

constexpr int SOS_NT = 256;          // threads per block = chunks per tile
constexpr int SOS_LEVELS = 8;        // log2(SOS_NT)
constexpr int SOS_LC = 256;          // samples per chunk (one thread)
constexpr int SOS_MAXSEC = 8;        // sections per launch group
constexpr int SOS_K2_NT = 256;       // threads in the tile-scan block
constexpr int SOS_K2_LEVELS = 8;

template <typename C, int NSEC> struct SosCoef { C c[NSEC][5]; };   // b0 b1 b2 -a1 -a2

template <typename C, int NSEC>
__device__ __forceinline__ C sos_step(const SosCoef<C, NSEC> &k, C (&z)[2 * NSEC], C v)
{
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        C xn = fma(k.c[s][0], v, z[2 * s]);
        z[2 * s] = fma(k.c[s][1], v, fma(k.c[s][3], xn, z[2 * s + 1]));
        z[2 * s + 1] = fma(k.c[s][2], v, k.c[s][4] * xn);
        v = xn;
    }
    return v;
}

__device__ __forceinline__ float  ch_get(const float &v, int)   { return v; }
__device__ __forceinline__ double ch_get(const double &v, int)  { return v; }
__device__ __forceinline__ float  ch_get(const float2 &v, int c)  { return c ? v.y : v.x; }
__device__ __forceinline__ double ch_get(const double2 &v, int c) { return c ? v.y : v.x; }
__device__ __forceinline__ void ch_set(float &v, int, float x)    { v = x; }
__device__ __forceinline__ void ch_set(double &v, int, double x)  { v = x; }
__device__ __forceinline__ void ch_set(float2 &v, int c, float x)   { if (c) v.y = x; else v.x = x; }
__device__ __forceinline__ void ch_set(double2 &v, int c, double x) { if (c) v.y = x; else v.x = x; }

// e <- e + P * o with P block-lower-triangular (section i never depends on a later section)
template <typename C, int D>
__device__ __forceinline__ void matvec_acc(C (&e)[D], const C *__restrict__ P, const C (&o)[D])
{
#pragma unroll
    for (int i = 0; i < D; ++i) {
        C acc = e[i];
#pragma unroll
        for (int k = 0; k <= (i | 1); ++k) acc = fma(P[i * D + k], o[k], acc);
        e[i] = acc;
    }
}

// Inclusive Kogge-Stone scan of the per-thread carries of ONE channel:
//   e_t <- sum_{u<=t} A^(LC*(t-u)) e_u .  mats = [SOS_LEVELS][D*D] in shared memory,
//   xch = SOS_NT*D scratch in shared memory.
template <typename C, int D>
__device__ void scan_carries(C (&e)[D], const C *mats, C *xch, int tid)
{
    // block-wide Kogge-Stone: every level goes through shared memory because the partner
    // tid - 2^j lives in the previous warp for the low lanes (a warp-shuffle phase would stop
    // at the warp boundary and leave the prefix incomplete).
    for (int j = 0; j < SOS_LEVELS; ++j) {
        const int off = 1 << j;
        __syncthreads();
#pragma unroll
        for (int d = 0; d < D; ++d) xch[d * SOS_NT + tid] = e[d];
        __syncthreads();
        if (tid >= off) {
            C o[D];
#pragma unroll
            for (int d = 0; d < D; ++d) o[d] = xch[d * SOS_NT + tid - off];
            matvec_acc<C, D>(e, mats + j * D * D, o);
        }
    }
}

template <typename S, int NSEC>
struct SosArgs {
    using C = typename Sample<S>::C;
    const S *x;
    S *y;
    int64_t n_in;        // samples readable from x
    int64_t n_rate;      // samples at the filter rate (n_in * L)
    int64_t n_out;       // samples written to y (n_rate / M)
    int32_t L, M;
    const C *mats;       // [SOS_LEVELS][D*D] device, A^(LC*2^j)
    C *aggr;             // [tiles][NCH][D]   tile carries (zero tile-start state)
    C *start;            // [tiles][NCH][D]   true state at tile start (from K2)
    C *carry;            // [tiles][NCH][D][SOS_NT] chunk carries
    C *zf;               // final state out [s][2][ch] or NULL
    int32_t d_real;      // 2 * (sections that are not identity padding)
    SosCoef<C, NSEC> k;
};

// ---- streaming access to one thread's chunk -----------------------------------------------------
// Input element of filter-rate index g (fused zero stuffing for L > 1).
template <typename S, int NSEC>
struct ChunkReader {
    using C = typename Sample<S>::C;
    const SosArgs<S, NSEC> &a;
    int64_t q;     // input index of the current filter-rate sample (L > 1)
    int r;         // phase within the zero-stuffing period
    __device__ ChunkReader(const SosArgs<S, NSEC> &a_, int64_t g0) : a(a_) {
        q = (a.L == 1) ? g0 : g0 / a.L;
        r = (a.L == 1) ? 0 : (int)(g0 - q * a.L);
    }
    __device__ __forceinline__ S next(int64_t g) {
        S v = zero_of(S());
        if (a.L == 1) {
            if (g < a.n_rate) v = a.x[g];
        } else {
            if (r == 0 && q < a.n_in) v = scale_of(a.x[q], (C)a.L);
            if (++r == a.L) { r = 0; ++q; }
        }
        return v;
    }
};

// ---- per-warp transposition buffers (fast path) ---------------------------------------------------
// A warp owns 32 consecutive chunks.  Sub-block sb = the sb-th 128-byte piece of every chunk: 32 rows of
// 128 B, each row contiguous in global memory.  cp.async moves it with fully coalesced 16-byte copies
// (8 lanes per row, 4 rows per instruction) into a per-warp shared-memory buffer with a 144-byte row
// pitch, so the per-lane 128-bit reads of "my row" are bank-conflict free.  Two buffers: the copy of
// sub-block sb+1 overlaps the recurrence over sub-block sb.  (Lane-strided 128-bit global accesses cost
// 32 L1 wavefronts per instruction; this path costs 4.)
constexpr int SOS_ROW_UNITS = 9;                 // 8 x 16 B of data + 16 B pad per row
constexpr int SOS_WBUF_BYTES = 2 * 32 * SOS_ROW_UNITS * 16;      // two stages per warp

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename S> struct VecOf {
    static constexpr int V = 16 / (int)sizeof(S);
    struct __align__(16) T { S v[16 / sizeof(S)]; };
};

template <typename S, int NSEC>
__global__ void __launch_bounds__(SOS_NT) sos_pass1_kernel(const SosArgs<S, NSEC> a)
{
    using C = typename Sample<S>::C;
    constexpr int NCH = Sample<S>::NCH;
    constexpr int D = 2 * NSEC;
    constexpr int V = VecOf<S>::V;
    using Vec = typename VecOf<S>::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *mats = reinterpret_cast<C *>(smem_raw);
    C *xch = mats + SOS_LEVELS * D * D;
    const int tid = threadIdx.x;
    const int64_t g0 = ((int64_t)blockIdx.x * SOS_NT + tid) * SOS_LC;

    for (int i = tid; i < SOS_LEVELS * D * D; i += SOS_NT) mats[i] = a.mats[i];

    C z[NCH][D];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int d = 0; d < D; ++d) z[c][d] = (C)0;

    const int lane = tid & 31, warp = tid >> 5;
    const int64_t wg0 = g0 - (int64_t)lane * SOS_LC;              // first sample of this warp's 32 chunks
    const bool fast = (a.L == 1) && (wg0 + 32 * (int64_t)SOS_LC <= a.n_rate) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
    if (fast) {
        Vec *wbuf = reinterpret_cast<Vec *>(smem_raw + sizeof(C) * ((size_t)SOS_LEVELS * D * D + (size_t)SOS_NT * D)) +
                    (size_t)warp * (SOS_WBUF_BYTES / 16);
        const Vec *wx = reinterpret_cast<const Vec *>(a.x + wg0);
        constexpr int UPR = SOS_LC / V;                  // 16-byte units per chunk (row pitch in global memory)
        constexpr int NSB = UPR / 8;                     // sub-blocks of 8 units (128 B) per chunk
        auto issue = [&](int sb, int stage) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = 4 * k + (lane >> 3), unit = lane & 7;
                cp_async16(wbuf + (stage * 32 + row) * SOS_ROW_UNITS + unit, wx + (size_t)row * UPR + sb * 8 + unit);
            }
            cp_async_commit();
        };
        issue(0, 0);
#pragma unroll 1
        for (int sb = 0; sb < NSB; ++sb) {
            const int stage = sb & 1;
            if (sb + 1 < NSB) { issue(sb + 1, stage ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            __syncwarp();
            const Vec *row = wbuf + (stage * 32 + lane) * SOS_ROW_UNITS;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const Vec vv = row[u];
#pragma unroll
                for (int e = 0; e < V; ++e)
#pragma unroll
                    for (int c = 0; c < NCH; ++c) sos_step<C, NSEC>(a.k, z[c], ch_get(vv.v[e], c));
            }
            __syncwarp();                                // buffer `stage` is free for sub-block sb+2
        }
    } else if (g0 < a.n_rate) {
        ChunkReader<S, NSEC> rd(a, g0);
#pragma unroll 1
        for (int i = 0; i < SOS_LC; ++i) {
            const S v = rd.next(g0 + i);
#pragma unroll
            for (int c = 0; c < NCH; ++c) sos_step<C, NSEC>(a.k, z[c], ch_get(v, c));
        }
    }
    // chunk carries -> workspace (coalesced over tid)
    C *cw = a.carry + (size_t)blockIdx.x * NCH * D * SOS_NT;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int d = 0; d < D; ++d) cw[(c * D + d) * SOS_NT + tid] = z[c][d];
    __syncthreads();                                   // mats staged
    // tile carry = last element of the inclusive scan
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        scan_carries<C, D>(z[c], mats, xch, tid);
        if (tid == SOS_NT - 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) a.aggr[((size_t)blockIdx.x * NCH + c) * D + d] = z[c][d];
        }
    }
}

template <typename S, int NSEC>
__global__ void __launch_bounds__(SOS_NT) sos_pass2_kernel(const SosArgs<S, NSEC> a)
{
    using C = typename Sample<S>::C;
    constexpr int NCH = Sample<S>::NCH;
    constexpr int D = 2 * NSEC;
    constexpr int V = VecOf<S>::V;
    using Vec = typename VecOf<S>::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *mats = reinterpret_cast<C *>(smem_raw);
    C *xch = mats + SOS_LEVELS * D * D;
    const int tid = threadIdx.x;
    const int64_t g0 = ((int64_t)blockIdx.x * SOS_NT + tid) * SOS_LC;

    for (int i = tid; i < SOS_LEVELS * D * D; i += SOS_NT) mats[i] = a.mats[i];

    C z[NCH][D];
    const C *cw = a.carry + (size_t)blockIdx.x * NCH * D * SOS_NT;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int d = 0; d < D; ++d) z[c][d] = cw[(c * D + d) * SOS_NT + tid];
    __syncthreads();

#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        C s0[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s0[d] = a.start[((size_t)blockIdx.x * NCH + c) * D + d];
        if (tid == 0) matvec_acc<C, D>(z[c], mats, s0);       // fold the tile's true start state
        scan_carries<C, D>(z[c], mats, xch, tid);             // z = state at END of chunk tid
        __syncthreads();
#pragma unroll
        for (int d = 0; d < D; ++d) xch[d * SOS_NT + tid] = z[c][d];
        __syncthreads();
#pragma unroll
        for (int d = 0; d < D; ++d) z[c][d] = (tid == 0) ? s0[d] : xch[d * SOS_NT + tid - 1];
    }

    // corrected pass: true start state -> outputs
    const int64_t last = a.n_rate - 1 - g0;            // position of the stream's final sample in this chunk
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t wg0 = g0 - (int64_t)lane * SOS_LC;
    const int64_t wlast = a.n_rate - 1 - wg0;          // final sample relative to the warp's range
    const bool fast = (a.L == 1) && (a.M == 1) && (wg0 + 32 * (int64_t)SOS_LC <= a.n_rate) &&
                      (a.zf == nullptr || wlast >= 32 * (int64_t)SOS_LC || wlast < 0) &&
                      (((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.y)) & 15) == 0);
    if (fast) {
        Vec *wbuf = reinterpret_cast<Vec *>(smem_raw + sizeof(C) * ((size_t)SOS_LEVELS * D * D + (size_t)SOS_NT * D)) +
                    (size_t)warp * (SOS_WBUF_BYTES / 16);
        const Vec *wx = reinterpret_cast<const Vec *>(a.x + wg0);
        Vec *wy = reinterpret_cast<Vec *>(a.y + wg0);
        constexpr int UPR = SOS_LC / V;
        constexpr int NSB = UPR / 8;
        auto issue = [&](int sb, int stage) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = 4 * k + (lane >> 3), unit = lane & 7;
                cp_async16(wbuf + (stage * 32 + row) * SOS_ROW_UNITS + unit, wx + (size_t)row * UPR + sb * 8 + unit);
            }
            cp_async_commit();
        };
        issue(0, 0);
#pragma unroll 1
        for (int sb = 0; sb < NSB; ++sb) {
            const int stage = sb & 1;
            if (sb + 1 < NSB) { issue(sb + 1, stage ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            __syncwarp();
            Vec *row = wbuf + (stage * 32 + lane) * SOS_ROW_UNITS;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const Vec vv = row[u];
                Vec o;
#pragma unroll
                for (int e = 0; e < V; ++e)
#pragma unroll
                    for (int c = 0; c < NCH; ++c) ch_set(o.v[e], c, sos_step<C, NSEC>(a.k, z[c], ch_get(vv.v[e], c)));
                row[u] = o;                              // outputs replace the inputs in the warp buffer
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {                // coalesced write-back: 8 lanes per 128-byte row
                const int r2 = 4 * k + (lane >> 3), unit = lane & 7;
                wy[(size_t)r2 * UPR + sb * 8 + unit] = wbuf[(stage * 32 + r2) * SOS_ROW_UNITS + unit];
            }
            __syncwarp();
        }
    } else if (g0 < a.n_rate) {
        ChunkReader<S, NSEC> rd(a, g0);
        // fused decimation: output index / phase of the current filter-rate sample
        int64_t mo = (a.M == 1) ? g0 : g0 / a.M;
        int mr = (a.M == 1) ? 0 : (int)(g0 - mo * a.M);
#pragma unroll 1
        for (int i = 0; i < SOS_LC; ++i) {
            const int64_t g = g0 + i;
            if (g >= a.n_rate) break;
            const S v = rd.next(g);
            S o;
#pragma unroll
            for (int c = 0; c < NCH; ++c) ch_set(o, c, sos_step<C, NSEC>(a.k, z[c], ch_get(v, c)));
            if (mr == 0 && mo < a.n_out) a.y[mo] = o;
            if (a.M == 1) { ++mo; } else if (++mr == a.M) { mr = 0; ++mo; }
            if (a.zf != nullptr && i == last) {
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int d = 0; d < D; ++d)
                        if (d < a.d_real) a.zf[d * NCH + c] = z[c][d];
            }
        }
    }
}

// ---- K2: scan of tile carries -----------------------------------------------------------
template <int D> struct TileScanMats { double m1[D * D]; double pj[SOS_K2_LEVELS][D * D]; };

template <int D>
__device__ __forceinline__ void matvec_full(double (&out)[D], const double *P, const double (&in)[D])
{
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fma(P[i * D + k], in[k], acc);
        out[i] = acc;
    }
}

template <typename C, int D>
__global__ void __launch_bounds__(SOS_K2_NT)
sos_tile_scan_kernel(const C *__restrict__ aggr, C *__restrict__ start, const C *__restrict__ zi,
                     int64_t n_tiles, int64_t run, int nch, int d_real, const TileScanMats<D> mm)
{
    extern __shared__ double xchd[];             // [D][SOS_K2_NT]
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)tid * run;
    const int64_t t1 = (t0 + run < n_tiles) ? t0 + run : n_tiles;
    for (int c = 0; c < nch; ++c) {
        double s[D], e[D], tmp[D];
        // initial state of the whole stream (scipy zi layout [s][2][ch] == [d][ch])
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = (zi != nullptr && d < d_real) ? (double)zi[d * nch + c] : 0.0;
        // local run carry (zero start, except run 0 which starts from zi)
#pragma unroll
        for (int d = 0; d < D; ++d) e[d] = (tid == 0) ? s[d] : 0.0;
        for (int64_t k = t0; k < t1; ++k) {
            matvec_full<D>(tmp, mm.m1, e);
#pragma unroll
            for (int d = 0; d < D; ++d) e[d] = tmp[d] + (double)aggr[(k * nch + c) * D + d];
        }
        // Kogge-Stone over runs
        for (int j = 0; j < SOS_K2_LEVELS; ++j) {
            const int off = 1 << j;
            __syncthreads();
#pragma unroll
            for (int d = 0; d < D; ++d) xchd[d * SOS_K2_NT + tid] = e[d];
            __syncthreads();
            if (tid >= off) {
                double o[D];
#pragma unroll
                for (int d = 0; d < D; ++d) o[d] = xchd[d * SOS_K2_NT + tid - off];
                matvec_full<D>(tmp, mm.pj[j], o);
#pragma unroll
                for (int d = 0; d < D; ++d) e[d] += tmp[d];
            }
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < D; ++d) xchd[d * SOS_K2_NT + tid] = e[d];
        __syncthreads();
        if (tid > 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) s[d] = xchd[d * SOS_K2_NT + tid - 1];
        }
        for (int64_t k = t0; k < t1; ++k) {
#pragma unroll
            for (int d = 0; d < D; ++d) start[(k * nch + c) * D + d] = (C)s[d];
            matvec_full<D>(tmp, mm.m1, s);
#pragma unroll
            for (int d = 0; d < D; ++d) s[d] = tmp[d] + (double)aggr[(k * nch + c) * D + d];
        }
        __syncthreads();
    }
}

// ---- host side -----------------------------------------------------------------------------
struct SosGroup {
    int nsec;                          // sections in this launch group incl. identity padding (1,2,4,6,8)
    int nsec_real;                     // sections that came from the user's sos
    double coef[SOS_MAXSEC][5];        // b0 b1 b2 -a1 -a2
    std::vector<double> A;             // D x D one-sample state transition
    std::vector<double> tileA;         // A^(SOS_NT*SOS_LC): one tile step
    float *mats_f32;                   // device [SOS_LEVELS][D*D]: A^(SOS_LC * 2^j)
    double *mats_f64;
    StcTables tc;                      // tensor-core single-pass kernel tables (float32 streams, sos_tc.cu)
    // tile-level scan matrices (A^T)^(run 2^j) per distinct `run` (= tiles per thread of the tile-scan block, a
    // function of the call length): computed once, then read-only.  shared_ptr-free: the map lives as long as the plan.
    mutable std::map<int64_t, std::vector<double>> k2_cache;
};

struct b200dsp_sos_plan_impl {
    int nsec;
    int sm_count;
    std::vector<SosGroup> groups;
    mutable std::mutex k2_mutex;       // guards the groups' k2_cache (first call of a new length only)
};

// 0 auto, 1 force the 3-kernel scan path, 2 force the tensor-core kernel whenever its tables exist
static thread_local int g_sos_variant = 0;

static void matmul(const std::vector<double> &a, const std::vector<double> &b, std::vector<double> &c, int D)
{
    std::vector<double> r((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) {
            double aik = a[i * D + k];
            if (aik == 0.0) continue;
            for (int j = 0; j < D; ++j) r[i * D + j] += aik * b[k * D + j];
        }
    c.swap(r);
}

static void matpow(const std::vector<double> &a, int64_t p, std::vector<double> &out, int D)
{
    std::vector<double> res((size_t)D * D, 0.0), base = a;
    for (int i = 0; i < D; ++i) res[i * D + i] = 1.0;
    while (p > 0) {
        if (p & 1) matmul(res, base, res, D);
        p >>= 1;
        if (p) matmul(base, base, base, D);
    }
    out.swap(res);
}

template <typename S, int NSEC>
static int run_group(const SosGroup &g, std::mutex *k2_mutex, const S *x, S *y, int64_t n_in, int64_t n_rate, int64_t n_out,
                     int32_t L, int32_t M, const void *zi, void *zf, unsigned char *ws,
                     cudaStream_t stream)
{
    using C = typename Sample<S>::C;
    constexpr int NCH = Sample<S>::NCH;
    constexpr int D = 2 * NSEC;
    constexpr int64_t T = (int64_t)SOS_NT * SOS_LC;
    const int64_t n_tiles = (n_rate + T - 1) / T;
    if (n_tiles > 2147483647LL) {
        set_error("sos: too many tiles");
        return B200DSP_E_UNSUPPORTED;
    }
    SosArgs<S, NSEC> a;
    a.x = x;
    a.y = y;
    a.n_in = n_in;
    a.n_rate = n_rate;
    a.n_out = n_out;
    a.L = L;
    a.M = M;
    a.mats = sizeof(C) == 4 ? reinterpret_cast<const C *>(g.mats_f32)
                            : reinterpret_cast<const C *>(g.mats_f64);
    size_t per_tile = (size_t)NCH * D * sizeof(C);
    size_t off = 0;
    a.aggr = reinterpret_cast<C *>(ws + off);
    off += ((size_t)n_tiles * per_tile + 255) & ~(size_t)255;
    a.start = reinterpret_cast<C *>(ws + off);
    off += ((size_t)n_tiles * per_tile + 255) & ~(size_t)255;
    a.carry = reinterpret_cast<C *>(ws + off);
    a.zf = static_cast<C *>(zf);
    a.d_real = 2 * g.nsec_real;
    for (int s = 0; s < NSEC; ++s)
        for (int q = 0; q < 5; ++q) a.k.c[s][q] = (C)g.coef[s][q];

    const size_t mats_bytes = sizeof(C) * (size_t)SOS_LEVELS * D * D;
    const size_t xch_bytes = sizeof(C) * (size_t)SOS_NT * D;
    const size_t smem1 = mats_bytes + xch_bytes + (size_t)(SOS_NT / 32) * SOS_WBUF_BYTES, smem3 = smem1;
    auto k1 = sos_pass1_kernel<S, NSEC>;
    auto k3 = sos_pass2_kernel<S, NSEC>;
    B200_CHECK_CUDA(allow_smem(k1, smem1));
    B200_CHECK_CUDA(allow_smem(k3, smem3));

    k1<<<(unsigned)n_tiles, SOS_NT, smem1, stream>>>(a);
    B200_CHECK_LAUNCH("sos_pass1_kernel");

    // tile-level scan matrices: m1 = A^T, pj = (A^T)^(run*2^j) -- from the plan's cache (host matrix powers are
    // computed once per distinct call length, not per call)
    TileScanMats<D> mm;
    const int64_t run = (n_tiles + SOS_K2_NT - 1) / SOS_K2_NT;
    memcpy(mm.m1, g.tileA.data(), sizeof(double) * D * D);
    {
        std::lock_guard<std::mutex> lock(*k2_mutex);
        auto it = g.k2_cache.find(run);
        if (it == g.k2_cache.end()) {
            std::vector<double> all((size_t)SOS_K2_LEVELS * D * D), p;
            matpow(g.tileA, run, p, D);
            for (int j = 0; j < SOS_K2_LEVELS; ++j) {
                memcpy(all.data() + (size_t)j * D * D, p.data(), sizeof(double) * D * D);
                if (j + 1 < SOS_K2_LEVELS) matmul(p, p, p, D);
            }
            if (g.k2_cache.size() > 64) g.k2_cache.clear();
            it = g.k2_cache.emplace(run, std::move(all)).first;
        }
        for (int j = 0; j < SOS_K2_LEVELS; ++j) memcpy(mm.pj[j], it->second.data() + (size_t)j * D * D, sizeof(double) * D * D);
    }
    auto k2 = sos_tile_scan_kernel<C, D>;
    size_t smem2 = sizeof(double) * (size_t)D * SOS_K2_NT;
    k2<<<1, SOS_K2_NT, smem2, stream>>>(a.aggr, a.start, static_cast<const C *>(zi), n_tiles, run, NCH, a.d_real, mm);
    B200_CHECK_LAUNCH("sos_tile_scan_kernel");

    k3<<<(unsigned)n_tiles, SOS_NT, smem3, stream>>>(a);
    B200_CHECK_LAUNCH("sos_pass2_kernel");
    return B200DSP_OK;
}

template <typename S>
static int run_group_nsec(const SosGroup &g, std::mutex *k2m, const S *x, S *y, int64_t n_in, int64_t n_rate,
                          int64_t n_out, int32_t L, int32_t M, const void *zi, void *zf,
                          unsigned char *ws, cudaStream_t st)
{
    switch (g.nsec) {
    case 1: return run_group<S, 1>(g, k2m, x, y, n_in, n_rate, n_out, L, M, zi, zf, ws, st);
    case 2: return run_group<S, 2>(g, k2m, x, y, n_in, n_rate, n_out, L, M, zi, zf, ws, st);
    case 4: return run_group<S, 4>(g, k2m, x, y, n_in, n_rate, n_out, L, M, zi, zf, ws, st);
    case 6: return run_group<S, 6>(g, k2m, x, y, n_in, n_rate, n_out, L, M, zi, zf, ws, st);
    case 8: return run_group<S, 8>(g, k2m, x, y, n_in, n_rate, n_out, L, M, zi, zf, ws, st);
    }
    set_error("sos: bad group size %d", g.nsec);
    return B200DSP_E_BADARG;
}

static size_t dtype_size(int dtype)
{
    switch (dtype) {
    case B200DSP_F32: return 4;
    case B200DSP_F64: return 8;
    case B200DSP_C64: return 8;
    case B200DSP_C128: return 16;
    }
    return 0;
}

// workspace = [scan area (aggr, start, carry) for the largest group] [tmp stream if >1 group]
static size_t sos_scan_area_bytes(int dtype, int64_t n_rate)
{
    const int lc = SOS_LC;
    const int nch = (dtype == B200DSP_C64 || dtype == B200DSP_C128) ? 2 : 1;
    const size_t csz = (dtype == B200DSP_F32 || dtype == B200DSP_C64) ? 4 : 8;
    const int64_t T = (int64_t)SOS_NT * lc;
    const int64_t n_tiles = (n_rate + T - 1) / T;
    const size_t per_tile = (size_t)nch * 2 * SOS_MAXSEC * csz;
    size_t a = ((size_t)n_tiles * per_tile + 255) & ~(size_t)255;
    return 2 * a + (size_t)n_tiles * per_tile * SOS_NT + 256;
}


// resample.cu
int resample_up_scaled(int dtype, const void *x, void *y, int64_t n, int32_t L, double gain, cudaStream_t st);
int resample_dn(int dtype, const void *x, void *y, int64_t n_out, int32_t M, cudaStream_t st);

constexpr int64_t SOS_STAGE_MIN = 1 << 16;     // full-rate staging pays off above this many filter-rate samples

static bool sos_needs_tmp(size_t ngroups, int64_t n_rate, int32_t L, int32_t M)
{
    return ngroups > 1 || ((L > 1 || M > 1) && n_rate >= SOS_STAGE_MIN);
}

// Long rate-changing calls run the cascade on the vectorised full-rate path: the (L x gain) zero-stuffed
// input or the undecimated output is staged once in the workspace by the streaming index-map kernels
// (an IIR has to run at the full rate anyway); short calls use the fused element-wise path.
static bool sos_tc_path(const b200dsp_sos_plan_impl *p, int dtype, int64_t n_rate, int32_t M)
{
    if (dtype != B200DSP_F32 || g_sos_variant == 1) return false;
    if (M > 1 && p->groups.size() > 2) return false;      // full-rate intermediates would need two scratch streams
    // Cascades with poles very close to the unit circle (more than one warm-up tile) always take the tensor-core
    // kernel when they can: its in-tile scan runs in float64 and the zero-state part is a direct sum, whereas the
    // float32 recurrence of the scan kernels loses the 1e-4 bar there (ten-band equaliser: 1.1e-4 vs 4e-6).
    bool force = g_sos_variant == 2;
    for (const SosGroup &g : p->groups)
        if (g.tc.ok && g.tc.warm_tiles > 1) force = true;
    for (const SosGroup &g : p->groups)
        if (!stc_usable(g.tc, n_rate, p->sm_count, force)) return false;
    return true;
}

// float32 streams whose cascade decays within a few tiles: one single-pass tensor-core launch per group of
// <= 8 sections, zero stuffing fused into the first group's tile load and decimation into the last group's stores
static int sos_run_tc(const b200dsp_sos_plan_impl *p, const float *x, float *y, int64_t n, int32_t L, int32_t M,
                      const void *zi, void *zf, unsigned char *ws, cudaStream_t st)
{
    const int64_t n_rate = n * L, n_out = n_rate / M;
    const size_t ng = p->groups.size();
    float *tmp = reinterpret_cast<float *>(ws);
    // A long interpolating call with a single group stages the (x L) zero-stuffed stream once in the workspace with
    // the streaming index-map kernel and runs the cascade on bulk-TMA tiles (the fused zero stuffing has the
    // converter warps fetch every tile themselves: fine for short calls, 3x slower on long ones).
    const float *src0 = x;
    int64_t n_in0 = n;
    int32_t L0 = L;
    if (L > 1 && ng == 1 && n_rate >= SOS_STAGE_MIN) {
        int rc = resample_up_scaled(B200DSP_F32, x, tmp, n, L, (double)L, st);
        if (rc != B200DSP_OK) return rc;
        src0 = tmp;
        n_in0 = n_rate;
        L0 = 1;
    }
    size_t sec0 = 0;
    for (size_t gi = 0; gi < ng; ++gi) {
        const SosGroup &g = p->groups[gi];
        const bool first = gi == 0, last = gi + 1 == ng;
        // ping-pong between y and the workspace so that the last group lands in y (never in place: a block's
        // warm-up tiles are read from its neighbour's range)
        auto out_of = [&](size_t k) -> float * { return ((ng - 1 - k) % 2 == 0 && (M == 1 || k + 1 == ng)) ? y : tmp; };
        const float *src = first ? src0 : out_of(gi - 1);
        float *dst = out_of(gi);
        const float *zig = zi ? static_cast<const float *>(zi) + sec0 * 2 : nullptr;
        float *zfg = zf ? static_cast<float *>(zf) + sec0 * 2 : nullptr;
        int rc = launch_sos_tc(g.tc, src, dst, first ? n_in0 : n_rate, n_rate, last ? n_out : n_rate, first ? L0 : 1,
                               last ? M : 1, zig, zfg, p->sm_count, st);
        if (rc != B200DSP_OK) return rc;
        sec0 += g.nsec_real;
    }
    return B200DSP_OK;
}

template <typename S>
static int sos_run(const b200dsp_sos_plan_impl *p, const S *x, S *y, int64_t n, int32_t L, int32_t M,
                   const void *zi, void *zf, unsigned char *ws, int dtype, cudaStream_t st)
{
    using C = typename Sample<S>::C;
    constexpr int NCH = Sample<S>::NCH;
    const int64_t n_rate = n * L;
    const int64_t n_out = n_rate / M;
    const size_t ng = p->groups.size();
    if constexpr (std::is_same<S, float>::value) {
        if (sos_tc_path(p, dtype, n_rate, M)) return sos_run_tc(p, x, y, n, L, M, zi, zf, ws, st);
    }
    S *tmp = nullptr;
    if (sos_needs_tmp(ng, n_rate, L, M))
        tmp = reinterpret_cast<S *>(ws + ((sos_scan_area_bytes(dtype, n_rate) + 255) & ~(size_t)255));
    const bool stage = (L > 1 || M > 1) && n_rate >= SOS_STAGE_MIN;
    const S *src0 = x;
    int64_t n_in0 = n;
    int32_t L0 = L, M_last = M;
    if (stage) {
        if (L > 1) {
            int rc = resample_up_scaled(dtype, x, tmp, n, L, (double)L, st);
            if (rc != B200DSP_OK) return rc;
            src0 = tmp;
            n_in0 = n_rate;
            L0 = 1;
        }
        M_last = 1;
    }
    size_t sec0 = 0;
    for (size_t gi = 0; gi < ng; ++gi) {
        const SosGroup &g = p->groups[gi];
        const bool first = gi == 0, last = gi + 1 == ng;
        const S *src = first ? src0 : tmp;
        S *dst = (last && !(stage && M > 1)) ? y : tmp;
        const C *zig = zi ? static_cast<const C *>(zi) + sec0 * 2 * NCH : nullptr;
        C *zfg = zf ? static_cast<C *>(zf) + sec0 * 2 * NCH : nullptr;
        int rc = run_group_nsec<S>(g, &p->k2_mutex, src, dst, first ? n_in0 : n_rate, n_rate,
                                   (last && M_last > 1) ? n_out : n_rate,
                                   first ? L0 : 1, last ? M_last : 1, zig, zfg, ws, st);
        if (rc != B200DSP_OK) return rc;
        sec0 += g.nsec_real;
    }
    if (stage && M > 1) return resample_dn(dtype, tmp, y, n_out, M, st);
    return B200DSP_OK;
}

}  // namespace b200dsp

using namespace b200dsp;

struct b200dsp_sos_plan : b200dsp_sos_plan_impl {};

extern "C" {

int b200dsp_sos_plan_create(const double *sos_host, int32_t nsec, b200dsp_sos_plan **plan)
{
    if (!sos_host || !plan || nsec < 1) {
        set_error("sos_plan_create: bad argument (nsec=%d)", nsec);
        return B200DSP_E_BADARG;
    }
    for (int s = 0; s < nsec; ++s)
        if (sos_host[6 * s + 3] != 1.0) {
            set_error("sos_plan_create: sos[%d][3] must be 1 (scipy _validate_sos)", s);
            return B200DSP_E_BADARG;
        }
    b200dsp_sos_plan *p = new b200dsp_sos_plan();
    p->nsec = nsec;
    p->sm_count = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    for (int s0 = 0; s0 < nsec; s0 += SOS_MAXSEC) {
        SosGroup g;
        g.nsec_real = (nsec - s0 < SOS_MAXSEC) ? nsec - s0 : SOS_MAXSEC;
        // kernels are instantiated for 1,2,4,6,8 sections; odd counts get an identity section
        // (b0 = 1, everything else 0: x_new = x, both states stay 0 -- exact)
        g.nsec = (g.nsec_real == 1) ? 1 : ((g.nsec_real + 1) & ~1);
        const int D = 2 * g.nsec;
        static const double ident[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 0.0};
        for (int s = 0; s < g.nsec; ++s) {
            const double *q = (s < g.nsec_real) ? sos_host + 6 * (s0 + s) : ident;
            g.coef[s][0] = q[0];
            g.coef[s][1] = q[1];
            g.coef[s][2] = q[2];
            g.coef[s][3] = -q[4];
            g.coef[s][4] = -q[5];
        }
        // one-sample zero-input state transition: column d = next state from unit state e_d
        g.A.assign((size_t)D * D, 0.0);
        for (int d = 0; d < D; ++d) {
            std::vector<double> z(D, 0.0);
            z[d] = 1.0;
            double v = 0.0;
            for (int s = 0; s < g.nsec; ++s) {
                double xn = g.coef[s][0] * v + z[2 * s];
                double z0 = g.coef[s][1] * v + g.coef[s][3] * xn + z[2 * s + 1];
                double z1 = g.coef[s][2] * v + g.coef[s][4] * xn;
                z[2 * s] = z0;
                z[2 * s + 1] = z1;
                v = xn;
            }
            for (int i = 0; i < D; ++i) g.A[i * D + d] = z[i];
        }
        g.mats_f32 = nullptr;
        g.mats_f64 = nullptr;
        cudaError_t e = cudaSuccess;
        {
            std::vector<double> pw;
            matpow(g.A, SOS_LC, pw, D);
            std::vector<double> all((size_t)SOS_LEVELS * D * D);
            for (int j = 0; j < SOS_LEVELS; ++j) {
                memcpy(all.data() + (size_t)j * D * D, pw.data(), sizeof(double) * D * D);
                matmul(pw, pw, pw, D);
            }
            g.tileA = pw;           // A^(SOS_LC * SOS_NT)
            std::vector<float> allf(all.begin(), all.end());
            e = cudaMalloc(&g.mats_f32, allf.size() * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&g.mats_f64, all.size() * sizeof(double));
            if (e == cudaSuccess) e = cudaMemcpy(g.mats_f32, allf.data(), allf.size() * sizeof(float), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(g.mats_f64, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice);
            // pageable-memory cudaMemcpy may return before the DMA completes; consumers use non-blocking streams
            if (e == cudaSuccess) e = cudaStreamSynchronize(0);
        }
        if (e == cudaSuccess && stc_build(g.coef, g.nsec, g.nsec_real, &g.tc) != B200DSP_OK) e = cudaErrorUnknown;
        p->groups.push_back(g);
        if (e != cudaSuccess) {
            set_error("sos_plan_create: %s", cudaGetErrorString(e));
            b200dsp_sos_plan_destroy(p);
            return B200DSP_E_CUDA;
        }
    }
    *plan = p;
    return B200DSP_OK;
}

void b200dsp_sos_plan_destroy(b200dsp_sos_plan *plan)
{
    if (!plan) return;
    for (auto &g : plan->groups) {
        cudaFree(g.mats_f32);
        cudaFree(g.mats_f64);
        stc_free(&g.tc);
    }
    delete plan;
}

int32_t b200dsp_sos_plan_nsec(const b200dsp_sos_plan *plan) { return plan ? plan->nsec : 0; }

size_t b200dsp_sos_workspace_bytes(const b200dsp_sos_plan *plan, int dtype, int64_t n, int32_t L, int32_t M)
{
    if (!plan || n < 0 || L < 1 || M < 1 || dtype_size(dtype) == 0) return 0;
    const int64_t n_rate = n * L;
    size_t b = (sos_scan_area_bytes(dtype, n_rate) + 255) & ~(size_t)255;
    if (sos_needs_tmp(plan->groups.size(), n_rate, L, M)) b += (size_t)n_rate * dtype_size(dtype) + 256;
    return b;
}

int b200dsp_sos_filter(const b200dsp_sos_plan *plan, int dtype, const void *x, void *y, int64_t n,
                       int32_t L, int32_t M, const void *zi, void *zf, void *ws, size_t ws_bytes,
                       void *stream)
{
    if (!plan || n < 0 || L < 1 || M < 1 || (L > 1 && M > 1) || (n > 0 && (!x || !y))) {
        set_error("sos_filter: bad argument (L=%d, M=%d)", L, M);
        return B200DSP_E_BADARG;
    }
    if (dtype_size(dtype) == 0) {
        set_error("sos_filter: bad dtype code %d", dtype);
        return B200DSP_E_DTYPE;
    }
    if (n == 0) return B200DSP_OK;
    if (!ws || ws_bytes < b200dsp_sos_workspace_bytes(plan, dtype, n, L, M)) {
        set_error("sos_filter: workspace too small (%zu < %zu)", ws_bytes,
                  b200dsp_sos_workspace_bytes(plan, dtype, n, L, M));
        return B200DSP_E_WORKSPACE;
    }
    unsigned char *w = static_cast<unsigned char *>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case B200DSP_F32: return sos_run<float>(plan, (const float *)x, (float *)y, n, L, M, zi, zf, w, dtype, st);
    case B200DSP_F64: return sos_run<double>(plan, (const double *)x, (double *)y, n, L, M, zi, zf, w, dtype, st);
    case B200DSP_C64: return sos_run<float2>(plan, (const float2 *)x, (float2 *)y, n, L, M, zi, zf, w, dtype, st);
    case B200DSP_C128: return sos_run<double2>(plan, (const double2 *)x, (double2 *)y, n, L, M, zi, zf, w, dtype, st);
    }
    return B200DSP_E_DTYPE;
}

int b200dsp_sos_filter_batch(const b200dsp_sos_plan *plan, int dtype, const void *x, void *y, int64_t rows, int64_t n,
                             int64_t x_row_stride, int64_t y_row_stride, void *ws, size_t ws_bytes, void *stream)
{
    if (!plan || rows < 0 || n < 0 || x_row_stride < n || y_row_stride < n || dtype_size(dtype) == 0) {
        set_error("sos_filter_batch: bad argument");
        return B200DSP_E_BADARG;
    }
    const size_t es = dtype_size(dtype);
    for (int64_t r = 0; r < rows && n > 0; ++r) {
        // rows run one after the other on the stream, so they can share one workspace
        const int rc = b200dsp_sos_filter(plan, dtype, static_cast<const char *>(x) + (size_t)r * x_row_stride * es,
                                          static_cast<char *>(y) + (size_t)r * y_row_stride * es, n, 1, 1, nullptr, nullptr,
                                          ws, ws_bytes, stream);
        if (rc != B200DSP_OK) return rc;
    }
    return B200DSP_OK;
}

void b200dsp_set_sos_variant(int variant) { g_sos_variant = variant; }

}  // extern "C"
