// fir_tc_real.cu -- float32 (real) FIR filter / polyphase up(L) / polyphase dn(M) on tcgen05.
//
// Same machinery as fir_tc2.cu (taps stationary in TMEM, fp16 hi/lo split around a per-tile power-of-two
// scale, sample streams in K-major SWIZZLE_128B shared memory read through row-shifted UMMA descriptors,
// bulk-TMA ring, in-place conversion, elect-issued UTCHMMA, fused epilogue) for the three real-valued
// entry points of multirate_FIR (reference: src/sk_dsp_comm/multirate_helper.py:104-127):
//
//   MODE_FILTER  y[i]      = sum_t b[t] x[i-t]                      1 stream, 5 k-blocks (<= 256 taps)
//   MODE_UP      y[L m + r] = sum_q g_r[q] x[m-q],  g_r[q] = L b[Lq+r]   1 stream, L tap matrices (phases) of
//                                                                   2 k-blocks, L accumulators per tile,
//                                                                   outputs interleaved by the epilogue
//   MODE_DN      y[m]      = sum_p sum_q g_p[q] f_p[m-q]            M forward phase streams f_p[m] = x[M m + p]
//                g_p[q]    = b[M q - p]                             de-interleaved by the converters, one
//                                                                   accumulator fed by all M streams (K-concat)
// Rate factors 2..4 with at most 64 (65 for dn) taps per phase, e.g. 256 taps at L = M = 4 (cfg3).
#include "tc_common.cuh"

namespace b200dsp {
namespace tcr {
using namespace tcx;

enum { MODE_FILTER = 1, MODE_UP = 2, MODE_DN = 3 };
constexpr int BK = 64;
constexpr int TILE_N = 64;                   // stream rows per tile
constexpr int TILE = TILE_N * BK;            // 4096 stream positions per tile
constexpr int STREAM_BYTES = 9 * 1024;
constexpr int NACC = 4;
constexpr int ACC_COL0 = 256;                // accumulators occupy TMEM columns 256..511
constexpr int N_EPI_WARPS = 4, MMA_WARP = 4, PROD_WARP = 5, CVT_WARP0 = 6, N_CVT_WARPS = 8;
constexpr int N_CVT = N_CVT_WARPS * 32;
constexpr int NTHREADS = (CVT_WARP0 + N_CVT_WARPS) * 32;     // 448
constexpr int INV_RING = 16;

template <int MODE, int P> struct RCfg {
    static constexpr int KB = (MODE == MODE_FILTER) ? 5 : 2;          // k-blocks per tap matrix
    static constexpr int HS = BK * (KB - 1);                          // stream halo (samples)
    static constexpr int ROWS = TILE_N + KB - 1;
    static constexpr int NSTREAM = (MODE == MODE_DN) ? P : 1;         // fp16 stream pairs (hi, lo)
    static constexpr int NPH = (MODE == MODE_FILTER) ? 1 : P;         // tap matrices in TMEM
    static constexpr int UNITS = (MODE == MODE_UP) ? P : 1;           // independent outputs sets per tile
    // dn with 3-4 phase streams splits the K-concatenation over two accumulators (summed by the epilogue):
    // half as many truncating fp32 adds at full magnitude per accumulator
    static constexpr int APU = (MODE == MODE_DN && P > 2) ? 2 : 1;    // accumulators per unit
    static constexpr int GROUPS = ROWS * BK / 4;                      // groups of 4 stream positions
    static constexpr int G_PER_THREAD = (GROUPS + N_CVT - 1) / N_CVT;
    static constexpr int RAW_BYTES = NSTREAM * ROWS * BK * 4;
    static constexpr int STAGE_BYTES = 2 * NSTREAM * STREAM_BYTES;
    static constexpr int NSTAGE_RAW = 228000 / STAGE_BYTES;
    static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
    static constexpr int SMEM_BAR_OFF = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 512 + 1024;
    static constexpr int A_COLS_PH = KB * 32;                         // TMEM columns per tap matrix
    static_assert(NPH * (KB * 32) <= ACC_COL0, "tap matrices must fit below the accumulators");
    static_assert((2 * NSTREAM * STREAM_BYTES) >= (NSTREAM * ROWS * BK * 4), "in-place conversion");
};

struct Args {
    const float *x;
    const float *hist;
    float *y;
    const uint4 *amat;        // device: [NPH][128 rows][KB*64 fp16]
    int64_t n_in;             // input samples
    int64_t n_m;              // stream positions to produce (filter: n, up: n, dn: n / M)
    int64_t n_tiles;
    int32_t hist_len;
    int32_t sb_exp;
    int32_t jmin;             // filter: first k-block with a non-zero tap
    int32_t dbg;
    int32_t M;                // <MODE_FILTER, 0> (filter with decimating stores): keep every M-th output
    int64_t n_dec;            //                                                    outputs kept
};

__device__ __forceinline__ float load_sample(const Args &a, int64_t g) {
    if (g >= 0) return (g < a.n_in) ? a.x[g] : 0.f;
    if (a.hist != nullptr) {
        int64_t h = (int64_t)a.hist_len + g;
        if (h >= 0) return a.hist[h];
    }
    return 0.f;
}

// first raw input sample of a tile and whether the tile can be fetched with one bulk copy
template <int MODE, int P>
__device__ __forceinline__ int64_t tile_g0(int64_t tile) {
    using C = RCfg<MODE, P>;
    return (MODE == MODE_DN) ? (int64_t)P * (tile * TILE - C::HS) : tile * TILE - C::HS;
}
template <int MODE, int P>
__device__ __forceinline__ bool tile_is_bulk(const Args &a, int64_t tile) {
    using C = RCfg<MODE, P>;
    const int64_t g0 = tile_g0<MODE, P>(tile);
    return g0 >= 0 && g0 + C::RAW_BYTES / 4 <= a.n_in && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
}

template <int MODE, int P, bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1) fir_tc_real_kernel(const Args a)
{
    using C = RCfg<MODE, P>;
    constexpr int KB = C::KB, NSTREAM = C::NSTREAM, NPH = C::NPH, UNITS = C::UNITS, NSTAGE = C::NSTAGE, APU = C::APU;
    constexpr int STAGE_BYTES = C::STAGE_BYTES, RAW_BYTES = C::RAW_BYTES, GROUPS = C::GROUPS;
    constexpr int G_PER_THREAD = C::G_PER_THREAD;
    constexpr uint32_t kIdesc = idesc_f16(128, TILE_N);

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    const uint32_t base = (raw_base + 1023u) & ~1023u;
    unsigned char *sm = smem_dyn + (base - raw_base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::SMEM_BAR_OFF);
    const uint32_t bar0 = base + C::SMEM_BAR_OFF;
    auto RAW_FULL = [&](int s) { return bar0 + 8u * s; };
    auto A_FULL = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto D_FULL = [&](int b) { return bar0 + 8u * (3 * NSTAGE + b); };
    auto D_EMPTY = [&](int b) { return bar0 + 8u * (3 * NSTAGE + NACC + b); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * NSTAGE + 2 * NACC);
    float *wmax = reinterpret_cast<float *>(bars + 3 * NSTAGE + 2 * NACC + 1);
    float *tile_inv = reinterpret_cast<float *>(bars + 3 * NSTAGE + 2 * NACC + 5);

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(RAW_FULL(s), 1);
            mbar_init(A_FULL(s), N_CVT_WARPS);
            mbar_init(A_EMPTY(s), 1);
        }
        for (int b = 0; b < NACC; ++b) {
            mbar_init(D_FULL(b), 1);
            mbar_init(D_EMPTY(b), N_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- tap matrices -> TMEM (stay for the whole kernel) ----
    if (warp < N_EPI_WARPS) {
        const int row = warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int ph = 0; ph < NPH; ++ph) {
            const uint4 *src = a.amat + ((size_t)ph * 128 + row) * (KB * BK * 2 / 16);
#pragma unroll 1
            for (int c0 = 0; c0 < C::A_COLS_PH; c0 += 16) {
                uint32_t v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u = src[c0 / 4 + q];
                    v[4 * q + 0] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
                }
                tmem_st16(taddr + ph * C::A_COLS_PH + c0, v);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int64_t first = blockIdx.x, step = gridDim.x;

    if (warp == PROD_WARP) {
        // =============================== producer (bulk TMA) ===============================
        int it = 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            mbar_wait(A_EMPTY(s), ph ^ 1u, 0);
            if (elect_one()) {
                if (tile_is_bulk<MODE, P>(a, tile)) {
                    mbar_arrive_expect_tx(RAW_FULL(s), RAW_BYTES);
                    bulk_g2s(base + s * STAGE_BYTES, a.x + tile_g0<MODE, P>(tile), RAW_BYTES, RAW_FULL(s));
                } else {
                    mbar_arrive(RAW_FULL(s));
                }
            }
            __syncwarp();
        }
    } else if (warp >= CVT_WARP0) {
        // =============================== converters ===============================
        const int ct = tid - CVT_WARP0 * 32;
        int it = 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            unsigned char *st = sm + s * STAGE_BYTES;
            mbar_wait(RAW_FULL(s), ph, 1);
            // group g = 4 consecutive positions of every stream = NSTREAM consecutive float4 of the raw tile
            float4 raw4[G_PER_THREAD][NSTREAM];
            const bool bulk = tile_is_bulk<MODE, P>(a, tile);
            const int64_t g0 = tile_g0<MODE, P>(tile);
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < G_PER_THREAD; ++i) {
                const int g = ct + i * N_CVT;
#pragma unroll
                for (int j = 0; j < NSTREAM; ++j) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g < GROUPS) {
                        const int f = g * NSTREAM + j;
                        if (bulk) {
                            v = reinterpret_cast<const float4 *>(st)[f];
                        } else {
                            v.x = load_sample(a, g0 + 4 * f + 0);
                            v.y = load_sample(a, g0 + 4 * f + 1);
                            v.z = load_sample(a, g0 + 4 * f + 2);
                            v.w = load_sample(a, g0 + 4 * f + 3);
                        }
                    }
                    raw4[i][j] = v;
                    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));
            if (lane == 0) wmax[warp - CVT_WARP0] = m;
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));        // all raw samples are in registers
            float bm = 0.f;
#pragma unroll
            for (int w = 0; w < N_CVT_WARPS; ++w) bm = fmaxf(bm, wmax[w]);
            int ex = 14;
            if (bm > 0.f && bm < 3.0e38f) (void)frexpf(bm, &ex);
            const int e = max(-110, min(110, 14 - ex));
            const float sx = ldexpf(1.0f, e);
            // element 4g of a stream: row g>>4, 16-byte chunk (g&15)>>1, byte (g&1)*8; +16 rows per iteration
            const uint32_t off0 = stream_off(4 * ct);
#pragma unroll
            for (int i = 0; i < G_PER_THREAD; ++i) {
                const int g = ct + i * N_CVT;
                if (g < GROUPS) {
                    const uint32_t off = off0 + (uint32_t)i * 2048u;
#pragma unroll
                    for (int p = 0; p < NSTREAM; ++p) {
                        // stream p takes flat elements u = 4*... : for NSTREAM = 1 the float4 itself; for dn the
                        // raw order is x[P m' + p], so position k of stream p is flat element k*NSTREAM + p
                        float v4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int u = k * NSTREAM + p;
                            const float4 &r = raw4[i][u >> 2];
                            v4[k] = ((u & 3) == 0 ? r.x : (u & 3) == 1 ? r.y : (u & 3) == 2 ? r.z : r.w) * sx;
                        }
                        const __half2 h01 = __floats2half2_rn(v4[0], v4[1]);
                        const __half2 h23 = __floats2half2_rn(v4[2], v4[3]);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(v4[0] - f01.x, v4[1] - f01.y);
                        const __half2 l23 = __floats2half2_rn(v4[2] - f23.x, v4[3] - f23.y);
                        uint2 hv, lv;
                        hv.x = *reinterpret_cast<const uint32_t *>(&h01);
                        hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                        lv.x = *reinterpret_cast<const uint32_t *>(&l01);
                        lv.y = *reinterpret_cast<const uint32_t *>(&l23);
                        *reinterpret_cast<uint2 *>(st + (2 * p + 0) * STREAM_BYTES + off) = hv;
                        *reinterpret_cast<uint2 *>(st + (2 * p + 1) * STREAM_BYTES + off) = lv;
                    }
                }
            }
            if (ct == 0) tile_inv[it % INV_RING] = ldexpf(1.0f, -e - a.sb_exp);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(A_FULL(s));
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================
        int it = 0;
        uint32_t unit = 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            mbar_wait(A_FULL(s), ph, 2);
            tc_fence_after();
            const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll 1
            for (int un = 0; un < UNITS; ++un, ++unit) {
                const uint32_t b0 = (unit * APU) % NACC;
#pragma unroll
                for (int h = 0; h < APU; ++h) mbar_wait(D_EMPTY((b0 + h) % NACC), (((unit * APU + h) / NACC) & 1u) ^ 1u, 3);
                tc_fence_after();
                if (elect_one()) {
                    uint32_t started[APU];
#pragma unroll
                    for (int h = 0; h < APU; ++h) started[h] = 0;
                    // all residual (lo) streams first -- the accumulators are still tiny -- then the hi streams
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const int part = 1 - pp;
#pragma unroll
                        for (int sp = 0; sp < NSTREAM; ++sp) {
                            const int h = (APU == 2 && sp >= (NSTREAM + 1) / 2) ? 1 : 0;
                            const uint32_t dd = tmem_base + ACC_COL0 + ((b0 + h) % NACC) * TILE_N;
                            const int mat = (MODE == MODE_UP) ? un : sp;            // tap matrix (phase)
                            const uint64_t bd0 = make_desc_sw128(st + (2 * sp + part) * STREAM_BYTES);
                            const uint32_t a0 = tmem_base + mat * C::A_COLS_PH;
#pragma unroll
                            for (int jj = 0; jj < KB; ++jj) {
                                constexpr int kOrder5[5] = {0, 1, 4, 3, 2};
                                const int j = (KB == 5) ? kOrder5[jj] : jj;
                                if (MODE == MODE_FILTER && j < a.jmin) continue;
#pragma unroll
                                for (int sl = 0; sl < 4; ++sl) {
                                    const uint64_t bd = bd0 + (uint64_t)(8 * j + 2 * sl);
                                    if (!DBG || !(a.dbg & 1)) umma_f16_ts(dd, a0 + (j * 4 + sl) * 8, bd, kIdesc, started[h]);
                                    started[h] = 1;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < APU; ++h) umma_commit(D_FULL((b0 + h) % NACC));
                    if (un == UNITS - 1) umma_commit(A_EMPTY(s));
                }
                __syncwarp();
            }
        }
    } else {
        // =============================== epilogue (warps 0-3) ===============================
        int it = 0;
        uint32_t unit = 0;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        const bool lo_row = lane >= 16;
        const int c = warp * 16 + (lane & 15);
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it, unit += UNITS) {
            constexpr int NA = UNITS * APU;                 // accumulators of this tile
            uint32_t dbase[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                const uint32_t u = unit * APU + k, b = u % NACC;
                mbar_wait(D_FULL(b), (u / NACC) & 1u, 4);
                dbase[k] = tmem_base + lane_sel + ACC_COL0 + b * TILE_N;
            }
            tc_fence_after();
            const float inv = tile_inv[it % INV_RING];
            const int64_t m_c = tile * TILE + c + (lo_row ? BK : 0);
            // filter with decimating stores: the lane's positions advance by 2 BK from m_c; (mq, rm) = divmod(m, M)
            constexpr bool FDEC = (MODE == MODE_FILTER && P == 0);
            int64_t mq = 0;
            int32_t rm = 0, dq = 0, drm = 0;
            if constexpr (FDEC) {
                mq = m_c / a.M;
                rm = (int32_t)(m_c - mq * a.M);
                dq = (2 * BK) / a.M;
                drm = (2 * BK) - dq * a.M;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < TILE_N; c0 += 16) {
                uint32_t d[NA][16];
#pragma unroll
                for (int k = 0; k < NA; ++k) tmem_ld16(dbase[k] + c0, d[k]);
                tmem_ld_wait();
                if (c0 + 16 == TILE_N) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < NA; ++k) mbar_arrive(D_EMPTY((unit * APU + k) % NACC));
                    }
                }
#pragma unroll
                for (int q = 0; q < 16; q += 2) {
                    float val[UNITS];
#pragma unroll
                    for (int un = 0; un < UNITS; ++un) {
                        float own_a = __uint_as_float(d[un * APU][q]), own_b = __uint_as_float(d[un * APU][q + 1]);
                        if (APU == 2) {
                            own_a += __uint_as_float(d[un * APU + APU - 1][q]);
                            own_b += __uint_as_float(d[un * APU + APU - 1][q + 1]);
                        }
                        const float recv = __shfl_xor_sync(0xffffffffu, lo_row ? own_a : own_b, 16);
                        val[un] = (recv + (lo_row ? own_b : own_a)) * inv;
                    }
                    const int64_t m = m_c + (int64_t)(c0 + q) * BK;          // stream position of this lane's output
                    if constexpr (FDEC) {
                        if (rm == 0 && mq < a.n_dec) a.y[mq] = val[0];
                        rm += drm;
                        mq += dq;
                        if (rm >= a.M) { rm -= a.M; ++mq; }
                    } else if (m < a.n_m && (!DBG || !(a.dbg & 4))) {
                        if (MODE != MODE_UP) {
                            a.y[m] = val[0];
                        } else if (P == 4) {
                            *reinterpret_cast<float4 *>(a.y + 4 * m) = make_float4(val[0], val[1 % UNITS], val[2 % UNITS], val[3 % UNITS]);
                        } else if (P == 2) {
                            *reinterpret_cast<float2 *>(a.y + 2 * m) = make_float2(val[0], val[1 % UNITS]);
                        } else {
#pragma unroll
                            for (int un = 0; un < UNITS; ++un) a.y[(int64_t)P * m + un] = val[un];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tcr

// ------------------------------------------------------------------------------------------ host
// Tap matrices [NPH][128][KB*64] fp16 (row rho = 32 w + l: l < 16 hi part of Toeplitz row c = 16 w + l,
// l >= 16 its fp16 residual).  Toeplitz row: T[c][kk] = g[c + HS - kk], HS = 64 (KB - 1).
//   filter: g = b                       up(L), phase r: g[q] = L b[L q + r]        dn(M), stream p: g[q] = b[M q - p]
int tcr_build(const double *taps, int ntaps, int mode, int P, unsigned char *out, int *sb_exp)
{
    using namespace tcr;
    const int KB = (mode == MODE_FILTER) ? 5 : 2, HS = BK * (KB - 1), KT = KB * BK;
    const int NPH = (mode == MODE_FILTER) ? 1 : P;
    auto g = [&](int ph, int q) -> double {
        if (q < 0) return 0.0;
        long t;
        double gain = 1.0;
        if (mode == MODE_FILTER) t = q;
        else if (mode == MODE_UP) { t = (long)P * q + ph; gain = (double)P; }
        else t = (long)P * q - ph;
        return (t >= 0 && t < ntaps) ? gain * taps[t] : 0.0;
    };
    // every non-zero tap must be reachable: q <= HS for all phases
    for (int ph = 0; ph < NPH; ++ph)
        for (int q = HS + 1; q < HS + 1 + ntaps; ++q)
            if (g(ph, q) != 0.0) return -1;
    double mx = 0.0;
    for (int ph = 0; ph < NPH; ++ph)
        for (int q = 0; q <= HS; ++q) mx = fmax(mx, fabs(g(ph, q)));
    int ex = 0;
    if (mx > 0.0) (void)frexp(mx, &ex);
    int e = 12 - ex;
    if (e > 60) e = 60;
    if (e < -60) e = -60;
    *sb_exp = e;
    __half *m = reinterpret_cast<__half *>(out);
    for (int ph = 0; ph < NPH; ++ph)
        for (int rho = 0; rho < 128; ++rho) {
            const int w = rho / 32, l = rho % 32, c = 16 * w + (l & 15);
            const bool lo = l >= 16;
            for (int kk = 0; kk < KT; ++kk) {
                double v = ldexp(g(ph, c + HS - kk), e);
                __half hi = __float2half_rn((float)v);
                __half lw = __float2half_rn((float)(v - (double)__half2float(hi)));
                m[((size_t)ph * 128 + rho) * KT + kk] = lo ? lw : hi;
            }
        }
    return 0;
}
int tcr_matrix_bytes(int mode, int P)
{
    const int KB = (mode == tcr::MODE_FILTER) ? 5 : 2;
    const int NPH = (mode == tcr::MODE_FILTER) ? 1 : P;
    return NPH * 128 * KB * 64 * 2;
}

template <int MODE, int P>
static int launch_tcr(tcr::Args a, int sm_count, cudaStream_t stream)
{
    using namespace tcr;
    a.n_tiles = (a.n_m + TILE - 1) / TILE;
    int64_t grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    if (a.dbg) {
        auto kern = fir_tc_real_kernel<MODE, P, true>;
        B200_CHECK_CUDA(allow_smem(kern, RCfg<MODE, P>::SMEM_TOTAL));
        kern<<<(unsigned)grid, NTHREADS, RCfg<MODE, P>::SMEM_TOTAL, stream>>>(a);
    } else {
        auto kern = fir_tc_real_kernel<MODE, P, false>;
        B200_CHECK_CUDA(allow_smem(kern, RCfg<MODE, P>::SMEM_TOTAL));
        kern<<<(unsigned)grid, NTHREADS, RCfg<MODE, P>::SMEM_TOTAL, stream>>>(a);
    }
    B200_CHECK_LAUNCH("fir_tc_real_kernel");
    return B200DSP_OK;
}

// mode: 1 filter, 2 up (P = L), 3 dn (P = M).  n = input samples.  mode 1 with P > 1: the filter kernel with
// decimating stores (dn by factors that have no phase-stream formulation here): y[o] = filter output P o.
int launch_fir_tc_real(int mode, int P, const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                       const void *amat_dev, int sb_exp, int ntaps, int sm_count, cudaStream_t stream)
{
    using namespace tcr;
    Args a;
    a.M = 1;
    a.n_dec = 0;
    a.x = static_cast<const float *>(x);
    a.hist = static_cast<const float *>(hist);
    a.y = static_cast<float *>(y);
    a.amat = static_cast<const uint4 *>(amat_dev);
    a.n_in = n;
    a.n_m = (mode == MODE_DN) ? n / P : n;
    a.n_tiles = 0;
    a.hist_len = hist_len;
    a.sb_exp = sb_exp;
    a.jmin = (mode == MODE_FILTER) ? ((256 + 1 - ntaps) / BK > 0 ? (256 + 1 - ntaps) / BK : 0) : 0;
    a.dbg = 0;
    if (const char *e = getenv("B200DSP_TC_DBG")) a.dbg = atoi(e);
    if (a.n_m <= 0) return B200DSP_OK;
    if (mode == MODE_FILTER && P > 1) {
        a.M = P;
        a.n_dec = n / P;
        if (a.n_dec <= 0) return B200DSP_OK;
        a.n_m = (a.n_dec - 1) * P + 1;              // nothing after the last kept output is computed
        return launch_tcr<MODE_FILTER, 0>(a, sm_count, stream);
    }
    if (mode == MODE_FILTER) return launch_tcr<MODE_FILTER, 1>(a, sm_count, stream);
    if (mode == MODE_UP) {
        if (P == 2) return launch_tcr<MODE_UP, 2>(a, sm_count, stream);
        if (P == 3) return launch_tcr<MODE_UP, 3>(a, sm_count, stream);
        if (P == 4) return launch_tcr<MODE_UP, 4>(a, sm_count, stream);
    }
    if (mode == MODE_DN) {
        if (P == 2) return launch_tcr<MODE_DN, 2>(a, sm_count, stream);
        if (P == 3) return launch_tcr<MODE_DN, 3>(a, sm_count, stream);
        if (P == 4) return launch_tcr<MODE_DN, 4>(a, sm_count, stream);
    }
    set_error("fir_tc_real: unsupported mode/factor %d/%d", mode, P);
    return B200DSP_E_UNSUPPORTED;
}

}  // namespace b200dsp
