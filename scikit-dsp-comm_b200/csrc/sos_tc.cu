// sos_tc.cu -- float32 biquad cascade (scipy.signal.sosfilt as called by multirate_IIR.filter/.up/.dn,
// reference src/sk_dsp_comm/multirate_helper.py:169-192) in ONE pass over HBM on tcgen05.
//
// The cascade of <= 8 sections is an LTI state-space system  s[n+1] = A s[n] + B x[n],  y[n] = C s[n] + D x[n]
// with 2*nsec states (direct-form-II-transposed, scipy's state order).  The stream is cut into chunks of
// 128 samples = two 64-sample rows ("even" row e, "odd" row o); a tile is 64 chunks (8192 samples):
//
//   zero-state outputs  Y_e = T0 X_e                 T0[r,k] = h[r-k] (k <= r), T1[r,k] = h[64+r-k]
//                       Y_o = T1 X_e + T0 X_o        h = impulse response of the whole cascade (128 values)
//   chunk carries       E   = Ka X_e + Kb X_o        Ka[:,k] = A^(127-k) B,  Kb[:,k] = A^(63-k) B
//       -> dense GEMMs on the tensor core: the constant matrices live in TENSOR MEMORY (fp16 hi/lo rows,
//          per-row power-of-two scales), the sample rows are the shared-memory operand (K-major SWIZZLE_128B,
//          fp16 hi/lo around a per-tile power-of-two scale), fp32 accumulators in TMEM.  The O(nsec) serial
//          recurrence per sample is gone: cost per sample is independent of the number of sections.
//   prefix scan         s_{c+1} = A^128 s_c + E_c    one warp, lane = two chunks, float64 Kogge-Stone with the
//                                                    constant matrices A^(256 * 2^j) in shared memory
//   correction          Y += O S                     O[r,:] = C A^r (zero-input response rows), S[:,c] = state at the start
//                                                    of chunk c: one more K=16 GEMM step into the SAME accumulator
//                                                    (fp16 hi/lo, three partial products), so the epilogue only scales
//                                                    and stores
//
// Every CTA owns a CONTIGUOUS run of tiles, so the carry between tiles never leaves the CTA and there is no
// grid-wide scan, no second pass and no inter-CTA traffic: x is read once, y written once (8 B/sample).  A
// block other than the first starts `warm_tiles` tiles early with zero state and discards those outputs: the
// plan picks warm_tiles so that ||A^(8192 warm_tiles)|| <= 1e-13, i.e. the ignored homogeneous term is below
// float64 resolution of the state (cascades that do not decay that fast take the 3-kernel scan path, sos_scan.cu).
//
// Roles (15 warps): 8 epilogue warps (TMEM -> registers -> hi/lo combine -> correction -> stores), 1 scan warp,
// 1 MMA issuer (elect.sync lane, 40 UTCHMMA per tile), 1 bulk-TMA producer, 4 converter warps.
#include "tc_common.cuh"
#include "sos_tc.cuh"
#include <math.h>
#include <type_traits>
#include <vector>

namespace b200dsp {
namespace stc {
using namespace tcx;

constexpr int BK = 64;                        // samples per stream row (one 128-byte swizzle row of fp16)
constexpr int NCHUNK = 64;                    // chunks per tile = GEMM N
constexpr int LC = 128;                       // samples per chunk
constexpr int TILE = NCHUNK * LC;             // 8192
static_assert(TILE == STC_TILE, "tile size");
constexpr int REGION = NCHUNK * BK * 2;       // one fp16 operand region: 64 rows x 128 B
constexpr int STAGE_BYTES = 4 * REGION;       // (even,odd) x (hi,lo); the raw fp32 tile has the same size
constexpr int NSTAGE = 3;                     // operand ring (fp16 regions, read by the MMAs)
constexpr int NRAW = 2;                       // raw ring (fp32 tiles written by bulk TMA, read by the converters)
constexpr int NACC = 2;
// TMEM columns: T_hi (128 rows x 128 k -> 64 columns) | T_lo (64) | Ka (32) | Kb (32) | O_hi (8) | O_lo (8) | accumulators
constexpr int COL_THI = 0, COL_TLO = 64, COL_KA = 128, COL_KB = 160, COL_OHI = 192, COL_OLO = 200;
constexpr int A_COLS = 208;
constexpr int ACC_COL0 = A_COLS;
constexpr int ACC_STRIDE = 2 * NCHUNK;        // Y | E
// warps: a warp reads TMEM lanes 32 * (warp % 4) .. +31, so the carry readers (lanes 0..31) sit at 8 and 12
// warp 8 (TMEM lanes 0..31) issues the TMA loads and the MMAs AND reads the carry accumulator back while the tensor
// pipe works on the tile's output MMAs; chain and scan warps need no TMEM access and sit on the other schedulers
// (scheduler = warp % 4), one converter warp per scheduler
constexpr int N_EPI_WARPS = 8, MMA_WARP = 8, CHAIN_WARP = 9, ZSCAN_WARP0 = 10, ZSCAN_WARP1 = 11, CVT_WARP0 = 12;
constexpr int N_CVT_WARPS = 4;
constexpr int N_CVT = N_CVT_WARPS * 32;
constexpr int NTHREADS = 16 * 32;             // 512 threads -> 128 registers per thread
constexpr int G_PER_THREAD = TILE / 4 / N_CVT;                // 16 float4 groups per converter thread
constexpr int E_RING = 16;
constexpr int EP = 20;                        // float pitch of a chunk's carry / start-state row in shared memory
constexpr int NSMAT = 7;                      // A^128, A^256, ... A^8192 (float64, column-major)
constexpr int NFOLD = 18;                     // float32 matrices: A^(1024 h) h<8 | A^(256 l) l<4 | A^128 | A^256 .. A^4096
constexpr int FOLD_PITCH = 260;               // floats per fold matrix (= 4 mod 32: distinct matrices, distinct banks)

constexpr int MAXWIN = 8;                     // tiles of input history that bound the state (window of the block scale)
constexpr int SOP_BYTES = NCHUNK * 128;       // chunk start states as an MMA operand: 64 rows x (16 hi | 16 lo | pad) fp16

// shared memory map (offsets from the 1024-aligned base)
constexpr int SM_RAW = NSTAGE * STAGE_BYTES;                  // [NRAW][STAGE_BYTES]      raw fp32 tiles
constexpr int SM_SOP = SM_RAW + NRAW * STAGE_BYTES;           // [2][SOP_BYTES]           K-major SWIZZLE_128B, 1024-aligned
// the rest depends on the in-tile scan type: float64 scans (Z64) need all seven float64 matrices and a 2-deep ring
// of carry buffers; float32 scans need only A^8192 in float64 (chain) but 18 float32 matrices, and get a 4-deep
// ring of carry buffers so that the scan warps never wait for the chain warp to hand a buffer back
template <bool Z64> struct Lay {
    static constexpr int ESM_DEPTH = Z64 ? 2 : 4;
    static constexpr int NSMAT_SM = Z64 ? NSMAT : 1;              // float32 scans: slot 0 = A^8192
    static constexpr int NFOLD_SM = Z64 ? 13 : NFOLD;
    static constexpr int SM_SMAT = SM_SOP + 2 * SOP_BYTES;        // double [NSMAT_SM][16*16]
    static constexpr int SM_FOLD = SM_SMAT + NSMAT_SM * 256 * 8;  // float [NFOLD_SM][FOLD_PITCH]
    static constexpr int SM_ESM = SM_FOLD + NFOLD_SM * FOLD_PITCH * 4;   // float [ESM_DEPTH][64][EP]  carries (warp 8 -> scan
                                                                  // warp), overwritten in place by the zero-state start states
    static constexpr int SM_AGG = SM_ESM + ESM_DEPTH * NCHUNK * EP * 4;  // double [ESM_DEPTH][16]  zero-state end state of a tile
    static constexpr int SM_BAR = SM_AGG + ESM_DEPTH * 16 * 8;
    static constexpr int NBAR = 2 * NRAW + 2 * NSTAGE + 6 * NACC + 3 * ESM_DEPTH;
    static constexpr int SM_MISC = SM_BAR + NBAR * 8;             // tmem slot, wmax[2][4], tile_e[16], tile_m[16]
    static constexpr int SMEM_TOTAL = SM_MISC + 256 + 1024;
    static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
};

struct Args {
    const float *x;
    float *y;
    const uint4 *amat;
    const double *smat;
    const float *fold;
    const float *rowinv;
    const float *zi;
    float *zf;
    int64_t n_in, n_rate, n_out;
    int64_t n_tiles;
    int32_t L, M;
    int32_t tiles_per_block, warm_tiles;
    int32_t e_t, d_real;
    int32_t xexp, win;        // block scale: window max of |x| over `win` tiles is scaled into [2^(xexp-1), 2^xexp)
    float sfac[16];           // state d enters the correction operand as s_d * sfac[d] * 2^e_x
    float ginv[16];           // 1 / (power-of-two bound of the input->state gain): |s_d| * ginv[d] <= max |x|
    int32_t nlev;             // Kogge-Stone levels whose matrix A^(256 * 2^j) is not negligible
    int32_t at_zero;          // A^8192 is negligible: the state entering a tile is the previous tile's aggregate
    int32_t dbg;              // bring-up (DBG build): bit0 skip MMAs, bit1 skip scan math, bit2 skip stores, bit3 print waits, bit4 skip correction
    double coef[STC_MAXSEC][5];
};

// filter-rate sample g of the (virtually zero-stuffed, x L) input
__device__ __forceinline__ float load_rate_sample(const Args &a, int64_t g) {
    if (a.L == 1) return (g < a.n_in) ? a.x[g] : 0.f;
    const int64_t q = g / a.L;
    return (g - q * a.L == 0 && q < a.n_in) ? a.x[q] * (float)a.L : 0.f;
}
__device__ __forceinline__ bool tile_is_bulk(const Args &a, int64_t tile) {
    return a.L == 1 && (tile + 1) * (int64_t)TILE <= a.n_in && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
}

// acc += P v, P block-lower-triangular (a section's state never depends on a later section), stored
// column-major (Pt[k * ND + i] = P[i][k]).  A dependent DFMA costs ~65 cycles on this part, so every row
// accumulates its even and odd k terms in two independent chains (k outer, i inner: 2 ND chains in flight);
// column pairs are 16-byte shared-memory loads.
template <int ND>
__device__ __forceinline__ void mv_acc(double (&acc)[ND], const double *__restrict__ Pt, const double (&v)[ND]) {
    double alt[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) alt[i] = 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
#pragma unroll
        for (int i = (k & ~1); i < ND; i += 2) {
            const double2 p = *reinterpret_cast<const double2 *>(Pt + k * ND + i);
            if (k & 1) {
                alt[i] = fma(p.x, v[k], alt[i]);
                alt[i + 1] = fma(p.y, v[k], alt[i + 1]);
            } else {
                acc[i] = fma(p.x, v[k], acc[i]);
                acc[i + 1] = fma(p.y, v[k], acc[i + 1]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ND; ++i) acc[i] += alt[i];
}
// float32 version for the fold (direct products of the float64 carry, no accumulation across chunks):
// out = P v, column-major P with pitch ND, 16-byte loads (rows above the block diagonal hold zeros)
template <int ND>
__device__ __forceinline__ void mv_acc(float (&out)[ND], const float *__restrict__ Pt, const float (&v)[ND]) {
#pragma unroll
    for (int k = 0; k < ND; ++k) {
#pragma unroll
        for (int i = (k & ~3); i < ND; i += 4) {
            const float4 p = *reinterpret_cast<const float4 *>(Pt + k * ND + i);
            out[i] = fmaf(p.x, v[k], out[i]);
            out[i + 1] = fmaf(p.y, v[k], out[i + 1]);
            out[i + 2] = fmaf(p.z, v[k], out[i + 2]);
            out[i + 3] = fmaf(p.w, v[k], out[i + 3]);
        }
    }
}
template <int ND>
__device__ __forceinline__ void mv_f32(float (&out)[ND], const float *__restrict__ Pt, const float (&v)[ND]) {
#pragma unroll
    for (int i = 0; i < ND; ++i) out[i] = 0.f;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
#pragma unroll
        for (int i = (k & ~3); i < ND; i += 4) {
            const float4 p = *reinterpret_cast<const float4 *>(Pt + k * ND + i);
            out[i] = fmaf(p.x, v[k], out[i]);
            out[i + 1] = fmaf(p.y, v[k], out[i + 1]);
            out[i + 2] = fmaf(p.z, v[k], out[i + 2]);
            out[i + 3] = fmaf(p.w, v[k], out[i + 3]);
        }
    }
}

#define STC_WAIT(bar, par, who)                                   \
    do {                                                          \
        if (DBG && (a.dbg & 8)) {                                 \
            const long long _t0 = clock64();                      \
            mbar_wait(bar, par, who);                             \
            wait_cycles += clock64() - _t0;                       \
        } else {                                                  \
            mbar_wait(bar, par, who);                             \
        }                                                         \
    } while (0)

template <int ND, bool Z64, bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1) sos_tc_kernel(const Args a)
{
    long long wait_cycles = 0, ph_a = 0, ph_b = 0, ph_c = 0;
    (void)ph_a; (void)ph_b; (void)ph_c;
    const long long t_start = DBG ? clock64() : 0;
    const uint64_t ns_start = DBG ? global_ns() : 0;
    constexpr uint32_t kIdesc = idesc_f16(128, NCHUNK);
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    const uint32_t base = (raw_base + 1023u) & ~1023u;
    unsigned char *sm = smem_dyn + (base - raw_base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    using LY = Lay<Z64>;
    constexpr int ESM_DEPTH = LY::ESM_DEPTH;
    constexpr int SM_SMAT = LY::SM_SMAT, SM_FOLD = LY::SM_FOLD, SM_ESM = LY::SM_ESM, SM_AGG = LY::SM_AGG,
                  SM_BAR = LY::SM_BAR, SM_MISC = LY::SM_MISC;
    const uint32_t bar0 = base + SM_BAR;
    constexpr int B0 = 2 * NRAW + 2 * NSTAGE;
    auto RAW_FULL = [&](int r) { return bar0 + 8u * r; };
    auto RAW_EMPTY = [&](int r) { return bar0 + 8u * (NRAW + r); };
    auto A_FULL = [&](int s) { return bar0 + 8u * (2 * NRAW + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NRAW + NSTAGE + s); };
    auto E_FULL = [&](int b) { return bar0 + 8u * (B0 + b); };
    auto E_EMPTY = [&](int b) { return bar0 + 8u * (B0 + NACC + b); };
    auto D_FULL = [&](int b) { return bar0 + 8u * (B0 + 2 * NACC + b); };
    auto D_EMPTY = [&](int b) { return bar0 + 8u * (B0 + 3 * NACC + b); };
    auto S_READY = [&](int b) { return bar0 + 8u * (B0 + 4 * NACC + b); };
    auto SOP_EMPTY = [&](int b) { return bar0 + 8u * (B0 + 5 * NACC + b); };
    // carry-buffer ring (ESM_DEPTH deep): indexed by tile % ESM_DEPTH
    auto Z_READY = [&](int e) { return bar0 + 8u * (B0 + 6 * NACC + e); };
    auto ESM_READY = [&](int e) { return bar0 + 8u * (B0 + 6 * NACC + ESM_DEPTH + e); };
    auto ESM_EMPTY = [&](int e) { return bar0 + 8u * (B0 + 6 * NACC + 2 * ESM_DEPTH + e); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + SM_MISC);
    float *wmax = reinterpret_cast<float *>(sm + SM_MISC + 16);
    int *tile_e = reinterpret_cast<int *>(sm + SM_MISC + 48);                    // wmax: [2][4] floats
    float *tile_m = reinterpret_cast<float *>(sm + SM_MISC + 48 + 4 * E_RING);
    double *smat = reinterpret_cast<double *>(sm + SM_SMAT);
    float *fold = reinterpret_cast<float *>(sm + SM_FOLD);
    float *esm = reinterpret_cast<float *>(sm + SM_ESM);
    double *aggsm = reinterpret_cast<double *>(sm + SM_AGG);

    if (tid == 0) {
        for (int r = 0; r < NRAW; ++r) {
            mbar_init(RAW_FULL(r), 1);
            mbar_init(RAW_EMPTY(r), N_CVT_WARPS);
        }
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(A_FULL(s), N_CVT_WARPS);
            mbar_init(A_EMPTY(s), 2);                      // carry MMAs (warp 8) + output MMAs (warp 0) have read the stage
        }
        for (int b = 0; b < NACC; ++b) {
            mbar_init(E_FULL(b), 1);
            mbar_init(E_EMPTY(b), 1);
            mbar_init(D_FULL(b), 1);
            mbar_init(D_EMPTY(b), N_EPI_WARPS);
            mbar_init(S_READY(b), 1);
            mbar_init(SOP_EMPTY(b), 1);
        }
        for (int e = 0; e < ESM_DEPTH; ++e) {
            mbar_init(Z_READY(e), 1);
            mbar_init(ESM_READY(e), 1);
            mbar_init(ESM_EMPTY(e), 1);
        }
        fence_barrier_init();
    }
    if (tid < E_RING) tile_m[tid] = 0.f;
    // the correction operand rows hold 32 of 64 fp16: clear the rest once (0 * garbage could be NaN)
    for (int i = tid; i < 2 * SOP_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm + SM_SOP)[i] = make_uint4(0, 0, 0, 0);
    // float64 scans: all seven matrices; float32 scans: only A^8192 (slot 6 of the plan's table -> slot 0 here)
    for (int i = tid; i < LY::NSMAT_SM * ND * ND; i += NTHREADS) smat[i] = a.smat[(Z64 ? 0 : 6 * ND * ND) + i];
    for (int i = tid; i < LY::NFOLD_SM * FOLD_PITCH; i += NTHREADS) fold[i] = a.fold[i];
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- constant matrices -> TMEM (stay for the whole kernel); lane = matrix row ----
    if (warp < 4) {
        const uint4 *src = a.amat + (size_t)(warp * 32 + lane) * (A_COLS / 4);
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < A_COLS; c0 += 16) {
            uint32_t v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u = src[c0 / 4 + q];
                v[4 * q + 0] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
            }
            tmem_st16(taddr + c0, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // this CTA's contiguous run of tiles: [t_begin, t_end), outputs stored from t_own on
    const int64_t t_own = (int64_t)blockIdx.x * a.tiles_per_block;
    const int64_t t_begin = (blockIdx.x == 0) ? 0 : t_own - a.warm_tiles;
    const int64_t t_end = (t_own + a.tiles_per_block < a.n_tiles) ? t_own + a.tiles_per_block : a.n_tiles;

    if (warp >= CVT_WARP0) {
        // =============================== converters ===============================
        // thread ct owns float4 groups g = ct + 128 i: stream row (ct >> 4) + 8 i, i.e. a fixed row parity
        const int cw = warp - CVT_WARP0;                 // converter warp index 0..3
        const int ct = cw * 32 + lane;
        const int parity = (ct >> 4) & 1, c_first = ct >> 5, colb = (ct & 15) >> 1, sub = (ct & 1) * 8;
        int it = 0;
        for (int64_t tile = t_begin; tile < t_end; ++tile, ++it) {
            const int r = it % NRAW, s = it % NSTAGE;
            const unsigned char *raw = sm + SM_RAW + r * STAGE_BYTES;
            unsigned char *st = sm + s * STAGE_BYTES;
            STC_WAIT(RAW_FULL(r), (uint32_t)(it / NRAW) & 1u, 1);
            if (!tile_is_bulk(a, tile)) {
                // edge / unaligned / zero-stuffed tile: the converters fetch it themselves into the raw layout
                float *rawf = reinterpret_cast<float *>(sm + SM_RAW + r * STAGE_BYTES);
                const int64_t g0 = tile * (int64_t)TILE;
                if (a.L == 1) {
#pragma unroll 1
                    for (int i = ct; i < TILE; i += N_CVT) rawf[i] = (g0 + i < a.n_in) ? a.x[g0 + i] : 0.f;
                } else {
                    // sample g0 + i is x[q] * L when (g0 + i) == q * L: walk the multiples of L inside the tile
#pragma unroll 1
                    for (int i = ct; i < TILE; i += N_CVT) rawf[i] = 0.f;
                    asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));
                    const int64_t q0 = (g0 + a.L - 1) / a.L;                 // first input sample inside the tile
                    const float gain = (float)a.L;
#pragma unroll 1
                    for (int64_t q = q0 + ct; q < a.n_in; q += N_CVT) {
                        const int64_t i = q * a.L - g0;
                        if (i >= TILE) break;
                        rawf[i] = a.x[q] * gain;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));
            }
            // pass 1 (rolled): tile maximum.  The raw tile stays in shared memory and is read again by pass 2 -- both
            // passes are short loops that stay in the instruction cache
            float m = 0.f;
#pragma unroll 4
            for (int i = 0; i < G_PER_THREAD; ++i) {
                const float4 v = reinterpret_cast<const float4 *>(raw)[ct + i * N_CVT];
                m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            // wmax is double buffered by tile parity: one named barrier per tile is enough
            if (lane == 0) wmax[(it & 1) * N_CVT_WARPS + cw] = m;
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));
            float bm = 0.f;
#pragma unroll
            for (int w = 0; w < N_CVT_WARPS; ++w) bm = fmaxf(bm, wmax[(it & 1) * N_CVT_WARPS + w]);
            // block scale from the largest |x| of the last `win` tiles (the cascade's memory): the chunk start states
            // of this tile are then bounded by the plan-time gain times that maximum, so they fit the fp16 operand of
            // the correction GEMM whatever the signal does
            float wm = bm;
#pragma unroll 1
            for (int j = 1; j < a.win; ++j)
                if (it - j >= 0) wm = fmaxf(wm, tile_m[(it - j) % E_RING]);
            if (blockIdx.x == 0 && a.zi != nullptr && it < a.win) {
                for (int k = 0; k < a.d_real; ++k) wm = fmaxf(wm, fabsf(a.zi[k]) * a.ginv[k]);
            }
            int ex = a.xexp;
            if (wm > 0.f && wm < 3.0e38f) (void)frexpf(wm, &ex);
            const int e = max(-100, min(100, a.xexp - ex));
            const float sx = ldexpf(1.0f, e);
            if (ct == 0) {
                tile_m[it % E_RING] = bm;
                tile_e[it % E_RING] = e;
            }
            // pass 2 (rolled): fp32 -> fp16 hi + fp16 residual, into the swizzled operand regions of stage s
            STC_WAIT(A_EMPTY(s), ((uint32_t)(it / NSTAGE) & 1u) ^ 1u, 10);
            // chunk c = c_first + 4 i: its swizzle phase c & 7 alternates between two values
            unsigned char *wbase = st + (2 * parity) * REGION + c_first * 128 + sub;
            const uint32_t sw0 = (uint32_t)((colb ^ (c_first & 7)) << 4), sw1 = (uint32_t)((colb ^ ((c_first + 4) & 7)) << 4);
#pragma unroll 2
            for (int i = 0; i < G_PER_THREAD; ++i) {
                const float4 rv = reinterpret_cast<const float4 *>(raw)[ct + i * N_CVT];
                const float v0 = rv.x * sx, v1 = rv.y * sx, v2 = rv.z * sx, v3 = rv.w * sx;
                const __half2 h01 = __floats2half2_rn(v0, v1);
                const __half2 h23 = __floats2half2_rn(v2, v3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
                const __half2 l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t *>(&h01);
                hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                lv.x = *reinterpret_cast<const uint32_t *>(&l01);
                lv.y = *reinterpret_cast<const uint32_t *>(&l23);
                unsigned char *w = wbase + i * 512 + ((i & 1) ? sw1 : sw0);
                *reinterpret_cast<uint2 *>(w) = hv;
                *reinterpret_cast<uint2 *>(w + REGION) = lv;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(RAW_EMPTY(r));
                mbar_arrive(A_FULL(s));
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================== bulk-TMA producer + MMA issuer ===============================
        // One warp does both single-lane jobs: the tile NSTAGE-1 ahead is fetched into the stage the previous
        // tile's MMAs release, i.e. the wait for that stage ends exactly when the tensor pipe is ready for this
        // tile's MMAs (no separate producer warp: 16 warps keep 128 registers per thread).
        auto produce = [&](int64_t ptile, int pit) {
            const int pr = pit % NRAW;
            STC_WAIT(RAW_EMPTY(pr), ((uint32_t)(pit / NRAW) & 1u) ^ 1u, 0);
            if (elect_one()) {
                if (tile_is_bulk(a, ptile)) {
                    mbar_arrive_expect_tx(RAW_FULL(pr), STAGE_BYTES);
                    bulk_g2s(base + SM_RAW + pr * STAGE_BYTES, a.x + ptile * TILE, STAGE_BYTES, RAW_FULL(pr));
                } else {
                    mbar_arrive(RAW_FULL(pr));
                }
            }
            __syncwarp();
        };
        // carry accumulator (TMEM lanes 0-15: fp16-hi rows of the states, 16-31: their residual rows) -> sum, unscale,
        // transpose through shared memory into one row per chunk for the scan warps
        const bool lo_row = lane >= 16;
        const int d = lane & 15;
        const float rowinv = a.rowinv[d];
        auto eread = [&](int eit) {
            const int eb = eit & 1;
            const uint32_t epb = (uint32_t)(eit >> 1) & 1u;
            const int ee = eit % ESM_DEPTH;
            STC_WAIT(ESM_EMPTY(ee), ((uint32_t)(eit / ESM_DEPTH) & 1u) ^ 1u, 13);
            STC_WAIT(E_FULL(eb), epb, 14);
            tc_fence_after();
            float *es = esm + ee * NCHUNK * EP;
            const float sc = ldexpf(rowinv, -tile_e[eit % E_RING]);
            const uint32_t eaddr = tmem_base + ACC_COL0 + eb * ACC_STRIDE + NCHUNK;     // TMEM lanes 0..31
#pragma unroll 1
            for (int c0 = 0; c0 < NCHUNK; c0 += 16) {        // rolled: keeps the role's code inside the instruction cache
                uint32_t v[16];
                tmem_ld16(eaddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; q += 2) {
                    const float own_a = __uint_as_float(v[q]), own_b = __uint_as_float(v[q + 1]);
                    const float recv = __shfl_xor_sync(0xffffffffu, lo_row ? own_a : own_b, 16);
                    const float val = (recv + (lo_row ? own_b : own_a)) * sc;
                    if (d < ND) es[(c0 + q + (lo_row ? 1 : 0)) * EP + d] = val;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(E_EMPTY(eb));
                mbar_arrive(ESM_READY(ee));
            }
            __syncwarp();
        };
        // carry MMAs of tile eit: rows = states (fp16 hi rows in TMEM lanes 0-15, residual rows in 16-31),
        // residual (lo) stream parts before the hi parts
        auto issue_e = [&](int eit) {
            const int s = eit % NSTAGE;
            const int eb = eit & 1;
            const uint32_t st = base + s * STAGE_BYTES;
            const uint64_t xe_hi = make_desc_sw128(st + 0 * REGION), xe_lo = make_desc_sw128(st + 1 * REGION);
            const uint64_t xo_hi = make_desc_sw128(st + 2 * REGION), xo_lo = make_desc_sw128(st + 3 * REGION);
            STC_WAIT(A_FULL(s), (uint32_t)(eit / NSTAGE) & 1u, 2);
            STC_WAIT(E_EMPTY(eb), ((uint32_t)(eit >> 1) & 1u) ^ 1u, 3);
            tc_fence_after();
            if (elect_one()) {
                if (!(DBG && (a.dbg & 1))) {
                    const uint32_t ea = tmem_base + ACC_COL0 + eb * ACC_STRIDE + NCHUNK;
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) umma_f16_ts(ea, tmem_base + COL_KA + sl * 8, xe_lo + (uint64_t)(2 * sl), kIdesc, sl > 0);
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) umma_f16_ts(ea, tmem_base + COL_KB + sl * 8, xo_lo + (uint64_t)(2 * sl), kIdesc, 1);
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) umma_f16_ts(ea, tmem_base + COL_KA + sl * 8, xe_hi + (uint64_t)(2 * sl), kIdesc, 1);
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) umma_f16_ts(ea, tmem_base + COL_KB + sl * 8, xo_hi + (uint64_t)(2 * sl), kIdesc, 1);
                }
                umma_commit(E_FULL(eb));         // commit from the issuing thread: it tracks THAT thread's MMAs
                umma_commit(A_EMPTY(s));
            }
            __syncwarp();
        };
        // The carry path (this warp -> scan warps -> chain warp) never waits for the output path (the output MMAs are
        // issued by epilogue warp 0).  Loads: a tile's stage must be requested before its A_FULL is awaited (blocking);
        // beyond that the ring is topped up opportunistically (non-blocking probe of the stage barrier).
        const int n_it = (int)(t_end - t_begin);
        int next_prod = 0;
        auto top_up = [&](int need) {
            while (next_prod < n_it) {
                const int pr = next_prod % NRAW;
                const uint32_t par = ((uint32_t)(next_prod / NRAW) & 1u) ^ 1u;
                if (next_prod > need && !mbar_test(RAW_EMPTY(pr), par)) break;
                produce(t_begin + next_prod, next_prod);
                ++next_prod;
            }
        };
        top_up(0);
        if (n_it > 0) issue_e(0);
        for (int it = 0; it < n_it; ++it) {
            top_up(it + 1);
            if (it + 1 < n_it) issue_e(it + 1);            // the tensor pipe works on the next tile's carries ...
            eread(it);                                     // ... while this warp reads the current ones back
        }
    } else if (warp == ZSCAN_WARP0 || warp == ZSCAN_WARP1) {
        // =============================== carries -> zero-state chunk start states ===============================
        // Two warps take alternate tiles (warp b owns accumulator buffer b).  Everything here is independent of the
        // state entering the tile, so it is OFF the serial chain between tiles; float64 throughout.
        // Z64: float64 arithmetic (plans whose cascade decays slowly: the state is ill-conditioned against the
        // output); otherwise float32 -- the scan covers one tile from zero state, nothing accumulates across tiles.
        using ZT = typename std::conditional<Z64, double, float>::type;
        const int b = (warp == ZSCAN_WARP0) ? 0 : 1;
        const ZT *MA128, *MKS;                            // A^128 and A^256, A^512 ... (column-major)
        int mstride;
        if constexpr (Z64) { MA128 = smat; MKS = smat + ND * ND; mstride = ND * ND; }
        else { MA128 = fold + 12 * FOLD_PITCH; MKS = fold + 13 * FOLD_PITCH; mstride = FOLD_PITCH; }
        int it = b;
        for (int64_t tile = t_begin + b; tile < t_end; tile += 2, it += 2) {
            const int ee = it % ESM_DEPTH;
            const uint32_t pe = (uint32_t)(it / ESM_DEPTH) & 1u;
            float *es = esm + ee * NCHUNK * EP;
            STC_WAIT(ESM_READY(ee), pe, 5);
            // lane l owns chunks 2l, 2l+1
            ZT q[ND], o[ND], zs[ND];
            const float *e0 = es + (2 * lane) * EP, *e1 = e0 + EP;
            // One rolled loop, one matvec body (instruction-cache footprint):
            //   step 0            q = e1 + A^128 e0                      zero-state end state of the chunk pair
            //   steps 1..nlev     q += A^(256 2^j) q[lane - 2^j]         Kogge-Stone over the 32 pairs
            //   step nlev+1       zs = shfl_up(q) = state at the start of chunk 2l;  q = e0 + A^128 zs  (chunk 2l+1)
            const int nsteps = a.nlev + 2;
#pragma unroll 1
            for (int st = 0; st < nsteps; ++st) {
                const ZT *Mt = MA128;
                bool active = true;
                if (st == 0) {
#pragma unroll
                    for (int k = 0; k < ND; k += 4) {
                        const float4 u0 = *reinterpret_cast<const float4 *>(e0 + k);
                        const float4 u1 = *reinterpret_cast<const float4 *>(e1 + k);
                        o[k] = u0.x; o[k + 1] = u0.y; o[k + 2] = u0.z; o[k + 3] = u0.w;
                        q[k] = u1.x; q[k + 1] = u1.y; q[k + 2] = u1.z; q[k + 3] = u1.w;
                    }
                } else if (st <= a.nlev) {
                    const int off = 1 << (st - 1);
#pragma unroll
                    for (int k = 0; k < ND; ++k) o[k] = __shfl_up_sync(0xffffffffu, q[k], off);
                    Mt = MKS + (st - 1) * mstride;
                    active = lane >= off;
                } else {
#pragma unroll
                    for (int k = 0; k < ND; ++k) {
                        o[k] = __shfl_up_sync(0xffffffffu, q[k], 1);
                        if (lane == 0) o[k] = (ZT)0;
                        zs[k] = o[k];
                    }
                    if (lane == 31) {
#pragma unroll
                        for (int k = 0; k < ND; ++k) aggsm[ee * 16 + k] = (double)q[k];
                    }
#pragma unroll
                    for (int k = 0; k < ND; k += 4) {
                        const float4 u0 = *reinterpret_cast<const float4 *>(e0 + k);
                        q[k] = u0.x; q[k + 1] = u0.y; q[k + 2] = u0.z; q[k + 3] = u0.w;
                    }
                }
                if (active && !(DBG && (a.dbg & 2))) mv_acc<ND>(q, Mt, o);
            }
#pragma unroll
            for (int k = 0; k < ND; ++k) o[k] = zs[k];
            {
                // the zero-state start states replace this lane's two carry rows (same rows, same lane: in place)
                float *d0 = es + (2 * lane) * EP, *d1 = d0 + EP;
#pragma unroll
                for (int k = 0; k < ND; k += 4) {
                    *reinterpret_cast<float4 *>(d0 + k) = make_float4((float)o[k], (float)o[k + 1], (float)o[k + 2], (float)o[k + 3]);
                    *reinterpret_cast<float4 *>(d1 + k) = make_float4((float)q[k], (float)q[k + 1], (float)q[k + 2], (float)q[k + 3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(Z_READY(ee));
        }
    } else if (warp == CHAIN_WARP) {
        // =============================== serial chain between tiles ===============================
        // s_in(t+1) = A^8192 s_in(t) + aggregate(t) in float64 (replicated in every lane), and the fold of s_in into
        // the tile's chunk start states:  S_c = Z_c + A^(128 c) s_in.  The fold is a direct float32 product of
        // the float64 carry (no accumulation): lane l = 4 h + m owns chunks 2l, 2l+1 and applies
        // A^(256 m) A^(1024 h), then A^128 for the odd chunk.
        double s_in[ND];
#pragma unroll
        for (int k = 0; k < ND; ++k)
            s_in[k] = (blockIdx.x == 0 && a.zi != nullptr && k < a.d_real) ? (double)a.zi[k] : 0.0;
        const float *Ph = fold + (lane >> 2) * FOLD_PITCH, *Qm = fold + (8 + (lane & 3)) * FOLD_PITCH;
        const float *F128 = fold + 12 * FOLD_PITCH;
        const double *MAT = smat + (Z64 ? 6 : 0) * ND * ND;            // A^8192
        int it = 0;
        for (int64_t tile = t_begin; tile < t_end; ++tile, ++it) {
            const int b = it & 1;
            const uint32_t pb = (uint32_t)(it >> 1) & 1u;
            // fold: h0 = A^(256 m) A^(1024 h) s_in, h1 = A^128 h0 -- one matvec body, three passes (code footprint)
            float h0[ND], h1[ND];
#pragma unroll
            for (int k = 0; k < ND; ++k) h1[k] = (float)s_in[k];
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                const float *Mt = (pass == 0) ? Ph : (pass == 1 ? Qm : F128);
#pragma unroll
                for (int k = 0; k < ND; ++k) { h0[k] = h1[k]; }
                mv_f32<ND>(h1, Mt, h0);
            }
            // now h1 = chunk 2l+1's share, h0 = chunk 2l's share (the input of the last pass)
            const int ee = it % ESM_DEPTH;
            STC_WAIT(Z_READY(ee), (uint32_t)(it / ESM_DEPTH) & 1u, 8);
            double nxt[ND];
#pragma unroll
            for (int k = 0; k < ND; ++k) nxt[k] = aggsm[ee * 16 + k];
            if (!a.at_zero) mv_acc<ND>(nxt, MAT, s_in);
            const float *d0 = esm + (ee * NCHUNK + 2 * lane) * EP, *d1 = d0 + EP;
#pragma unroll
            for (int k = 0; k < ND; k += 4) {
                const float4 z0 = *reinterpret_cast<const float4 *>(d0 + k), z1 = *reinterpret_cast<const float4 *>(d1 + k);
                h0[k] += z0.x; h0[k + 1] += z0.y; h0[k + 2] += z0.z; h0[k + 3] += z0.w;     // true start state of chunk 2l
                h1[k] += z1.x; h1[k + 1] += z1.y; h1[k + 2] += z1.z; h1[k + 3] += z1.w;     //                     chunk 2l+1
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ESM_EMPTY(ee));           // carries / start states / aggregate of this buffer consumed
            // the two chunk rows of the correction operand: fp16 hi (k 0..15) and residual (k 16..31) of s_d * sfac[d] * 2^e_x
            STC_WAIT(SOP_EMPTY(b), pb ^ 1u, 17);
            {
                const float px = ldexpf(1.0f, tile_e[it % E_RING]);
                unsigned char *sop = sm + SM_SOP + b * SOP_BYTES;
                float hv[ND];
#pragma unroll
                for (int k = 0; k < ND; ++k) hv[k] = h0[k];
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const int c = 2 * lane + hh;
                    if (hh) {
#pragma unroll
                        for (int k = 0; k < ND; ++k) hv[k] = h1[k];
                    }
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int k = 0; k < 16; k += 2) {
                        float v0 = 0.f, v1 = 0.f;
                        if (k < ND) {
                            v0 = hv[k] * (a.sfac[k] * px);
                            v1 = hv[k + 1] * (a.sfac[k + 1] * px);
                            v0 = fminf(fmaxf(v0, -60000.f), 60000.f);          // never reached inside the plan's gain bound
                            v1 = fminf(fmaxf(v1, -60000.f), 60000.f);
                        }
                        const __half2 h = __floats2half2_rn(v0, v1);
                        const float2 f = __half22float2(h);
                        const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
                        hi[k >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                        lo[k >> 1] = *reinterpret_cast<const uint32_t *>(&l);
                    }
                    unsigned char *row = sop + c * 128;
                    const int sw = c & 7;
                    *reinterpret_cast<uint4 *>(row + ((0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(row + ((1 ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<uint4 *>(row + ((2 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4 *>(row + ((3 ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(S_READY(b));

            // final state of the stream: advance the start state of the chunk that holds sample n_rate
            if (a.zf != nullptr && tile == a.n_tiles - 1) {
                const int64_t rem = a.n_rate - tile * (int64_t)TILE;          // 1 .. 8192 samples of this tile are real
                const int cs = (int)(rem / LC);                                // chunk that holds position n_rate
                const int r = (int)(rem - (int64_t)cs * LC);
                if ((cs == NCHUNK && lane == 31) || (cs < NCHUNK && lane == (cs >> 1))) {
                    double z[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) z[k] = 0.0;
#pragma unroll
                    for (int k = 0; k < ND; ++k) z[k] = (cs == NCHUNK) ? nxt[k] : (double)((cs & 1) ? h1[k] : h0[k]);
                    const int64_t g0 = tile * (int64_t)TILE + (int64_t)cs * LC;
#pragma unroll 1
                    for (int i = 0; i < r; ++i) {
                        double v = (double)load_rate_sample(a, g0 + i);
#pragma unroll
                        for (int sct = 0; sct < ND / 2; ++sct) {
                            const double xn = fma(a.coef[sct][0], v, z[2 * sct]);
                            z[2 * sct] = fma(a.coef[sct][1], v, fma(a.coef[sct][3], xn, z[2 * sct + 1]));
                            z[2 * sct + 1] = fma(a.coef[sct][2], v, a.coef[sct][4] * xn);
                            v = xn;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < ND; ++k)
                        if (k < a.d_real) a.zf[k] = (float)z[k];
                }
            }
#pragma unroll
            for (int k = 0; k < ND; ++k) s_in[k] = nxt[k];
        }
    } else {
        // =============================== epilogue (warps 0-7) ===============================
        // TMEM lane = position inside the 128-sample chunk, column = chunk: a warp's store instruction writes 32
        // consecutive samples (128 bytes)
        const int qd = warp & 3, half = warp >> 2;
        const int rho = qd * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        const int colbase = half * 32;
        // epilogue warp 0 also issues the output MMAs of tile it+2 as soon as tile it has left its accumulator: zero-state
        // outputs, rows = the 128 positions of a chunk, K = the chunk's 128 samples (even row then odd row); three
        // partial products, smallest first: T_hi x_lo, T_lo x_hi, T_hi x_hi
        auto issue_y = [&](int yit) {
            const int s = yit % NSTAGE;
            const int yb = yit & 1;
            const uint32_t st = base + s * STAGE_BYTES;
            const uint32_t acc = tmem_base + ACC_COL0 + yb * ACC_STRIDE;
            const uint64_t xe_hi = make_desc_sw128(st + 0 * REGION), xe_lo = make_desc_sw128(st + 1 * REGION);
            const uint64_t xo_hi = make_desc_sw128(st + 2 * REGION), xo_lo = make_desc_sw128(st + 3 * REGION);
            STC_WAIT(D_EMPTY(yb), ((uint32_t)(yit >> 1) & 1u) ^ 1u, 4);
            STC_WAIT(A_FULL(s), (uint32_t)(yit / NSTAGE) & 1u, 12);
            tc_fence_after();
            if (elect_one()) {
                if (!(DBG && (a.dbg & 1))) {
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl)
                        umma_f16_ts(acc, tmem_base + COL_THI + sl * 8, (sl < 4 ? xe_lo : xo_lo) + (uint64_t)(2 * (sl & 3)), kIdesc, sl > 0);
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl)
                        umma_f16_ts(acc, tmem_base + COL_TLO + sl * 8, (sl < 4 ? xe_hi : xo_hi) + (uint64_t)(2 * (sl & 3)), kIdesc, 1);
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl)
                        umma_f16_ts(acc, tmem_base + COL_THI + sl * 8, (sl < 4 ? xe_hi : xo_hi) + (uint64_t)(2 * (sl & 3)), kIdesc, 1);
                }
                umma_commit(A_EMPTY(s));
            }
            __syncwarp();
        };
        // ... and the correction  Y += O S  of a tile once the chain warp has written its chunk start states; the commit
        // after it covers the tile's output MMAs as well (same issuing thread, same accumulator)
        auto issue_corr = [&](int cit) {
            const int cb = cit & 1;
            const uint32_t acc = tmem_base + ACC_COL0 + cb * ACC_STRIDE;
            const uint64_t sd = make_desc_sw128(base + SM_SOP + cb * SOP_BYTES);
            STC_WAIT(S_READY(cb), (uint32_t)(cit >> 1) & 1u, 8);
            tc_fence_after();
            if (elect_one()) {
                if (!(DBG && (a.dbg & 1))) {
                    umma_f16_ts(acc, tmem_base + COL_OLO, sd, kIdesc, 1);                       // O_lo S_hi
                    umma_f16_ts(acc, tmem_base + COL_OHI, sd + 2, kIdesc, 1);                   // O_hi S_lo
                    umma_f16_ts(acc, tmem_base + COL_OHI, sd, kIdesc, 1);                       // O_hi S_hi
                }
                umma_commit(D_FULL(cb));
                umma_commit(SOP_EMPTY(cb));
            }
            __syncwarp();
        };
        const int n_it = (int)(t_end - t_begin);
        int corr_next = 0;                                // next tile whose correction has not been issued (warp 0)
        for (int it = -2; it < n_it; ++it) {
            if (warp == 0) {
                // keep the tensor pipe's queue in the order the data becomes ready: a correction that can go now goes
                // before the next output MMAs (it is three MMAs and unblocks all eight epilogue warps)
                if (it >= 0 && corr_next == it) { issue_corr(it); ++corr_next; }
            }
            if (it < 0) {
                if (warp == 0 && it + 2 < n_it) issue_y(it + 2);
                continue;
            }
            const int64_t tile = t_begin + it;
            const int b = it & 1;
            const uint32_t pb = (uint32_t)(it >> 1) & 1u;
            STC_WAIT(D_FULL(b), pb, 7);
            tc_fence_after();
            long long tp0 = DBG ? clock64() : 0;
            const uint32_t taddr = tmem_base + lane_sel + ACC_COL0 + b * ACC_STRIDE + colbase;
            const float inv = ldexpf(1.0f, -tile_e[it % E_RING] - a.e_t);
            const int64_t tile0 = tile * (int64_t)TILE;
            // samples of this tile that exist and are to be stored (0 for warm-up tiles)
            int limit = 0;
            if (tile >= t_own && !(DBG && (a.dbg & 4))) limit = (int)((a.n_rate - tile0 < TILE) ? a.n_rate - tile0 : TILE);
            int p = colbase * LC + rho;                    // position inside the tile
            float *yp = a.y + tile0 + p;
            // decimating stores keep positions g with g % M == 0: (m, r) = (g / M, g % M) walked without divisions
            int64_t m = 0;
            int r = 0, step_m = 0, step_r = 0;
            if (a.M > 1) {
                const int64_t g = tile0 + p;
                m = g / a.M;
                r = (int)(g - m * a.M);
                step_m = LC / a.M;
                step_r = LC - step_m * a.M;
            }
            // rolled over groups of 8 chunks: the loop body stays resident in the instruction cache
#pragma unroll 1
            for (int c0 = 0; c0 < 32; c0 += 8) {
                uint32_t dv[8];
                tmem_ld8(taddr + c0, dv);
                tmem_ld_wait();
                if (c0 == 24) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(D_EMPTY(b));
                }
                float acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = __uint_as_float(dv[c]) * inv;
                if (a.M == 1) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (p + c * LC < limit) yp[c * LC] = acc[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (r == 0 && p + c * LC < limit && m < a.n_out) a.y[m] = acc[c];
                        m += step_m;
                        r += step_r;
                        if (r >= a.M) { r -= a.M; ++m; }
                    }
                }
                p += 8 * LC;
                yp += 8 * LC;
            }
            __syncwarp();
            if (DBG) ph_b += clock64() - tp0;
            if (warp == 0) {
                // tile it+1's correction first if its states are already there, then the output MMAs of tile it+2
                if (it + 1 < n_it && corr_next == it + 1 && mbar_test(S_READY((it + 1) & 1), (uint32_t)((it + 1) >> 1) & 1u)) {
                    issue_corr(it + 1);
                    ++corr_next;
                }
                if (it + 2 < n_it) issue_y(it + 2);
            }
        }
    }
    if (DBG && (a.dbg & 8) && blockIdx.x == 1 && lane == 0)
        printf("stc warp %2d total %lld (%llu ns) waits %lld phase a %lld b %lld c %lld\n", warp, clock64() - t_start,
               (unsigned long long)(global_ns() - ns_start), wait_cycles, ph_a, ph_b, ph_c);
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace stc

// ------------------------------------------------------------------------------------------ host
namespace {
using Mat = std::vector<double>;
void mm(const Mat &a, const Mat &b, Mat &c, int D) {
    Mat r((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) {
            const double aik = a[i * D + k];
            if (aik == 0.0) continue;
            for (int j = 0; j < D; ++j) r[i * D + j] += aik * b[k * D + j];
        }
    c.swap(r);
}
double maxabs(const Mat &a) {
    double m = 0.0;
    for (double v : a) m = fmax(m, fabs(v));
    return m;
}
int scale_exp(double mx) {
    if (!(mx > 0.0)) return 0;
    int ex = 0;
    (void)frexp(mx, &ex);
    int e = 12 - ex;
    return e > 100 ? 100 : (e < -100 ? -100 : e);
}
}  // namespace

int stc_build(const double (*coef)[5], int nsec, int nsec_real, StcTables *t)
{
    using namespace stc;
    t->ok = false;
    if (nsec < 1 || nsec > STC_MAXSEC) return B200DSP_OK;
    const int D = 2 * nsec;
    const int ND = (D + 3) & ~3;
    // one step of the cascade in float64: z <- next state, returns the output
    auto step = [&](std::vector<double> &z, double v) {
        for (int s = 0; s < nsec; ++s) {
            const double xn = coef[s][0] * v + z[2 * s];
            const double z0 = coef[s][1] * v + coef[s][3] * xn + z[2 * s + 1];
            const double z1 = coef[s][2] * v + coef[s][4] * xn;
            z[2 * s] = z0;
            z[2 * s + 1] = z1;
            v = xn;
        }
        return v;
    };
    Mat A((size_t)ND * ND, 0.0);
    std::vector<double> Cv(ND, 0.0), Bv(ND, 0.0);
    for (int dd = 0; dd < D; ++dd) {
        std::vector<double> z(D, 0.0);
        z[dd] = 1.0;
        Cv[dd] = step(z, 0.0);
        for (int i = 0; i < D; ++i) A[i * ND + dd] = z[i];
    }
    double Dd;
    {
        std::vector<double> z(D, 0.0);
        Dd = step(z, 1.0);
        for (int i = 0; i < D; ++i) Bv[i] = z[i];
    }
    // w_k = A^k B (k < 128),  o_n = C A^n (n < 128),  h[0] = D, h[k] = C A^(k-1) B
    std::vector<std::vector<double>> w(LC, std::vector<double>(ND, 0.0)), orow(LC, std::vector<double>(ND, 0.0));
    w[0] = Bv;
    orow[0] = Cv;
    for (int k = 1; k < LC; ++k) {
        for (int i = 0; i < ND; ++i) {
            double s = 0.0, r = 0.0;
            for (int j = 0; j < ND; ++j) {
                s += A[i * ND + j] * w[k - 1][j];
                r += orow[k - 1][j] * A[j * ND + i];
            }
            w[k][i] = s;
            orow[k][i] = r;
        }
    }
    std::vector<double> h(LC);
    h[0] = Dd;
    for (int k = 1; k < LC; ++k) {
        double s = 0.0;
        for (int j = 0; j < ND; ++j) s += Cv[j] * w[k - 1][j];
        h[k] = s;
    }
    // scan matrices A^128, A^256 ... A^8192 (the last one is also the tile matrix) and the warm-up length
    Mat P = A, AT;
    for (int i = 0; i < 7; ++i) mm(P, P, P, ND);          // A^128
    std::vector<double> smat((size_t)NSMAT * ND * ND);
    for (int j = 0; j < NSMAT; ++j) {
        for (int i = 0; i < ND; ++i)
            for (int k = 0; k < ND; ++k) smat[(size_t)j * ND * ND + k * ND + i] = P[i * ND + k];    // column-major
        AT = P;
        mm(P, P, P, ND);
    }
    for (double v : smat)
        if (!std::isfinite(v)) return B200DSP_OK;
    int warm = 1;
    {
        Mat Q = AT;                                        // A^8192
        while (maxabs(Q) * ND > 1e-13 && warm < MAXWIN) {
            mm(Q, AT, Q, ND);
            ++warm;
        }
        if (maxabs(Q) * ND > 1e-13) return B200DSP_OK;     // decays too slowly: scan kernels
    }
    // Kogge-Stone levels that matter: A^(256 * 2^j) below float64 resolution of the state contributes nothing
    int nlev = 0;
    for (int j = 0; j < 5; ++j) {
        double m = 0.0;
        for (int i = 0; i < ND * ND; ++i) m = fmax(m, fabs(smat[(size_t)(1 + j) * ND * ND + i]));
        if (m * ND >= 1e-17) nlev = j + 1;
    }
    const int at_zero = (maxabs(AT) * ND < 1e-17) ? 1 : 0;
    // float32 matrices (column-major, pitch ND): A^(1024 h) h < 8 | A^(256 m) m < 4 | A^128 | A^256 ... A^4096
    std::vector<float> foldm((size_t)NFOLD * FOLD_PITCH, 0.f);
    {
        auto put = [&](int slot, const Mat &M) {
            for (int i = 0; i < ND; ++i)
                for (int k = 0; k < ND; ++k) foldm[(size_t)slot * FOLD_PITCH + k * ND + i] = (float)M[i * ND + k];
        };
        auto from_smat = [&](int j) {
            Mat M((size_t)ND * ND);
            for (int i = 0; i < ND; ++i)
                for (int k = 0; k < ND; ++k) M[i * ND + k] = smat[(size_t)j * ND * ND + k * ND + i];
            return M;
        };
        Mat I((size_t)ND * ND, 0.0);
        for (int i = 0; i < ND; ++i) I[i * ND + i] = 1.0;
        const Mat A128 = from_smat(0), A256 = from_smat(1), A1024 = from_smat(3);
        Mat Pw = I;
        for (int hh = 0; hh < 8; ++hh) {
            put(hh, Pw);
            mm(Pw, A1024, Pw, ND);
        }
        Pw = I;
        for (int m = 0; m < 4; ++m) {
            put(8 + m, Pw);
            mm(Pw, A256, Pw, ND);
        }
        put(12, A128);
        for (int j = 0; j < 5; ++j) put(13 + j, from_smat(1 + j));
    }
    double hmax = 0.0;
    for (double v : h) hmax = fmax(hmax, fabs(v));
    const int e_t = scale_exp(hmax);
    int e_row[16];
    for (int dd = 0; dd < 16; ++dd) {
        double m = 0.0;
        if (dd < ND)
            for (int k = 0; k < LC; ++k) m = fmax(m, fabs(w[k][dd]));
        e_row[dd] = scale_exp(m);
    }
    // Correction operand scales.  g_d = sum_k |(A^k B)_d| bounds |s_d| <= g_d max|x| (over the cascade's memory =
    // the window the block scale looks at), so with f_d = ceil(log2 g_d) the operand s_d 2^-f_d never exceeds the
    // window maximum; the rows of O carry 2^f_d back.  e_o puts the largest |O_d 2^f_d| just under 2^14.
    int f_exp[16];
    {
        std::vector<double> g(ND, 0.0), wk = Bv, wn(ND);
        const long steps = (long)warm * STC_TILE;
        for (long k = 0; k < steps; ++k) {
            for (int i = 0; i < ND; ++i) g[i] += fabs(wk[i]);
            for (int i = 0; i < ND; ++i) {
                double acc = 0.0;
                for (int j = 0; j < ND; ++j) acc += A[i * ND + j] * wk[j];
                wn[i] = acc;
            }
            wk.swap(wn);
        }
        for (int dd = 0; dd < 16; ++dd) {
            int ex = 0;
            if (dd < ND && g[dd] > 0.0) (void)frexp(g[dd], &ex);     // g < 2^ex
            f_exp[dd] = ex;
        }
    }
    double omax = 0.0;
    for (int n = 0; n < LC; ++n)
        for (int dd = 0; dd < ND; ++dd) omax = fmax(omax, fabs(ldexp(orow[n][dd], f_exp[dd])));
    int e_o = 0;
    if (omax > 0.0) {
        int ex = 0;
        (void)frexp(omax, &ex);
        e_o = 14 - ex;
    }
    // |s_d| sfac_d 2^e_x <= max|x| 2^e_x 2^(e_t - e_o) < 2^(xexp + e_t - e_o) must stay below 2^15 (one bit spare)
    int xexp = 14 - (e_t - e_o);
    if (xexp > 12) xexp = 12;
    if (xexp < 2) return B200DSP_OK;                       // states too large against the outputs for the fp16 operand
    // TMEM image, one row per lane (192 columns = 384 fp16):
    //   T_hi | T_lo : lane rho = position inside the chunk, k = sample of the chunk: h[rho - k] (k <= rho), fp16 hi / residual
    //   Ka | Kb     : lanes 0-15 = fp16 hi of state row d = lane, lanes 16-31 = its residual (summed by the scan warp)
    std::vector<__half> img((size_t)128 * 2 * A_COLS, __float2half_rn(0.f));
    for (int rho = 0; rho < 128; ++rho) {
        __half *row = img.data() + (size_t)rho * 2 * A_COLS;
        for (int kk = 0; kk < LC; ++kk) {
            const double v = (kk <= rho) ? ldexp(h[rho - kk], e_t) : 0.0;
            const __half hi = __float2half_rn((float)v);
            row[2 * COL_THI + kk] = hi;
            row[2 * COL_TLO + kk] = __float2half_rn((float)(v - (double)__half2float(hi)));
        }
        for (int k = 0; k < 16; ++k) {
            const double v = (k < ND) ? ldexp(orow[rho][k], f_exp[k] + e_o) : 0.0;
            const __half hi = __float2half_rn((float)v);
            row[2 * COL_OHI + k] = hi;
            row[2 * COL_OLO + k] = __float2half_rn((float)(v - (double)__half2float(hi)));
        }
        const int dd = rho & 15;
        const bool lo = (rho & 16) != 0;
        for (int kk = 0; kk < BK; ++kk)
            for (int ab = 0; ab < 2; ++ab) {
                double v = 0.0;
                if (rho < 32 && dd < ND) v = ldexp(w[(ab == 0 ? LC - 1 : BK - 1) - kk][dd], e_row[dd]);
                const __half hi = __float2half_rn((float)v);
                const __half lw = __float2half_rn((float)(v - (double)__half2float(hi)));
                row[2 * (ab == 0 ? COL_KA : COL_KB) + kk] = lo ? lw : hi;
            }
    }
    std::vector<float> rowinv(16, 1.f);
    for (int dd = 0; dd < 16; ++dd) rowinv[dd] = (float)ldexp(1.0, -e_row[dd]);

    t->nd = ND;
    t->d_real = 2 * nsec_real;
    t->e_t = e_t;
    t->warm_tiles = warm;
    t->nlev = nlev;
    t->at_zero = at_zero;
    t->xexp = xexp;
    for (int dd = 0; dd < 16; ++dd) {
        t->sfac[dd] = (float)ldexp(1.0, e_t - e_o - f_exp[dd]);
        t->ginv[dd] = (float)ldexp(1.0, -f_exp[dd]);
    }
    memset(t->coef, 0, sizeof(t->coef));
    for (int s = 0; s < nsec; ++s)
        for (int q = 0; q < 5; ++q) t->coef[s][q] = coef[s][q];
    cudaError_t e = cudaMalloc(&t->amat, img.size() * sizeof(__half));
    if (e == cudaSuccess) e = cudaMalloc(&t->smat, smat.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&t->rowinv, rowinv.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&t->fold, foldm.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(t->fold, foldm.data(), foldm.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->amat, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->smat, smat.data(), smat.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->rowinv, rowinv.data(), rowinv.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        set_error("sos_plan_create (tensor-core tables): %s", cudaGetErrorString(e));
        stc_free(t);
        return B200DSP_E_CUDA;
    }
    t->ok = true;
    return B200DSP_OK;
}

void stc_free(StcTables *t)
{
    cudaFree(t->amat);
    cudaFree(t->smat);
    cudaFree(t->rowinv);
    cudaFree(t->fold);
    t->fold = nullptr;
    t->amat = nullptr;
    t->smat = nullptr;
    t->rowinv = nullptr;
    t->ok = false;
}

static void stc_geometry(const StcTables &t, int64_t n_rate, int sm_count, int64_t *n_tiles, int *nblk, int *tpb)
{
    const int64_t nt = (n_rate + STC_TILE - 1) / STC_TILE;
    // a block must be longer than its warm-up, or the discarded work shows (long streams: >= 200 tiles per block)
    const int64_t min_tpb = 4 * (int64_t)t.warm_tiles;
    int64_t blocks = nt / min_tpb;
    if (blocks > sm_count) blocks = sm_count;
    if (blocks < 1) blocks = 1;
    const int64_t per = (nt + blocks - 1) / blocks;
    *n_tiles = nt;
    *tpb = (int)per;
    *nblk = (int)((nt + per - 1) / per);
}

bool stc_usable(const StcTables &t, int64_t n_rate, int sm_count, bool force)
{
    if (!t.ok || n_rate < 1) return false;
    if (force) return true;
    if (n_rate < (int64_t)1 << 17) return false;       // measured: faster than the scan kernels from 2^18 samples on
    int64_t nt;
    int nblk, tpb;
    stc_geometry(t, n_rate, sm_count, &nt, &nblk, &tpb);
    (void)nblk;
    (void)sm_count;
    return nt < ((int64_t)1 << 40);
}

int launch_sos_tc(const StcTables &t, const float *x, float *y, int64_t n_in, int64_t n_rate, int64_t n_out,
                  int32_t L, int32_t M, const float *zi, float *zf, int sm_count, cudaStream_t stream)
{
    using namespace stc;
    Args a;
    a.x = x;
    a.y = y;
    a.amat = static_cast<const uint4 *>(t.amat);
    a.smat = t.smat;
    a.fold = t.fold;
    a.nlev = t.nlev;
    a.at_zero = t.at_zero;
    a.xexp = t.xexp;
    a.win = t.warm_tiles + 1 > 8 ? 8 : t.warm_tiles + 1;
    memcpy(a.sfac, t.sfac, sizeof(a.sfac));
    memcpy(a.ginv, t.ginv, sizeof(a.ginv));
    a.rowinv = t.rowinv;
    a.zi = zi;
    a.zf = zf;
    a.n_in = n_in;
    a.n_rate = n_rate;
    a.n_out = n_out;
    a.L = L;
    a.M = M;
    int nblk, tpb;
    stc_geometry(t, n_rate, sm_count, &a.n_tiles, &nblk, &tpb);
    a.tiles_per_block = tpb;
    a.warm_tiles = t.warm_tiles;
    a.e_t = t.e_t;
    a.d_real = t.d_real;
    a.dbg = 0;
    if (const char *e = getenv("B200DSP_STC_DBG")) a.dbg = atoi(e);
    // slowly decaying cascades (more than one warm-up tile) keep the in-tile scan in float64
    bool z64 = t.warm_tiles > 1;
    if (const char *e = getenv("B200DSP_STC_Z64")) z64 = atoi(e) != 0;
    memcpy(a.coef, t.coef, sizeof(a.coef));
#define STC_LAUNCH(NDV)                                                                  \
    {                                                                                    \
        auto kern = a.dbg ? (z64 ? sos_tc_kernel<NDV, true, true> : sos_tc_kernel<NDV, false, true>)      \
                          : (z64 ? sos_tc_kernel<NDV, true, false> : sos_tc_kernel<NDV, false, false>);   \
        const int smem_total = z64 ? Lay<true>::SMEM_TOTAL : Lay<false>::SMEM_TOTAL;     \
        B200_CHECK_CUDA(allow_smem(kern, smem_total));                                   \
        kern<<<(unsigned)nblk, NTHREADS, smem_total, stream>>>(a);                       \
    }
    switch (t.nd) {
    case 4: STC_LAUNCH(4) break;
    case 8: STC_LAUNCH(8) break;
    case 12: STC_LAUNCH(12) break;
    case 16: STC_LAUNCH(16) break;
    default:
        set_error("sos_tc: bad state count %d", t.nd);
        return B200DSP_E_BADARG;
    }
#undef STC_LAUNCH
    B200_CHECK_LAUNCH("sos_tc_kernel");
    return B200DSP_OK;
}

}  // namespace b200dsp
