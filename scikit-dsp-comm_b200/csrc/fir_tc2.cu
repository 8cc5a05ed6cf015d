// fir_tc2.cu -- complex64 FIR on tcgen05, "taps-stationary" formulation (second generation).
//
// fir_tc.cu showed that the block-Toeplitz GEMM is limited by the tensor core's SHARED-MEMORY operand
// fetch (~64 B/clk/SM: 96 cycles per 128x64x16 MMA instead of 32; ncu: tensor pipe 33 % active).
// Here the tap matrix -- the operand that never changes -- lives in TENSOR MEMORY for the whole
// kernel, so only the sample stream is fetched from shared memory:
//
//     D[rho, r] = sum_{kk<320} A[rho, kk] * B[r, kk]
//       A (TMEM, 128 lanes x 160 columns, written once per CTA with tcgen05.st):
//           row rho = (c, part): Toeplitz row T[c][kk] = b[c + 256 - kk] * 2^sb, part in {fp16 hi, fp16 lo};
//           rows are interleaved so that lanes l and l+16 of one warp hold hi and lo of the same c
//       B (shared memory, K-major SWIZZLE_128B, never materialised as a Hankel matrix):
//           row r = the fp16 sample stream starting at sample t0 - 256 + 64 r: k-block j of the UMMA
//           descriptor simply starts 128*j bytes later in the SAME stream (one swizzle row = 64 fp16)
//       D (TMEM accumulators, fp32): lanes = (c, part), columns = r  ->  y[t0 + 64 r + c]
//
// One MMA (M=128, N=96, K=16) reads 3 KB from shared memory and produces both the b_hi*x and b_lo*x products.
// Per complex sample: 2 channels x 2 sample parts (x_lo first, then x_hi) x 20 k-slices = 80 MMAs per 96-row tile;
// all four partial products are accumulated (the epilogue adds the hi and lo lanes): y = D / (s_x s_b).
//
// Data movement: a producer warp streams raw complex64 tiles (51 KB for 96 rows) into a 4-deep shared-memory
// ring with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx); 8 converter warps turn a stage IN PLACE
// into four swizzled fp16 streams (re/im x hi/lo) around a per-tile power-of-two scale; one elected lane of the
// MMA warp consumes it; tcgen05.commit recycles the stage.  A ring of three accumulators lets two epilogue
// warpgroups (TMEM -> registers -> shuffle-combine hi/lo lanes -> coalesced complex64 stores) overlap the next
// channel's MMAs.  Tile heights 64 / 80 / 128 are kept as tuning variants.
#include "tc_common.cuh"

namespace b200dsp {
namespace tc2 {
using namespace tcx;

constexpr int BK = 64;                       // fp16 per 128-byte swizzle row = outputs per GEMM column block
constexpr int NKB = 5;                       // k-blocks (256 taps + 64) / 64
constexpr int KTOT = NKB * BK;               // 320
template <int TN> struct Cfg {
    static constexpr int TILE_N = TN;                    // stream rows (GEMM N) per tile: 64 or 128
    static constexpr int TILE = TILE_N * BK;             // complex outputs per tile
    static constexpr int ROWS = TILE_N + NKB - 1;        // stream rows staged per tile
    static constexpr int TILE_IN = ROWS * BK;            // complex samples staged per tile
    static constexpr int RAW_BYTES = TILE_IN * 8;
    static constexpr int STREAM_BYTES = ((ROWS * 128 + 1023) / 1024) * 1024;
    static constexpr int STAGE_BYTES = 4 * STREAM_BYTES; // >= RAW_BYTES: converted in place
    static constexpr int NSTAGE = (TN == 64) ? 6 : (TN == 80 ? 5 : (TN == 96 ? 4 : 3));
    static constexpr int SMEM_BAR_OFF = NSTAGE * STAGE_BYTES;
    static constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 512 + 1024;
    static constexpr int ACC_BUF_COLS = TILE_N;          // one fp32 accumulator (128 lanes x TILE_N stream rows) per (tile, channel)
    static constexpr int NACC = (TN == 64) ? 4 : (TN == 80 ? 4 : (TN == 96 ? 3 : 2));   // accumulator ring depth (160 + NACC*TN <= 512)
    static constexpr int F4_PER_TILE = TILE_IN / 2;
    static constexpr int F4_PER_THREAD = (F4_PER_TILE + 255) / 256;
    static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    static_assert((4 * STREAM_BYTES) >= (TILE_IN * 8), "in-place conversion: stage must hold the raw tile");
};
constexpr int HALO = (NKB - 1) * BK;         // 256

constexpr int A_COLS = KTOT / 2;             // 160 TMEM columns of packed fp16 pairs

constexpr int N_EPI_WARPS = 8;               // warps 0-3 and 14-17: a warp reads TMEM lanes 32*(warp%4)..+31; the
                                             // two groups split the accumulator columns between them
constexpr int EPI2_WARP0 = 14;
constexpr int MMA_WARP = 4;
constexpr int PROD_WARP = 5;
constexpr int CVT_WARP0 = 6;
constexpr int N_CVT_WARPS = 8;
constexpr int N_CVT = N_CVT_WARPS * 32;      // 256
constexpr int NTHREADS = (EPI2_WARP0 + 4) * 32;              // 576
constexpr int INV_RING = 16;

// bring-up instrumentation: cycles spent inside a wait, accumulated per call site
template <bool DBG>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int who, long long &acc) {
    if constexpr (DBG) {
        long long t0 = clock64();
        mbar_wait(bar, parity, who);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity, who);
    }
}
struct Args {
    const float2 *x;
    const float2 *hist;
    float2 *y;
    const uint4 *amat;        // device: 128 rows x 320 fp16 (row = TMEM lane), row-major
    int64_t n;
    int64_t n_tiles;
    int32_t hist_len;
    int32_t sb_exp;
    int32_t jmin;             // first k-block that holds a non-zero tap (short filters skip the rest)
    int32_t dbg;              // bring-up: bit0 skip MMAs, bit1 skip conversion, bit2 skip stores
    int32_t M;                // DEC kernels: keep every M-th output (multirate_FIR.dn), y[o] = filter output M o
    int64_t n_m;              // DEC kernels: outputs kept
    int32_t L;                // UP kernels: the stream is x zero-stuffed by L and scaled by L (multirate_FIR.up);
    int64_t n_in;             //             n = n_in * L is the stream length, x / hist hold input-rate samples
};

__device__ __forceinline__ float2 load_sample(const Args &a, int64_t g) {
    if (g >= 0) return (g < a.n) ? a.x[g] : make_float2(0.f, 0.f);
    if (a.hist != nullptr) {
        int64_t h = (int64_t)a.hist_len + g;
        if (h >= 0) return a.hist[h];
    }
    return make_float2(0.f, 0.f);
}
// UP: stream position g (may be negative: history) of the zero-stuffed, L-scaled input
__device__ __forceinline__ float2 load_up_sample(const Args &a, int64_t g) {
    int64_t m = g / a.L;
    if (m * a.L != g) {
        if (g >= 0 || (m - 1) * a.L != g) return make_float2(0.f, 0.f);
        --m;                                           // (unreachable: a multiple of L divides exactly)
    }
    float2 v = make_float2(0.f, 0.f);
    if (m >= 0) {
        if (m < a.n_in) v = a.x[m];
    } else if (a.hist != nullptr && (int64_t)a.hist_len + m >= 0) {
        v = a.hist[(int64_t)a.hist_len + m];
    }
    return make_float2(v.x * (float)a.L, v.y * (float)a.L);
}
// UP: the input samples x[m0 .. m1) whose positions m L fall inside the staged range of a tile travel compact: one
// bulk copy of the 16-byte granules [ma, me) that hold them
struct CompactTile { int64_t ma; int32_t bytes; };
template <int TN>
__device__ __forceinline__ bool tile_is_compact(const Args &a, int64_t tile, CompactTile &c) {
    const int64_t g0 = tile * Cfg<TN>::TILE - HALO;
    if (g0 < 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) != 0) return false;
    const int64_t m0 = (g0 + a.L - 1) / a.L, m1 = (g0 + Cfg<TN>::TILE_IN + a.L - 1) / a.L;
    const int64_t ma = m0 & ~(int64_t)1, me = (m1 + 1) & ~(int64_t)1;
    if (me > a.n_in) return false;
    c.ma = ma;
    c.bytes = (int32_t)(me - ma) * 8;
    return true;
}
template <int TN>
__device__ __forceinline__ bool tile_is_bulk(const Args &a, int64_t tile) {
    const int64_t g0 = tile * Cfg<TN>::TILE - HALO;
    return g0 >= 0 && g0 + Cfg<TN>::TILE_IN <= a.n && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
}
// MODE 0 filter | 1 decimating stores (dn) | 2 zero-stuffed input (up).
// ZLO: the x_lo pass runs against a second copy of the tap matrix whose b_lo rows are zero -- the x_lo*b_lo product is
// numerically free to drop (tools/tc2_numerics.py), and multiplying zeros costs the tensor core less energy: the kernel
// is limited by the clock the chip sustains under this load, and this is worth 3.8 % (1.089 -> 1.048 ms).
template <int TN, bool DBG, int MODE = 0, bool ZLO = false>
__global__ void __launch_bounds__(NTHREADS, 1) fir_tc2_kernel(const Args a)
{
    using C = Cfg<TN>;
    constexpr bool DEC = MODE == 1, UP = MODE == 2;
    constexpr int TILE_N = C::TILE_N, TILE = C::TILE, TILE_IN = C::TILE_IN, RAW_BYTES = C::RAW_BYTES;
    constexpr int STREAM_BYTES = C::STREAM_BYTES, STAGE_BYTES = C::STAGE_BYTES, NSTAGE = C::NSTAGE;
    // ZLO: a second copy of the tap matrix (b_lo rows zeroed) sits in TMEM columns 160..319 for the x_lo pass; two
    // accumulators instead of three (320 + 2 * 96 = 512 columns)
    constexpr int SMEM_BAR_OFF = C::SMEM_BAR_OFF, ACC_BUF_COLS = C::ACC_BUF_COLS, NACC = ZLO ? 2 : C::NACC;
    constexpr int ACC_COL0 = ZLO ? 2 * A_COLS : A_COLS;
    static_assert(ACC_COL0 + NACC * ACC_BUF_COLS <= 512, "TMEM columns");
    constexpr int F4_PER_TILE = C::F4_PER_TILE, F4_PER_THREAD = C::F4_PER_THREAD;
    constexpr uint32_t kIdesc = C::kIdesc;
    (void)TILE_IN;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    const uint32_t base = (raw_base + 1023u) & ~1023u;
    unsigned char *sm = smem_dyn + (base - raw_base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + SMEM_BAR_OFF);
    const uint32_t bar0 = base + SMEM_BAR_OFF;
    // raw_full[s]=s, a_full[s]=6+s, a_empty[s]=12+s, d_full[b]=18+b, d_empty[b]=20+b
    auto RAW_FULL = [&](int s) { return bar0 + 8u * s; };
    auto A_FULL = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto D_FULL = [&](int b) { return bar0 + 8u * (3 * NSTAGE + b); };
    auto D_EMPTY = [&](int b) { return bar0 + 8u * (3 * NSTAGE + NACC + b); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * NSTAGE + 2 * NACC);
    float *wmax = reinterpret_cast<float *>(bars + 3 * NSTAGE + 2 * NACC + 1);          // 8 floats
    float *tile_inv = reinterpret_cast<float *>(bars + 3 * NSTAGE + 2 * NACC + 5);      // INV_RING floats

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

    // ---- tap matrix -> TMEM (A operand, stays for the whole kernel) ----
    if (warp < 4) {
        const int row = warp * 32 + lane;                                  // TMEM lane == matrix row
        const uint4 *src = a.amat + (size_t)row * (2 * KTOT * 2 / 16);     // 80 uint4 per row: taps | taps with b_lo rows zero
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < (ZLO ? 2 * A_COLS : A_COLS); c0 += 16) {
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

    const int64_t first = blockIdx.x, step = gridDim.x;

    if (warp == PROD_WARP) {
        // =============================== producer (bulk TMA) ===============================
        int it = 0;
        long long w0 = 0, t_start = DBG ? clock64() : 0;
        const uint64_t ns_start = DBG ? global_ns() : 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            mbar_wait_t<DBG>(A_EMPTY(s), ph ^ 1u, 0, w0);
            if (elect_one()) {
                CompactTile ctile;
                if (UP) {
                    if (tile_is_compact<TN>(a, tile, ctile)) {
                        mbar_arrive_expect_tx(RAW_FULL(s), (uint32_t)ctile.bytes);
                        bulk_g2s(base + s * STAGE_BYTES, a.x + ctile.ma, (uint32_t)ctile.bytes, RAW_FULL(s));
                    } else {
                        mbar_arrive(RAW_FULL(s));
                    }
                } else if (tile_is_bulk<TN>(a, tile)) {
                    mbar_arrive_expect_tx(RAW_FULL(s), RAW_BYTES);
                    bulk_g2s(base + s * STAGE_BYTES, a.x + (tile * TILE - HALO), RAW_BYTES, RAW_FULL(s));
                } else {
                    mbar_arrive(RAW_FULL(s));       // edge tile: converters fetch it with guarded loads
                }
            }
            __syncwarp();
        }
        if (DBG && (a.dbg & 8) && blockIdx.x == 1 && lane == 0)
            printf("tc2 producer: tiles %d total %lld wait_a_empty %lld  (%llu ns: SM clock %.0f MHz)\n", it, clock64() - t_start, w0,
                   (unsigned long long)(global_ns() - ns_start), 1e3 * (double)(clock64() - t_start) / (double)(global_ns() - ns_start));
    } else if (warp >= CVT_WARP0 && warp < CVT_WARP0 + N_CVT_WARPS) {
        // =============================== converters ===============================
        const int ct = tid - CVT_WARP0 * 32;
        int it = 0;
        long long w0 = 0, t_start = DBG ? clock64() : 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            unsigned char *st = sm + s * STAGE_BYTES;
            mbar_wait_t<DBG>(RAW_FULL(s), ph, 1, w0);
            float4 raw4[F4_PER_THREAD];
            const bool bulk = !UP && tile_is_bulk<TN>(a, tile);
            const int64_t g0 = tile * TILE - HALO;
            // UP, compact tile: stream positions g0 + 2 f (+1) hold x[q] L when r == 0, (q, r) = divmod(position, L)
            // advancing by 2 N_CVT positions per iteration
            CompactTile ctile;
            const bool compact = UP && tile_is_compact<TN>(a, tile, ctile);
            int32_t qi[2] = {0, 0}, ri[2] = {0, 0}, dq = 0, drm = 0;
            if (UP && compact) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int64_t pos = g0 + 2 * ct + j, q = pos / a.L;
                    ri[j] = (int32_t)(pos - q * a.L);
                    qi[j] = (int32_t)(q - ctile.ma);
                }
                dq = (2 * N_CVT) / a.L;
                drm = (2 * N_CVT) - dq * a.L;
            }
#pragma unroll
            for (int i = 0; i < F4_PER_THREAD; ++i) {
                const int f = ct + i * N_CVT;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f < F4_PER_TILE) {
                    if (UP) {
                        if (compact) {
                            const float2 *comp = reinterpret_cast<const float2 *>(st);
                            const float gain = (float)a.L;
                            float2 e[2];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const bool hit = ri[j] == 0;
                                const float2 cv = comp[hit ? qi[j] : 0];
                                e[j] = hit ? make_float2(cv.x * gain, cv.y * gain) : make_float2(0.f, 0.f);
                                ri[j] += drm;
                                qi[j] += dq;
                                if (ri[j] >= a.L) { ri[j] -= a.L; ++qi[j]; }
                            }
                            v = make_float4(e[0].x, e[0].y, e[1].x, e[1].y);
                        } else {
                            const float2 s0 = load_up_sample(a, g0 + 2 * f), s1 = load_up_sample(a, g0 + 2 * f + 1);
                            v = make_float4(s0.x, s0.y, s1.x, s1.y);
                        }
                    } else if (bulk) {
                        v = reinterpret_cast<const float4 *>(st)[f];
                    } else {
                        float2 s0 = load_sample(a, g0 + 2 * f), s1 = load_sample(a, g0 + 2 * f + 1);
                        v = make_float4(s0.x, s0.y, s1.x, s1.y);
                    }
                }
                raw4[i] = v;
            }
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < F4_PER_THREAD; ++i)
                m = fmaxf(fmaxf(m, fmaxf(fabsf(raw4[i].x), fabsf(raw4[i].y))), fmaxf(fabsf(raw4[i].z), fabsf(raw4[i].w)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));        // wmax free; (in place) nobody writes yet
            if (lane == 0) wmax[warp - CVT_WARP0] = m;
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));        // every thread holds its raw samples in registers
            float bm = 0.f;
#pragma unroll
            for (int w = 0; w < N_CVT_WARPS; ++w) bm = fmaxf(bm, wmax[w]);
            int ex = 14;
            if (bm > 0.f && bm < 3.0e38f) (void)frexpf(bm, &ex);
            int e = max(-110, min(110, 14 - ex));
            const float sx = ldexpf(1.0f, e);
            if (!DBG || !(a.dbg & 2)) {
                // element e = 2f lives at row f>>5, 16-byte chunk (f&31)>>2: the swizzled offset of this
                // thread's pair advances by exactly 8 rows (1024 B) per iteration
                const uint32_t off0 = stream_off(2 * ct);
#pragma unroll
                for (int i = 0; i < F4_PER_THREAD; ++i) {
                    const int f = ct + i * N_CVT;
                    if (f < F4_PER_TILE) {
                        const float4 v = raw4[i];
                        const float re0 = v.x * sx, im0 = v.y * sx, re1 = v.z * sx, im1 = v.w * sx;
                        const __half2 rh = __floats2half2_rn(re0, re1);
                        const __half2 ih = __floats2half2_rn(im0, im1);
                        const float2 rhf = __half22float2(rh), ihf = __half22float2(ih);
                        // lo = true residual (not rescaled): products land in the same accumulator as hi*hi
                        const __half2 rl = __floats2half2_rn(re0 - rhf.x, re1 - rhf.y);
                        const __half2 il = __floats2half2_rn(im0 - ihf.x, im1 - ihf.y);
                        const uint32_t off = off0 + (uint32_t)i * 1024u;
                        *reinterpret_cast<__half2 *>(st + 0 * STREAM_BYTES + off) = rh;
                        *reinterpret_cast<__half2 *>(st + 1 * STREAM_BYTES + off) = rl;
                        *reinterpret_cast<__half2 *>(st + 2 * STREAM_BYTES + off) = ih;
                        *reinterpret_cast<__half2 *>(st + 3 * STREAM_BYTES + off) = il;
                    }
                }
            }
            if (ct == 0) tile_inv[it % INV_RING] = ldexpf(1.0f, -e - a.sb_exp);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(A_FULL(s));
        }
        if (DBG && (a.dbg & 8) && blockIdx.x == 1 && ct == 0)
            printf("tc2 converter: tiles %d total %lld wait_raw_full %lld\n", it, clock64() - t_start, w0);
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================
        // The whole warp runs the loop (uniform control flow); one elected lane issues, so ptxas emits the
        // UTCHMMA sequence without a per-lane serialisation loop (measured: ~76 -> ~16 cycles per issue).
        int it = 0;
        uint32_t unit = 0;                             // (tile, channel) counter -> accumulator buffer
        long long w0 = 0, w1 = 0, t_start = DBG ? clock64() : 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it % NSTAGE;
            const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
            mbar_wait_t<DBG>(A_FULL(s), ph, 2, w0);
            tc_fence_after();
            const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch, ++unit) {
                const uint32_t b = unit % NACC;
                mbar_wait_t<DBG>(D_EMPTY(b), ((unit / NACC) & 1u) ^ 1u, 3, w1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dd = tmem_base + ACC_COL0 + b * ACC_BUF_COLS;
                    // (M = 64 MMAs for the x_lo pass -- b_hi rows only, dropping the numerically free x_lo*b_lo
                    //  product -- do not fit this layout: an M = 64 accumulator keeps columns N/2.. in lanes +16 of
                    //  each quarter and wants the A rows duplicated there, where the b_lo rows live.  Measured with the
                    //  wrong result: same cycle count, SM clock 1457 instead of 1356 MHz, 8 % faster -- the kernel is
                    //  limited by the clock the chip sustains under this tensor load, not by issue slots.)
                    // The x_lo stream goes FIRST: its products are ~2^-11 of the result, so the accumulator is
                    // still tiny while they are added and the tensor core's truncating fp32 adds cost nothing;
                    // then x_hi with the small (outer) tap blocks before the large centre ones.
                    uint32_t started = 0;
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const int part = 1 - pp;
                        // descriptor of (stream, k-block 0, slice 0); only the 14-bit start field changes
                        const uint64_t bd0 = make_desc_sw128(st + (2 * ch + part) * STREAM_BYTES);
#pragma unroll
                        for (int jj = 0; jj < NKB; ++jj) {
                            constexpr int kOrder[NKB] = {0, 1, 4, 3, 2};
                            const int j = kOrder[jj];
                            if (j < a.jmin) continue;                  // all-zero tap block (filter shorter than 256)
#pragma unroll
                            for (int sl = 0; sl < 4; ++sl) {
                                const uint64_t bd = bd0 + (uint64_t)(8 * j + 2 * sl);     // +128 j + 32 sl bytes (>>4)
                                const uint32_t at = tmem_base + ((ZLO && part == 1) ? A_COLS : 0) + (j * 4 + sl) * 8;
                                if (!DBG || !(a.dbg & 1)) umma_f16_ts(dd, at, bd, kIdesc, started);
                                started = 1;
                            }
                        }
                    }
                    umma_commit(D_FULL(b));
                    if (ch == 1) umma_commit(A_EMPTY(s));
                }
                __syncwarp();
            }
        }
        if (DBG && (a.dbg & 8) && blockIdx.x == 1 && lane == 0)
            printf("tc2 mma: tiles %d total %lld wait_a_full %lld wait_d_empty %lld\n", it, clock64() - t_start, w0, w1);
    } else {
        // =============================== epilogue (warps 0-3 and 14-17) ===============================
        // Lane l < 16 holds row b_hi[c], lane l+16 row b_lo[c] (c = 16*q + l, q = warp % 4 = TMEM lane quarter):
        // y = (D[l] + D[l+16]) * inv.  Hi-row lanes finish even stream rows, lo-row lanes the odd ones, so all
        // 32 lanes store.  The two warp groups take the lower / upper half of the accumulator columns.
        int it = 0;
        uint32_t unit = 0;
        const int q4 = warp & 3, grp = (warp < 4) ? 0 : 1;
        const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
        const bool lo_row = lane >= 16;
        const int c = q4 * 16 + (lane & 15);
        constexpr int NCHUNK = TILE_N / 16;
        const int ch_lo = grp == 0 ? 0 : (NCHUNK + 1) / 2, ch_hi = grp == 0 ? (NCHUNK + 1) / 2 : NCHUNK;
        long long w0 = 0, t_start = DBG ? clock64() : 0;
        constexpr int CPG = (NCHUNK + 1) / 2;               // column chunks per warp group
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it, unit += 2) {
            const uint32_t b_re = unit % NACC, b_im = (unit + 1) % NACC;
            const uint32_t d_re = tmem_base + lane_sel + ACC_COL0 + b_re * ACC_BUF_COLS;
            const uint32_t d_im = tmem_base + lane_sel + ACC_COL0 + b_im * ACC_BUF_COLS;
            const int64_t g_c = tile * TILE + c + (lo_row ? BK : 0);
            // DEC: this lane's outputs g advance by 2 BK from g_c + 16 BK ch_lo; (gq, rm) = divmod(g, M) follows along
            int64_t gq = 0;
            int32_t rm = 0, dq = 0, drm = 0;
            if constexpr (DEC) {
                const int64_t g_first = g_c + (int64_t)ch_lo * 16 * BK;
                gq = g_first / a.M;
                rm = (int32_t)(g_first - gq * a.M);
                dq = (2 * BK) / a.M;
                drm = (2 * BK) - dq * a.M;
            }
            // real-part accumulator: pull this warp's columns into registers and hand the TMEM slot back
            // immediately -- the MMA warp can refill it while the imaginary part is still being computed
            mbar_wait_t<DBG>(D_FULL(b_re), (unit / NACC) & 1u, 4, w0);
            tc_fence_after();
            uint32_t dr[CPG][16];
#pragma unroll
            for (int k = 0; k < CPG; ++k)
                if (ch_lo + k < ch_hi) tmem_ld16(d_re + (ch_lo + k) * 16, dr[k]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(D_EMPTY(b_re));
            mbar_wait_t<DBG>(D_FULL(b_im), ((unit + 1) / NACC) & 1u, 4, w0);
            tc_fence_after();
            const float inv = tile_inv[it % INV_RING];
#pragma unroll
            for (int k = 0; k < CPG; ++k) {
                const int ck = ch_lo + k;
                if (ck < ch_hi) {
                    const int c0 = ck * 16;
                    uint32_t di[16];
                    tmem_ld16(d_im + c0, di);
                    tmem_ld_wait();
                    if (ck + 1 == ch_hi) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(D_EMPTY(b_im));
                    }
#pragma unroll
                    for (int q = 0; q < 16; q += 2) {
                        const float sr = __uint_as_float(lo_row ? dr[k][q] : dr[k][q + 1]);
                        const float si = __uint_as_float(lo_row ? di[q] : di[q + 1]);
                        const float rr = __shfl_xor_sync(0xffffffffu, sr, 16);
                        const float ri = __shfl_xor_sync(0xffffffffu, si, 16);
                        const float vr = (rr + __uint_as_float(lo_row ? dr[k][q + 1] : dr[k][q])) * inv;
                        const float vi = (ri + __uint_as_float(lo_row ? di[q + 1] : di[q])) * inv;
                        if constexpr (DEC) {
                            if (rm == 0 && gq < a.n_m) a.y[gq] = make_float2(vr, vi);
                            rm += drm;
                            gq += dq;
                            if (rm >= a.M) { rm -= a.M; ++gq; }
                        } else {
                            const int64_t g = g_c + (int64_t)(c0 + q) * BK;
                            if ((!DBG || !(a.dbg & 4)) && g < a.n) a.y[g] = make_float2(vr, vi);
                        }
                    }
                }
            }
        }
        if (DBG && (a.dbg & 8) && blockIdx.x == 1 && tid == 0)
            printf("tc2 epilogue: tiles %d total %lld wait_d_full %lld\n", it, clock64() - t_start, w0);
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc2

// ------------------------------------------------------------------------------------------ host
// 128 x 320 fp16 row-major tap matrix for the TMEM A operand.  Row rho = 32 w + l:
//   l < 16 : hi part of Toeplitz row c = 16 w + l;   l >= 16 : lo part of row c = 16 w + l - 16
int tc2_build_tap_matrix(const double *taps, int ntaps, unsigned char *out, int *sb_exp)
{
    using namespace tc2;
    if (ntaps > HALO) return -1;
    double mx = 0.0;
    for (int i = 0; i < ntaps; ++i) mx = fmax(mx, fabs(taps[i]));
    int ex = 0;
    if (mx > 0.0) (void)frexp(mx, &ex);
    int e = 12 - ex;
    if (e > 60) e = 60;
    if (e < -60) e = -60;
    *sb_exp = e;
    __half *m = reinterpret_cast<__half *>(out);
    for (int rho = 0; rho < 128; ++rho) {
        const int w = rho / 32, l = rho % 32;
        const int c = 16 * w + (l & 15);
        const bool lo = l >= 16;
        for (int kk = 0; kk < KTOT; ++kk) {
            int t = c + HALO - kk;
            double v = (t >= 0 && t < ntaps) ? ldexp(taps[t], e) : 0.0;
            __half hi = __float2half_rn((float)v);
            __half lw = __float2half_rn((float)(v - (double)__half2float(hi)));   // true residual
            m[(size_t)rho * 2 * KTOT + kk] = lo ? lw : hi;
            m[(size_t)rho * 2 * KTOT + KTOT + kk] = lo ? __float2half_rn(0.f) : hi;      // x_lo pass: b_hi rows only
        }
    }
    return 0;
}
int tc2_matrix_bytes() { return 128 * 2 * tc2::KTOT * 2; }

template <int TN, bool DBG, int MODE = 0, bool ZLO = false>
static int launch_tc2_cfg(tc2::Args a, int64_t n, int sm_count, cudaStream_t stream)
{
    using namespace tc2;
    a.n_tiles = (n + Cfg<TN>::TILE - 1) / Cfg<TN>::TILE;
    auto kern = fir_tc2_kernel<TN, DBG, MODE, ZLO>;
    B200_CHECK_CUDA(allow_smem(kern, Cfg<TN>::SMEM_TOTAL));
    int64_t grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    kern<<<(unsigned)grid, NTHREADS, Cfg<TN>::SMEM_TOTAL, stream>>>(a);
    B200_CHECK_LAUNCH("fir_tc2_kernel");
    return B200DSP_OK;
}

int launch_fir_tc2(const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                   const void *amat_dev, int sb_exp, int ntaps, int tile_rows, int sm_count, cudaStream_t stream,
                   int32_t M, int32_t L)
{
    using namespace tc2;
    Args a;
    a.M = M;
    a.n_m = n / M;
    a.L = L;
    a.n_in = n;
    if (L > 1) n *= L;                                 // the kernel's stream is the zero-stuffed one
    a.x = static_cast<const float2 *>(x);
    a.hist = static_cast<const float2 *>(hist);
    a.y = static_cast<float2 *>(y);
    a.amat = static_cast<const uint4 *>(amat_dev);
    a.n = n;
    a.n_tiles = 0;
    a.hist_len = hist_len;
    a.sb_exp = sb_exp;
    // T[c][kk] = b[c + 256 - kk] is non-zero only for kk >= 257 - ntaps: skip the leading all-zero k-blocks
    a.jmin = (HALO + 1 - ntaps) / BK;
    if (a.jmin < 0) a.jmin = 0;
    a.dbg = 0;
    if (const char *e = getenv("B200DSP_TC_DBG")) a.dbg = atoi(e);
    // dn(M): the filter runs at the full rate (the tensor work per input sample is what it is) and the epilogue
    // stores every M-th output; only the samples up to the last kept output are processed
    if (M > 1) return launch_tc2_cfg<96, false, 1, true>(a, (a.n_m - 1) * M + 1, sm_count, stream);
    if (L > 1) return launch_tc2_cfg<96, false, 2, true>(a, n, sm_count, stream);
    if (tile_rows == 97) return launch_tc2_cfg<96, false, 0, true>(a, n, sm_count, stream);
    if (a.dbg) {
        if (tile_rows == 128) return launch_tc2_cfg<128, true>(a, n, sm_count, stream);
        if (tile_rows == 96) return launch_tc2_cfg<96, true>(a, n, sm_count, stream);
        if (tile_rows == 80) return launch_tc2_cfg<80, true>(a, n, sm_count, stream);
        return launch_tc2_cfg<64, true>(a, n, sm_count, stream);
    }
    if (tile_rows == 128) return launch_tc2_cfg<128, false>(a, n, sm_count, stream);
    if (tile_rows == 96) return launch_tc2_cfg<96, false>(a, n, sm_count, stream);
    if (tile_rows == 80) return launch_tc2_cfg<80, false>(a, n, sm_count, stream);
    return launch_tc2_cfg<64, false>(a, n, sm_count, stream);
}

}  // namespace b200dsp
