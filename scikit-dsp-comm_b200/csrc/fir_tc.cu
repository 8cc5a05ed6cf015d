// fir_tc.cu -- complex64 FIR as a block-Toeplitz GEMM on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Why: a 256-tap direct form needs 512 real FMA per complex sample; the CUDA cores top out at
// ~18 % of the HBM roofline.  The contraction  y[64r+c] = sum_t b[t] x[64r+c-t]
// is a genuine dense GEMM once outputs are grouped in rows of 64:
//
//     D[r, c] = sum_{kk<320} A[r, kk] * T[kk, c],   A[r, kk] = x[t0 - 256 + 64 r + kk],
//                                                   T[kk, c] = b[c + 256 - kk]  (80 % dense)
//
// The Hankel operand A is NEVER materialised.  Row r+1 of A is row r shifted by 64 samples, i.e. by
// exactly one 128-byte swizzle row of fp16.  The sample stream is therefore written ONCE into shared
// memory in the canonical K-major SWIZZLE_128B layout (row pitch 128 B), and the UMMA shared-memory
// descriptor of k-block j simply starts 128*j bytes later: 132 stream rows serve all five k-blocks
// of a 128-row tile.
//
// Precision (fp32 in / fp32 out, target |err| <= 1e-6 * max|y|): every fp32 sample and tap is split
// into two fp16 values, v*s = hi + lo*2^-11 (s = power-of-two block scale, per tile for the samples,
// per plan for the taps), and three MMAs are issued: D1 += hi*hi, D2 += hi*lo + lo*hi (the lo*lo term is
// 2^-22 relative and dropped).  y = (D1 + 2^-11 D2) / (s_x s_b), accumulated in fp32 in TMEM.
//
// Warp roles (one persistent CTA per SM, 13 warps):
//   warps 0-3   epilogue: tcgen05.ld TMEM -> registers -> scaled, interleaved complex64 -> global
//   warp  4     MMA issuer (one elected lane) + TMEM allocator
//   warps 5-12  converters: global fp32 (re,im) -> block max -> fp16 hi/lo, de-interleaved, swizzled
// Pipelines: 2 shared-memory stages (converter <-> MMA), 2 TMEM accumulator stages (MMA <-> epilogue),
// all hand-offs through mbarriers; tcgen05.commit releases stages.
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace b200dsp {
namespace tc {

constexpr int BK = 64;                      // fp16 per 128-byte swizzle row == outputs per GEMM row
constexpr int NKB = 5;                      // k-blocks: (256 taps + 64) / 64
constexpr int TILE_M = 128;                 // GEMM rows (TMEM lanes) per tile
constexpr int TILE = TILE_M * BK;           // 8192 complex outputs per tile
constexpr int HALO = (NKB - 1) * BK;        // 256 samples before the tile
constexpr int ROWS = TILE_M + NKB - 1;      // 132 stream rows per tile
constexpr int TILE_IN = ROWS * BK;          // 8448 complex samples staged per tile
constexpr int STREAM_BYTES = 17 * 1024;     // 132*128 = 16896 -> padded to a 1 KB multiple
constexpr int STAGE_BYTES = 4 * STREAM_BYTES;   // re_hi, re_lo, im_hi, im_lo
constexpr int B_KB_BYTES = BK * 128;        // one k-block of the 64 x 320 tap matrix: 8 KB
constexpr int B_BYTES = NKB * B_KB_BYTES;   // 40 KB per matrix (hi, lo)
constexpr int SMEM_B_OFF = 0;
constexpr int SMEM_A_OFF = 2 * B_BYTES;                     // 80 KB
constexpr int SMEM_BAR_OFF = SMEM_A_OFF + 2 * STAGE_BYTES;  // + 136 KB
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 256 + 1024;       // barriers/scratch + alignment slack

constexpr int N_EPI_WARPS = 4;
constexpr int MMA_WARP = 4;
constexpr int N_CVT_WARPS = 8;
constexpr int CVT_THREAD0 = (MMA_WARP + 1) * 32;            // 160
constexpr int N_CVT = N_CVT_WARPS * 32;                     // 256
constexpr int NTHREADS = CVT_THREAD0 + N_CVT;               // 416
constexpr int F4_PER_TILE = TILE_IN / 2;                    // 4224 float4 (2 complex each)
constexpr int F4_PER_THREAD = (F4_PER_TILE + N_CVT - 1) / N_CVT;   // 17

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must end in a trap, never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    printf("b200dsp fir_tc: mbarrier timeout (role %d, block %d)\n", who, (int)blockIdx.x);
    __trap();
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every previously issued MMA of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 (1024 B: next 8 rows)
//   [46,48) version = 1 | [49,52) base offset | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=b=F16 (0), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BK >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

struct Args {
    const float2 *x;
    const float2 *hist;
    float2 *y;
    const uint4 *bmat;        // device: [B_hi | B_lo], swizzled fp16, 2*B_BYTES
    int64_t n;
    int64_t n_tiles;
    int32_t hist_len;
    int32_t sb_exp;           // taps were scaled by 2^sb_exp before the split
    int32_t bo_mode;          // 1: descriptor base_offset = (start>>7)&7, 0: always 0 (A/B test)
    uint32_t kb_order;        // k-block issue order, 4 bits per step (accumulate small tap blocks first)
    int32_t dbg;              // bring-up only: bit0 skip MMAs, bit1 skip conversion, bit2 skip stores
};

__device__ __forceinline__ float2 load_sample(const Args &a, int64_t g) {
    if (g >= 0) return (g < a.n) ? a.x[g] : make_float2(0.f, 0.f);
    if (a.hist != nullptr) {
        int64_t h = (int64_t)a.hist_len + g;
        if (h >= 0) return a.hist[h];
    }
    return make_float2(0.f, 0.f);
}

// byte offset of fp16 element e (0..TILE_IN) inside one swizzled stream
__device__ __forceinline__ uint32_t stream_off(int e) {
    const int row = e >> 6, col = e & 63;
    return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1));
}

__device__ __forceinline__ void split_h(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * 2048.0f);
}

__global__ void __launch_bounds__(NTHREADS, 1) fir_tc_kernel(const Args a)
{
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned base (SWIZZLE_128B atoms repeat every 1024 B)
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_dyn + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // barriers + scratch
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + SMEM_BAR_OFF);
    const uint32_t bar0 = base + SMEM_BAR_OFF;
    // indices: a_full[2]=0,1  a_empty[2]=2,3  d_full[2]=4,5  d_empty[2]=6,7
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    float *wmax = reinterpret_cast<float *>(bars + 9);           // 8 floats
    float *tile_inv = reinterpret_cast<float *>(bars + 13);      // 4 slots (tile it -> slot it&3), converter -> epilogue

    // ---- one-time setup ----
    if (tid == 0) {
        mbar_init(BAR(0), N_CVT_WARPS); mbar_init(BAR(1), N_CVT_WARPS);
        mbar_init(BAR(2), 1); mbar_init(BAR(3), 1);
        mbar_init(BAR(4), 1); mbar_init(BAR(5), 1);
        mbar_init(BAR(6), N_EPI_WARPS); mbar_init(BAR(7), N_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 512);
    // tap matrices -> shared memory (generic proxy), visible to the async proxy after the fence
    {
        uint4 *dst = reinterpret_cast<uint4 *>(sm + SMEM_B_OFF);
        for (int i = tid; i < 2 * B_BYTES / 16; i += NTHREADS) dst[i] = a.bmat[i];
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = blockIdx.x, step = gridDim.x;

    if (warp >= MMA_WARP + 1) {
        // =============================== converters ===============================
        const int ct = tid - CVT_THREAD0;
        float4 raw4[F4_PER_THREAD];
        auto load_tile = [&](int64_t tile) {
            const int64_t g0 = tile * TILE - HALO;
            const bool interior = (g0 >= 0) && (g0 + TILE_IN <= a.n) &&
                                  ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
#pragma unroll
            for (int i = 0; i < F4_PER_THREAD; ++i) {
                const int f = ct + i * N_CVT;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f < F4_PER_TILE) {
                    if (interior) {
                        v = __ldg(reinterpret_cast<const float4 *>(a.x + g0) + f);
                    } else {
                        float2 s0 = load_sample(a, g0 + 2 * f), s1 = load_sample(a, g0 + 2 * f + 1);
                        v = make_float4(s0.x, s0.y, s1.x, s1.y);
                    }
                }
                raw4[i] = v;
            }
        };
        if (first < a.n_tiles) load_tile(first);
        int it = 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it & 1;
            // ---- block max -> power-of-two scale putting max|x| in [2^13, 2^14) ----
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < F4_PER_THREAD; ++i)
                m = fmaxf(fmaxf(m, fmaxf(fabsf(raw4[i].x), fabsf(raw4[i].y))), fmaxf(fabsf(raw4[i].z), fabsf(raw4[i].w)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));        // previous tile's wmax readers are done
            if (lane == 0) wmax[warp - (MMA_WARP + 1)] = m;
            asm volatile("bar.sync 1, %0;" ::"n"(N_CVT));
            float bm = 0.f;
#pragma unroll
            for (int w = 0; w < N_CVT_WARPS; ++w) bm = fmaxf(bm, wmax[w]);
            int ex = 0;
            if (bm > 0.f && bm < 3.0e38f) (void)frexpf(bm, &ex); else ex = 14;
            int e = 14 - ex;
            e = max(-110, min(110, e));
            const float sx = ldexpf(1.0f, e);
            // ---- wait until the MMAs that read this stage two tiles ago are done ----
            mbar_wait(BAR(2 + s), ((it >> 1) & 1) ^ 1, 1);
            unsigned char *st = sm + SMEM_A_OFF + s * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < F4_PER_THREAD; ++i) {
                const int f = ct + i * N_CVT;
                if (f < F4_PER_TILE && !(a.dbg & 2)) {
                    const float4 v = raw4[i];
                    __half rh0, rl0, ih0, il0, rh1, rl1, ih1, il1;
                    split_h(v.x * sx, rh0, rl0);
                    split_h(v.y * sx, ih0, il0);
                    split_h(v.z * sx, rh1, rl1);
                    split_h(v.w * sx, ih1, il1);
                    const uint32_t off = stream_off(2 * f);
                    *reinterpret_cast<__half2 *>(st + 0 * STREAM_BYTES + off) = __halves2half2(rh0, rh1);
                    *reinterpret_cast<__half2 *>(st + 1 * STREAM_BYTES + off) = __halves2half2(rl0, rl1);
                    *reinterpret_cast<__half2 *>(st + 2 * STREAM_BYTES + off) = __halves2half2(ih0, ih1);
                    *reinterpret_cast<__half2 *>(st + 3 * STREAM_BYTES + off) = __halves2half2(il0, il1);
                }
            }
            if (ct == 0) tile_inv[it & 3] = ldexpf(1.0f, -e - a.sb_exp);
            fence_proxy_async();                   // generic-proxy writes -> visible to UMMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(0 + s));
            // prefetch the next tile's raw samples; the latency hides behind this tile's MMAs
            if (tile + step < a.n_tiles) load_tile(tile + step);
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================
        // Uniform control flow for the whole warp; one elected lane issues (no per-lane UTCHMMA loop).
        int it = 0;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it & 1;
            mbar_wait(BAR(0 + s), (it >> 1) & 1, 2);              // operands staged
            mbar_wait(BAR(6 + s), ((it >> 1) & 1) ^ 1, 3);        // accumulators drained
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_base = base + SMEM_A_OFF + s * STAGE_BYTES;
                const uint32_t b_hi = base + SMEM_B_OFF, b_lo = b_hi + B_BYTES;
                const uint32_t d_stage = tmem_base + (uint32_t)s * 256u;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const uint32_t a_hi = a_base + (2 * ch) * STREAM_BYTES, a_lo = a_hi + STREAM_BYTES;
                    const uint32_t d1 = d_stage + ch * 128, d2 = d1 + 64;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t ad0 = make_desc((pass == 2) ? a_lo : a_hi, 0);
                        const uint64_t bd0 = make_desc((pass == 1) ? b_lo : b_hi, 0);
                        const uint32_t dd = (pass == 0) ? d1 : d2;
#pragma unroll
                        for (int jj = 0; jj < NKB; ++jj) {
                            constexpr int kOrder[NKB] = {0, 1, 4, 3, 2};
                            const int j = kOrder[jj];
#pragma unroll
                            for (int sl = 0; sl < 4; ++sl) {
                                const uint64_t ad = ad0 + (uint64_t)(8 * j + 2 * sl);
                                const uint64_t bd = bd0 + (uint64_t)((B_KB_BYTES >> 4) * j + 2 * sl);
                                const uint32_t acc = (pass == 2) ? 1u : ((jj | sl) ? 1u : 0u);
                                if (!(a.dbg & 1)) umma_f16(dd, ad, bd, kIdesc, acc);
                            }
                        }
                    }
                }
                umma_commit(BAR(2 + s));        // stage s may be overwritten
                umma_commit(BAR(4 + s));        // accumulators of this tile are complete
            }
            __syncwarp();
        }
    } else {
        // =============================== epilogue ===============================
        int it = 0;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int64_t tile = first; tile < a.n_tiles; tile += step, ++it) {
            const int s = it & 1;
            mbar_wait(BAR(4 + s), (it >> 1) & 1, 4);
            tc_fence_after();
            const float inv = tile_inv[it & 3];
            const float inv_lo = inv * (1.0f / 2048.0f);
            const uint32_t d_stage = tmem_base + (uint32_t)s * 256u + lane_base;
            const int64_t g_row = tile * TILE + (int64_t)(warp * 32 + lane) * BK;
#pragma unroll 1
            for (int c0 = 0; c0 < BK; c0 += 16) {
                uint32_t r1[16], r2[16], i1[16], i2[16];
                tmem_ld16(d_stage + 0 + c0, r1);
                tmem_ld16(d_stage + 64 + c0, r2);
                tmem_ld16(d_stage + 128 + c0, i1);
                tmem_ld16(d_stage + 192 + c0, i2);
                tmem_ld_wait();
                float2 out[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    out[q].x = fmaf(__uint_as_float(r2[q]), inv_lo, __uint_as_float(r1[q]) * inv);
                    out[q].y = fmaf(__uint_as_float(i2[q]), inv_lo, __uint_as_float(i1[q]) * inv);
                }
                const int64_t g = g_row + c0;
                if (a.dbg & 4) continue;
                if (g + 16 <= a.n && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0)) {
                    float4 *dst = reinterpret_cast<float4 *>(a.y + g);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        dst[q] = make_float4(out[2 * q].x, out[2 * q].y, out[2 * q + 1].x, out[2 * q + 1].y);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (g + q < a.n) a.y[g + q] = out[q];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(6 + s));
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------ host
// Build the two 64 x 320 Toeplitz tap matrices (fp16 hi / lo) in the K-major SWIZZLE_128B layout the
// kernel copies verbatim into shared memory.  Element (n = c, kk) = b[c + 256 - kk] * 2^sb_exp.
int tc_build_tap_matrices(const double *taps, int ntaps, unsigned char *out /* 2*B_BYTES */, int *sb_exp)
{
    using namespace tc;
    if (ntaps > (NKB - 1) * BK) return -1;
    double mx = 0.0;
    for (int i = 0; i < ntaps; ++i) mx = fmax(mx, fabs(taps[i]));
    int ex = 0;
    if (mx > 0.0) (void)frexp(mx, &ex);
    int e = 12 - ex;                         // max|b| * 2^e in [2^11, 2^12)
    if (e > 60) e = 60;
    if (e < -60) e = -60;
    *sb_exp = e;
    memset(out, 0, 2 * B_BYTES);
    for (int c = 0; c < BK; ++c)
        for (int kk = 0; kk < NKB * BK; ++kk) {
            int t = c + (NKB - 1) * BK - kk;
            double v = (t >= 0 && t < ntaps) ? ldexp(taps[t], e) : 0.0;
            __half hi = __float2half_rn((float)v);
            double rem = (v - (double)__half2float(hi)) * 2048.0;
            __half lo = __float2half_rn((float)rem);
            int j = kk / BK, kq = kk % BK;
            size_t off = (size_t)j * B_KB_BYTES + (size_t)c * 128 + (size_t)((((kq >> 3) ^ (c & 7))) << 4) + (size_t)((kq & 7) << 1);
            memcpy(out + off, &hi, 2);
            memcpy(out + B_BYTES + off, &lo, 2);
        }
    return 0;
}

int tc_matrix_bytes() { return 2 * tc::B_BYTES; }
int tc_max_taps() { return (tc::NKB - 1) * tc::BK; }

int launch_fir_tc(const void *x, const void *hist, void *y, int64_t n, int32_t hist_len,
                  const void *bmat_dev, int sb_exp, int bo_mode, int sm_count, cudaStream_t stream)
{
    using namespace tc;
    Args a;
    a.x = static_cast<const float2 *>(x);
    a.hist = static_cast<const float2 *>(hist);
    a.y = static_cast<float2 *>(y);
    a.bmat = static_cast<const uint4 *>(bmat_dev);
    a.n = n;
    a.n_tiles = (n + TILE - 1) / TILE;
    a.hist_len = hist_len;
    a.sb_exp = sb_exp;
    a.bo_mode = bo_mode;
    a.kb_order = 0x23410;      // k-blocks 0,1,4,3,2: the large centre taps are accumulated last
    a.dbg = 0;
    if (const char *e = getenv("B200DSP_TC_ORDER")) a.kb_order = (uint32_t)strtoul(e, nullptr, 16);
    if (const char *e = getenv("B200DSP_TC_DBG")) a.dbg = atoi(e);
    B200_CHECK_CUDA(allow_smem(fir_tc_kernel, SMEM_TOTAL));
    int64_t grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    fir_tc_kernel<<<(unsigned)grid, NTHREADS, SMEM_TOTAL, stream>>>(a);
    B200_CHECK_LAUNCH("fir_tc_kernel");
    return B200DSP_OK;
}

}  // namespace b200dsp
