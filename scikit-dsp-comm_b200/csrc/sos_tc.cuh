// sos_tc.cuh -- interface between the SOS plan (sos_scan.cu) and the tensor-core cascade kernel (sos_tc.cu).
#pragma once
#include "common.cuh"

namespace b200dsp {

constexpr int STC_MAXSEC = 8;            // sections per pass (16 states)
constexpr int STC_TILE = 8192;           // samples per tile

// Device-resident tables of one group of <= 8 sections (built once in b200dsp_sos_plan_create).
struct StcTables {
    bool ok = false;                     // false: cascade does not decay fast enough (or build failed) -> scan kernels
    int nd = 0;                          // states, padded to a multiple of 4 (4, 8, 12, 16)
    int d_real = 0;                      // 2 * real sections
    int e_t = 0;                         // power-of-two scale of the impulse-response matrices
    int warm_tiles = 0;                  // tiles a block runs ahead of its first output (state warm-up)
    int nlev = 0;                        // scan levels whose matrix A^(256 * 2^j) is above float64 resolution
    int at_zero = 0;                     // A^8192 negligible
    int xexp = 12;                       // block scale: window max of |x| -> [2^(xexp-1), 2^xexp)
    float sfac[16];                      // correction operand scale per state (x 2^e_x at run time)
    float ginv[16];                      // 2^-f_d: |s_d| ginv[d] <= max |x| over the cascade's memory
    double coef[STC_MAXSEC][5];          // b0 b1 b2 -a1 -a2 (final-state tail recurrence)
    void *amat = nullptr;                // [128 lanes][416 fp16]  (T_hi | T_lo | Ka | Kb), TMEM image
    double *smat = nullptr;              // [6][nd*nd]    A^128, A^256, ... A^4096, column-major
    float *rowinv = nullptr;             // [16]          2^-e_d of the carry rows
    float *fold = nullptr;               // [13][260]     float32 A^(1024 h) | A^(256 m) | A^128, column-major
};

// coef: nsec x {b0 b1 b2 -a1 -a2}.  Allocates the device tables (cudaMalloc + synchronous copies).
int stc_build(const double (*coef)[5], int nsec, int nsec_real, StcTables *t);
void stc_free(StcTables *t);

// true if the call can take the tensor-core kernel (long enough for the warm-up scheme)
bool stc_usable(const StcTables &t, int64_t n_rate, int sm_count, bool force);

int launch_sos_tc(const StcTables &t, const float *x, float *y, int64_t n_in, int64_t n_rate, int64_t n_out,
                  int32_t L, int32_t M, const float *zi, float *zf, int sm_count, cudaStream_t stream);

}  // namespace b200dsp
