// Host build of csrc/seqik_core.cuh for CPU-side algorithm tests (tests/ only).
// TEST INFRASTRUCTURE: compiled by tests/hostsim_build.py with g++, loaded via ctypes by
// tests; the product package never loads it.
#include "seqik_core.cuh"
#include <cstdint>

using namespace seqik;

template <typename R>
static void run_chain(const R* pose, int64_t n_frame, const R* seg, const R* lb, const R* ub,
                      const R* null_sq, const R* seed, R* angles, R* fk, int32_t* nfev, int32_t* status,
                      int stage_mask) {
    ChainParams<R> P;
    for (int i = 0; i < 4; ++i) { P.seg[i] = seg[i]; P.null_sq[i] = null_sq[i]; }
    for (int i = 0; i < 7; ++i) { P.lb[i] = lb[i]; P.ub[i] = ub[i]; }
    R ang[7];
    for (int i = 0; i < 7; ++i) ang[i] = seed[i];
    for (int64_t t = 0; t < n_frame; ++t) {
        FrameStats fs;
        solve_frame<R>(P, pose + t * 15, ang, fk ? fk + t * 27 : nullptr, &fs, stage_mask);
        for (int i = 0; i < 7; ++i) angles[t * 7 + i] = ang[i];
        if (nfev) for (int s = 0; s < 4; ++s) { nfev[t * 4 + s] = fs.nfev[s]; status[t * 4 + s] = fs.status[s]; }
    }
}

extern "C" {
void hostsim_chain_f32(const float* pose, int64_t n_frame, const float* seg, const float* lb, const float* ub,
                       const float* null_sq, const float* seed, float* angles, float* fk, int32_t* nfev,
                       int32_t* status, int stage_mask) {
    run_chain<float>(pose, n_frame, seg, lb, ub, null_sq, seed, angles, fk, nfev, status, stage_mask);
}
void hostsim_chain_f64(const double* pose, int64_t n_frame, const double* seg, const double* lb, const double* ub,
                       const double* null_sq, const double* seed, double* angles, double* fk, int32_t* nfev,
                       int32_t* status, int stage_mask) {
    run_chain<double>(pose, n_frame, seg, lb, ub, null_sq, seed, angles, fk, nfev, status, stage_mask);
}
}
