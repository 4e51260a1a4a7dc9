// seqik_generic.cu -- the generic (single-target, 7-DOF) leg-IK kernel (sm_100a) behind seqik_leg_solve_generic_f32.
//
// Replaces LegInvKinGeneric.run_ik_and_fk / calculate_ik_stage (seqikpy/leg_inverse_kinematics.py:473-613) on the chain of
// KinematicChainGeneric (seqikpy/kinematic_chain.py:424-532).  A chain = one leg of one trial; its frames are solved
// serially (the warm start of leg_inverse_kinematics.py:524), chains are independent.
//
// Mapping: ONE LANE PER CHAIN.  A generic solve is a 7-variable problem whose trust-region subproblem is eleven
// symmetric 3x3 solves per evaluation (seqik_generic.cuh); its vectors are short (7) and its many reductions would cost
// more as warp shuffles than they do as seven dependent FMAs, so a chain is not split across lanes.  All lanes run
// GenericSolve::trip() -- one function evaluation -- in a warp-convergent loop and sit at different (frame, iteration)
// positions of their own chains, so a slow solve (the reference needs 25 - 330 evaluations per frame) delays only its
// own chain.  Chains are first spread over warps (one warp per SM sub-partition while that is possible), then warps are
// filled: a warp instruction costs the same for 1 or 32 active lanes.  No tensor cores: scalar 3x3 / 7-vector algebra.
//
// Two instantiations: float32 (the throughput path) and float64 (B200's FP64 pipe runs at half the FP32 rate, and a lane-
// per-chain solve is latency-bound anyway).  The reference's generic solve crawls along zig-zagging 1-D minimisations whose
// candidate choice amplifies rounding; in float64 the device agrees with scipy solve by solve as often as scipy agrees
// with itself, in float32 a few percent less often.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seqik.h"
#include "seqik_common.h"
#include "seqik_generic.cuh"
#include "seqik_generic_group.cuh"

using namespace seqik;

template <typename R>
struct GenArgs {
    const R* pose; int64_t pose_cs, pose_fs; int target_row;
    const R* params;
    R* angles; int64_t ang_cs, ang_fs;
    R* fk; int64_t fk_cs, fk_fs;
    const R* warm; int64_t warm_cs;
    int32_t* status; uint32_t* nfev;
    int64_t n_chain, n_frame;
};

template <typename R>
__global__ void __launch_bounds__(32) leg_solve_generic_kernel(GenArgs<R> a, int cpw) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const int64_t c = (int64_t)blockIdx.x * cpw + lane;
    const bool owner = lane < cpw && c < a.n_chain;
    const int64_t cc = owner ? c : 0;
    const R* prm = a.params + cc * SEQIK_CHAIN_PARAM_FLOATS;
    const R* pk = a.pose + cc * a.pose_cs;                        // key points of the frame to open next
    R* pa = a.angles + cc * a.ang_cs;
    R* pf = a.fk ? a.fk + cc * a.fk_cs : nullptr;
    const int n_frame = (int)a.n_frame;
    const int trow = 3 * a.target_row;
    const R null_sq = __ldg(prm + 25);

    GenericSolve<R> S;
    {
        const R* seed = a.warm ? a.warm + cc * a.warm_cs : prm + 18;
#pragma unroll
        for (int i = 0; i < GEN_DOF; ++i) S.x[i] = seed[i];
    }
    S.status = ST_GTOL;
    int t = owner ? 0 : n_frame;
    bool solving = false;
    uint32_t nf = 0; int worst = ST_GTOL;
    Vec3<R> o = {R(0), R(0), R(0)};
    // key points of the next frame, loaded one frame ahead (the loads are in flight during the current solve)
    Vec3<R> ko = o, kt = o;
    if (t < n_frame) { ko = {__ldg(pk), __ldg(pk + 1), __ldg(pk + 2)}; kt = {__ldg(pk + trow), __ldg(pk + trow + 1), __ldg(pk + trow + 2)}; pk += a.pose_fs; }

    while (__any_sync(full, t < n_frame)) {
        if (t < n_frame && !solving) {                                   // open frame t
            o = ko;
            S.start(prm, Vec3<R>{kt.x - ko.x, kt.y - ko.y, kt.z - ko.z}, null_sq);
            if (t + 1 < n_frame) { ko = {__ldg(pk), __ldg(pk + 1), __ldg(pk + 2)}; kt = {__ldg(pk + trow), __ldg(pk + trow + 1), __ldg(pk + trow + 2)}; pk += a.pose_fs; }
            solving = true;
        }
        if (solving) S.trip();                                           // one function evaluation
        if (solving && S.done()) {                                       // close frame t: angles + the 9 joint rows
            nf += (uint32_t)S.nfev;
            if (S.status == ST_MAXFEV && worst > ST_MAXFEV) worst = ST_MAXFEV;
            if (S.status == ST_NONFINITE) worst = ST_NONFINITE;
#pragma unroll
            for (int i = 0; i < GEN_DOF; ++i) pa[i] = S.x[i];
            pa += a.ang_fs;
            if (pf) {
                Vec3<R> org[3], claw;
                S.joints(org, &claw);
#pragma unroll
                for (int r = 0; r < 4; ++r) { pf[3 * r] = o.x; pf[3 * r + 1] = o.y; pf[3 * r + 2] = o.z; }
                const Vec3<R> rows[5] = {org[0], org[0], org[1], org[2], claw};
#pragma unroll
                for (int r = 0; r < 5; ++r) { pf[12 + 3 * r] = rows[r].x + o.x; pf[13 + 3 * r] = rows[r].y + o.y; pf[14 + 3 * r] = rows[r].z + o.z; }
                pf += a.fk_fs;
            }
            solving = false; ++t;
        }
    }
    if (owner) {
        if (a.status) a.status[c] = worst == ST_NONFINITE ? -1 : worst;
        if (a.nfev) a.nfev[c] = nf;
    }
}

// Second mapping (round 2; the default while a batch fits the GPU at once): a chain spread over eight lanes
// (seqik_generic_group.cuh), up to four chains per warp.
// Warp-convergent loop as above: every group runs one GroupSolve::trip() per iteration; the branches inside a trip are uniform
// within a group, and groups reconverge after them.
template <typename R>
__global__ void __launch_bounds__(32) leg_solve_generic_group_kernel(GenArgs<R> a, int gpw) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x, grp = lane >> 3;
    const int64_t c = (int64_t)blockIdx.x * gpw + grp;
    const bool owner = grp < gpw && c < a.n_chain;
    const int64_t cc = owner ? c : 0;
    const R* prm = a.params + cc * SEQIK_CHAIN_PARAM_FLOATS;
    const R* pk = a.pose + cc * a.pose_cs;
    R* pa = a.angles + cc * a.ang_cs;
    R* pf = a.fk ? a.fk + cc * a.fk_cs : nullptr;
    const int n_frame = (int)a.n_frame;
    const int trow = 3 * a.target_row;

    GroupSolve<R> S;
    S.init(prm, lane);
    {
        const R* seed = a.warm ? a.warm + cc * a.warm_cs : prm + 18;
        if (S.real) S.x = seed[S.gl];
    }
    int t = owner ? 0 : n_frame;
    bool solving = false;
    uint32_t nf = 0; int worst = ST_GTOL;
    Vec3<R> o = {R(0), R(0), R(0)};
    Vec3<R> ko = o, kt = o;                                            // key points of the next frame, one frame ahead
    if (t < n_frame) { ko = {__ldg(pk), __ldg(pk + 1), __ldg(pk + 2)}; kt = {__ldg(pk + trow), __ldg(pk + trow + 1), __ldg(pk + trow + 2)}; pk += a.pose_fs; }

    while (__any_sync(full, t < n_frame)) {
        if (t < n_frame && !solving) {                                   // open frame t
            o = ko;
            S.start(Vec3<R>{kt.x - ko.x, kt.y - ko.y, kt.z - ko.z});
            if (t + 1 < n_frame) { ko = {__ldg(pk), __ldg(pk + 1), __ldg(pk + 2)}; kt = {__ldg(pk + trow), __ldg(pk + trow + 1), __ldg(pk + trow + 2)}; pk += a.pose_fs; }
            solving = true;
        }
        if (solving) S.trip();                                           // one function evaluation of this group's chain
        if (solving && S.done()) {                                       // close frame t: angles + the 9 joint rows
            nf += (uint32_t)S.nfev;
            if (S.status == ST_MAXFEV && worst > ST_MAXFEV) worst = ST_MAXFEV;
            if (S.status == ST_NONFINITE) worst = ST_NONFINITE;
            if (S.real) pa[S.gl] = S.x;
            pa += a.ang_fs;
            if (pf) {
                Vec3<R> org[3], claw;
                S.template chain<false>(S.x, org, &claw, (Vec3<R>*)nullptr);
                // lane r of the group writes row r (rows 0-3 origin, 4-5 CTr, 6 FTi, 7 TiTa), lane 0 also row 8 (claw)
                Vec3<R> row = {R(0), R(0), R(0)};
                if (S.gl == 4 || S.gl == 5) row = org[0];
                if (S.gl == 6) row = org[1];
                if (S.gl == 7) row = org[2];
                pf[3 * S.gl] = row.x + o.x; pf[3 * S.gl + 1] = row.y + o.y; pf[3 * S.gl + 2] = row.z + o.z;
                if (S.gl == 0) { pf[24] = claw.x + o.x; pf[25] = claw.y + o.y; pf[26] = claw.z + o.z; }
                pf += a.fk_fs;
            }
            solving = false; ++t;
        }
    }
    if (owner && S.gl == 0) {
        if (a.status) a.status[c] = worst == ST_NONFINITE ? -1 : worst;
        if (a.nfev) a.nfev[c] = nf;
    }
}

template <typename R>
static int launch_generic(const char* me, const R* pose, int64_t pose_chain_stride, int64_t pose_frame_stride, int32_t target_row,
                          const R* params, R* angles, int64_t ang_chain_stride, int64_t ang_frame_stride,
                          R* fk, int64_t fk_chain_stride, int64_t fk_frame_stride, const R* warm, int64_t warm_chain_stride,
                          int32_t* status, uint32_t* nfev, int64_t n_chain, int64_t n_frame, uint32_t flags, void* stream) {
    if (n_chain < 0 || n_frame < 0) return seqik_fail(SEQIK_EINVAL, "%s: negative size", me);
    if (n_chain == 0 || n_frame == 0) return SEQIK_OK;
    if (!pose || !params || !angles) return seqik_fail(SEQIK_EINVAL, "%s: pose, params and angles must not be NULL", me);
    if (target_row < 1) return seqik_fail(SEQIK_EINVAL, "%s: target_row must be >= 1 (row 0 is the Thorax-Coxa origin)", me);
    if (pose_frame_stride < 3 * ((int64_t)target_row + 1) || ang_frame_stride < 7 || (fk && fk_frame_stride < 27))
        return seqik_fail(SEQIK_EINVAL, "%s: frame stride smaller than the innermost block", me);
    if (n_frame > 2147483647LL) return seqik_fail(SEQIK_EINVAL, "%s: too many frames", me);
    if (flags & ~((0x3Fu << SEQIK_FLAG_CPW_SHIFT) | SEQIK_FLAG_SCHED_MASK)) return seqik_fail(SEQIK_EINVAL, "%s: unknown flag bits", me);
    const uint32_t sched = (flags & SEQIK_FLAG_SCHED_MASK) >> SEQIK_FLAG_SCHED_SHIFT;      // 0 automatic (by batch size), 1 a lane per chain, 2 eight lanes per chain
    if (sched > 2) return seqik_fail(SEQIK_EINVAL, "%s: unknown schedule", me);
    GenArgs<R> a;
    a.pose = pose; a.pose_cs = pose_chain_stride; a.pose_fs = pose_frame_stride; a.target_row = (int)target_row;
    a.params = params;
    a.angles = angles; a.ang_cs = ang_chain_stride; a.ang_fs = ang_frame_stride;
    a.fk = fk; a.fk_cs = fk_chain_stride; a.fk_fs = fk_frame_stride;
    a.warm = warm; a.warm_cs = warm_chain_stride;
    a.status = status; a.nfev = nfev; a.n_chain = n_chain; a.n_frame = n_frame;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    // Measured (profiles/r02_generic_kernel.jsonl): eight lanes per chain win while the whole batch is resident at once --
    // float64 6 000 chains 282 against 410 ms (183 registers and no spills against 255 and 480 B), float32 71 against 82 ms --
    // and lose beyond (60 000 float32 chains: 444 against 273 ms: a lane per chain keeps 320 chains per SM in flight, the
    // groups 64); the evaluation stays a latency chain either way (the butterfly sums cost what the seven-term sums did).
    cudaFuncAttributes at;
    int warps_sm = 16;
    if (cudaFuncGetAttributes(&at, leg_solve_generic_group_kernel<R>) == cudaSuccess && at.numRegs > 0) {
        warps_sm = 65536 / (((at.numRegs + 7) / 8 * 8) * 32);
        warps_sm = (warps_sm / 4) * 4;                                 // registers are partitioned over the four schedulers
        if (warps_sm > 32) warps_sm = 32;
        if (warps_sm < 4) warps_sm = 4;
    }
    // (float64: the one-lane kernel spills and holds 8 warps per SM, so the groups still win at 1.3 waves: 6 000 chains)
    const bool group = sched == 2 || (sched == 0 && n_chain <= (sizeof(R) == 8 ? 6LL : 4LL) * warps_sm * n_sm);
    if (group) {
        // fill the warps early: fewer warps at different places of the (long) evaluation code run faster than more
        int gpw = (int)((n_chain + 4LL * n_sm - 1) / (4LL * n_sm));
        gpw = gpw < 1 ? 1 : (gpw > 4 ? 4 : gpw);
        const uint32_t forced_g = (flags >> SEQIK_FLAG_CPW_SHIFT) & 0x3F;       // tuning / tests: chains per warp, capped at 4 here
        if (forced_g) gpw = forced_g > 4 ? 4 : (int)forced_g;
        const int64_t grid_g = (n_chain + gpw - 1) / gpw;
        leg_solve_generic_group_kernel<R><<<(unsigned)grid_g, 32, 0, (cudaStream_t)stream>>>(a, gpw);
        return seqik_check_launch(me);
    }
    int cpw = (int)((n_chain + 4LL * n_sm - 1) / (4LL * n_sm));
    cpw = cpw < 1 ? 1 : (cpw > 32 ? 32 : cpw);
    const uint32_t forced = (flags >> SEQIK_FLAG_CPW_SHIFT) & 0x3F;         // tuning / tests
    if (forced) cpw = (int)forced;
    if (cpw < 1 || cpw > 32) return seqik_fail(SEQIK_EINVAL, "%s: chains per warp must be 1..32", me);
    const int64_t grid = (n_chain + cpw - 1) / cpw;
    leg_solve_generic_kernel<R><<<(unsigned)grid, 32, 0, (cudaStream_t)stream>>>(a, cpw);
    return seqik_check_launch(me);
}

extern "C" int seqik_leg_solve_generic_f32(const float* pose, int64_t pose_chain_stride, int64_t pose_frame_stride, int32_t target_row,
                                           const float* params,
                                           float* angles, int64_t ang_chain_stride, int64_t ang_frame_stride,
                                           float* fk, int64_t fk_chain_stride, int64_t fk_frame_stride,
                                           const float* warm, int64_t warm_chain_stride,
                                           int32_t* status, uint32_t* nfev,
                                           int64_t n_chain, int64_t n_frame, uint32_t flags, void* stream) {
    return launch_generic<float>("seqik_leg_solve_generic_f32", pose, pose_chain_stride, pose_frame_stride, target_row, params,
                                 angles, ang_chain_stride, ang_frame_stride, fk, fk_chain_stride, fk_frame_stride,
                                 warm, warm_chain_stride, status, nfev, n_chain, n_frame, flags, stream);
}

extern "C" int seqik_leg_solve_generic_f64(const double* pose, int64_t pose_chain_stride, int64_t pose_frame_stride, int32_t target_row,
                                           const double* params,
                                           double* angles, int64_t ang_chain_stride, int64_t ang_frame_stride,
                                           double* fk, int64_t fk_chain_stride, int64_t fk_frame_stride,
                                           const double* warm, int64_t warm_chain_stride,
                                           int32_t* status, uint32_t* nfev,
                                           int64_t n_chain, int64_t n_frame, uint32_t flags, void* stream) {
    return launch_generic<double>("seqik_leg_solve_generic_f64", pose, pose_chain_stride, pose_frame_stride, target_row, params,
                                  angles, ang_chain_stride, ang_frame_stride, fk, fk_chain_stride, fk_frame_stride,
                                  warm, warm_chain_stride, status, nfev, n_chain, n_frame, flags, stream);
}
