// seqik_solver.cu -- the sequential leg-IK solver kernels (sm_100a) behind seqik_leg_solve_f32.
//
// Work decomposition.  A chain = one leg of one trial.  Inside a chain the reference's data dependences are
// kept exactly (leg_inverse_kinematics.py:259-282): the stage-s solve of frame t starts from the stage-s
// result of frame t-1 (warm start, :272) and from the frame built by stages 1..s-1 of frame t (frozen links).
// That dependence graph is a 4 x n_frame wavefront: solve(s+1, t) and solve(s, t+1) are independent.
//
//   schedule 1, "lane per chain"        one lane runs stages 1..4 of frame t, then frame t+1, ...
//                                       Fewest lanes per chain: the throughput schedule for very many chains.
//   schedule 2, "stage pipeline"        four adjacent lanes own one chain, lane s runs stage s+1 of every frame and
//                                       hands the frame (3x3 orientation + pivot) to lane s+1 through a small
//                                       shared-memory ring.  Chain latency drops from the sum of the four stages'
//                                       evaluations per frame to the slowest stage's.  The schedule for the
//                                       benchmark configurations (6 000 - 7 500 chains per GPU leave a B200 mostly
//                                       idle under schedule 1, whose run time is one chain's latency).
//
// Both run the same per-lane arithmetic (seqik_core.cuh: StageSolve::init / restart / trip), one function evaluation
// per loop trip, in a warp-convergent loop; lanes sit at different (frame, stage) positions ("decoupled" trips).
// Schedule 2 carries a solve from frame to frame (restart: no trigonometry), which is also what lets the closed-form
// warm step end most solves with their first evaluation (SEQIK_FLAG_CLOSED_FORM); schedule 1 starts every solve afresh.
// No tensor cores: the work is scalar FP32 2x2 / 3x3 algebra (BASELINE.json north_star).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seqik.h"
#include "seqik_common.h"
#include "seqik_core.cuh"

using namespace seqik;

struct LegArgs {
    const float* pose; int64_t pose_cs, pose_fs;
    const float* affine; const float* params;
    float* angles; int64_t ang_cs, ang_fs;
    float* fk; int64_t fk_cs, fk_fs;
    const float* warm; int64_t warm_cs;
    int32_t* status; uint32_t* nfev;
    int64_t n_chain, n_frame;
    int stage_mask, gn_mask;
    int fk_joints;                          // SEQIK_FLAG_FK_JOINTS: fk holds the 4 joint rows only ([4][3] per leg-frame)
};

// alignment map applied on load (AlignPose.align_leg, alignment.py:471-485)
struct LoadMap {
    float fx, fy, fz, sc, tx, ty, tz; bool on;
    __device__ __forceinline__ void init(const float* affine, int64_t c) {
        on = affine != nullptr; fx = fy = fz = 0.f; sc = 1.f; tx = ty = tz = 0.f;
        if (on) { const float* q = affine + c * 8; fx = q[0]; fy = q[1]; fz = q[2]; sc = q[3]; tx = q[4]; ty = q[5]; tz = q[6]; }
    }
    __device__ __forceinline__ Vec3<float> apply(Vec3<float> v, int row) const {
        if (on) {
            if (row == 0) v = {tx, ty, tz};
            else v = {(v.x - fx) * sc + tx, (v.y - fy) * sc + ty, (v.z - fz) * sc + tz};
        }
        return v;
    }
};

// ---------------------------------------------------------------------------------------------
// schedule 1: one lane per chain
// ---------------------------------------------------------------------------------------------
// Global-memory IO policy of one chain.  Key points are read with plain (L1-cached) loads: a chain's
// frames are contiguous (60 B apart), so consecutive frames share 128 B lines.
struct DevIO {
    const float* pose; int64_t fs;          // base of this chain, frame stride
    const float* prm;                       // 32 floats
    float* ang; int64_t ang_fs;
    float* fk; int64_t fk_fs; bool fk_joints;
    LoadMap map;

    __device__ __forceinline__ Vec3<float> kp(int64_t t, int row) const {
        const float* p = pose + t * fs + row * 3;
        return map.apply({__ldg(p), __ldg(p + 1), __ldg(p + 2)}, row);
    }
    __device__ __forceinline__ void put_angles(int64_t t, const float* a, int i0, int i1) const {
        float* p = ang + t * ang_fs;
#pragma unroll
        for (int i = 0; i < 7; ++i) if (i >= i0 && i < i1) p[i] = a[i];
    }
    __device__ __forceinline__ float angle_in(int64_t t, int i) const { return ang[t * ang_fs + i]; }
    __device__ __forceinline__ void put_fk(int64_t t, int row, const Vec3<float>& v) const {
        if (fk_joints) { if (row < 5) return; row -= 5; }       // joints-only layout: rows 5..8 of the full one
        if (fk) { float* p = fk + t * fk_fs + row * 3; p[0] = v.x; p[1] = v.y; p[2] = v.z; }
    }
    __device__ __forceinline__ float seg(int i) const { return __ldg(prm + i); }
    __device__ __forceinline__ float lb(int i) const { return __ldg(prm + 4 + i); }
    __device__ __forceinline__ float ub(int i) const { return __ldg(prm + 11 + i); }
    __device__ __forceinline__ float null_sq(int i) const { return __ldg(prm + 25 + i); }
};

__global__ void __launch_bounds__(32) leg_solve_lane_kernel(LegArgs a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_chain) return;
    DevIO io;
    io.pose = a.pose + c * a.pose_cs; io.fs = a.pose_fs;
    io.prm = a.params + c * SEQIK_CHAIN_PARAM_FLOATS;
    io.ang = a.angles + c * a.ang_cs; io.ang_fs = a.ang_fs;
    io.fk = a.fk ? a.fk + c * a.fk_cs : nullptr; io.fk_fs = a.fk_fs; io.fk_joints = a.fk_joints != 0;
    io.map.init(a.affine, c);
    float seed[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) seed[i] = a.warm ? a.warm[c * a.warm_cs + i] : io.prm[18 + i];
    ChainRunner<float, DevIO> run;
    run.start(io, a.n_frame, seed, a.stage_mask, a.gn_mask);
    while (!run.finished()) run.step();
    if (a.status) a.status[c] = run.worst_status == ST_NONFINITE ? -1 : run.worst_status;
    if (a.nfev) { uint32_t* nf = a.nfev + c * 4; nf[0] = run.nf0; nf[1] = run.nf1; nf[2] = run.nf2; nf[3] = run.nf3; }
}

// ---------------------------------------------------------------------------------------------
// schedule 2: stage pipeline -- four adjacent lanes own one chain, lane s runs stage s+1 of every frame
// ---------------------------------------------------------------------------------------------
// One warp = `cpw` chains (1, 2, 4 or 8) x 4 stage lanes; lanes >= 4 cpw idle.  All lanes run the SAME straight-line
// code (StageSolve::trip is branch-free, the rotation kind is a per-lane select), so a warp iteration costs one trip
// whatever mix of positions its lanes are in.  Stage s hands frame t (orientation after its own rotation + the
// next pivot, 12 floats) to stage s+1 through a shared-memory ring of PIPE_DEPTH frames; progress counters travel
// by warp shuffle.  No block-level synchronisation: a block is one warp.
constexpr int PIPE_DEPTH = 4;               // frames a stage may run ahead of the next one
constexpr int PIPE_SLOT = 12;               // 3x3 frame (columns) + pivot
constexpr int PIPE_CHAINS = 8;              // chains per warp (maximum)

// kFull: all four stages solved and the nine-row FK layout written (the benchmark configurations and the dict API):
// the frozen-stage, partial-stage and output-layout decisions are compiled out.  Same arithmetic either way.
template <bool kFull>
__global__ void __launch_bounds__(32) leg_solve_pipe_kernel(LegArgs a, int cpw, int gate_period, int trip_period) {
    // hand-off ring: [stage][frame slot x 12 floats (+ 1 pad row)][chain].  Bank of an element = (8 (k + stage) + chain) mod 32:
    // the pad row shifts each stage's plane by 8 banks, so the 4 x 8 lanes of a warp, which all touch the same k at once,
    // hit 32 different banks (without it the four stage lanes of a chain collided: half of the kernel's shared-memory
    // wavefronts were bank-conflict replays, ncu l1tex__data_bank_conflicts_pipe_lsu_mem_shared).  [3] = identity
    // (stage 1's input)
    __shared__ float ring[4][PIPE_DEPTH * PIPE_SLOT + 1][PIPE_CHAINS];
    __shared__ float kpbuf[6][32];                                     // prefetched key points, one column per lane
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const int s = lane & 3;                                   // this lane's stage (0..3)
    const int cw = lane >> 2;                                 // chain within the warp
    const int64_t c = (int64_t)blockIdx.x * cpw + cw;
    int lo = 0, hi = 3;
    if (!kFull) {
        while (lo < 3 && !((a.stage_mask >> lo) & 1)) ++lo;
        while (hi > 0 && !((a.stage_mask >> hi) & 1)) --hi;
    }
    const bool owner = cw < cpw && c < a.n_chain;
    const bool live = owner && s <= hi;                       // lanes of stages after the last requested one idle
    const bool frozen = !kFull && s < lo;                     // DOFs read from the angles buffer, not solved
    const int n_frame = (int)a.n_frame;

    // per-lane constants of (chain, stage)
    const int64_t cc = owner ? c : 0;
    const float* prm = a.params + cc * SEQIK_CHAIN_PARAM_FLOATS;
    const float* pose = a.pose + cc * a.pose_cs;
    float* ang = a.angles + cc * a.ang_cs;
    float* fk = (kFull || a.fk) ? a.fk + cc * a.fk_cs : nullptr;
    const bool fk_joints = !kFull && a.fk_joints;
    LoadMap map; map.init(a.affine, cc);
    const int kind = (s == 0) ? KIND_XY : KIND_ZY;
    const float seg = __ldg(prm + s);
    const int ia = 2 * s, ib = (s == 3) ? 6 : 2 * s + 1;
    const float inf = Num<float>::inf();
    const float lb0 = (s == 3 || frozen) ? -inf : __ldg(prm + 4 + ia), ub0 = (s == 3 || frozen) ? inf : __ldg(prm + 11 + ia);
    const float lb1 = frozen ? -inf : __ldg(prm + 4 + ib), ub1 = frozen ? inf : __ldg(prm + 11 + ib);
    const float null_sq = frozen ? 0.f : __ldg(prm + 25 + s);
    const int n_full = (s == 0) ? 4 : (s == 1) ? 6 : (s == 2) ? 8 : 9;
    const int gn = stage_mode(a.gn_mask, s);                                  // StageSolve mode: Gauss-Newton, skip-confirm, Newton
    const bool esc = (a.gn_mask >> 4) & 1;
    const float has_a = (s == 3) ? 0.f : 1.f;
    const float* seed = a.warm ? a.warm + cc * a.warm_cs : prm + 18;
    float xa = (s == 3) ? 0.f : seed[ia], xb = seed[ib];                           // warm start, frame to frame

    // running output / input pointers of the frame this lane works on (advanced per frame: no 64-bit index math per access)
    float* pa_a = ang + ((s == 3) ? 6 : ia); float* pa_b = ang + ib;             // stage 4 has one DOF: both point at it
    float* pf = fk ? fk + 3 * s : nullptr;
    const float* pnext = pose;       // frame whose key points are prefetched next

    StageSolve<float> S;
    S.set_problem(kind, seg, has_a, null_sq, n_full, gn);
    if (live && !frozen) S.set_limit_trig(lb0, ub0);
    Mat3<float> A = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    Vec3<float> piv = {0.f, 0.f, 0.f}, o = {0.f, 0.f, 0.f}, rel = {0.f, 0.f, 0.f};
    int t = 0;                       // frame this lane works on
    int started = 0, done = 0;       // frames whose hand-off was consumed / produced by this lane
    bool solving = false;
    uint32_t nf = 0; int worst = ST_GTOL;
    S.status = ST_GTOL;
    // stage 1 starts every frame from the identity frame at the origin: it reads them from a constant ring slot like
    // the other stages read their producer's, so that the hand-off read is the same instruction for all lanes
    if (lane < PIPE_CHAINS)
        for (int d = 0; d < PIPE_DEPTH; ++d)
            for (int k = 0; k < PIPE_SLOT; ++k) ring[3][d * PIPE_SLOT + k][lane] = (k == 0 || k == 4 || k == 8) ? 1.f : 0.f;
    __syncwarp(full);
    const int sp = (s + 3) & 3;      // ring this lane reads from
    // key points of frame t (origin + this stage's target, 6 floats), fetched one frame ahead with cp.async into
    // shared memory: the copy is in flight during the previous solve and does not hold a register scoreboard
    // (plain loads made the first dependent instruction of every open wait a full DRAM latency)
    const uint32_t kp_dst = (uint32_t)__cvta_generic_to_shared(&kpbuf[0][lane]);
    auto prefetch = [&]() {
        const float* p = pnext;
        const float* q = p + 3 * (s + 1);
        pnext += a.pose_fs;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(kp_dst), "l"(p) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(kp_dst + 128), "l"(p + 1) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(kp_dst + 256), "l"(p + 2) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(kp_dst + 384), "l"(q) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(kp_dst + 512), "l"(q + 1) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(kp_dst + 640), "l"(q + 2) : "memory");
    };
    if (live && n_frame > 0) prefetch();
    bool carried = false;            // S holds the previous frame's solve of this (chain, stage)

    for (int gate_ctr = 1, trip_ctr = trip_period; __any_sync(full, live && t < n_frame);) {
        if (--gate_ctr == 0) {                          // open/close phases only every gate_period-th iteration (warp-uniform)
        gate_ctr = gate_period;
        __syncwarp(full);            // ring reads of the previous open phase are complete before a slot is written again
        const int started_next = __shfl_sync(full, started, (lane + 1) & 31);   // consumer's progress (lane + 1)
        // ---- optional singularity escape: a solve that ended on sin b = 0 may continue from a closed-form candidate
        if (esc && live && solving && !frozen && S.done() && S.escape_possible()) S.escape();
        // ---- close the converged solve: outputs + hand-off to the next stage (needs a free ring slot)
        if (live && t < n_frame && solving && S.done() && (s == hi || t < started_next + PIPE_DEPTH)) {
            if (!frozen) {
                xa = S.x0; xb = S.angle_b(); nf += (uint32_t)S.nfev;
                worst = (S.status == ST_MAXFEV && worst > ST_MAXFEV) ? ST_MAXFEV : worst;
                worst = (S.status == ST_NONFINITE) ? ST_NONFINITE : worst;
                *pa_a = (s == 3) ? xb : xa;
                *pa_b = xb;
            }
            // joint position = pivot + A w(x) = target + A f   (q = A^T rel, f = w - q)
            const Vec3<float> Af = mul(A, S.res());
            const Vec3<float> np_ = {(piv.x + rel.x) + Af.x, (piv.y + rel.y) + Af.y, (piv.z + rel.z) + Af.z};
            if (kFull || fk) {
                // rows 0-3 repeat the origin, 4 and 5 are both the Coxa-Femur joint: lane s writes origin row s and its
                // own joint row(s), which spreads the 27 floats of a frame over the four lanes
                const Vec3<float> jw = {np_.x + o.x, np_.y + o.y, np_.z + o.z};
                if (fk_joints) {                                                         // joints-only layout: row s = this lane's joint
                    pf[0] = jw.x; pf[1] = jw.y; pf[2] = jw.z;
                } else {
                    pf[0] = o.x; pf[1] = o.y; pf[2] = o.z;                               // row s
                    pf[15] = jw.x; pf[16] = jw.y; pf[17] = jw.z;                         // row 5 + s
                    if (s == 0) { pf[12] = jw.x; pf[13] = jw.y; pf[14] = jw.z; }         // row 4
                    if (!kFull && s == hi && hi < 3) for (int r = 1; r < 4 - hi; ++r) { pf[3 * r] = o.x; pf[3 * r + 1] = o.y; pf[3 * r + 2] = o.z; }
                }
                pf += a.fk_fs;
            }
            pa_a += a.ang_fs; pa_b += a.ang_fs;
            if (s < hi) {
                const Mat3<float> B = rotate_frame_sel(A, kind, S.sa, S.ca, S.sin_b(), S.cos_b());   // lanes mix kinds: no branch
                float (*q)[PIPE_CHAINS] = &ring[s][(t & (PIPE_DEPTH - 1)) * PIPE_SLOT];
                q[0][cw] = B.c0.x; q[1][cw] = B.c0.y; q[2][cw] = B.c0.z; q[3][cw] = B.c1.x; q[4][cw] = B.c1.y; q[5][cw] = B.c1.z;
                q[6][cw] = B.c2.x; q[7][cw] = B.c2.y; q[8][cw] = B.c2.z; q[9][cw] = np_.x; q[10][cw] = np_.y; q[11][cw] = np_.z;
            }
            solving = false; ++t; done = t;
        }
        __syncwarp(full);            // ring writes above are visible to the reads below
        const int done_prev = __shfl_sync(full, done, (lane + 31) & 31);        // producer's progress (lane - 1)
        // ---- open the next solve when the previous stage has published this frame
        if (live && t < n_frame && !solving && (s == 0 || t < done_prev)) {
            {
                const float (*q)[PIPE_CHAINS] = &ring[sp][(t & (PIPE_DEPTH - 1)) * PIPE_SLOT];
                A.c0 = {q[0][cw], q[1][cw], q[2][cw]}; A.c1 = {q[3][cw], q[4][cw], q[5][cw]}; A.c2 = {q[6][cw], q[7][cw], q[8][cw]};
                piv = {q[9][cw], q[10][cw], q[11][cw]};
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            const Vec3<float> ko = {kpbuf[0][lane], kpbuf[1][lane], kpbuf[2][lane]};
            const Vec3<float> kt = {kpbuf[3][lane], kpbuf[4][lane], kpbuf[5][lane]};
            if (t + 1 < n_frame) prefetch();
            o = map.apply(ko, 0);
            const Vec3<float> k = map.apply(kt, s + 1);
            rel = {(k.x - o.x) - piv.x, (k.y - o.y) - piv.y, (k.z - o.z) - piv.z};
            const Vec3<float> q3 = mulT(A, rel);
            // a carried solve continues from its own final iterate; every SEQIK_RESYNC frames (and at the first frame of
            // a call) the iterate is re-derived from the angles.  Same code but for the trigonometry (restart's `fresh`).
            const bool fresh = !(carried && !frozen && (t & (SEQIK_RESYNC - 1)) != 0);
            if (fresh) {
                if (frozen) { xa = (s == 3) ? 0.f : *pa_a; xb = *pa_b; }
                S.set_iterate(xa, xb);
                carried = true;
            }
            S.restart(q3, lb0, ub0, lb1, ub1, fresh, !frozen && (t > 0 || a.warm != nullptr));   // frozen DOFs stay where they are
            if (frozen) S.status = ST_GTOL;
            solving = true; started = t + 1;
        }
        }   // gate
        // ---- one function evaluation (every trip_period-th iteration, warp-uniform)
        if (--trip_ctr == 0) {
            trip_ctr = trip_period;
            if (live && solving && !S.done()) S.trip();
        }
    }
    // per-chain statistics
    const int w1 = min(worst, __shfl_xor_sync(full, worst, 1));
    const int w2 = min(w1, __shfl_xor_sync(full, w1, 2));
    if (owner) {
        if (a.nfev) a.nfev[c * 4 + s] = nf;
        if (a.status && s == 0) a.status[c] = w2 == ST_NONFINITE ? -1 : w2;
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int seqik_leg_solve_f32(const float* pose, int64_t pose_chain_stride, int64_t pose_frame_stride,
                                   const float* affine, const float* params,
                                   float* angles, int64_t ang_chain_stride, int64_t ang_frame_stride,
                                   float* fk, int64_t fk_chain_stride, int64_t fk_frame_stride,
                                   const float* warm, int64_t warm_chain_stride,
                                   int32_t* status, uint32_t* nfev,
                                   int64_t n_chain, int64_t n_frame, uint32_t stage_mask, uint32_t flags, void* stream) {
    if (n_chain < 0 || n_frame < 0) return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: negative size");
    if (n_chain == 0 || n_frame == 0) return SEQIK_OK;
    if (!pose || !params || !angles) return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: pose, params and angles must not be NULL");
    {
        uint32_t m = stage_mask;
        while (m && !(m & 1u)) m >>= 1;
        if (stage_mask == 0 || stage_mask > 0xF || (m & (m + 1)) != 0)
            return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: stage_mask must be a contiguous run of bits within 0xF");
    }
    const bool fk_joints = (flags & SEQIK_FLAG_FK_JOINTS) != 0;
    if (pose_frame_stride < 15 || ang_frame_stride < 7 || (fk && fk_frame_stride < (fk_joints ? 12 : 27)))
        return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: frame stride smaller than the innermost block");
    if (n_frame > 2147483647LL) return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: too many frames");
    uint32_t sched = (flags & SEQIK_FLAG_SCHED_MASK) >> SEQIK_FLAG_SCHED_SHIFT;
    if (sched > 2) return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: unknown schedule");
    // automatic: the stage pipeline (measured faster than one lane per chain from 600 to 60 000 chains)
    if (sched == 0) sched = 2;
    LegArgs a;
    a.pose = pose; a.pose_cs = pose_chain_stride; a.pose_fs = pose_frame_stride;
    a.affine = affine; a.params = params;
    a.angles = angles; a.ang_cs = ang_chain_stride; a.ang_fs = ang_frame_stride;
    a.fk = fk; a.fk_cs = fk_chain_stride; a.fk_fs = fk_frame_stride;
    a.warm = warm; a.warm_cs = warm_chain_stride;
    a.status = status; a.nfev = nfev; a.n_chain = n_chain; a.n_frame = n_frame;
    a.fk_joints = fk_joints ? 1 : 0;
    a.stage_mask = (int)stage_mask; a.gn_mask = (int)(flags & 0xFF);   // bits 0-3 Gauss-Newton mode per stage, 4 escape, 5 skip-confirm, 6 Newton, 7 closed-form warm step
    if (sched == 1) {
        const int64_t grid = (n_chain + 31) / 32;
        leg_solve_lane_kernel<<<(unsigned)grid, 32, 0, (cudaStream_t)stream>>>(a);
    } else {
        // chains per warp (measured): spread the chains over one warp per SM sub-partition (4 x 148) while
        // that is possible, then fill the warps -- up to 5 chains (20 lanes) while the warps still fit the SMs in about two
        // waves (a warp runs the iteration block for all its lanes whenever ONE lane needs it, so less-than-full warps
        // execute fewer instructions per chain: 6 000 / 7 500 / 15 000 chains are 3 / 6 / 18 % faster at 4 - 5 chains per
        // warp than at 8), 6 beyond that (18 000 - 60 000 chains: 6 is 15 - 22 % faster than 8; `scripts/sweep.py big`).
        // Small batches run at pure chain latency whatever the packing.
        int dev = 0, n_sm = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        int cpw = (int)((n_chain + 4LL * n_sm - 1) / (4LL * n_sm));
        // (with the iterating flag sets nearly every lane needs the iteration block anyway: full warps are best there)
        const int cap = !(flags & SEQIK_FLAG_CLOSED_FORM) ? PIPE_CHAINS : (n_chain <= 27LL * 4 * n_sm) ? 5 : 6;
        cpw = cpw < 1 ? 1 : (cpw > cap ? cap : cpw);
        const uint32_t forced = (flags >> SEQIK_FLAG_CPW_SHIFT) & 0x3F;     // tuning / tests
        if (forced) cpw = (int)forced;
        if (cpw < 1 || cpw > PIPE_CHAINS) return seqik_fail(SEQIK_EINVAL, "seqik_leg_solve_f32: chains per warp must be 1..8");
        const int64_t grid = (n_chain + cpw - 1) / cpw;
        // Period of the open/close phases.  A phase costs the warp about two trips whatever the number of lanes that take
        // part, so it pays to let nearly all lanes of the warp finish their solves and then close / open together: every
        // 6th iteration for the reference's iterates (3 - 6 trips per solve; config 3, ms per 1000 frames at period
        // 1 / 2 / 4 / 6 / 8 = 5.8 / 4.4 / 3.9 / 3.6 / 3.6), every 4th with Newton steps (2 - 3 trips), every 2nd with the
        // closed-form warm step every iteration for up to 7 000 chains (latency regime: a lane then handles one frame per
        // iteration) and every 3rd beyond (throughput regime: fewer instructions; 60 000 chains: 6.1 / 5.7 / 5.2 ms at
        // period 1 / 2 / 3, `profiles/r01_ab_merged_phase.txt`).  Scheduling only: results are unchanged.
        const uint32_t gate_sel = (flags >> SEQIK_FLAG_GATE_SHIFT) & 0xFu;    // 0 auto, else the period in iterations
        const int gate_period = gate_sel ? (int)gate_sel
                              : (flags & SEQIK_FLAG_CLOSED_FORM) ? (n_chain <= 7000 ? 1 : 3) : (flags & SEQIK_FLAG_NEWTON) ? 4 : 6;
        const uint32_t trip_sel = (flags >> SEQIK_FLAG_TRIP_SHIFT) & 0x7u;
        const int trip_period = trip_sel ? (int)trip_sel : 1;
        if (stage_mask == 0xF && fk && !fk_joints) leg_solve_pipe_kernel<true><<<(unsigned)grid, 32, 0, (cudaStream_t)stream>>>(a, cpw, gate_period, trip_period);
        else leg_solve_pipe_kernel<false><<<(unsigned)grid, 32, 0, (cudaStream_t)stream>>>(a, cpw, gate_period, trip_period);
    }
    return seqik_check_launch("seqik_leg_solve_f32");
}
