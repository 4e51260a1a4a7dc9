"""Batched, multi-trial front end of the leg-IK hot path (BASELINE.json configs 3-5).

The reference handles one recording per ``LegInvKinSeq`` object and parallelises, at most, over the six
legs with ``multiprocessing.Pool`` (examples/example_leg_inv_kinematics_parallel.py:143-193).  Here many
trials are solved by one kernel launch: every (trial, leg) pair is an independent chain.

Layout (chain-major, float32): ``pose[trial, leg, frame, 5, 3]`` -> ``angles[trial, leg, frame, 7]``,
``fk[trial, leg, frame, 9, 3]``.  ``synthetic.to_chains`` converts from the frame-major
``(trial, frame, leg, 5, 3)`` layout.

Multi-GPU: trials are split into contiguous shards, one per rank (``shard_range``); there is no
collective on the data path -- chains never exchange data -- and results are gathered on the host once.
"""
from typing import Dict, Optional, Sequence

import numpy as np

from . import _native as N
from . import engine

RESYNC_FRAMES = 32      # SEQIK_RESYNC of csrc/seqik_core.cuh


def bind_to_gpu_numa(device_index: int) -> Optional[Sequence[int]]:
    """Pin the calling process to the CPUs NVML reports as local to GPU ``device_index`` (before pinned host buffers are
    allocated, so that they land on the GPU's NUMA node).  With one process per GPU this keeps the host side of the
    PCIe copies off the inter-socket link.  Returns the CPU list used, or None when NVML / affinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception:
        return None


def shard_range(n_trial: int, rank: int, world_size: int):
    """Contiguous, balanced [lo, hi) trial range of ``rank`` (first ``n_trial % world_size`` ranks get one more)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(int(n_trial), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chain_param_table(kinematic_chain_class, initial_angles: Dict[str, Dict[str, np.ndarray]], legs: Sequence[str],
                      n_trial: int) -> np.ndarray:
    """(n_trial * n_leg, 32) float32 constant rows (include/seqik.h) -- the same six rows repeated per trial."""
    row = np.stack([kinematic_chain_class.pack_chain_params(leg, initial_angles[leg]) for leg in legs])
    return np.tile(row, (int(n_trial), 1)).astype(np.float32)


class BatchedLegIK:
    """Solver session for ``n_trial`` trials x ``len(legs)`` legs x ``n_frame`` frames on one device.

    Device buffers and pinned host result buffers are allocated once and reused by every call.
    """

    def __init__(self, kinematic_chain_class, initial_angles, legs: Sequence[str], n_trial: int, n_frame: int,
                 device="cuda", want_fk: bool = True, schedule: int = N.SCHED_AUTO, host_buffers: bool = True,
                 chains_per_warp: int = 0, fk_layout: str = "full", flags: int = N.FLAG_DEFAULT, wire: str = "same",
                 expand_threads: int = 0):
        """``flags``: solver flags (include/seqik.h; ``N.FLAG_REFERENCE_ITERATES`` walks the reference's own iterates).
        ``wire="joints"`` (with ``fk_layout="full"``): ``solve_host`` still returns the reference's 9-row FK, but only the
        four joint rows cross the host link (76 instead of 136 result bytes per leg-frame); rows 0-3 are rebuilt on the host
        from the origin row of the host-resident pose and row 4 from row 5 by ``expand_threads`` host threads
        (``seqik_fk_expand_host_f32``), chunk by chunk while later chunks are still in flight.  Bit-identical results.
        ``fk_layout``: "full" = the reference's 9 rows per leg-frame; "joints" = only the four rows that carry
        information (rows 5..8: rows 0-3 of the full layout repeat the input origin, row 4 repeats row 5), which cuts
        the device->host result from 136 to 76 bytes per leg-frame -- the end-to-end call is bound by that copy."""
        torch = N.require_cuda()
        N.load_library()
        self.torch = torch
        self.legs = list(legs)
        self.n_trial, self.n_leg, self.n_frame = int(n_trial), len(self.legs), int(n_frame)
        self.n_chain = self.n_trial * self.n_leg
        self.device = torch.device(device)
        self.schedule = schedule
        self.chains_per_warp = chains_per_warp
        self.flags = int(flags)
        self.params = torch.from_numpy(chain_param_table(kinematic_chain_class, initial_angles, self.legs, n_trial)).to(self.device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.d_pose = torch.empty((self.n_chain, self.n_frame, 5, 3), **f32)
        self.d_angles = torch.empty((self.n_chain, self.n_frame, 7), **f32)
        if fk_layout not in ("full", "joints"):
            raise ValueError(f"fk_layout must be 'full' or 'joints', got {fk_layout!r}")
        if wire not in ("same", "joints"):
            raise ValueError(f"wire must be 'same' or 'joints', got {wire!r}")
        self.wire_joints = wire == "joints" and fk_layout == "full" and want_fk
        self.out_rows = 9 if fk_layout == "full" else 4               # rows of the fk the caller receives
        if self.wire_joints:
            fk_layout = "joints"                                       # what the device computes and the link carries
        self.fk_layout = fk_layout
        self.fk_rows = 9 if fk_layout == "full" else 4
        self.d_fk = torch.empty((self.n_chain, self.n_frame, self.fk_rows, 3), **f32) if want_fk else None
        self.h_angles = self.h_fk = self.h_wire = None
        import os
        self.expand_threads = int(expand_threads) if expand_threads else max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
        if host_buffers:
            self.h_angles = torch.empty((self.n_chain, self.n_frame, 7), dtype=torch.float32, pin_memory=True)
            self.h_fk = torch.empty((self.n_chain, self.n_frame, self.out_rows, 3), dtype=torch.float32, pin_memory=True) if want_fk else None
        self.h_origin = self.d_origin = None
        if self.wire_joints:
            self.h_wire = torch.empty((self.n_chain, self.n_frame, 4, 3), dtype=torch.float32, pin_memory=True)
            # the origin rows travel compact next to the joint rows (12 B per leg-frame): the host then rebuilds rows 0-3 from
            # contiguous memory instead of walking its 60-byte-per-frame pose
            self.d_origin = torch.empty((self.n_chain, self.n_frame, 3), **f32)
            self.h_origin = torch.empty((self.n_chain, self.n_frame, 3), dtype=torch.float32, pin_memory=True)
        self.status = self.nfev = None
        self._copy_streams = None
        self.default_chunks = 8
        self.launches_per_call = 1

    @property
    def leg_frames(self) -> int:
        return self.n_chain * self.n_frame

    def solve_device(self, pose=None, affine=None, want_stats: bool = True):
        """One pass of the hot path over device-resident pose (default: the session's own pose buffer).
        Asynchronous on the current stream.  Returns (angles, fk) device tensors (session buffers)."""
        pose = self.d_pose if pose is None else pose.reshape(self.n_chain, self.n_frame, 5, 3)
        _, _, self.status, self.nfev = engine.leg_solve(pose, self.params, affine=affine, angles=self.d_angles, fk=self.d_fk,
                                                        want_fk=self.d_fk is not None, schedule=self.schedule,
                                                        chains_per_warp=self.chains_per_warp, flags=self.flags,
                                                        want_stats=want_stats, fk_layout=self.fk_layout)
        return self.d_angles, self.d_fk

    def solve_host(self, pose_host, synchronize: bool = True, n_chunks: Optional[int] = None, out=None):
        """Host (pinned) pose (n_trial, n_leg, n_frame, 5, 3) float32 -> pinned host (angles, fk).
        ``out``: optional (angles, fk) pinned host tensors of this session's shapes to receive the results instead of the
        session's own buffers (``MultiGpuLegIK`` passes each device its slice of ONE result tensor).

        The frames are cut into ``n_chunks`` ranges.  While range k is solved (warm-started from the last frame of
        range k-1: bit-identical to one launch over all frames), range k+1 is copied host->device and the results of
        range k-1 device->host on two copy streams, so PCIe traffic in both directions overlaps the kernel.
        """
        torch, lib = self.torch, N.load_library()
        h_angles, h_fk = (self.h_angles, self.h_fk) if out is None else out
        if h_angles is None:
            raise RuntimeError("session was created with host_buffers=False: pass out=(angles, fk)")
        if h_angles.numel() != self.n_chain * self.n_frame * 7 or not h_angles.is_contiguous() or h_angles.dtype != torch.float32:
            raise ValueError("out angles must be a contiguous float32 tensor of (n_trial, n_leg, n_frame, 7)")
        if self.d_fk is not None and (h_fk is None or h_fk.numel() != self.n_chain * self.n_frame * 3 * self.out_rows or not h_fk.is_contiguous()):
            raise ValueError(f"out fk must be a contiguous float32 tensor of (n_trial, n_leg, n_frame, {self.out_rows}, 3)")
        src = pose_host if isinstance(pose_host, torch.Tensor) else torch.from_numpy(pose_host)
        if src.dtype != torch.float32 or not src.is_contiguous() or src.numel() != self.n_chain * self.n_frame * 15:
            raise ValueError("pose_host must be a contiguous float32 array of (n_trial, n_leg, n_frame, 5, 3)")
        if n_chunks is None:
            n_chunks = self.default_chunks
        n_chunks = max(1, min(int(n_chunks), self.n_frame))
        main = torch.cuda.current_stream(self.device)
        if self._copy_streams is None:
            self._copy_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        s_in, s_out = self._copy_streams
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        F = self.n_frame
        # chunk starts on multiples of 32 frames (the kernel's resync period): then chunking is bit-identical
        # to one launch over all frames
        bounds = sorted({0, F} | {min(F, RESYNC_FRAMES * round(F * k / n_chunks / RESYNC_FRAMES)) for k in range(1, n_chunks)})

        arrived = []

        def copy2d(dst, src_, row_floats, t0, t1, direction, stream):
            off = 4 * row_floats * t0
            N.check(lib.seqik_memcpy2d_async(dst.data_ptr() + off, 4 * row_floats * F, src_.data_ptr() + off, 4 * row_floats * F,
                                             4 * row_floats * (t1 - t0), self.n_chain, direction, stream.cuda_stream),
                    "seqik_memcpy2d_async")
        with torch.cuda.device(self.device):
            for k in range(len(bounds) - 1):
                t0, t1 = bounds[k], bounds[k + 1]
                if t1 == t0:
                    continue
                copy2d(self.d_pose, src, 15, t0, t1, 1, s_in)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
                main.wait_event(ev_in)
                engine.leg_solve(self.d_pose, self.params, angles=self.d_angles, fk=self.d_fk, want_fk=self.d_fk is not None,
                                 schedule=self.schedule, chains_per_warp=self.chains_per_warp, want_stats=False, frames=(t0, t1),
                                 fk_layout=self.fk_layout, flags=self.flags)
                if self.wire_joints:
                    N.check(lib.seqik_origin_rows_f32(self.d_pose.data_ptr(), F * 15, 15, self.d_origin.data_ptr(), F * 3, self.n_chain,
                                                      t0, t1, main.cuda_stream), "seqik_origin_rows_f32")
                ev_k = torch.cuda.Event()
                ev_k.record(main)
                s_out.wait_event(ev_k)
                if self.wire_joints:                                   # what the host expansion needs first
                    copy2d(self.h_wire, self.d_fk, 12, t0, t1, 2, s_out)
                    copy2d(self.h_origin, self.d_origin, 3, t0, t1, 2, s_out)
                    ev_o = torch.cuda.Event()
                    ev_o.record(s_out)
                    arrived.append((ev_o, t0, t1))
                elif self.d_fk is not None:
                    copy2d(h_fk, self.d_fk, 3 * self.fk_rows, t0, t1, 2, s_out)
                copy2d(h_angles, self.d_angles, 7, t0, t1, 2, s_out)
            main.wait_stream(s_out)
        self.launches_per_call = len(bounds) - 1
        # joints-only wire format: the 9-row layout is rebuilt on the host, chunk by chunk as the copies land (later chunks
        # are still being solved / copied meanwhile); this part of the call is synchronous by nature
        for ev_o, t0, t1 in arrived:
            ev_o.synchronize()
            N.check(lib.seqik_fk_expand_host_f32(self.h_wire.data_ptr(), F * 12, 12, self.h_origin.data_ptr(), F * 3, 3, h_fk.data_ptr(), F * 27, 27,
                                                 self.n_chain, t0, t1, self.expand_threads), "seqik_fk_expand_host_f32")
        if synchronize:
            main.synchronize()
        return h_angles, h_fk

    def mean_fk_error(self, pose=None) -> float:
        """Mean over chains, frames and the 4 distal joints of |fk[5..8] - pose[1..4]| in mm.
        Reduction of RESULTS for reporting (torch ops on the device tensors), not part of the solve."""
        pose = self.d_pose if pose is None else pose.reshape(self.n_chain, self.n_frame, 5, 3)
        joints = self.d_fk if self.fk_layout == "joints" else self.d_fk[:, :, 5:9, :]      # (wire="joints": the device holds joints)
        d = joints - pose[:, :, 1:5, :]
        return float(d.square().sum(-1).sqrt().mean())


class MultiGpuLegIK:
    """The hot path over every visible GPU of one host, in one process: contiguous trial shards (``shard_range``), one
    ``BatchedLegIK`` session + stream set + issuing thread per device, NO collective -- chains never exchange data -- and
    the results of all devices land in ONE pinned host tensor, gathered once.  The multi-GPU analogue of the reference's
    only parallel driver, ``multiprocessing.Pool(6).starmap`` over legs followed by ``dict.update``
    (examples/example_leg_inv_kinematics_parallel.py:143-160,186-193): split, run, merge once.

    Results are bit-identical to a single-device ``BatchedLegIK`` over the same trials (chains are independent; tested).
    """

    def __init__(self, kinematic_chain_class, initial_angles, legs: Sequence[str], n_trial: int, n_frame: int,
                 devices: Optional[Sequence] = None, want_fk: bool = True, fk_layout: str = "full", flags: int = N.FLAG_DEFAULT):
        torch = N.require_cuda()
        N.load_library()
        self.torch = torch
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        self.devices = [torch.device(d) for d in devices]
        if not self.devices:
            raise N.SeqIKNativeError("MultiGpuLegIK needs at least one CUDA device")
        self.n_trial, self.n_leg, self.n_frame = int(n_trial), len(list(legs)), int(n_frame)
        self.fk_rows = 9 if fk_layout == "full" else 4
        world = len(self.devices)
        self.shards = [shard_range(self.n_trial, r, world) for r in range(world)]
        self.sessions = [None if hi == lo else
                         BatchedLegIK(kinematic_chain_class, initial_angles, legs, hi - lo, n_frame, device=d, want_fk=want_fk,
                                      host_buffers=False, fk_layout=fk_layout, flags=flags)
                         for d, (lo, hi) in zip(self.devices, self.shards)]
        self.h_angles = torch.empty((self.n_trial, self.n_leg, self.n_frame, 7), dtype=torch.float32, pin_memory=True)
        self.h_fk = (torch.empty((self.n_trial, self.n_leg, self.n_frame, self.fk_rows, 3), dtype=torch.float32, pin_memory=True)
                     if want_fk else None)
        self.status = None

    @property
    def leg_frames(self) -> int:
        return self.n_trial * self.n_leg * self.n_frame

    def solve_host(self, pose_host, n_chunks: Optional[int] = None):
        """Pinned host pose (n_trial, n_leg, n_frame, 5, 3) float32 -> pinned host (angles, fk) of all trials.  Each device
        runs the chunked copy/solve/copy pipeline of ``BatchedLegIK.solve_host`` on its shard, issued by its own thread
        (ctypes and the CUDA runtime release the GIL); returns when every device is done."""
        import threading
        torch = self.torch
        src = pose_host if isinstance(pose_host, torch.Tensor) else torch.from_numpy(pose_host)
        if src.dtype != torch.float32 or not src.is_contiguous() or src.numel() != self.leg_frames * 15:
            raise ValueError("pose_host must be a contiguous float32 array of (n_trial, n_leg, n_frame, 5, 3)")
        src = src.view(self.n_trial, self.n_leg, self.n_frame, 5, 3)
        errors = []

        def work(sess, dev, lo, hi):
            try:
                with torch.cuda.device(dev):
                    sess.solve_host(src[lo:hi], synchronize=True, n_chunks=n_chunks,
                                    out=(self.h_angles[lo:hi], None if self.h_fk is None else self.h_fk[lo:hi]))
            except Exception as exc:            # re-raised in the caller's thread
                errors.append(exc)
        threads = [threading.Thread(target=work, args=(s_, d, lo, hi))
                   for s_, d, (lo, hi) in zip(self.sessions, self.devices, self.shards) if s_ is not None]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return self.h_angles, self.h_fk


class FusedPipeline:
    """AlignPose + LegInvKinSeq + HeadInverseKinematics of many trials in one device-resident pass (BASELINE config 5).

    raw leg key points -> per-chain alignment statistics (series, radix-select mid-quantiles, affine rows) -> the
    solver with the affine map applied on load (no aligned copy of the pose is written) -> angles + FK;
    raw antenna key points + thorax -> head alignment rows -> head/antenna angles with the map applied on load.
    Equivalent to running the three drop-in classes one after the other for every trial (tests/test_gpu_parity.py).
    """

    def __init__(self, kinematic_chain_class, initial_angles, legs: Sequence[str], body_template, body_size,
                 n_trial: int, n_frame: int, device="cuda", include_claw: bool = False, with_head: bool = True):
        torch = N.require_cuda()
        self.torch = torch
        self.legs = list(legs)
        self.n_trial, self.n_frame = int(n_trial), int(n_frame)
        self.include_claw = bool(include_claw)
        self.with_head = with_head
        self.session = BatchedLegIK(kinematic_chain_class, initial_angles, legs, n_trial, n_frame, device=device,
                                    host_buffers=False)
        dev = self.session.device
        size = body_size
        rows = []
        for leg in self.legs:
            length = size[leg] - (0.0 if include_claw else size[f"{leg}_Tarsus"])
            rows.append(list(np.asarray(body_template[f"{leg}_Coxa"], dtype=float)) + [length])
        self.leg_consts = torch.from_numpy(np.tile(np.asarray(rows, dtype=np.float32), (self.n_trial, 1))).to(dev)
        if with_head:
            from .head_inverse_kinematics import HeadInverseKinematics
            dummy = {"R_head": np.zeros((1, 2, 3)), "L_head": np.ones((1, 2, 3)), "Neck": np.zeros((1, 1, 3))}
            hk = HeadInverseKinematics(dummy, body_template, log_level="ERROR", device=str(dev))
            self.rest = torch.tensor([[float(hk.rest_head_pitch[0]), float(hk.rest_antenna_pitch[0])]] * self.n_trial,
                                     dtype=torch.float32, device=dev)
            self.neck = torch.from_numpy(np.tile(np.asarray(body_template["Neck"], dtype=np.float32), (self.n_trial, 1))).to(dev)
            self.head_consts = {
                side: torch.from_numpy(np.tile(np.asarray(list(body_template[f"{side}_Antenna_base"])
                                                          + [size["Antenna_mid_thorax"], size["Antenna"]], dtype=np.float32),
                                               (self.n_trial, 1))).to(dev) for side in ("R", "L")}

    def run(self, raw_legs, r_head=None, l_head=None, thorax=None):
        """raw_legs (n_trial, n_leg, n_frame, 5, 3); r_head, l_head (n_trial, n_frame, 2, 3); thorax (n_trial, n_frame, k, 3):
        float32 CUDA tensors.  Returns a dict of device tensors (asynchronous on the current stream)."""
        sess = self.session
        pose = raw_legs.reshape(sess.n_chain, sess.n_frame, 5, 3)
        leg_aff = engine.leg_affine(pose, self.leg_consts, include_claw=self.include_claw)
        angles, fk = sess.solve_device(pose, affine=leg_aff, want_stats=True)
        out = {"angles": angles.view(self.n_trial, sess.n_leg, sess.n_frame, 7),
               "fk": fk.view(self.n_trial, sess.n_leg, sess.n_frame, 9, 3), "leg_affine": leg_aff.view(self.n_trial, sess.n_leg, 8)}
        if self.with_head and r_head is not None:
            aff_r, _ = engine.head_affine(r_head, thorax, self.head_consts["R"])
            aff_l, _ = engine.head_affine(l_head, thorax, self.head_consts["L"])
            out["head_angles"] = engine.head_angles(r_head, l_head, self.neck, self.rest, affine_r=aff_r, affine_l=aff_l)
            out["head_affine_r"], out["head_affine_l"] = aff_r, aff_l
        return out
