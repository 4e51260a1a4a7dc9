"""Per-lane simulator of the stage-pipeline kernel's loop (csrc/seqik_solver.cu: leg_solve_pipe_kernel): which lanes
close / open / evaluate in which iteration, for a given period of the open/close phases, driven by the REAL number of
trips of every (chain, stage, frame) solve taken from the host build of the solver core (tests/hostsim).  CPU only.

    python scripts/schedule_sim.py [flags-hex [periods...]]        e.g.  python scripts/schedule_sim.py 3f 1 2 4 6 8

Cost model: an iteration costs T when any lane evaluates and P when any lane closes or opens (the warp executes the
block for all its lanes).  T = 409 ns, P = 770 ns fit the measured config-3 times at periods 1 / 2 / 4 of the
reference-iterates flag set to within 1 % (5.63 / 4.18 / 3.73 ms per 1000 frames) and predicted the optimum at 6
(3.41 predicted, 3.58 measured) before it was measured; the same tool said that deferring the evaluation block or
vote-based gates would not pay.  A development aid, not a product path."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

NF, DEPTH = 1000, 4


def trips_of(flags, n_trial=3):
    """(chains, frames, 4) trips per solve = evaluations - 1, from the carried host emulation of the kernel's lanes."""
    import hostsim_build as H
    from seqikpy_b200 import synthetic as S
    from seqikpy_b200.kinematic_chain import KinematicChainSeq
    size, bounds, init = S.chain_constants()
    chain = KinematicChainSeq(bounds, list(S.LEGS), size)
    out = []
    for tr in range(n_trial):
        pose = S.make_trials([tr], NF)[0]
        for li, leg in enumerate(S.LEGS):
            _, _, nfev = H.run_carried_f32(pose[:, li], chain.pack_chain_params(leg, init[leg]), flags)
            out.append(nfev - 1)
    return np.stack(out)


def simulate(trips, period, t_trip=409.0, t_phase=770.0, depth=DEPTH):
    """One warp of len(trips) chains x 4 stage lanes.  Returns iterations, phases and modelled ns per frame."""
    n = len(trips)
    t = np.zeros((n, 4), int); started = np.zeros((n, 4), int); done = np.zeros((n, 4), int)
    solving = np.zeros((n, 4), bool); left = np.zeros((n, 4), int)
    it = cost = phases = 0
    while (t < NF).any():
        if it % period == 0:
            nxt = np.roll(started, -1, axis=1)
            room = np.ones((n, 4), bool); room[:, :3] = t[:, :3] < nxt[:, :3] + depth
            close = solving & (left == 0) & (t < NF) & room
            t[close] += 1; done[close] = t[close]; solving[close] = False
            prev = np.roll(done, 1, axis=1)
            ready = np.ones((n, 4), bool); ready[:, 1:] = t[:, 1:] < prev[:, 1:]
            opn = (~solving) & (t < NF) & ready
            for c, s in zip(*np.where(opn)):
                left[c, s] = trips[c, t[c, s], s]
            solving[opn] = True; started[opn] = t[opn] + 1
            if close.any() or opn.any():
                cost += t_phase; phases += 1
        go = solving & (left > 0)
        if go.any():
            cost += t_trip
        left[go] -= 1
        it += 1
    return {"iterations_per_frame": it / NF, "phases_per_frame": phases / NF, "ns_per_frame": cost / NF}


if __name__ == "__main__":
    flags = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0x3F
    periods = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 6, 8]
    tr = trips_of(flags)
    print(f"flags {flags:#x}: mean evaluations per solve by stage {np.round(tr.mean((0, 1)) + 1, 2)}")
    for p in periods:
        r = [simulate(tr[g * 8:(g + 1) * 8], p) for g in range(len(tr) // 8)]
        print(f"period {p}: " + ", ".join(f"{k} {np.mean([x[k] for x in r]):.2f}" for k in r[0]))
