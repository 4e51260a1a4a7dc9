"""Schedule 2 vs 3 on REAL data: the bundled grooming trial (2 legs x 6000 frames; 3-4 % of its frames replay through the
serial solver in schedule 3) tiled to many chains. Dev tool (GPU box)."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from seqikpy_b200 import data as D, engine
from seqikpy_b200.kinematic_chain import KinematicChainSeq

g = dict(np.load(ROOT / "tests/golden/grooming_leg.npz"))
ch = KinematicChainSeq(D.BOUNDS, ["RF", "LF"], None)
prm = torch.from_numpy(np.stack([ch.pack_chain_params(l, D.INITIAL_ANGLES[l]) for l in ("RF", "LF")]).astype(np.float32)).cuda()
gp = torch.from_numpy(np.ascontiguousarray(g["pose"], dtype=np.float32)).cuda()
for tiles in (1, 300, 1500):
    pose = gp.repeat(tiles, 1, 1, 1); params = prm.repeat(tiles, 1)
    n = pose.shape[0]
    ang = torch.empty((n, 6000, 7), device="cuda"); fk = torch.empty((n, 6000, 9, 3), device="cuda")
    for sched, variant in ((2, 0), (3, 1), (3, 2)):
        for _ in range(2):
            engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=sched, block_variant=variant, want_stats=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=sched, block_variant=variant, want_stats=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"data": "grooming trial tiled", "chains": n, "frames": 6000, "sched": sched, "variant": variant, "ms": round(ms, 3), "Glf_s": round(n * 6000 / ms / 1e6, 3)}), flush=True)
