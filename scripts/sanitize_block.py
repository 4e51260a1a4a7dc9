"""Small run of the block schedule (both kernels, bulk and plain IO paths, a replay-heavy real recording) for compute-sanitizer."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from seqikpy_b200 import data as D, engine, synthetic as S
from seqikpy_b200.batch import chain_param_table
from seqikpy_b200.kinematic_chain import KinematicChainSeq

size, bounds, init = S.chain_constants()
chain = KinematicChainSeq(bounds, list(S.LEGS), size)
pose = S.to_chains(torch.from_numpy(S.make_trials(range(2), 1000)).cuda())[:, :200].contiguous()
params = torch.from_numpy(chain_param_table(chain, init, S.LEGS, 2)).cuda()
for variant in (1, 2):                                                             # lean and robust kernel
    for kw in ({}, {"fk_layout": "joints"}, {"want_fk": False}):
        engine.leg_solve(pose, params, schedule=3, block_variant=variant, **kw)
engine.leg_solve(pose[:, :97].contiguous(), params, schedule=3)                    # ragged last block: plain loads / stores
ang = torch.zeros((12, 200, 7), device="cuda"); fk = torch.zeros((12, 200, 9, 3), device="cuda")
engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=3, frames=(0, 64))
engine.leg_solve(pose, params, angles=ang, fk=fk, schedule=3, frames=(64, 200))   # warm-started range
g = dict(np.load(ROOT / "tests/golden/grooming_leg.npz"))
ch = KinematicChainSeq(D.BOUNDS, ["RF", "LF"], None)
prm = torch.from_numpy(np.stack([ch.pack_chain_params(l, D.INITIAL_ANGLES[l]) for l in ("RF", "LF")]).astype(np.float32)).cuda()
gp = torch.from_numpy(np.ascontiguousarray(g["pose"][:, :640], dtype=np.float32)).cuda()
engine.leg_solve(gp, prm, schedule=3, block_variant=1)                             # replays (LF frames 280-304), lean kernel
a, f, st, nf = engine.leg_solve(gp, prm, schedule=3, block_variant=2)              # ... and the robust one
torch.cuda.synchronize()
print("ok", int(nf.sum()), st.tolist())
