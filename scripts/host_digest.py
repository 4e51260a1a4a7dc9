"""SHA-256 over the outputs of the host build of the solver core (frame-by-frame solves, the lane state machine, the
carried solves, float64) for the flag sets that must NOT change when the optional steps are edited (0x3F, 0x0F, 0x00,
0x2F): a refactor of csrc/seqik_core.cuh that is meant to be arithmetic-neutral has to reproduce the digest.

    python scripts/host_digest.py        (compare with the value printed before the edit)

Development aid (CPU only)."""
import hashlib
import sys

import numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import hostsim_build as H
from seqikpy_b200 import synthetic as S
from seqikpy_b200.kinematic_chain import KinematicChainSeq
size,bounds,init = S.chain_constants()
chain = KinematicChainSeq(bounds, list(S.LEGS), size)
h=hashlib.sha256(); tot=0
for tr in range(2):
    pose = S.make_trials([tr],1000)[0]
    for li,leg in enumerate(S.LEGS):
        row = chain.pack_chain_params(leg, init[leg])
        for gm in (0x3F, 0x0F, 0x00, 0x2F):
            n = 1000 if gm==0x3F else 150
            ang,fk,nfev,st = H.solve_chain(pose[:n,li], row[0:4], row[4:11], row[11:18], row[25:29], row[18:25], gn_mask=gm)
            for a in (ang,fk,nfev,st): h.update(a.tobytes())
            tot+=int(nfev.sum())
        ang,fk,nf,steps = H.run_runner_f32(pose[:,li], row[0:4], row[4:11], row[11:18], row[25:29], row[18:25], gn_mask=0x3F)
        for a in (ang,fk,nf): h.update(a.tobytes())
        h.update(str(steps).encode())
        for gm in (0x3F,0x2F):
            ang,fk,nfev = H.run_carried_f32(pose[:,li], row, gm)
            for a in (ang,fk,nfev): h.update(a.tobytes())
            tot+=int(nfev.sum())
        # f64
        ang,fk,nfev,st = H.solve_chain(pose[:100,li], row[0:4], row[4:11], row[11:18], row[25:29], row[18:25], dtype=np.float64, gn_mask=0x3F)
        for a in (ang,fk,nfev,st): h.update(a.tobytes())
print(h.hexdigest(), tot)
