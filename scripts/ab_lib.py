"""A/B of alternative builds of the library (dev tool, GPU box): python scripts/ab_lib.py libA.so libB.so ...
Per library: kernel time at the benchmark sizes for each gate setting, and a SHA-256 of the solver's outputs over
96 chains x 1000 frames in every kernel variant (all stages / partial stages with frozen DOFs / joints-only FK /
no FK / lane-per-chain schedule / chunked frames / fused alignment map), so that a refactor can be checked bit for bit."""
import hashlib, json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 2 or (len(sys.argv) == 2 and not os.environ.get("SEQIK_LIB_CHILD")):
    for lib in sys.argv[1:]:
        env = dict(os.environ, SEQIK_LIB_CHILD="1", SEQIK_LIB=lib)
        r = subprocess.run([sys.executable, __file__, lib], env=env, capture_output=True, text=True)
        print("==", lib); print(r.stdout.strip()); print(r.stderr.strip()[-2000:])
    sys.exit(0)
sys.path.insert(0, str(ROOT))
from seqikpy_b200 import _native as N
N.LIB_PATH = Path(os.environ["SEQIK_LIB"]).resolve()
sys.argv = [sys.argv[0]]
import runpy
import torch
ns = runpy.run_path(str(ROOT / "scripts" / "sweep.py"), run_name="sweep")
engine, base, S, chain, init = ns["engine"], ns["base"], ns["S"], ns["chain"], ns["init"]
from seqikpy_b200.batch import chain_param_table
params = torch.from_numpy(chain_param_table(chain, init, S.LEGS, 16)).cuda()


def digest(*ts):
    h = hashlib.sha256()
    for t in ts:
        if t is not None:
            h.update(t.detach().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


out = {}
ang, fk, st, nf = engine.leg_solve(base, params)
out["full"] = digest(ang, fk, st, nf)
a1, f1, _, _ = engine.leg_solve(base, params, schedule=1)
out["lane"] = digest(a1, f1)
a2, f2, _, _ = engine.leg_solve(base, params, stages=(1, 2))
out["stages12"] = digest(a2, f2)
a3 = a2.clone()
a3, f3, _, _ = engine.leg_solve(base, params, stages=(3, 4), angles=a3)
out["stages34_frozen"] = digest(a3, f3)
a4, f4, _, _ = engine.leg_solve(base, params, fk_layout="joints")
out["joints"] = digest(a4, f4)
a5, _, _, _ = engine.leg_solve(base, params, want_fk=False)
out["nofk"] = digest(a5)
a6 = torch.empty_like(ang); f6 = torch.empty_like(fk)
for t0 in range(0, 1000, 160):
    engine.leg_solve(base, params, angles=a6, fk=f6, frames=(t0, min(t0 + 160, 1000)))
out["chunked160"] = digest(a6, f6)
out["chunked_equals_full"] = bool(torch.equal(a6, ang) and torch.equal(f6, fk))
aff = torch.zeros((96, 8), device="cuda"); aff[:, 3] = 1.05; aff[:, 0:3] = base[:, 0, 0, :]; aff[:, 4:7] = base[:, 0, 0, :]
a7, f7, _, _ = engine.leg_solve(base, params, affine=aff)
out["affine"] = digest(a7, f7)
for cpw in (1, 3, 8):
    a8, f8, _, _ = engine.leg_solve(base, params, chains_per_warp=cpw)
    out[f"cpw{cpw}_equals_full"] = bool(torch.equal(a8, ang) and torch.equal(f8, fk))
print(json.dumps(out))
for n_trial in (100, 1000, 1250, 10000):
    for gate in (0, 1, 2, 3):
        ns["run"](n_trial, 1000 if n_trial == 1000 else 500, 2, 0, gate=gate)
