"""A/B kernel timing of alternative builds of the library (dev tool, GPU box): python scripts/ab_lib.py libA.so libB.so ..."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 2 or (len(sys.argv) == 2 and not os.environ.get("SEQIK_LIB_CHILD")):
    for lib in sys.argv[1:]:
        env = dict(os.environ, SEQIK_LIB_CHILD="1", SEQIK_LIB=lib)
        print(lib, subprocess.run([sys.executable, __file__, lib], env=env, capture_output=True, text=True).stdout.strip())
    sys.exit(0)
sys.path.insert(0, str(ROOT))
from seqikpy_b200 import _native as N
N.LIB_PATH = Path(os.environ["SEQIK_LIB"]).resolve()
sys.argv = [sys.argv[0]]
import runpy
ns = runpy.run_path(str(ROOT / "scripts" / "sweep.py"), run_name="sweep")
for n_trial in (100, 1000, 1250, 10000):
    ns["run"](n_trial, 500, 2, 0)
