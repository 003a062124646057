"""Regenerate tests/golden from the reference itself (run here, where /root/reference exists):

    python tests/golden/make_golden.py

Writes compact .npz dumps (fallback fixtures for boxes without oracle/_ref) and
busse1a_scalars.json (the SURVEY.md section 4 first-golden values, re-measured).
"""
import json
import re
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import pyoracle as po  # noqa: E402
from citcomcu_b200 import inputfile  # noqa: E402
from conftest import CASES  # noqa: E402

HERE = Path(__file__).resolve().parent
SKIP = ("GNX", "GDA", "Node_map", "XX1", "XX2", "XX3")

for name in ("busse_l3", "tdepv_l3_tight"):
    txt, nsteps, kat = CASES[name]()
    dumps, err = po.run_harness(txt, tempfile.mkdtemp(), nsteps=nsteps, kat=kat)
    d = dumps[0]
    keep = {k: d[k] for k in d.entries if not any(k.endswith(s) for s in SKIP)}
    np.savez_compressed(HERE / f"{name}.npz", **keep)
    print(name, sum(v.nbytes for v in keep.values()) / 1e6, "MB raw")

dumps, err = po.run_harness(inputfile.busse1a(levels=5, maxstep=1), tempfile.mkdtemp(), nsteps=0)
m = re.search(r"initial residue of momentum equation (\S+) (\d+)", err)
last = [l for l in err.split("\n") if l.startswith("AhatP")][-1]
loops = int(re.search(r"after \((\d+)\) pressure loops", err).group(1))
v = float(re.search(r"with v (\S+)", last).group(1))
p = float(re.search(r" p (\S+) dp/p", last).group(1))
json.dump(dict(v_res=float(m.group(1)), neq=int(m.group(2)), pressure_loops=loops, v=v, p=p,
               source="oracle/_ref/ref_harness on inputfile.busse1a(levels=5), step 0"),
          open(HERE / "busse1a_scalars.json", "w"), indent=1)
print(open(HERE / "busse1a_scalars.json").read())
