import sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import po
from test_gpu_markers import _mk_text, _setup_rank, MK_TEXT, _rows
from citcomcu_b200 import decomp
from citcomcu_b200.problem import CartesianProblem
from citcomcu_b200.stokes import context_from_problem
text = _mk_text()
dumps, err = po.run_harness(text, Path(tempfile.mkdtemp(prefix="ccu_mk2_")), nsteps=2, marker_kat=True, nproc=2, timeout=300)
nproc = MK_TEXT["nproc"]
ctxs = []
for r, d in enumerate(dumps):
    me = decomp.me_loc_of(r, nproc)
    ctx = context_from_problem(CartesianProblem(text, me_loc=me), communicator=False)
    _setup_rank(ctx, d, nproc, me)
    ctxs.append(ctx)
dt = np.float32(dumps[0]["mk_doubles"][6])
for corrector, tag, key in ((0, "euler", "XMCpred"), (1, "rk", "XMC")):
    outs = [ctx.markers_step_export(dt, corrector) for ctx in ctxs]
    print(tag, "sent", [int(o[0].sum()) for o in outs], [np.nonzero(o[0])[0].tolist() for o in outs])
    for r, ctx in enumerate(ctxs):
        ctx.markers_import_finish(corrector, outs[1 - r][1])
    for r, (ctx, d) in enumerate(zip(ctxs, dumps)):
        m = ctx.markers_download()
        n_ref = int(d[f"mk_{tag}_nmarkers"][0])
        ours = np.stack([m[key][a] for a in range(3)], 1)
        ref = np.stack([d[f"mk_{tag}_{key}{a + 1}"] for a in range(3)], 1)
        so = set(map(tuple, ours)); sr = set(map(tuple, ref))
        print(" rank", r, "n ours/ref", ours.shape[0], n_ref, "only ours", len(so - sr), "only ref", len(sr - so))
        for p in list(so - sr)[:5]: print("   ours-only", p)
        for p in list(sr - so)[:5]: print("   ref-only ", p)
        if tag == "rk":
            xo = np.stack([m["XMCpred"][a] for a in range(3)], 1)
            print("   x range ours", ours[:,0].min(), ours[:,0].max(), " XP", d["mk_XP1"][[0,-1]])
