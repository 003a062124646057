import subprocess, os, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from citcomcu_b200 import inputfile
wd = Path(tempfile.mkdtemp())
(wd / "out").mkdir(); (wd / "dump").mkdir()
(wd / "in.input").write_text(inputfile.busse1a(levels=4, maxstep=6))
env = dict(os.environ, CCU_MPI_NP="1", LD_PRELOAD=str(ROOT / "dropin/libcitcomcu_dropin.so"))
r = subprocess.run([str(ROOT / "oracle/_ref/ref_harness"), "dump", "in.input", "dump", "3"], cwd=wd, env=env, capture_output=True, text=True)
print("rc", r.returncode)
print(r.stderr[-1500:])
print(r.stdout[-500:])
