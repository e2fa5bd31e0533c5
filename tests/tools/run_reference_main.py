"""Dry run of the UNMODIFIED reference main.py (evaluation only) on a synthetic GroZi-format dataset, with the test-only yacs /
matplotlib stand-ins on sys.path.  usage: python tests/tools/run_reference_main.py <workdir> <hook: 0|1> [KEY VALUE ...]
<workdir> receives a copy of the reference's `os2d` package + main.py (the reference derives its data path from the package
location, os2d/utils/utils.py:13-16) and the dataset under <workdir>/data/."""
import os
import runpy
import shutil
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
work, hook = sys.argv[1], sys.argv[2] == "1"
opts = sys.argv[3:]
src = None
for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isdir(os.path.join(cand, "os2d", "modeling")) and os.path.isfile(os.path.join(cand, "main.py")):
        src = cand
        break
assert src is not None, "the reference (os2d package + main.py) is neither at /root/reference nor at baseline/_ref"
os.makedirs(work, exist_ok=True)
if not os.path.isdir(os.path.join(work, "os2d")):
    shutil.copytree(os.path.join(src, "os2d"), os.path.join(work, "os2d"))
    shutil.copy(os.path.join(src, "main.py"), os.path.join(work, "main.py"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
sys.path.insert(0, ROOT)
sys.path.insert(0, work)
import _synthetic_grozi  # noqa: E402
if not os.path.isfile(os.path.join(work, "data", "grozi", "classes", "grozi.csv")):
    _synthetic_grozi.make(os.path.join(work, "data"))
warnings.filterwarnings("ignore")
import torch  # noqa: E402
torch.backends.cudnn.allow_tf32 = False      # both arms in fp32, so that they can be compared at the parity bar
torch.backends.cuda.matmul.allow_tf32 = False
if not torch.cuda.is_available():
    torch.cuda.synchronize = lambda *a, **k: None       # the reference calls it unguarded (os2d/engine/evaluate.py:312,332)
if hook:
    import os2d_b200.install as install
    install.install()
sys.argv = [os.path.join(work, "main.py")] + opts
os.chdir(work)
runpy.run_path(os.path.join(work, "main.py"), run_name="__main__")
print("MAIN_DRY_RUN_DONE hook=%d" % int(hook))
