#!/bin/bash
# The one offline install of the UNMODIFIED reference (bench contract): /root/reference has no setup.py and is read-only,
# so its `os2d` package is copied to /tmp, given a three-line setup.py, and pip-installed into baseline/_ref/ (git-ignored,
# travels to the GPU box with the gpurun snapshot).  bench.py --impl reference / reference-gpu import it from there.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${1:-/root/reference}
TMP=$(mktemp -d)
cp -r "$SRC/os2d" "$TMP/os2d"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup, find_packages
setup(name="os2d-reference", version="0", packages=find_packages())
PY
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP" 2>&1 | tail -2
cp "$SRC/main.py" "$ROOT/baseline/_ref/main.py"     # the reference's entry script, for the hooked dry run (tests/test_gpu_main_dry_run.py)
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
