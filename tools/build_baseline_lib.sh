#!/bin/bash
# Build os2d_b200/libos2d_b200_base.so from the csrc of a git revision (default HEAD) for in-process A/B timing (tools/gpu_ab.py).
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REV" os2d_b200/csrc include | tar -x -C "$TMP"
make -C "$TMP/os2d_b200/csrc" >/dev/null 2>&1
cp "$TMP/os2d_b200/libos2d_b200.so" "$ROOT/os2d_b200/libos2d_b200_base.so"
rm -rf "$TMP"
echo "built os2d_b200/libos2d_b200_base.so from $REV"
