#!/bin/sh
# Build-container helper: place the UNMODIFIED reference scripts under baseline/_ref/ (git-ignored, travels
# with gpurun) so tests/test_gpu_scripts.py can run the reference's own train.py / main.py against this
# repo's modules on the GPU box.  Nothing is copied into tracked files.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${TNERF_REFERENCE:-/root/reference/src}"
mkdir -p "$ROOT/baseline/_ref/src"
cp "$SRC/train.py" "$SRC/main.py" "$SRC/make_gif.py" "$ROOT/baseline/_ref/src/"
echo "staged $(ls "$ROOT/baseline/_ref/src" | wc -l) reference scripts in baseline/_ref/src"
