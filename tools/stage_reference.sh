#!/bin/sh
# Build-container helper: place the UNMODIFIED reference files where the checker side can find them on the GPU box
# (both directories are git-ignored and travel with gpurun; nothing is copied into tracked files):
#   baseline/_ref/src/  train.py main.py make_gif.py          -- the reference's own scripts, run against THIS repo's modules
#                                                                (tests/test_gpu_scripts.py)
#   oracle/_ref/src/    rays sampling encoding nerf volume utils data camera (+ tiny_nerf_min.py)
#                                                             -- the reference's op modules: the CPU arm of bench.py
#                                                                (`--impl reference`, cpu_baseline.kind = "reference") and the
#                                                                cross-check of oracle/oracle.py (tests/test_oracle_golden.py)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${TNERF_REFERENCE:-/root/reference/src}"
mkdir -p "$ROOT/baseline/_ref/src" "$ROOT/oracle/_ref/src"
cp "$SRC/train.py" "$SRC/main.py" "$SRC/make_gif.py" "$ROOT/baseline/_ref/src/"
cp "$SRC/rays.py" "$SRC/sampling.py" "$SRC/encoding.py" "$SRC/nerf.py" "$SRC/volume.py" "$SRC/utils.py" "$SRC/data.py" "$SRC/camera.py" "$SRC/tiny_nerf_min.py" "$ROOT/oracle/_ref/src/"
echo "staged $(ls "$ROOT/baseline/_ref/src" | wc -l) reference scripts in baseline/_ref/src, $(ls "$ROOT/oracle/_ref/src" | wc -l) reference modules in oracle/_ref/src"
