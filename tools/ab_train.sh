#!/bin/sh
# developer A/B on the GPU box: time the fused fwd+bwd call (tools/hot_cold.py, last line = 124 rotating sets x 100 calls) under
# each environment setting given as an argument, e.g.  tools/ab_train.sh TNERF_TRAIN_ORDER=0 TNERF_TRAIN_ORDER=1
for kv in "$@"; do
  echo "== $kv"
  env $kv python tools/hot_cold.py 2>&1 | tail -3
done
