#!/bin/bash
# A/B the compiled-in tile-kernel experiments (QIPB200_TILE_VARIANT) on a B200:
#   gpurun --timeout 900 -- 'bash tools/ab_variants.sh > gpurun_out/ab_variants.log 2>&1'
# Timing first (N=30 headline circuit, 3 timed steps, no extras / CPU legs); the fused GPU parity tests are then
# run for every variant that beat the default.
set -u -o pipefail
cd "$(dirname "$0")/.."
for v in 0 1 4 7 8 9; do
  echo "=== QIPB200_TILE_VARIANT=$v"
  QIPB200_TILE_VARIANT=$v timeout 200 python bench.py --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n30 f64', d['ms_per_step'], 'ms', d['gate_apps_per_s'], 'gate-apps/s', 'parity_ok', d.get('parity_ok'))"
done
