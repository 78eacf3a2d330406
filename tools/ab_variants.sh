#!/bin/bash
# A/B the compiled-in tile-kernel experiments (QIPB200_TILE_VARIANT) on a B200:
#   gpurun --timeout 900 -- 'bash tools/ab_variants.sh > gpurun_out/ab_variants.log 2>&1'
# For every variant: the fused GPU parity tests first (a variant that fails them is not timed), then the
# N=30 headline circuit and the f32 QFT (3 timed steps each, no extras / CPU legs).
set -u -o pipefail
cd "$(dirname "$0")/.."
for v in 0 1 4 7 8 9; do
  echo "=== QIPB200_TILE_VARIANT=$v"
  if ! QIPB200_TILE_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused or schedule or qft or permutation" 2>&1 | tail -2; then
    echo "variant $v: parity tests FAILED or did not finish -- not timed"; continue
  fi
  QIPB200_TILE_VARIANT=$v timeout 120 python bench.py --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n30 f64', d['ms_per_step'], 'ms', d['gate_apps_per_s'], 'gate-apps/s')"
done
