#!/bin/bash
# ncu captures of round 2 (one B200; run under gpurun; summaries are made on the CPU box with tools/ncu_summary.py):
#   gpurun --timeout 1500 -- 'bash tools/profile_r2.sh > gpurun_out/profile_r2.log 2>&1'
set -u
cd "$(dirname "$0")/.."
NCU="ncu --clock-control none"
B="python bench.py --no-extras --steps 1 --warmup 1"
# launch list of the fused N=28 circuit (per-launch device times, cold-cache: compare shares)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches_n28_fused.csv $B --n-local 28 --depth 10 > /dev/null
# the generated pass kernel (dominant kernel of the product path) and, for comparison, the interpreter kernel
$NCU --set full --import-source on -k regex:qip_pass -s 12 -c 2 -o gpurun_out/r2_qip_pass_n28 $B --n-local 28 --depth 10 > /dev/null
QIPB200_JIT=off $NCU --set full --import-source on -k regex:k_tile_pass -s 12 -c 1 -o gpurun_out/r2_k_tile_pass_n28 $B --n-local 28 --depth 10 > /dev/null
# f32 QFT pass (configs[2]) on the generated kernel
$NCU --set full -k regex:qip_pass -s 8 -c 1 -o gpurun_out/r2_qip_pass_qft_f32_n28 $B --n-local 28 --workload qft --dtype f32 > /dev/null
# per-gate kernels: CNOT (k_exchange), diagonal T (k_diag), dense 1-qubit, dense 4-qubit, dense 5-qubit
$NCU --set full -k regex:k_exchange -s 2 -c 1 -o gpurun_out/r2_k_exchange_n28 $B --n-local 28 --depth 3 --no-fusion > /dev/null
$NCU --set full -k regex:k_diag -s 2 -c 1 -o gpurun_out/r2_k_diag_n28 $B --n-local 28 --depth 3 --no-fusion > /dev/null
$NCU --set full --kernel-name-base demangled -k "regex:k_dense<double, 4" -s 2 -c 1 -o gpurun_out/r2_k_dense4_n26 $B --n-local 26 --workload dense4 --depth 6 --no-fusion > /dev/null
ls -la gpurun_out/*.ncu-rep
