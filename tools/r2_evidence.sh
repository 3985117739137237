#!/usr/bin/env bash
# Round-2 evidence on one B200 (run after `python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline`
# has exited 0 without ncu): launch list of the default bench command, and one --set full capture of the three
# kernels of a CG iteration at the default workload's size (Nel=4096, R=2).
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_evidence_plain.json 2> gpurun_out/r2_evidence_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
python tools/spmv_probe.py --nel 4096 --r 2 --dt 0.001 --cg > gpurun_out/r2_probe_4096p2.json || exit 1
WAVE_CG_FUSED=0 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_spmv_st|k_cg_update|k_cg_direction' --launch-skip 30 --launch-count 3 \
    -o gpurun_out/r2_cg_kernels_4096p2 -f python tools/spmv_probe.py --nel 4096 --r 2 --dt 0.001 --cg --reps 2 \
    > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_cg_kernels_4096p2.ncu-rep gpurun_out/r2_launches_default.csv
