#!/usr/bin/env bash
# ncu --set full capture of the stencil SpMV (and the two CG vector kernels) on a matrix far larger than L2:
#   tools/ncu_stencil.sh [nel] [r] [tag]
# Run only after `python tools/spmv_probe.py --nel .. --r .. --cg` has exited 0 without ncu.
set -euo pipefail
nel="${1:-2048}"; r="${2:-2}"; tag="${3:-r2}"
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
python tools/spmv_probe.py --nel "$nel" --r "$r" --dt 0.002 --cg > "gpurun_out/probe_${tag}.json"
cat "gpurun_out/probe_${tag}.json"
WAVE_CG_FUSED=0 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_spmv_st' --launch-skip 2 --launch-count 1 \
    -o "gpurun_out/stencil_${tag}" -f python tools/spmv_probe.py --nel "$nel" --r "$r" --dt 0.002 --no-init --reps 4 \
    > "gpurun_out/ncu_stencil_${tag}.log" 2>&1 || tail -5 "gpurun_out/ncu_stencil_${tag}.log"
ls -la gpurun_out/stencil_${tag}.ncu-rep
