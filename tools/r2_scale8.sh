#!/usr/bin/env bash
# Round-2 8-GPU session: multi-rank parity at 8 ranks, the default strong-scaling bench at N=8, BASELINE
# configs[4] at full size (Jacobi and multigrid over strips), weak scaling of 1 M-DoF strips.
set -u
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
run() { N=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
        --master-port $((29500 + RANDOM % 400)) bench.py --gpus "$N" "$@"; }
(timeout 420 python -m pytest tests/test_gpu_multi.py -x -q -k "8]" 2>&1 | tail -4)
run 8 --steps 10 --warmup 3 > gpurun_out/r2_scale_n8.json 2> gpurun_out/r2_scale_n8.err
run 8 --workload c5-traveling-newmark-8192-p2 --steps 3 --warmup 3 > gpurun_out/r2_c5_n8_jacobi.json 2> gpurun_out/r2_c5_n8_jacobi.err
run 8 --workload c5-traveling-newmark-8192-p2 --precond mg --steps 3 --warmup 3 > gpurun_out/r2_c5_n8_mg.json 2> gpurun_out/r2_c5_n8_mg.err
run 8 --workload c2-standing-newmark-1024-p1 --steps 20 --warmup 3 > gpurun_out/r2_c2weak_n8.json 2> gpurun_out/r2_c2weak_n8.err
WAVE_CG_FUSED=1 run 8 --workload c2-standing-newmark-1024-p1 --steps 20 --warmup 3 > gpurun_out/r2_c2weak_n8_fused.json 2> gpurun_out/r2_c2weak_n8_fused.err
for f in r2_scale_n8 r2_c5_n8_jacobi r2_c5_n8_mg r2_c2weak_n8 r2_c2weak_n8_fused; do
  python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "its", d["run"]["cg_its_per_step"], "cg ms/it", d["cg"]["ms_per_iteration"], d["run"]["cg_path"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
done
