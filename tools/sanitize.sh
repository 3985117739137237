#!/usr/bin/env bash
# compute-sanitizer over one tiny pass of the hot path (one tool per call, as the profiling guide asks):
#   tools/sanitize.sh memcheck|racecheck|initcheck|synccheck [out-dir]
# Runs __graft_entry__.smoke() (P2 Newmark, Nel=24, 5 steps, oracle-checked) under the tool and keeps the
# log under gpurun_out/ (copy the summary to profiles/).  Round 2: compute-sanitizer is closed on the GPU pool
# ("runs under it have left GPUs needing a reset"), so this recipe could not be executed there.
set -euo pipefail
tool="${1:-memcheck}"
out="${2:-gpurun_out}"
mkdir -p "$out"
cd "$(dirname "${BASH_SOURCE[0]}")/.."
timeout 900 compute-sanitizer --tool "$tool" --error-exitcode 1 --log-file "$out/sanitizer_${tool}.log" \
    python -c "import __graft_entry__ as g; g.smoke()"
tail -5 "$out/sanitizer_${tool}.log"
