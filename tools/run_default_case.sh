#!/usr/bin/env bash
# BASELINE configs[0]: the reference's default CPU-runnable case (parameters/sine-membrane.json as shipped,
# theta = 0.5, `mpirun -np 4 main-theta`) through the drop-in executable, whole-process wall time.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W=$(mktemp -d)
mkdir -p "$W/build" "$W/parameters"
python - "$ROOT" "$W" <<'PY'
import sys
sys.path.insert(0, sys.argv[1] + "/nmpde-wave-equation_b200")
from wavegpu.problems import problem, write_json
write_json(sys.argv[2] + "/parameters/sine-membrane.json", problem("sine-membrane", Save_Solution=False))
PY
cd "$W/build"
START=$(date +%s.%N)
"$ROOT/nmpde-wave-equation_b200/bin/main-theta" ../parameters/sine-membrane.json > run.log
END=$(date +%s.%N)
grep -E "Number of DoFs|Simulation completed|Elapsed time|Total CG" run.log
echo "whole-process wall time: $(python -c "print(round($END-$START,3))") s"
echo "energy.csv (last rows):"; tail -2 ../results/theta-sine-membrane/*/energy.csv
ls ../results/theta-sine-membrane/*/
