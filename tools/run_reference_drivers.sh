#!/usr/bin/env bash
# Run the reference's sweep drivers (unmodified copies in REF/scripts, REF/parameters) against the executables
# in BIN with this repository's launcher, on a small grid, and collect the tables they write in OUT.
#   tools/run_reference_drivers.sh BIN REF OUT [per-driver time limit in seconds]
# BIN = nmpde-wave-equation_b200/bin on a GPU box (the product), tests/_build where there is no GPU (host
# classes on the test double of the C ABI).  REF is a scratch copy of the reference's scripts/ and parameters/
# made at run time (never committed).  See tools/reference_scripts_report.py.
set -u
BIN="$(cd "$1" && pwd)"; REF="$(cd "$2" && pwd)"; OUT="$3"; LIMIT="${4:-60}"
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$OUT"; OUT="$(cd "$OUT" && pwd)"
work="$(mktemp -d)"
run() { # name, driver arguments...
    local name="$1"; shift
    local t0=$SECONDS
    timeout "$LIMIT" python "$here/reference_scripts_report.py" run --reference "$REF" --work "$work" --bin "$BIN" \
        --tight-cg --script "$name" -- "$@" > "$OUT/$name.log" 2>&1
    echo "$name: exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
}
# SMALL=1: fewer runs (every run is a process that creates a CUDA context)
if [ "${SMALL:-0}" = 1 ]; then
    run convergence_sweep.py --nprocs 4 --nel 10 20 --r 1 2 --dt 0.05 0.01 --schemes theta-0.5 newmark-0.25 newmark-0.00
    run dissipation_dispersion_sweep.py --nprocs 4 --dt 0.1 0.05 --schemes theta-1.0 newmark-0.25
else
    run convergence_sweep.py --nprocs 4 --nel 10 20 --r 1 2 --dt 0.05 0.01 --schemes theta-0.5 theta-1.0 newmark-0.25 newmark-0.00
    run dissipation_dispersion_sweep.py --nprocs 4 --dt 0.15 0.1 0.05 --schemes theta-0.5 theta-1.0 newmark-0.25
fi
cp "$work"/build/*.csv "$OUT"/ 2>/dev/null
# partial tables of a driver that ran out of time: what the executables themselves appended
for d in "$work"/results/*; do
    [ -f "$d/convergence.csv" ] && cp "$d/convergence.csv" "$OUT/$(basename "$d")-convergence.csv"
done
cp -r "$work"/build/convergence-logs "$OUT"/ 2>/dev/null
ls "$OUT"
