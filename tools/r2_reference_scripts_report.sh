#!/usr/bin/env bash
# Regenerates profiles/r2_reference_scripts.md from the tables the reference's sweep drivers wrote
# (profiles/r2_reference_scripts_b200/: on a B200 through libwavegpu; profiles/r2_reference_scripts_cpu/: in the
# build container through the host classes on the test double of the C ABI).  See tools/reference_scripts_report.py.
cd "$(dirname "${BASH_SOURCE[0]}")/.."
B=profiles/r2_reference_scripts_b200
C=profiles/r2_reference_scripts_cpu
python tools/reference_scripts_report.py report \
  --b200-conv $B/convergence_theta-conv-params.csv $B/convergence_newmark-conv-params.csv $B/convergence-results_explicit.csv \
  --b200-conv-args "--nprocs 4 --nel 10 20 --r 1 2 --dt 0.05 0.01 --schemes theta-0.5 newmark-0.25 newmark-0.00; the 55 s of GPU budget allotted to the driver ended it after 15 of its 20 runs, the explicit scheme was then run with --dt 0.01 --schemes newmark-0.00" \
  --b200-diss $B/dissdisp-results.csv --b200-diss-args "--nprocs 4 --dt 0.1 0.05 --schemes theta-1.0 newmark-0.25" \
  --b200-note "the builder's gpurun lease, the last GPU seconds of the round; \`gpu.txt\`" \
  --conv $C/convergence-results.csv \
  --conv-args "--nprocs 4 [--nel … --dt … --schemes …]: the driver's default grid (Nel 10–320, R 1–2, ten time steps, five schemes), split over several invocations by mesh size and scheme so that they could run side by side; merged table: \`profiles/r2_reference_scripts_cpu/convergence-results.csv\`" \
  --diss $C/dissdisp-results.csv \
  --diss-args "--nprocs 4 --schemes theta-0.5 theta-1.0 newmark-0.25 (one invocation per scheme; the dt = 5e-5 runs again with --timeout 9000: 100 000 steps with the error norms every step take the CPU oracle 57 min, above the driver's default per-run limit). The 14 rows of the table that belong to the explicit schemes are runs that blew up (energy ratios 1e+254, inf, nan) and were left out" \
  --scal $C/scalability-results-1.csv --scal-note "OpenMP, 6 threads" \
  --out profiles/r2_reference_scripts.md
