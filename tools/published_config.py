"""Run the reference's published strong-scaling configuration end to end through the drop-in
executable and report WHOLE-PROCESS wall time, the methodology of scripts/scalability_sweep.py:166-169
(standing mode, Nel=640, R=1, Dt=8e-5, T=0.05, logging and VTU off; min of 3 repeats).
Published: Newmark 1/4: 296.3 s (1 rank), 27.6 s (16 ranks), 20.0 s (32 ranks);
theta=0.5: 624.9 s / 55.0 s / 37.1 s on 2x Xeon Gold 6238R with Trilinos ML AMG.

    python tools/published_config.py            # on a B200 box
"""
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))
from wavegpu.problems import problem, write_json  # noqa: E402

PUBLISHED = {"main-newmark": {"1": 296.3, "16": 27.6, "32": 20.0}, "main-theta": {"1": 624.9, "16": 55.0, "32": 37.1}}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        (d / "build").mkdir()
        (d / "parameters").mkdir()
        p = problem("standing-mode-wsol", Nel="640", R="1", Dt="8e-05", T="0.05", Theta="0.5", Beta="0.25", Gamma="0.5",
                    Save_Solution=False, Enable_Logging=False, Log_Every=0)
        write_json(d / "parameters" / "scal-params.json", p)
        for exe in ("main-newmark", "main-theta"):
            times, steps, loop = [], None, None
            for _ in range(3):
                t0 = time.time()
                r = subprocess.run([str(ROOT / "nmpde-wave-equation_b200" / "bin" / exe), "../parameters/scal-params.json"],
                                   cwd=d / "build", capture_output=True, text=True)
                times.append(time.time() - t0)
                assert r.returncode == 0, r.stdout[-2000:]
                for line in r.stdout.splitlines():
                    if line.startswith("Simulation completed:"):
                        steps = int(line.split()[2])
                    if line.startswith("Elapsed time:"):
                        loop = float(line.split()[2])
            n = 641 * 641
            out[exe] = {"whole_process_s_min_of_3": min(times), "loop_only_s_last": loop, "steps": steps,
                        "dof_steps_per_s_whole_process": n * steps / min(times),
                        "published_whole_process_s": PUBLISHED[exe],
                        "speedup_vs_published_1_rank": PUBLISHED[exe]["1"] / min(times),
                        "speedup_vs_published_32_ranks": PUBLISHED[exe]["32"] / min(times)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
