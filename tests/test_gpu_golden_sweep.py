"""Every known-answer row of the reference's result tables through libwavegpu (121 convergence rows,
15 dissipation rows; about 40 s on a B200)."""
import importlib.util
import json
import os
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_all_golden_rows_on_the_gpu(capsys):
    spec = importlib.util.spec_from_file_location("golden_sweep_gpu", ROOT / "tools" / "golden_sweep_gpu.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
    out = json.loads(capsys.readouterr().out)
    if os.environ.get("WAVE_GOLDEN_OUT"):  # keep the measured deviations (profiles/)
        Path(os.environ["WAVE_GOLDEN_OUT"]).write_text(json.dumps(out, indent=1))
    assert out["convergence_rows"] == 121 and out["dissdisp_rows"] == 15
    # 7 printed digits in the reference's tables, for P1 and P2 alike
    for key in ("worst_rel_dev_r1", "worst_rel_dev_r2"):
        assert out[key]["rel_L2"] < 1e-6 and out[key]["rel_H1"] < 1e-6, out
    assert out["energy_ratio_bit_identical_rows"] >= 14 and out["worst_energy_ratio_dev"] < 5e-6
