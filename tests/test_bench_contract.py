"""bench.py's output contract, checked on the arm that runs without a GPU (--impl reference): stdout is exactly one JSON
line with the keys the driver reads; anything else the run prints goes to stderr."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import oracle

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not oracle.ref_available(), reason="reference binaries not built (oracle/build_ref.sh)")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "genome_pairs_ani_per_sec" and d["higher_is_better"] is True
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
