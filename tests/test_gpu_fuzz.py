"""Short randomised soak on the device (tools/fuzz_bucket.py, tools/fuzz_all.py): random sizes, key distributions, alignments,
schedules and key types against torch.sort.  The long form (100 s per seed) is run by hand and recorded in profiles/README.md;
it found two bugs the fixed-size tests had missed (item-table end marker, typed keys of a big bucket inside a fitting item)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,key", [("fuzz_bucket.py", "fuzz"), ("fuzz_all.py", "fuzz_all")])
def test_randomised_soak(built_lib, tool, key):
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), "8", "12345"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    last = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(last)
    assert p.returncode == 0 and res[key] == "ok", (last, p.stderr[-2000:])
    assert sum(res["cases"].values()) > 50 if isinstance(res["cases"], dict) else res["cases"] > 50
