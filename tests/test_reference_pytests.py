"""The reference's own pytest files for the chain-simulator path, unmodified, against this repository's
`mic_eq_core` door (tools/run_reference_pytests.py; the CPU oracle behind the door: no GPU in the build container).
Only where the reference tree exists; the fast subset (the full set takes minutes)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not Path("/root/reference/python/tests/test_eq_filter_types.py").exists(),
                    reason="reference tree not present (only in the build container)")
def test_reference_test_files_pass_against_the_product_door():
    run = subprocess.run([sys.executable, str(ROOT / "tools" / "run_reference_pytests.py"), "--fast"],
                         capture_output=True, text=True, timeout=600)
    tail = run.stdout.strip().splitlines()[-1] if run.stdout.strip() else run.stderr[-400:]
    assert run.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
    assert int(tail.split(" passed")[0].split()[-1]) >= 20, tail
