"""GPU: the randomised fuzzers under tests/tools/ with a fixed seed and a bounded number of cases, each in its own
process (they exit non-zero at the first case outside tolerance).  Forward / gradient against the oracle over random
(dims, T, P, precision); measurement_norm over all its paths and mask kinds; the GroupNorm kernels; large shapes against
the fp32 CUDA-core path; input layouts; the host API at the default precision."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tests", "tools")


@pytest.mark.gpu
@pytest.mark.parametrize("tool,args", [
    ("fuzz_shapes.py", ["40", "101"]),
    ("fuzz_dps.py", ["25", "102"]),
    ("fuzz_gn.py", ["60", "103"]),
    ("fuzz_large.py", ["12", "104"]),
    ("fuzz_api.py", ["20", "105"]),
    ("fuzz_layouts.py", []),
    ("fuzz_unet.py", ["6", "106"]),
])
def test_fuzzer(tool, args):
    r = subprocess.run([sys.executable, os.path.join(TOOLS, tool)] + args, capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-15:])
    assert r.returncode == 0 and ("all ok" in r.stdout or "worst forward error" in r.stdout), tail
