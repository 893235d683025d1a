"""compute-sanitizer tier: tools/initcheck.py -- three small BASELINE-shaped echelonizations, twice
each so that the memory pool recycles blocks -- under memcheck, initcheck and racecheck.  The summaries land in
gpurun_out/sanitizer_<tool>.txt (tools/summarize_profiles.py copies them to profiles/).

memcheck and racecheck must report zero errors.  initcheck is run with --track-unused-memory no and must report zero
errors as well: every word a kernel reads was written by the library first (cudaMemsetAsync or a kernel)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool", ["memcheck", "initcheck", "racecheck"])
def test_compute_sanitizer_reports_no_error(tool):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    env = dict(os.environ, SCALE="0.35", REPS="2", SPASM_B200_BULK_MB="1000000")
    cmd = [exe, "--tool", tool, "--print-limit", "20", "--error-exitcode", "86", sys.executable, os.path.join(ROOT, "tools", "initcheck.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    text = out.stdout + out.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"sanitizer_{tool}.txt"), "w") as f:
        f.write(" ".join(cmd) + "\n" + text[-20000:])
    m = re.search(r"ERROR SUMMARY: (\d+) error", text) or re.search(r"RACECHECK SUMMARY: (\d+) hazard", text)
    assert m, text[-3000:]
    assert int(m.group(1)) == 0 and out.returncode == 0, text[-3000:]
    assert text.count(" rank ") >= 6, "the workload did not run to the end:\n" + text[-2000:]
