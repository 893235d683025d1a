"""One-shot cost of the relinked reference tool (a fresh process per matrix, like a user of tools/rank):
writes the config-2 matrix as SMS text, runs oracle/_ref/b200_rank on it with the library's phase trace, prints the
wall clock of the process and the tool's own "done in ... s" line (the first call of a process)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spasm_b200 import synthetic
path = "/tmp/config2.sms"
if not os.path.exists(path):
    with open(path, "wb") as f:
        f.write(synthetic.config2().to_sms())
for tool in sys.argv[1:] or ["b200_rank"]:
    exe = os.path.join(ROOT, "oracle", "_ref", tool)
    env = dict(os.environ, SPASM_B200_TRACE="1")
    for rep in range(2):
        t0 = time.perf_counter()
        with open(path, "rb") as f:
            r = subprocess.run([exe], stdin=f, capture_output=True, env=env)
        dt = time.perf_counter() - t0
        lines = r.stderr.decode(errors="replace").replace("\r", "\n").splitlines()
        keep = [l for l in lines if "trace]" in l and ("core/" in l or "upload" in l or "assemble" in l or "context" in l) or "done in" in l or "rank =" in l]
        print(f"{tool} run {rep}: process wall {dt:.3f} s, exit {r.returncode}")
        for l in keep:
            print("   ", l.strip())
