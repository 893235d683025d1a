"""Measured dense int8 tensor-core peak of this GPU ("builder must measure").

cuBLASLt int8 x int8 -> int32 through torch._int_mm at 8192^3 (2*N^3 integer ops): best of 10 single launches
("burst") and the average of back-to-back launches over ~4 s ("sustained"), CUDA events on torch's stream.
Writes profiles/int8_peak.json (copied from gpurun_out/ by the round-end script)."""
import json
import os
import sys
import time

import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
a = torch.randint(-128, 127, (N, N), dtype=torch.int8, device=dev)
b = torch.randint(-128, 127, (N, N), dtype=torch.int8, device=dev)
for _ in range(5):
    c = torch._int_mm(a, b)
torch.cuda.synchronize()
ops = 2.0 * N ** 3
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = torch._int_mm(a, b)
    e1.record()
    e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = max(10, int(4000.0 / best))
e0.record()
for _ in range(reps):
    c = torch._int_mm(a, b)
e1.record()
e1.synchronize()
sustained = e0.elapsed_time(e1) / reps
out = {"int8_tops_burst": ops / best / 1e9, "int8_tops_sustained": ops / sustained / 1e9, "N": N, "ms_burst": best,
       "ms_sustained": sustained, "reps_sustained": reps, "how": "torch._int_mm (cuBLASLt int8 -> int32), CUDA events",
       "gpu": torch.cuda.get_device_name(0), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/int8_peak.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
