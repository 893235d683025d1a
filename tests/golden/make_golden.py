"""Regenerates tests/golden/golden.json and tests/golden/fixtures.npz.

Run in the build container only (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

For every fixture matrix of the reference's test-suite (tests/Matrix/*.sms, the inputs of its
echelonize / kernel / schur tests, tests/CMakeLists.txt:64-100) and a list of moduli, and for small instances of the
five BASELINE configs, it runs the REFERENCE ITSELF (oracle/_ref/libspasm_ref.so: the C sources of
/root/reference/src compiled in place with one OpenMP thread, FFPACK boundary restated) and records the
canonical artefacts of SURVEY.md section 8c:
    rank, sha256(pivot column set), sha256(canonical RREF), sha256(canonical kernel basis), kernel dimension.
The fixture inputs themselves are stored as triplets in fixtures.npz so that the GPU box (which has no
/root/reference) can replay them.
"""
import glob
import json
import os
import sys

os.environ["OMP_NUM_THREADS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import util  # noqa: E402
from spasm_b200 import synthetic  # noqa: E402

MODULI = [3, 257, 42013, 65537, 189812507, 4294967291]     # reference: tests/CMakeLists.txt:46-53 (+ the tools' default)
SYNTHETIC = [
    ("config1", 0.05, {}), ("config1", 0.15, {}), ("config2T", 0.02, {}), ("config2T", 0.05, {}),
    ("config3", 0.01, {"sparsity_threshold": 0.01}), ("config3", 0.03, {"sparsity_threshold": 0.01}),
    ("config4", 0.004, {}), ("config4", 0.02, {}), ("config5", 0.03, {}), ("config5", 0.08, {}),
    ("config1", 0.04, {"enable_tall_and_skinny": False}), ("config2T", 0.03, {"dense_block_size": 64}),
    ("config4", 0.01, {"enable_greedy_pivot_search": False}),
]


def synth(name, scale):
    if name == "config2T":
        return synthetic.config2(scale).transposed()
    return synthetic.CONFIGS[name](scale)


def main():
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    cases = []
    arrays = {}
    try:
        for path in sorted(glob.glob("/root/reference/tests/Matrix/*.sms")):
            base = os.path.basename(path)[:-4]
            if base == "trefethen_2000":
                continue                          # commented out of the reference's own list (tests/CMakeLists.txt:86)
            t = util.load_sms(path, 257)
            arrays[f"{base}.i"], arrays[f"{base}.j"], arrays[f"{base}.x"] = t.i, t.j, t.x
            arrays[f"{base}.shape"] = np.array([t.n, t.m], np.int64)
            for p in MODULI:
                tp = t.with_prime(p)
                exp = util.run_reference(tp) if (t.n > 0 and t.m > 0) else None
                cases.append({"kind": "fixture", "name": base, "prime": p, "opts": {}, "expected": exp})
        for name, scale, opts in SYNTHETIC:
            t = synth(name, scale)
            cases.append({"kind": "synthetic", "name": name, "scale": scale, "prime": t.prime, "opts": opts,
                          "expected": util.run_reference(t, **opts)})
    finally:
        os.dup2(saved, 2)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "oracle/_ref (OMP_NUM_THREADS=1)", "cases": cases}, f, indent=1)
    print(f"{len(cases)} cases written")


if __name__ == "__main__":
    main()
