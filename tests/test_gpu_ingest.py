"""GPU-side ingest: spasm_triplet_load with the entry lines parsed on the device
(csrc/gpu/ingest.cu) against the host loop of csrc/host/io.c (the reference's line-by-line semantics,
src/spasm_io.c:59-159 + spasm_add_entry, src/spasm_triplet.c:7-24): same triplets in the same order, same dimensions,
same digest, same diagnostics."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from spasm_b200 import synthetic

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libc():
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    return libc


def load(product, path, prime, mode, monkeypatch, want_hash=True):
    monkeypatch.setenv("SPASM_B200_INGEST", mode)
    libc = _libc()
    f = libc.fopen(str(path).encode(), b"r")
    digest = (C.c_ubyte * 32)()
    T = product.spasm_triplet_load(f, prime, digest if want_hash else None)
    libc.fclose(f)
    t = T.contents
    nz = int(t.nz)
    out = {"n": t.n, "m": t.m, "nz": nz, "nzmax": int(t.nzmax),
           "i": np.ctypeslib.as_array(t.i, shape=(max(nz, 1),))[:nz].copy(),
           "j": np.ctypeslib.as_array(t.j, shape=(max(nz, 1),))[:nz].copy(),
           "x": (np.ctypeslib.as_array(t.x, shape=(max(nz, 1),))[:nz].copy() if t.x else None),
           "hash": bytes(digest).hex()}
    product.spasm_triplet_free(T)
    return out


def same(a, b):
    assert (a["n"], a["m"], a["nz"], a["hash"]) == (b["n"], b["m"], b["nz"], b["hash"])
    assert (a["i"] == b["i"]).all() and (a["j"] == b["j"]).all()
    if a["x"] is None:
        assert b["x"] is None
    else:
        assert (a["x"] == b["x"]).all()


@pytest.mark.parametrize("name,scale", [("config2", 0.3), ("config1", 1.0), ("config5", 0.5)])
def test_sms_parsed_on_the_gpu(product, tmp_path, monkeypatch, name, scale):
    t = synthetic.CONFIGS[name](scale)
    text = t.to_sms()
    path = tmp_path / "a.sms"
    path.write_bytes(text)
    host_side = load(product, path, t.prime, "host", monkeypatch)
    gpu_side = load(product, path, t.prime, "gpu", monkeypatch)
    same(gpu_side, host_side)
    assert gpu_side["hash"] == hashlib.sha256(text).hexdigest()
    assert gpu_side["nz"] > 0
    # the default (size threshold) takes the GPU for a file of this size
    monkeypatch.delenv("SPASM_B200_INGEST", raising=False)
    assert len(text) >= 1 << 20 or name != "config2"


def test_awkward_text(product, tmp_path, monkeypatch):
    """signs, blanks, tabs, huge values, multiples of p (dropped), entries beyond the declared shape (it grows), text
    after the terminator (ignored with a warning), no newline at the end of the file, pattern-only load (prime = -1)"""
    p = 42013
    lines = ["3 4 M", "1 1 5", "  2\t3   -7 ", "3 4 42013", "3 2 84026", "1 2 +9", "7 9 123456789012345", "2 2 -42014", "1 4 0",
             "5 5 9223372036854775807", "0 0 0", "this is not an entry", "1 1 1"]
    path = tmp_path / "b.sms"
    path.write_text("\n".join(lines))                    # no final newline
    for prime in (p, -1):
        h = load(product, path, prime, "host", monkeypatch)
        g = load(product, path, prime, "gpu", monkeypatch)
        same(g, h)
        assert g["n"] == 7 and g["m"] == 9
    g = load(product, path, p, "gpu", monkeypatch)
    assert g["nz"] == 6 and g["x"].tolist() == [5, -7, 9, 123456789012345 % p if 123456789012345 % p <= p // 2 else 123456789012345 % p - p, -1,
                                                9223372036854775807 % p if 9223372036854775807 % p <= p // 2 else 9223372036854775807 % p - p]


def test_matrixmarket_parsed_on_the_gpu(product, tmp_path, monkeypatch):
    t = synthetic.config1(0.2)
    path = tmp_path / "c.mtx"
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n% a comment\n%another\n")
        f.write(f"{t.n} {t.m} {len(t.i)}\n")
        for i, j, x in zip(t.i.tolist(), t.j.tolist(), t.x.tolist()):
            f.write(f"{i + 1} {j + 1} {x}\n")
        f.write("trailing garbage\n")
    h = load(product, path, t.prime, "host", monkeypatch)
    g = load(product, path, t.prime, "gpu", monkeypatch)
    same(g, h)
    assert g["nzmax"] == h["nzmax"] == len(t.i)


@pytest.mark.parametrize("bad,message", [("1 1\n", b"parse error line 3"), ("\n", b"parse error line 3"), ("1 x 3\n", b"parse error line 3"),
                                         ("1 1 " + "7" * 1100 + "\n", b"line 3 too long"), (None, b"premature end of file")])
def test_diagnostics_are_the_host_ones(tmp_path, bad, message):
    """a malformed file aborts with the reference's message and line number, whichever side parses it (run in a child
    process: errx exits)"""
    tool = os.path.join(ROOT, "oracle", "_ref", "b200_rank")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/b200_rank not built")
    body = "5 5 M\n1 1 1\n2 2 1\n" + (bad if bad is not None else "") + "3 3 1\n" + ("0 0 0\n" if bad is not None else "")
    path = tmp_path / "bad.sms"
    path.write_text(body)
    for mode in ("host", "gpu"):
        env = dict(os.environ, SPASM_B200_INGEST=mode, OMP_NUM_THREADS="1")
        with open(path, "rb") as f:
            r = subprocess.run([tool], stdin=f, capture_output=True, env=env, timeout=120)
        assert r.returncode == 1, (mode, r.stderr[-300:])
        assert message in r.stderr, (mode, r.stderr[-300:])


def test_ingest_rate(product, tmp_path, monkeypatch):
    """not an assertion on speed, a record: seconds of spasm_triplet_load on the BASELINE config 2 text, host and GPU"""
    import time
    t = synthetic.config2(1.0)
    text = t.to_sms()
    path = tmp_path / "big.sms"
    path.write_bytes(text)
    out = {}
    for mode in ("host", "gpu", "gpu"):
        t0 = time.perf_counter()
        r = load(product, path, t.prime, mode, monkeypatch, want_hash=False)
        out[mode] = time.perf_counter() - t0
        assert r["nz"] == len(t.i)
    print(f"\n[ingest] config 2 text: {len(text) / 1e6:.1f} MB, host {out['host'] * 1e3:.1f} ms, gpu {out['gpu'] * 1e3:.1f} ms "
          f"({len(text) / out['gpu'] / 1e9:.2f} GB/s)")
