"""The golden-vector exchange path: format, checker script and the Rust-side patch.

CPU: the checker passes on a dump in the patch's file set written by the oracle; a corrupted word is caught; the patch
still applies to the reference sources when they are present (build container only).  GPU (-m gpu): the same dump through
the CUDA path on both arithmetic paths."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import tfhe_research_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import check_reference_dump as crd  # noqa: E402


def test_checker_on_oracle_written_dump(tmp_path):
    d = str(tmp_path)
    crd.self_test_dump(T, d)
    assert sorted(os.listdir(d)) == sorted(f + ".bin" for f in crd.FILES)
    assert crd.check(d, want_gpu=False, quiet=True) == 0
    # header layout is the documented one: magic, version, kind, 10 x u32 params, 2 x f64, u64 count
    raw = open(os.path.join(d, "pbs_in.bin"), "rb").read()
    assert raw[:8] == b"TFHEB200" and int.from_bytes(raw[8:12], "little") == 1 and int.from_bytes(raw[12:16], "little") == 1
    assert np.frombuffer(raw[16:56], dtype="<u4").tolist() == [2, 9, 4, 1, 2, 32, 4, 5, 4, 6]
    assert int.from_bytes(raw[72:80], "little") == (len(raw) - 80) // 4 == 8 * 5
    # one flipped output bit must be reported
    kind, p, w = T.load_words(os.path.join(d, "gate_or_out.bin"))
    w[3] ^= 1
    T.save_words(os.path.join(d, "gate_or_out.bin"), kind, p, w)
    assert crd.check(d, want_gpu=False, quiet=True) == 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/src") or shutil.which("git") is None, reason="reference sources not present")
def test_patch_applies_to_reference_sources(tmp_path):
    dst = tmp_path / "ref"
    shutil.copytree("/root/reference", dst, ignore=shutil.ignore_patterns(".git"))
    subprocess.run(["git", "init", "-q"], cwd=dst, check=True)
    r = subprocess.run(["git", "apply", "--check", os.path.join(ROOT, "rust", "reference_dump.patch")], cwd=dst, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_checker_gpu_paths(tmp_path):
    d = str(tmp_path)
    crd.self_test_dump(T, d)
    assert crd.check(d, want_gpu=True, quiet=True) == 0
