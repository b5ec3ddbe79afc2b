"""Live pin of the golden fixtures (CPU tier, only where /root/reference exists -- i.e. in the build container, never on
the GPU box): oracle/make_golden.py drives the UNMODIFIED reference (oracle/ref_harness.py: /root/reference/src imported
under the no-op matplotlib / scipy `.A` / skimage + osqp shims) through its own control flow again, into a scratch
directory, and every array of every committed tests/golden/*.npz must come out bit for bit.  This is what ties the
committed fixtures -- and through tests/test_oracle_cpu.py the C oracle -- to the reference's code rather than to a file
somebody once generated."""
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference checkout exists only in the build container")
def test_golden_fixtures_regenerate_bit_for_bit(tmp_path):
    env = dict(os.environ, MPC_GOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(REPO, "oracle", "make_golden.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    names = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    assert names and sorted(os.listdir(tmp_path)) == names
    for name in names:
        a, b = np.load(os.path.join(GOLDEN, name), allow_pickle=False), np.load(os.path.join(tmp_path, name), allow_pickle=False)
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            if a[k].dtype.kind in "US":     # provenance strings (dates, versions) may differ
                continue
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k], equal_nan=True), (name, k)
