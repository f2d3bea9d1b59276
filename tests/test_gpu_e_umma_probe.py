"""GPU: pin the tcgen05 shared-memory descriptor semantics the convolution kernel relies on
(K-major, no swizzle, rows 16 B apart, 8-row groups SBO apart, K chunks LBO apart; start addresses
that are only 16-byte aligned; SBO = 144 B) with exact integer-valued bf16 data.  Runs in a
subprocess: a faulting probe must not poison this process' CUDA context."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_umma_descriptor_probes():
    p = subprocess.run([sys.executable, os.path.join(HERE, "oth_umma_probe.py")], capture_output=True, text=True,
                       timeout=300, cwd=os.path.dirname(HERE))
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-2000:]
    bad = [l for l in lines if l.get("rc") != 0 or not l.get("match")]
    assert not bad, bad
    assert len(lines) == 8
