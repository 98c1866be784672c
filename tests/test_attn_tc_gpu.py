"""The tcgen05 / TMEM / TMA attention forward (csrc/vb_attn_tc.cu, opt-in with VB_ATTN_TC=1) against fp32 torch attention on the
same 16-bit inputs, in a subprocess (the switch is read once per process). Shapes: the four of the model, ragged key counts
(27: not a multiple of 16), fp16 and bf16 operands; the backward still runs on the mma.sync kernels."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

_CHILD = r"""
import json, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from _gpu_util import attn_case
out = []
for args in [(4, 12, 36, 36, 64, False), (4, 8, 100, 100, 128, False), (4, 8, 36, 100, 128, True), (4, 8, 100, 36, 128, True),
             (3, 12, 27, 27, 64, False), (3, 8, 101, 27, 128, True), (3, 8, 27, 101, 128, True), (2, 8, 128, 128, 128, False), (5, 8, 17, 2, 64, True)]:
    for fp16 in (False, True):
        errs, _ = attn_case(*args, fp16=fp16)
        out.append([list(args), fp16, errs])
print("RESULT " + json.dumps(out))
"""


def test_tcgen05_attention_forward_matches_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VB_ATTN_TC="1")
    r = subprocess.run([sys.executable, "-c", _CHILD, root], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    for args, fp16, errs in json.loads(line[7:]):
        assert errs["lse"] < 1e-5, (args, fp16, errs)
        assert errs["O"] < (2e-3 if fp16 else 1e-2), (args, fp16, errs)
        assert max(errs.values()) < 3e-2, (args, fp16, errs)      # the backward consumes the tc forward's O and lse
