"""Programmatic dependent launch (csrc/dx_rt.h): every kernel is launched with programmatic stream serialization and waits
(griddepcontrol.wait) before its first global access.  A kernel that touched memory ahead of that wait would race with
the kernel in front; the same steps launched plainly (DX_NO_PDL=1) must give the same results."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, name, env_extra):
    out = str(tmp_path / name)
    env = dict(os.environ); env.pop("DX_NO_PDL", None); env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pdl_worker.py"), out], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


def test_dependent_launch_changes_no_result(tmp_path):
    a = _run(tmp_path, "pdl.pt", {})
    b = _run(tmp_path, "plain.pt", {"DX_NO_PDL": "1"})
    # greedy decode: every product is a single-writer store -> bit-identical
    assert torch.equal(a["adj"], b["adj"]) and torch.equal(a["params"], b["params"])
    for n in (128, 3000):
        # loss terms: forward products never race; the edge-head bias / weight-gradient atomics only feed gradients
        for x, y in zip(a["loss5_%d" % n].tolist(), b["loss5_%d" % n].tolist()):
            assert abs(x - y) <= 1e-6 * abs(y) + 1e-9, (n, x, y)
        ga, gb = a["g_%d" % n], b["g_%d" % n]
        # gradients: equal up to the summation order of the atomic accumulations (fp32 noise), far inside the 1e-4 tolerance
        rel = (ga - gb).abs().max().item() / gb.abs().max().item()
        assert rel <= 2e-6, (n, rel)
