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
    env = dict(os.environ)
    for k in ("DX_NO_PDL", "DX_HEADS_BATCH_MAX", "DX_NO_P1_BATCH", "DX_X3K_ROWS"):
        env.pop(k, None)
    env.update(env_extra)
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


def test_small_batch_schedule_equals_large_batch_schedule(tmp_path):
    """Training steps of <= 4096 graphs run the node-independent work of the teacher-forced decoder once over 6B rows (the
    parameter heads of nodes 1..6, their first propagates and self-loop heads, the weight gradients of what reads the
    finished node states) and let products of up to 1024 rows split their reduction over a cluster (csrc/dx_decoder.cu
    heads_batched / p1_batched, dx_gemm.h few_rows_for_batch).  DX_HEADS_BATCH_MAX=0 sends the same batches down the
    per-node schedule every larger batch (the benchmark's 65536 patches) takes: same loss terms and gradients up to
    summation order, so the oracle parity of the small-batch tests carries over to the large-batch code path."""
    a = _run(tmp_path, "batched.pt", {})
    b = _run(tmp_path, "pernode.pt", {"DX_HEADS_BATCH_MAX": "0"})
    assert torch.equal(a["adj"], b["adj"]) and torch.equal(a["params"], b["params"])     # inference takes neither
    for n in (128, 3000):
        for x, y in zip(a["loss5_%d" % n].tolist(), b["loss5_%d" % n].tolist()):
            assert abs(x - y) <= 2e-6 * abs(y) + 1e-9, (n, x, y)
        ga, gb = a["g_%d" % n], b["g_%d" % n]
        rel = (ga - gb).abs().max().item() / gb.abs().max().item()
        print("small-batch schedule vs per-node schedule, B=%d: max gradient difference %.3g of the largest gradient" % (n, rel))
        assert rel <= 1e-5, (n, rel)
