"""N-rank NCCL data-parallel step == 1-rank step on the concatenated batch.  Needs >= 2 visible GPUs
(skipped otherwise; `tools/gpu_dp.sh` runs it under `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["3xtf32", "fp32"])
def test_nccl_ranks_equal_single_process_on_the_concatenated_batch(tmp_path, precision):
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n_gpu >= 4 else 2
    n_global = 2048
    out = str(tmp_path / "dp.pt")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_nccl_worker.py"), out, str(n_global),
           precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = torch.load(out)
    assert res["world"] == world and res["same_on_all_ranks"] and res["replicas_identical"]
    # the same step in this process: rank 0's weights, the whole batch, the same noise
    from dxvae_b200 import DXVAE
    from dxvae_b200.dxdata import voices_to_batch
    from dxvae_b200.synth import random_voices
    from dxvae_b200.train import Trainer
    m = DXVAE(); m.verbose = False; m.precision = precision
    m._ensure_flat()
    m._flat.copy_(res["w0"].cuda())
    tr = Trainer(m, lr=1e-3)
    pool = voices_to_batch(random_voices(n_global, seed=3))
    eps = torch.randn(n_global, 128, generator=torch.Generator().manual_seed(5))
    d = m._prepare(pool)
    loss5 = tr.grad_step(d, eps.cuda(), n_global).cpu()
    g = tr.g.cpu()
    for a, b in zip(res["loss5"].tolist(), loss5.tolist()):
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-7, (res["loss5"], loss5)
    for n, p in m.named_parameters():
        lo = (p.data_ptr() - m._flat.data_ptr()) // 4
        a, b = res["g"][lo:lo + p.numel()], g[lo:lo + p.numel()]
        rel = (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)
        assert rel <= 1e-4, (n, rel)
