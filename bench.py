#!/usr/bin/env python
"""Benchmark of the DX-VAE hot path on B200 (contract: one JSON line on rank 0).

Workload (BASELINE.json config 5, the data-parallel one, at N GPUs; per-GPU work fixed):
  one "step" = one ELBO training step on a micro-batch of synthetic 6-operator patch graphs:
  batcher (pack + level schedule) -> fused encode + teacher-forced loss + hand-written backward
  -> NCCL all-reduce of the flat gradient (N>1) -> AdamW.  metric = patches/sec.

  value      device-timed, inputs (graph-format tensors) already resident in HBM
  e2e        same step driven from pinned HOST buffers through the public API: H2D of the
             batch inside the timed region, D2H of the 5 loss terms
  roofline   the dominant GEMM kernel class (tcgen05 TF32 by default), per-launch CUDA events
  --precision tf32 (default): dense products on the tcgen05 tensor cores, looser stated tolerance
              fp32: FFMA kernels, reference-tolerance parity (also reported in extra)
  cpu_baseline  the oracle port of the reference (torch CPU, all host cores) on a bounded sample
  extra      the other BASELINE configs, N=1 only: batched encode (cfg3) and greedy decode (cfg4) at micro-batch size and
             end to end at 1 M patches from host buffers (voices -> latents; z -> .syx bytes), decode also at Dexed edge
             density, the B=128 training step (cfg2) and the FP32 FFMA training path
`--impl reference` times that CPU port alone (the reference itself is Python+DGL and cannot
travel to the GPU box).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "patches/sec ELBO train step"
UNIT = "patches/s"
F_TRAIN = 789.67e6        # algorithmic FLOP per patch, fwd+bwd (SURVEY §8d / BASELINE.md §3)
F_ENC, F_DEC = 28.46e6, 234.76e6


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tensor=p["bf16_tflops_sustained"], tensor_burst=p["bf16_tflops"],
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_train_sample(n_patches, steps, warmup, seed=0):
    """The oracle port of the reference's train step (encode + loss + backward + AdamW) on the
    host cores.  Returns (patches_per_s, seconds_per_step, threads)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dxvae_oracle as O
    from dxvae_b200.synth import random_voices
    from dxvae_b200.algo import DX_ALGO
    torch.set_num_threads(os.cpu_count())
    v = random_voices(n_patches, seed)
    Xs, Ps, A = [], [], torch.zeros(n_patches, 7, 7)
    for i in range(n_patches):
        X, P, s, d = O.make_graph(v[i])
        Xs.append(X); Ps.append(P)
        A[i, s, d] = 1.0
    X, P = torch.stack(Xs), torch.stack(Ps)
    torch.manual_seed(0)
    o = O.OracleDXVAE()
    opt = torch.optim.AdamW(o.parameters(), lr=1e-3)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        eps = torch.randn(n_patches, 128)
        opt.zero_grad()
        mu, sd = o.encode(X, A)
        loss = o.loss(mu, sd, X, P, A, eps)[0]
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_patches / sec, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_patches
    pps, sec, thr = cpu_train_sample(n, args.steps, args.warmup)
    sample = "%d synthetic patches per step (oracle port of model.py:374-386 on torch CPU; the reference's Python/DGL " \
             "loops are vectorised in the port, so this is an upper bound on the reference's own speed)" % n
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg5: ELBO train step on synthetic patch graphs (CPU sample)", "micro_batch": n},
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dxvae_b200 import DXVAE, _lib
    from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch
    from dxvae_b200.synth import random_voices
    from dxvae_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if "DX_NCCL_DEBUG" in os.environ:                 # default: NCCL silent, stdout is the one JSON line
            os.environ["NCCL_DEBUG"] = os.environ["DX_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=120))
    L = _lib.require_cuda()
    M = args.micro_batch
    K, W = args.steps, args.warmup
    NPOOL = 4

    torch.manual_seed(0)
    model = DXVAE()
    model.verbose = False
    model.precision = args.precision
    model._ensure_flat()
    tr = Trainer(model, lr=1e-3, w=(2.0, 5.0, 0.01))
    voices = random_voices(NPOOL * M, seed=1000 + rank)
    pool = voices_to_batch(voices)                      # graph-format tensors resident in HBM
    host = pool.cpu()
    hX = host.X.pin_memory(); hP = host.params.pin_memory(); hA = host.adj.pin_memory()

    def device_step(i):
        lo = (i % NPOOL) * M
        sub = DXGraphBatch(pool.X[lo:lo + M], pool.params[lo:lo + M], pool.adj[lo:lo + M])
        d = model._prepare(sub)                          # batcher: pack + device level schedule
        eps = torch.empty(M, 128, device="cuda").normal_()
        loss5 = tr.grad_step(d, eps, M * world)          # fused fwd+bwd (+ NCCL all-reduce)
        tr.apply()                                       # AdamW
        return loss5

    def host_step(i):
        lo = (i % NPOOL) * M
        sub = DXGraphBatch(hX[lo:lo + M].to("cuda", non_blocking=True), hP[lo:lo + M].to("cuda", non_blocking=True),
                           hA[lo:lo + M].to("cuda", non_blocking=True))
        d = model._prepare(sub)
        eps = torch.empty(M, 128, device="cuda").normal_()
        loss5 = tr.grad_step(d, eps, M * world)
        tr.apply()
        return loss5.cpu()                               # D2H of the step's result

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for i in range(steps):
            out = fn(i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), _lib.launch_count() - n0, out

    for i in range(W):
        device_step(i)
    with ClockSampler(local) as cs:
        ms, launches, last = timed(device_step, K)
    clocks = cs.summary()
    value = M * world * K / (ms * 1e-3)

    for i in range(max(1, W // 2)):
        host_step(i)
    ms_e2e, _, _ = timed(host_step, K)
    e2e = M * world * K / (ms_e2e * 1e-3)
    h2d = M * (7 * 27 * 4 + 7 * 21 * 4 + 8)

    # ---- roofline of the dominant kernel family: per-launch events on the launching stream
    # (every rank runs the pass so the collectives inside the step stay matched; rank 0 reports)
    roof = None
    extra = {}
    pk = peaks()
    L.dxvae_prof_begin(4096 * max(1, K))
    for i in range(K):
        device_step(i)
    msv = (ctypes.c_double * 3)(); flv = (ctypes.c_double * 3)(); nv = (ctypes.c_longlong * 3)()
    L.dxvae_prof_end(msv, flv, nv)
    names = ["dx::k_gemm<128,128,8,8> (fp32 FFMA GEMM family)", "dx::k_gemm<64,64,4,4> (fp32 FFMA, small tiles)",
             "dx::k_tc_gemm (tcgen05 kind::tf32, TMA-fed, TMEM accumulators)"]
    dom = max(range(3), key=lambda c: msv[c])
    if msv[dom] > 0:
        ach = flv[dom] / (msv[dom] * 1e-3) / 1e12
        ffma_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12
        traffic = None
        try:    # DRAM bytes per launch of this kernel family from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "tc_gemm_traffic.json")))
            if dom == 2:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": names[dom], "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor"], "traffic": traffic,
                "traffic_source": "mean dram__bytes_read+write of the launches in the committed ncu --set full capture "
                                  "(profiles/r01b_tc_gemm_ncu.txt); `achieved` averages every launch of the timed steps",
                "peak_source": pk["source"] + ", bf16 sustained",
                "launches_per_step": nv[dom] / K, "avg_launch_ms": msv[dom] / max(1, nv[dom]),
                "share_of_step": msv[dom] / K / (ms / K), "fp32_ffma_peak_tflops": ffma_peak,
                "classes": [{"kernel": names[c], "ms_per_step": msv[c] / K, "tflops": (flv[c] / (msv[c] * 1e-3) / 1e12)
                             if msv[c] > 0 else None, "launches_per_step": nv[c] / K} for c in range(3)],
                "step_algorithmic_tflops": value / world * F_TRAIN / 1e12}
    if world > 1:
        dist.barrier()

    # ---- the other configs, briefly (N=1 only): cfg3 encode, cfg4 decode, cfg2 B=128 train
    if rank == 0 and world == 1 and not args.no_extra:
        with torch.no_grad():
            ne = min(NPOOL * M, 32768)
            gb = DXGraphBatch(pool.X[:ne], pool.params[:ne], pool.adj[:ne])
            model.encode(gb); torch.cuda.synchronize()
            t0 = time.perf_counter(); model.encode(gb); torch.cuda.synchronize()
            extra["encode_patches_per_s"] = ne / (time.perf_counter() - t0)
            model.encode_precision = "tf32"
            model.encode(gb); torch.cuda.synchronize()
            t0 = time.perf_counter(); model.encode(gb); torch.cuda.synchronize()
            extra["encode_tf32_patches_per_s"] = ne / (time.perf_counter() - t0)
            model.encode_precision = "3xtf32"
            model.encode(gb); torch.cuda.synchronize()
            t0 = time.perf_counter(); model.encode(gb); torch.cuda.synchronize()
            extra["encode_3xtf32_patches_per_s"] = ne / (time.perf_counter() - t0)
            model.encode_precision = "fp32"
            z = torch.randn(16384, 128, device="cuda")
            gdec = model.decode(z); torch.cuda.synchronize()
            t0 = time.perf_counter(); model.decode(z); torch.cuda.synchronize()
            extra["decode_patches_per_s"] = 16384 / (time.perf_counter() - t0)
            # greedy decode re-propagates only the graphs that gain an edge at a step: its speed depends on how many
            # edges the (here randomly initialised) model decides, so report that next to the number
            am = gdec.adj.cpu().numpy().view(np.uint64)
            extra["decode_mean_edges_per_graph"] = float(np.mean([bin(int(a)).count("1") for a in am[:4096]]))
            model.decode_precision = "3xtf32"
            model.decode(z); torch.cuda.synchronize()
            t0 = time.perf_counter(); model.decode(z); torch.cuda.synchronize()
            extra["decode_3xtf32_patches_per_s"] = 16384 / (time.perf_counter() - t0)
            model.decode_precision = "fp32"
            # the same decode with the edge-head biases shifted until the model decides as many edges as a Dexed patch
            # has (DX_ALGO: 7.3 on average); there is no trained checkpoint in the reference tree to take them from
            named = dict(model.named_parameters())
            eb, sb = named["h_to_edge.2.bias"], named["h_to_edge_self.2.bias"]
            eb0, sb0 = eb.data.clone(), sb.data.clone()
            lo_b, hi_b = -2.0, 2.0
            for _ in range(10):
                mid = 0.5 * (lo_b + hi_b)
                eb.data.copy_(eb0 + mid); sb.data.copy_(sb0 + mid)
                am = model.decode(z[:2048]).adj.cpu().numpy().view(np.uint64)
                me = float(np.mean([bin(int(a)).count("1") for a in am]))
                lo_b, hi_b = (lo_b, mid) if me > 7.3 else (mid, hi_b)
            model.decode(z); torch.cuda.synchronize()
            t0 = time.perf_counter(); gd2 = model.decode(z); torch.cuda.synchronize()
            extra["decode_dexed_density_patches_per_s"] = 16384 / (time.perf_counter() - t0)
            am = gd2.adj.cpu().numpy().view(np.uint64)
            extra["decode_dexed_density_mean_edges_per_graph"] = float(np.mean([bin(int(a)).count("1") for a in am[:4096]]))
            model.decode_precision = "3xtf32"
            model.decode(z); torch.cuda.synchronize()
            t0 = time.perf_counter(); model.decode(z); torch.cuda.synchronize()
            extra["decode_dexed_density_3xtf32_patches_per_s"] = 16384 / (time.perf_counter() - t0)
            model.decode_precision = "fp32"
            eb.data.copy_(eb0); sb.data.copy_(sb0)
        # cfg3 / cfg4 at their full size (1 M patches), end to end from HOST buffers: packed voices (128 B/patch) ->
        # on-device _make_graph -> encode -> latents on the host;  z on the host -> greedy decode -> .syx bytes on the host
        from dxvae_b200.dxdata import graph_to_syx_bytes
        nfull = args.full_patches
        if nfull > 0:
            hv = torch.from_numpy(random_voices(nfull, seed=7)).pin_memory()
            hz = torch.randn(nfull, 128, generator=torch.Generator().manual_seed(0)).pin_memory()
            mu_h = torch.empty(nfull, 128).pin_memory(); sd_h = torch.empty(nfull, 128).pin_memory()
            with torch.no_grad():
                for prec in ("fp32", "3xtf32"):
                    model.encode_precision = prec
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    q = model.encode(voices_to_batch(hv))
                    mu_h.copy_(q.loc, non_blocking=True); sd_h.copy_(q.scale, non_blocking=True)
                    torch.cuda.synchronize()
                    extra["cfg3_encode_%d_e2e_%s_patches_per_s" % (nfull, prec)] = nfull / (time.perf_counter() - t0)
                model.encode_precision = "fp32"
                for prec in ("fp32", "3xtf32"):
                    model.decode_precision = prec
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    syx = graph_to_syx_bytes(model.decode(hz))
                    extra["cfg4_decode_%d_to_syx_e2e_%s_patches_per_s" % (nfull, prec)] = nfull / (time.perf_counter() - t0)
                    assert len(syx) == 8 + 128 * nfull
                model.decode_precision = "fp32"
            del hv, hz, q, mu_h, sd_h, syx
        idx = list(range(128))
        for _ in range(3):
            tr.step(pool, idx)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            tr.step(pool, idx)
        torch.cuda.synchronize()
        extra["train_b128_patches_per_s"] = 1280 / (time.perf_counter() - t0)
        if args.precision != "fp32":        # the FP32 FFMA path (reference-tolerance parity) on the same workload
            model.precision = "fp32"
            device_step(0); torch.cuda.synchronize(); t0 = time.perf_counter()
            for i in range(2):
                device_step(i)
            torch.cuda.synchronize()
            extra["fp32_path_patches_per_s"] = 2 * M / (time.perf_counter() - t0)
            model.precision = args.precision

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        pps, sec, thr = cpu_train_sample(args.cpu_patches, 3, 1)
        cpu = {"value": pps, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": "%d synthetic patches x 3 train steps (+1 warm-up), oracle port on torch CPU" % args.cpu_patches}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.precision == "fp32" else "tf32 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": "cfg5 (the config the metric's 1/2/4/8-GPU patches/sec is quoted on), one GPU's share: "
                                       "data-parallel ELBO train step on synthetic 6-operator patch graphs; cfg2/3/4 "
                                       "figures are in `extra`",
                           "precision": args.precision,
                           "micro_batch_per_gpu": M, "global_batch": M * world, "parallelism": "dp%d" % world,
                           "optimizer": "AdamW lr=1e-3", "l2": "inputs cycle over a %d-graph pool; the step's %.1f GB "
                           "activation workspace is far larger than L2" %
                           (NPOOL * M, L.dxvae_workspace_bytes(2, M) / 1e9)},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 20,
                        "ms_per_step": ms_e2e / K},
                "roofline": roof, "cpu_baseline": cpu, "loss": float(last[0]), "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--micro-batch", type=int, default=32768)
    ap.add_argument("--cpu-patches", type=int, default=2048)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--full-patches", type=int, default=1 << 20, help="size of the cfg3/cfg4 end-to-end extras (0 skips them)")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
