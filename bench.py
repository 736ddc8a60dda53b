#!/usr/bin/env python
"""Benchmark of the HTR-VT hot path on B200 (driver contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our sm_100a path
  python bench.py --impl reference ...                           the reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): one HTR-VT IAM-shape training step = encoder forward (train mode,
span masking 0.4/8) + CTC loss + backward of every trainable parameter, batch 128 per GPU, 1x64x512
line images, 80 classes, bf16 tensor-core math with fp32 masters, driven through the reference's own call
sequence (model_v1/train.py:21-30: model(image, mask_ratio, max_span, use_masking=True) -> .float() ->
permute -> log_softmax -> criterion(...).mean() -> backward).  N > 1: batch sharding, NCCL gradient
all-reduce inside backward ("scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB_CLS, IMG_H, IMG_W = 80, 64, 512
MASK_RATIO, MAX_SPAN = 0.4, 8
# algorithmic forward FLOPs per image: stem 30.558 G + linears 7.263 G + attention 0.201 G
GFLOP_FWD_PER_IMG = 38.02


def synth_batch(B, seed):
    import numpy as np
    import torch
    rs = np.random.RandomState(seed)
    img = torch.from_numpy(rs.rand(B, 1, IMG_H, IMG_W).astype("float32"))
    lens = rs.randint(16, 65, size=B).astype("int32")
    tg = rs.randint(1, NB_CLS, size=int(lens.sum())).astype("int32")
    return img, torch.from_numpy(tg), torch.from_numpy(lens)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="htrvt_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [v.strip() for v in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_rate(steps, warmup, B=8, seed=0):
    """The reference algorithm (CPU oracle port of model_v1 forward/backward + nn.CTCLoss) on the host cores,
    on a bounded sample of the workload: B images per step.  Returns (img/s, seconds per step, threads)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import htrvt_oracle as O
    torch.manual_seed(123)
    sd = O.init_state_dict(NB_CLS, [IMG_H, IMG_W], seed=123)
    img, tg, tl = synth_batch(B, seed)
    times = []
    for i in range(warmup + steps):
        mask = O.draw_span_mask(IMG_W // 4, MASK_RATIO, MAX_SPAN)
        t0 = time.perf_counter()
        O.train_step(sd, img, tg, tl, mask)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    return B / per, per, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 8
    rate, per, threads = cpu_reference_rate(args.steps, max(args.warmup, 1), B=B)
    line = {
        "impl": "reference", "metric": "line images/sec (train step)", "value": rate, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "v1_train_step_b128_64x512_c80", "sample": "%d images per step on the host CPU" % B},
        "cpu_baseline": {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                         "sample": "oracle port of model_v1 fwd+bwd+CTCLoss, %d images/step, %d steps" % (B, args.steps)},
        "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from importlib import import_module
    import __graft_entry__ as ge
    ge.ensure_built()
    import htrvt_b200 as h
    ops = import_module("htr-vt_b200.ops")
    H = import_module("htr-vt_b200.model.HTR_VT")

    B = args.batch
    torch.manual_seed(123)
    model = H.create_model(NB_CLS, [IMG_H, IMG_W]).to(dev).train()
    if world > 1:
        model.enable_data_parallel()
    criterion = h.CTCLoss(reduction="none", zero_infinity=True).to(dev)
    img_h, tg_h, tl_h = synth_batch(B, seed=rank)
    img_h, tg_h, tl_h = img_h.pin_memory(), tg_h.pin_memory(), tl_h.pin_memory()
    img_d, tg_d, tl_d = img_h.to(dev), tg_h.to(dev), tl_h.to(dev)
    params = [p for p in model.parameters() if p.requires_grad]

    def compute_loss(image, text, length):
        # model_v1/train.py:21-30 with the drop-in model / criterion
        preds = model(image, MASK_RATIO, MAX_SPAN, use_masking=True)
        preds = preds.float()
        preds_size = torch.full((B,), preds.size(1), dtype=torch.int32, device=dev)
        preds = preds.permute(1, 0, 2).log_softmax(2)
        return criterion(preds, text, preds_size, length).mean()

    def step_resident():
        for p in params:
            p.grad = None
        loss = compute_loss(img_d, tg_d, tl_d)
        loss.backward()
        return loss

    def step_e2e():
        for p in params:
            p.grad = None
        image = img_h.to(dev, non_blocking=True)
        text = tg_h.to(dev, non_blocking=True)
        length = tl_h.to(dev, non_blocking=True)
        loss = compute_loss(image, text, length)
        loss.backward()
        return loss.item()                       # D2H read of the step's result, every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = ops.launch_count()
    total_ms = timed(step_resident, args.steps)
    launches = ops.launch_count() - n0
    clk = clocks.stop() if rank == 0 else None
    step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel family (tcgen05 tap-GEMM), one extra profiled step -----------
    ops.PROFILE = []
    step_resident()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    by = {}
    for name, fl, a, b in prof:
        d = by.setdefault(name, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += fl
        d[2] += 1
    gemm_names = ("gemm_tn", "gemm_nn", "linear_wgrad", "conv_fwd", "conv_dgrad", "conv_wgrad")
    g_ms = sum(by[n][0] for n in gemm_names if n in by)
    g_fl = sum(by[n][1] for n in gemm_names if n in by)
    g_n = sum(by[n][2] for n in gemm_names if n in by)
    all_ms = sum(v[0] for v in by.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    for cand in (os.path.join(ROOT, "MEASURED_PEAKS.json"),):
        if os.path.exists(cand):
            peaks = json.load(open(cand))
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md sustained)"
    ach_tf = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    ms_step = total_ms / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)
    h2d = img_h.numel() * 4 + tg_h.numel() * 4 + tl_h.numel() * 4
    line = {
        "metric": "line images/sec (train step)", "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "v1_train_step_b128_64x512_c80", "batch_per_gpu": B, "img": [1, IMG_H, IMG_W],
                   "nb_cls": NB_CLS, "mask_ratio": MASK_RATIO, "max_span": MAX_SPAN,
                   "parallelism": "dp%d" % world,
                   "l2": "per-step working set (activations ~7 GB) >> 126 MB L2; no flush needed",
                   "algorithmic_tflop_per_step": 3 * GFLOP_FWD_PER_IMG * B / 1e3},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": ach_tf / peak_tf if peak_tf else None, "traffic": None,
                     "kernel": "tapgemm_kernel (all conv / linear fwd, dgrad, wgrad launches of one step)",
                     "launches": g_n, "ms_in_step": g_ms, "share_of_step": g_ms / all_ms if all_ms else None,
                     "peak_source": peak_src},
        "breakdown_ms": {k: round(v[0], 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, per, threads = cpu_reference_rate(2, 1, B=8)
        line["cpu_baseline"] = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                                "sample": "oracle port of model_v1 fwd+bwd+CTCLoss on 8 images/step, 2 timed steps"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="images per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
