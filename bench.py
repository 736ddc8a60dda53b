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
                                          "-lms", "50", "-i", str(self.gpu)], stdout=self.fh,
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
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would
    # quietly turn the multi-threaded CPU arm into a single-threaded one
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
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


def extra_metrics(torch, dev, h, ops, H, model, peaks):
    """The other BASELINE.json metrics / configs, measured on the same box right after the headline run (1 GPU):
    CTC loss+grad us/batch with its HBM roofline fraction, inference (eval forward + greedy decode) img/s at the
    per-GPU share of config 4 (512 lines), and the windowed wide-line variant's training step (config 5)."""
    from importlib import import_module
    out = {}
    hbm = float(peaks.get("hbm_gbs", 6500.0))

    def timed_ms(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # ---- CTC loss + gradient, B=128, T=128, C=80, fused log-softmax (one launch); 16 rotating buffers (168 MB of
    # logits + gradients > 126 MB L2) so no iteration finds its data in L2
    B, T, C = 128, 128, NB_CLS
    g = torch.Generator(device="cpu").manual_seed(0)
    bufs = [torch.randn(B, T, C, generator=g).to(dev) for _ in range(16)]
    _, tg, tl = synth_batch(B, 0)
    tg, tl_d = tg.to(dev), tl.to(dev)
    mtl = int(tl.max())
    state = {"i": 0}

    def ctc_once():
        x = bufs[state["i"] % len(bufs)]
        state["i"] += 1
        ops.ctc_loss_grad(x, tg, None, tl_d, layout="btc", is_logprob=False, want_grad=True, max_target_len=mtl,
                          grad_scale_const=1.0 / B)

    ms = timed_ms(ctc_once, 64)
    bytes_ctc = 2.0 * B * T * C * 4 + float(tg.numel()) * 4 + 12.0 * B
    out["ctc_loss_grad"] = {"us_per_batch": ms * 1e3, "batch": B, "T": T, "C": C, "algorithmic_bytes": bytes_ctc,
                            "achieved_gbs": bytes_ctc / (ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                            "frac_of_hbm_roofline": bytes_ctc / (ms * 1e-3) / 1e9 / hbm,
                            "note": "T sequential alpha/beta steps bound this kernel by latency, not bytes; the "
                                    "recursion runs in the linear domain on the FP64 pipe (fallbacks to log space: %d)"
                                    % ops.lib().htrvt_ctc_fallback_count()}
    del bufs
    # the same kernel when the batch fills the chip (one CTA per sequence, 148 SMs): throughput, not latency
    Bl = 4096
    xl = torch.randn(Bl, T, C, device=dev)
    _, tgl, tll = synth_batch(Bl, 3)
    tgl, tll_d, mtl_l = tgl.to(dev), tll.to(dev), int(tll.max())
    ms = timed_ms(lambda: ops.ctc_loss_grad(xl, tgl, None, tll_d, layout="btc", is_logprob=False, want_grad=True,
                                            max_target_len=mtl_l, grad_scale_const=1.0 / Bl), 10)
    bl = 2.0 * Bl * T * C * 4 + float(tgl.numel()) * 4 + 12.0 * Bl
    out["ctc_loss_grad"]["batch_4096"] = {"us_per_batch": ms * 1e3, "achieved_gbs": bl / (ms * 1e-3) / 1e9,
                                          "frac_of_hbm_roofline": bl / (ms * 1e-3) / 1e9 / hbm}
    del xl

    # ---- inference: eval forward + greedy decode (argmax + collapse on device, one D2H copy), 512 lines per GPU
    Bi = 512
    conv = h.CTCLabelConverter("".join(chr(33 + i) for i in range(NB_CLS - 1)))
    img = synth_batch(Bi, 1)[0].to(dev)
    model.eval()

    def infer_once():
        with torch.no_grad():
            preds = model(img)
            return conv.decode_logits(preds.float())

    ms = timed_ms(infer_once, 5, warm=2)
    out["inference"] = {"img_per_s": Bi / (ms * 1e-3), "ms_per_batch": ms, "batch_per_gpu": Bi,
                        "what": "eval-mode forward + greedy CTC decode to Python strings (BASELINE config 4 share)"}
    # validation metrics on device: Levenshtein of the decoded ids against the label ids
    try:
        with torch.no_grad():
            preds = model(img).float()
        ids, lens = h.greedy_decode(preds, NB_CLS)
        _, tgv, tlv = synth_batch(Bi, 4)
        tgv = tgv.to(dev)
        ms_ed = timed_ms(lambda: h.cer_from_ids(ids, lens, tgv, tlv), 20)
        out["cer_device"] = {"us_per_batch": ms_ed * 1e3, "pairs": Bi,
                             "what": "edit distances of 512 decoded id rows vs label ids, one launch (valid.py:49-55)"}
    except Exception as e:
        out["cer_device"] = {"error": repr(e)[:200]}
    # ---- LM-rescoring evaluation (model_window/test_with_kenlm.py:25-59): K-best paths of 512 lines in one launch +
    # one D2H copy + the host-side string building / scorer calls, against the reference's per-frame Python beam
    try:
        class _LenLM(object):
            def score(self, text):
                return -0.1 * len(text)

        lp = torch.randn(IMG_W // 4, Bi, NB_CLS, device=dev).log_softmax(2)
        ms_k = timed_ms(lambda: ops.ctc_kbest_paths(lp, 5), 20)
        ms_b = timed_ms(lambda: h.beam_search_with_lm_batch(lp, conv, _LenLM(), beam_size=5), 3, warm=1)
        out["kbest_beam"] = {"kernel_us_per_batch": ms_k * 1e3, "with_host_strings_ms_per_batch": ms_b, "lines": Bi,
                             "beam": 5, "what": "K-best CTC paths on device + LM pick on host (test_with_kenlm.py:25-59)"}
        del lp
    except Exception as e:
        out["kbest_beam"] = {"error": repr(e)[:200]}
    model.train()
    del img

    # ---- training step fed with the loader's uint8 lines: 4x fewer PCIe bytes, conversion +
    # padding + input LayerNorm in one kernel; same timed region as the headline e2e (H2D + step + loss.item())
    try:
        import numpy as np
        Bu = 128
        u8_h = torch.from_numpy(np.random.RandomState(6).randint(0, 256, size=(Bu, 1, IMG_H, IMG_W)).astype("uint8")).pin_memory()
        _, tgu, tlu = synth_batch(Bu, 6)
        tgu, tlu = tgu.pin_memory(), tlu.pin_memory()
        uparams = [p for p in model.parameters() if p.requires_grad]

        def step_u8():
            for p in uparams:
                p.grad = None
            preds = model(u8_h.to(dev, non_blocking=True), MASK_RATIO, MAX_SPAN, use_masking=True)
            loss = h.ctc_loss_from_logits(preds.float(), tgu.to(dev, non_blocking=True), tlu).mean()
            loss.backward()
            return loss.item()

        ms = timed_ms(step_u8, 5, warm=2)
        out["e2e_uint8_input"] = {"img_per_s": Bu / (ms * 1e-3), "ms_per_step": ms,
                                  "h2d_bytes_per_step": u8_h.numel() + tgu.numel() * 4,
                                  "what": "train step from pinned uint8 line images through model.forward(uint8)"}
        for p in uparams:
            p.grad = None
    except Exception as e:
        out["e2e_uint8_input"] = {"error": repr(e)[:200]}

    # ---- the reference's full training iteration (model_v1/train.py:113-126): SAM(AdamW) = two forward/backward
    # passes around first_step / second_step + the EMA update, with the multi-tensor optimizer kernels
    try:
        SAMm = import_module("htr-vt_b200.utils.sam")
        Um = import_module("htr-vt_b200.utils.utils")
        Bs = 128
        imgs, tgs, tls = [t.to(dev) for t in synth_batch(Bs, 2)]
        opt = SAMm.SAM(model.parameters(), torch.optim.AdamW, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.5)
        ema = Um.ModelEma(model, 0.9999)
        crit = h.CTCLoss(reduction="none", zero_infinity=True)
        ps = torch.full((Bs,), IMG_W // 4, dtype=torch.int32, device=dev)

        def fb():
            preds = model(imgs, MASK_RATIO, MAX_SPAN, use_masking=True).float()
            loss = crit(preds.permute(1, 0, 2).log_softmax(2), tgs, ps, tls).mean()
            loss.backward()

        def sam_iter():
            opt.zero_grad()
            fb()
            opt.first_step(zero_grad=True)
            fb()
            opt.second_step(zero_grad=True)
            model.zero_grad()
            ema.update(model, num_updates=10)

        def opt_only():
            for p in model.parameters():
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)
            opt.first_step(zero_grad=False)
            opt.second_step(zero_grad=False)
            ema.update(model, num_updates=10)

        ms = timed_ms(sam_iter, 5, warm=2)
        ms_opt = timed_ms(opt_only, 5, warm=1)
        nparam = sum(p.numel() for p in model.parameters() if p.requires_grad)
        opt_bytes = nparam * (4 + 16 + 36 + 12)            # norm + climb + restore/AdamW + EMA, fp32 streams
        out["sam_iteration"] = {"img_per_s": Bs / (ms * 1e-3), "ms_per_iteration": ms, "batch_per_gpu": Bs,
                                "optimizer_ms": ms_opt, "optimizer_algorithmic_bytes": opt_bytes,
                                "optimizer_gbs": opt_bytes / (ms_opt * 1e-3) / 1e9,
                                "what": "2x(fwd+bwd+CTC) + SAM first/second step (fused AdamW) + EMA, train.py:113-126"}
        del opt, ema
        for p in model.parameters():
            p.grad = None
    except Exception as e:
        out["sam_iteration"] = {"error": repr(e)[:200]}

    # ---- windowed variant (model_window), 64x1024 lines, T = 256, 90 classes, labels up to 200: training step with
    # the reference's train-mode regularisers on (dropout 0.1 / attention dropout 0.05 / DropPath <= 0.1)
    try:
        Wm = import_module("htr-vt_b200.model_window.HTR_VT")
        import numpy as np
        Bw, Ww, Cw = 128, 1024, 90
        torch.manual_seed(123)
        wm = Wm.create_model(Cw, [IMG_H, Ww]).to(dev).train()
        rs = np.random.RandomState(5)
        imgw = torch.from_numpy(rs.rand(Bw, 1, IMG_H, Ww).astype("float32")).to(dev)
        lens = rs.randint(64, 201, size=Bw).astype("int32")
        tgw = torch.from_numpy(rs.randint(1, Cw, size=int(lens.sum())).astype("int32")).to(dev)
        tlw = torch.from_numpy(lens)
        wparams = [p for p in wm.parameters() if p.requires_grad]

        def win_step():
            for p in wparams:
                p.grad = None
            preds = wm(imgw, MASK_RATIO, MAX_SPAN, use_masking=True)
            loss = h.ctc_loss_from_logits(preds.float(), tgw, tlw).mean()
            loss.backward()

        ms = timed_ms(win_step, 3, warm=2)
        out["window_train_step"] = {"img_per_s": Bw / (ms * 1e-3), "ms_per_step": ms, "batch_per_gpu": Bw,
                                    "img": [1, IMG_H, Ww], "nb_cls": Cw, "T": Ww // 4,
                                    "what": "model_window fwd+bwd+CTC (BASELINE config 5)"}
        del wm, imgw
    except Exception as e:                                   # never lose the headline line to an extra
        out["window_train_step"] = {"error": repr(e)[:200]}
    torch.cuda.empty_cache()
    return out


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
        # NCCL writes its version banner / debug lines to stdout while the communicator comes up; rank 0's stdout must
        # carry ONE JSON line, so file descriptor 1 points at stderr until the first collective has completed
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
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
    gemm_names = ("gemm_tn", "gemm_nn", "linear_wgrad", "conv_fwd", "conv_dgrad", "conv_wgrad", "conv_wgrad_acc",
                  "conv_wgrad_acc_t", "conv_wgrad_acc_w")
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
    traffic = None          # DRAM bytes of the tap-GEMM launches of one step, from the committed ncu counters (profiles/)
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["tapgemm_dram_bytes_per_step"]
    except Exception:
        pass
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
                     "frac": ach_tf / peak_tf if peak_tf else None, "traffic": traffic,
                     "traffic_note": "ncu dram bytes (read + write) summed over the family's launches of one step; "
                                     "algorithmic = flops, the operands are re-read from L2",
                     "kernel": "tapgemm_kernel (all conv / linear fwd, dgrad, wgrad launches of one step)",
                     "launches": g_n, "ms_in_step": g_ms, "share_of_step": g_ms / all_ms if all_ms else None,
                     "peak_source": peak_src},
        "breakdown_ms": {k: round(v[0], 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])},
    }
    if world == 1 and not args.no_extras:
        line["extra"] = extra_metrics(torch, dev, h, ops, H, model, peaks)
    if world == 1 and not args.no_cpu_baseline:
        rate, per, threads = cpu_reference_rate(20, 1, B=8)          # ~10 s of host work
        line["cpu_baseline"] = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                                "sample": "oracle port of model_v1 fwd+bwd+CTCLoss on 8 images/step, 20 timed steps"}
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
    ap.add_argument("--no-extras", action="store_true", help="skip the CTC / inference / window-variant side metrics")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
