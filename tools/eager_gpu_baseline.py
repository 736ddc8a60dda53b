"""GPU-side baseline (last row): the reference's graph run by torch eager on the B200 with the stock
library kernels (cuDNN convs, cuBLAS GEMMs, ATen LayerNorm / softmax / native CTC) - "the existing Blackwell path".

/root/reference does not travel to the GPU box, so the graph comes from the functional restatement in oracle/ (pinned
to the reference by tests/test_oracle.py).  This is a MEASUREMENT TOOL next to bench.py's cpu_baseline leg: nothing in
the product imports it, and none of the repo's kernels run here.

Usage (GPU box):  python tools/eager_gpu_baseline.py [B]   -> one JSON line per mode
Modes: fp32 (the reference's own arithmetic: cuDNN TF32 convs allowed, fp32 matmuls) and bf16 autocast.
Timed like bench.py: 3 warm-ups, CUDA events around K steps of fwd + CTC loss + bwd (batch B, 1x64x512, C = 80),
plus eval forward + torch.max + the reference-style per-element Python decode loop on a bounded sample.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import bench  # noqa: E402
import htrvt_oracle as O  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    sd0 = O.init_state_dict(bench.NB_CLS, [bench.IMG_H, bench.IMG_W], seed=123)
    img, tg, tl = bench.synth_batch(B, 0)
    img, tg_d, tl_d = img.to(dev), tg.to(dev), tl.to(dev)
    T = bench.IMG_W // 4
    torch.manual_seed(0)
    mask = O.draw_span_mask(T, bench.MASK_RATIO, bench.MAX_SPAN).to(dev)
    in_len = torch.full((B,), T, dtype=torch.int32, device=dev)

    def make_leaves():
        sd = {k: v.to(dev) for k, v in sd0.items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k and k != "pos_embed":
                sd[k] = v.clone().requires_grad_(True)
        return sd

    def train_step(sd, autocast):
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits = O.forward(sd, img, mask=mask, training=True)
        lp = logits.float().permute(1, 0, 2).log_softmax(2)
        prev = torch.backends.cudnn.enabled
        torch.backends.cudnn.enabled = False                     # model_v1/train.py:26
        loss = F.ctc_loss(lp, tg_d, in_len, tl_d, blank=0, reduction="none", zero_infinity=True).mean()
        torch.backends.cudnn.enabled = prev
        loss.backward()
        return loss

    for mode, autocast in (("fp32_tf32conv", False), ("bf16_autocast", True)):
        sd = make_leaves()
        for _ in range(3):
            train_step(sd, autocast)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = train_step(sd, autocast)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # inference: eval forward + torch.max + the reference's per-element decode loop (valid.py:40-42,
        # utils.py:72-86) on a bounded sample of 8 lines (the loop syncs ~3 times per frame)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for _ in range(2):
                O.forward(sd, img, training=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                pe = O.forward(sd, img, training=False)
            e1.record()
            torch.cuda.synchronize()
            ms_fwd = e0.elapsed_time(e1) / steps
        idx = pe.float().permute(1, 0, 2).log_softmax(2).max(2)[1].transpose(1, 0).contiguous()
        nl = 8
        t0 = time.perf_counter()
        for b in range(nl):
            t = idx[b]
            chars = []
            for i in range(T):                                   # CUDA-scalar comparisons, as the reference does
                if t[i] != 0 and (not (i > 0 and t[i - 1] == t[i])) and t[i] < bench.NB_CLS:
                    chars.append(int(t[i]))
        dec_ms_per_line = (time.perf_counter() - t0) * 1e3 / nl
        print(json.dumps({"impl": "torch_eager_gpu", "mode": mode, "batch": B, "train_step_ms": ms,
                          "train_img_per_s": B / ms * 1e3, "eval_forward_ms": ms_fwd,
                          "eval_forward_img_per_s": B / ms_fwd * 1e3,
                          "reference_style_decode_ms_per_line": dec_ms_per_line,
                          "infer_img_per_s_with_reference_decode": B / (ms_fwd + dec_ms_per_line * B) * 1e3,
                          "loss": float(loss), "torch": torch.__version__,
                          "what": "oracle graph on cuda through torch eager (cuDNN / cuBLAS / ATen CTC), none of "
                                  "this repo's kernels"}))
        del sd
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
