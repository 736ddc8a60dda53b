"""How exact is the tensor pipe's fp32 accumulation?  bf16 x bf16 products are exact in fp32, so the only error of
out_fp32 = x_bf16 @ w_bf16^T is the accumulation.  Prints max / rms error relative to the rms output against a
float64 product of the SAME bf16 values, for our tcgen05 kernel and for cuBLAS (torch.matmul, fp32 output).
Developer aid for the fp32-parity mode (split-bf16 operands need a true fp32 accumulator)."""
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import htrvt_b200  # noqa: F401
    ops = import_module("htr-vt_b200.ops")
    torch.manual_seed(0)
    for K in (768, 3072, 6912):
        for positive in (False, True):
            M, N = 512, 256
            x = torch.randn(M, K, device="cuda")
            w = torch.randn(N, K, device="cuda")
            if positive:
                x, w = x.abs(), w.abs()              # no cancellation: a truncating accumulator shows as a bias
            x, w = x.bfloat16(), w.bfloat16()
            ref = x.double() @ w.double().t()
            out = torch.empty(M, N, device="cuda", dtype=torch.float32)
            ops.gemm_tn(x, w, out)
            lib = torch.matmul(x.float(), w.float().t())          # fp32 SIMT/TF32-off reference
            torch.backends.cuda.matmul.allow_tf32 = False
            scale = ref.abs().max()
            e = (out.double() - ref)
            el = (lib.double() - ref)
            print("K=%5d positive=%d  ours: max %.3e mean-signed %.3e | torch fp32: max %.3e mean-signed %.3e" %
                  (K, positive, float(e.abs().max() / scale), float(e.mean() / scale), float(el.abs().max() / scale),
                   float(el.mean() / scale)))


if __name__ == "__main__":
    main()
