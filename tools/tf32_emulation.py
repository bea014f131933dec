"""CPU emulation of the 3xTF32 GEMM arithmetic (cvt.rna.tf32 hi / lo split, a_lo b_hi + a_hi b_lo + a_hi b_hi with exact
products and fp32-or-better accumulation) inside the oracle's Linear layers, forward and backward, to separate the
representation error of the split from anything else the GPU kernels do (tools/ only; DESIGN.md accuracy note)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from conftest import rel_err  # noqa: E402
from oracle import dpi_oracle as O  # noqa: E402


def tf32_rna(x: torch.Tensor) -> torch.Tensor:
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def split(x):
    hi = tf32_rna(x)
    lo = tf32_rna(x - hi)
    return hi.double(), lo.double()


def mm3(a, b):
    """a [M, K] @ b [K, N] with the three-term split; products and sums in fp64 (better than the tensor core's fp32
    accumulator, so only the representation error remains), result rounded to fp32."""
    ah, al = split(a)
    bh, bl = split(b)
    return (al @ bh + ah @ bl + ah @ bh).float()


def rz32(t64: torch.Tensor) -> torch.Tensor:
    """fp64 -> fp32 rounding toward zero."""
    f = t64.float()
    over = f.double().abs() > t64.abs()
    return torch.where(over, torch.nextafter(f, torch.zeros_like(f)), f)


def mm3_rz(a, b):
    """Same three terms, but accumulated term by term in an fp32 accumulator that rounds toward zero after every
    addition (an upper bound on what a truncating tensor-core accumulator would do; products of TF32 values are
    exact in fp32)."""
    ah, al = split(a)
    bh, bl = split(b)
    acc = torch.zeros(a.shape[0], b.shape[1], dtype=torch.float32)
    for k in range(a.shape[1]):
        for x, y in ((al, bh), (ah, bl), (ah, bh)):
            acc = rz32(acc.double() + x[:, k:k + 1] * y[k:k + 1, :])
    return acc


MM = mm3


class Linear3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias):
        ctx.save_for_backward(x, w)
        return MM(x, w.t().contiguous()) + bias

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        return MM(g, w), MM(g.t().contiguous(), x), g.sum(0)


def main():
    D, NF, B, seqfrac = 2, 16, 512, 1 / 128
    sd = O.init_state(D, NF, seqfrac=seqfrac, seed=31, randomize=True)
    torch.manual_seed(8)
    z = torch.randn(B, D)

    def run(dtype, emulate):
        s = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        for k in O.param_keys(s):
            s[k].requires_grad_(True)
        orig = F.linear
        if emulate:
            F.linear = lambda x, w, b=None: Linear3.apply(x, w, b)
        try:
            loss, aux = O.toy_loss(s, z.to(dtype), kind=1, diversity=1.0, train=True)
            loss.backward()
        finally:
            F.linear = orig
        return s

    s64 = run(torch.float64, False)
    s32 = run(torch.float32, False)
    global MM
    s3x = run(torch.float32, True)
    MM = mm3_rz
    s3z = run(torch.float32, True)
    for name, s in (("CPU fp32", s32), ("CPU fp32 + emulated 3xTF32 Linear layers (exact accumulation)", s3x),
                    ("CPU fp32 + emulated 3xTF32, fp32 accumulator rounding toward zero", s3z)):
        errs = sorted(((rel_err(s[k].grad, s64[k].grad), k) for k in O.param_keys(sd)), reverse=True)
        print(name, "vs fp64:")
        for e, k in errs[:4]:
            print("   %.3e  %s" % (e, k))
        for k in ("blocks.15.coupling2.net.0.bias", "blocks.15.actnorm2.loc", "blocks.0.coupling2.net.0.bias"):
            print("   %.3e  %s" % (rel_err(s[k].grad, s64[k].grad), k))


if __name__ == "__main__":
    main()
