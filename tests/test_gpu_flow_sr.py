"""Sample-resident small-batch pass (dpi_b200/csrc/flow_sr.cu: one cluster of 16 CTAs, 4 samples each, weights streamed
from L2) for the reference's default flow RealNVP(1024, n_flow) at batch <= 64 (DPI_interferometry.py:127, :60): against
the CPU oracle on the same seeded inputs and against the feature-split fp32 kernel it replaces. Tolerance: 1e-4 relative
(north_star; fp32 FFMA throughout)."""
import os

import pytest
import torch

from conftest import RTOL, rel_err
from test_gpu_flow_small import _check, _oracle, _run

pytestmark = pytest.mark.gpu

ON = {"DPI_B200_SR": "1"}
OFF = {"DPI_B200_SR": "0"}


def _active(D, NF, B, env, seqfrac=4):
    from dpi_b200 import _lib
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        m = RealNVP(D, NF, seqfrac=seqfrac).cuda()
        return int(_lib.load().dpi_flow_sr_active(m._handle(torch.device("cuda", torch.cuda.current_device())), B))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_sample_resident_path_is_selected_by_shape():
    assert _active(1024, 2, 64, {"DPI_B200_SR": ""}) == 0      # opt-in: off unless asked for
    assert _active(1024, 2, 64, ON) == 1
    assert _active(1024, 2, 32, ON) == 1
    assert _active(1024, 2, 33, ON) == 0            # batch not a multiple of 4 -> feature-split kernel
    assert _active(1024, 2, 128, ON) == 0           # tcgen05 GEMM path
    assert _active(4096, 1, 64, ON) == 0            # C2: 13.6 MB of weights per stage, every CTA would stream them all
    assert _active(1024, 2, 64, OFF) == 0


@pytest.mark.parametrize("D,NF,B,seqfrac", [(1024, 2, 64, 4), (1024, 3, 32, 4), (1024, 2, 8, 4), (1024, 2, 52, 4),
                                            (256, 3, 64, 4), (64, 2, 8, 1), (1024, 8, 64, 4)])
def test_sample_resident_pass_vs_oracle(D, NF, B, seqfrac):
    """Samples, logdet, input gradient, every parameter gradient, BatchNorm running statistics (through the eval-mode
    pass that follows), invertibility."""
    _check(_run(D, NF, B, seqfrac, ON))


def test_sample_resident_pass_at_c1_size():
    """C1 itself (16 flows, batch 64): samples and logdet against the oracle. The gradients of this randomised 32-stage
    chain split fp32 implementations into two camps (one LeakyReLU pre-activation sits within rounding of zero and its
    slope flips: the feature-split kernel differs from the CPU oracle by 1.4e-2 on this seed), so they must agree with
    ONE of the two other fp32 implementations to 1e-4 / 3e-4."""
    a = _run(1024, 16, 64, 4, ON)
    c = _run(1024, 16, 64, 4, OFF)
    sd, zo, xo, ldo, xe, lde = _oracle(a)
    assert rel_err(a["x"], xo) < RTOL and rel_err(a["ld"], ldo) < RTOL
    assert rel_err(a["xe"], xe) < RTOL and rel_err(a["lde"], lde) < RTOL
    vs_oracle = rel_err(a["gz"], zo.grad) < RTOL
    assert vs_oracle or rel_err(a["gz"], c["gz"]) < RTOL, (rel_err(a["gz"], zo.grad), rel_err(a["gz"], c["gz"]))
    for k in a["grads"]:
        ref = sd[k].grad if vs_oracle else c["grads"][k]
        assert rel_err(a["grads"][k], ref) < 3 * RTOL, k


def test_sample_resident_pass_no_batchnorm():
    _check(_run(1024, 2, 32, 4, ON, bn=False), bn=False)


def test_sample_resident_pass_matches_feature_split_kernel_and_is_deterministic():
    a = _run(1024, 3, 64, 4, ON)
    b = _run(1024, 3, 64, 4, ON)
    c = _run(1024, 3, 64, 4, OFF)
    assert torch.equal(a["x"], b["x"]) and torch.equal(a["ld"], b["ld"]) and torch.equal(a["gz"], b["gz"])
    for k in a["grads"]:
        assert torch.equal(a["grads"][k], b["grads"][k]), k      # fixed-order reductions: bit-reproducible
    assert rel_err(a["x"], c["x"]) < 1e-5 and rel_err(a["ld"], c["ld"]) < 1e-5
    assert rel_err(a["xe"], c["xe"]) < 1e-5
    for k in a["bn_after"]:
        assert rel_err(a["bn_after"][k].float(), c["bn_after"][k].float()) < 1e-5, k
