"""Cluster-resident tcgen05 pass (dpi_b200/csrc/flow_cl.cu) for the reference's default flow RealNVP(1024, n_flow)
(DPI_interferometry.py:127): against the CPU oracle on the same seeded inputs, and against the fp32 small-batch
kernel it replaces. Tolerance: 1e-4 relative (north_star; the GEMMs run 3xTF32 on the tensor cores, stated in
DESIGN.md section 2 to stay inside the fp32 bar)."""
import pytest
import torch

from conftest import RTOL, rel_err
from test_gpu_flow_small import _check, _run

pytestmark = pytest.mark.gpu

ON = {"DPI_B200_CL": "1"}
OFF = {"DPI_B200_CL": "0"}


def _active(D, NF, B, env):
    import os
    from dpi_b200 import _lib
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        m = RealNVP(D, NF).cuda()
        return int(_lib.load().dpi_flow_cl_active(m._handle(torch.device("cuda", torch.cuda.current_device())), B))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_cluster_path_is_selected_for_the_reference_shape_only():
    assert _active(1024, 2, 64, ON) & 1
    assert _active(1024, 2, 32, ON) & 1
    assert _active(1024, 2, 33, ON) == 0       # batch not a multiple of 4 -> fp32 small-batch kernel
    assert _active(1024, 2, 128, ON) == 0      # tcgen05 GEMM path
    assert _active(256, 2, 64, ON) == 0
    assert _active(1024, 2, 64, OFF) == 0
    assert _active(1024, 2, 64, {"DPI_B200_CL": ""}) == 0      # opt-in: off unless asked for


@pytest.mark.parametrize("NF,B", [(2, 64), (3, 32), (2, 8), (2, 52), (8, 64)])
def test_cluster_pass_vs_oracle(NF, B):
    """Samples, logdet, input gradient, every parameter gradient, BatchNorm running statistics (through the eval-mode
    pass that follows), invertibility - at C1's own layer shapes."""
    _check(_run(1024, NF, B, 4, ON))


def test_cluster_pass_at_c1_size():
    """C1 itself (16 flows, batch 64): samples and logdet against the oracle. The gradients of this randomised 32-stage
    chain are not comparable between ANY two fp32 implementations (a LeakyReLU kink flips: the fp32 small-batch kernel
    differs from the CPU oracle by the same 1.4e-2), so they are compared with the fp32 kernel, which consumes the
    activations this pass saved."""
    from test_gpu_flow_small import _oracle
    a = _run(1024, 16, 64, 4, ON)
    c = _run(1024, 16, 64, 4, OFF)
    sd, zo, xo, ldo, xe, lde = _oracle(a)
    assert rel_err(a["x"], xo) < RTOL and rel_err(a["ld"], ldo) < RTOL
    assert rel_err(a["xe"], xe) < RTOL and rel_err(a["lde"], lde) < RTOL
    assert rel_err(a["gz"], c["gz"]) < RTOL
    for k in a["grads"]:
        assert rel_err(a["grads"][k], c["grads"][k]) < 3 * RTOL, k


def test_cluster_pass_no_batchnorm():
    _check(_run(1024, 2, 32, 4, ON, bn=False), bn=False)


def test_cluster_pass_matches_small_kernel_and_is_deterministic():
    a = _run(1024, 3, 64, 4, ON)
    b = _run(1024, 3, 64, 4, ON)
    c = _run(1024, 3, 64, 4, OFF)
    assert torch.equal(a["x"], b["x"]) and torch.equal(a["ld"], b["ld"]) and torch.equal(a["gz"], b["gz"])
    for k in a["grads"]:
        assert torch.equal(a["grads"][k], b["grads"][k]), k      # fixed-order reductions: bit-reproducible
    assert rel_err(a["x"], c["x"]) < RTOL and rel_err(a["ld"], c["ld"]) < RTOL
    assert rel_err(a["xe"], c["xe"]) < RTOL
    for k in a["bn_after"]:
        assert rel_err(a["bn_after"][k].float(), c["bn_after"][k].float()) < RTOL, k
