"""GPU: the first building block of the Glow path (not built yet as a whole): the coupling-net
convolutions as tcgen05 GEMMs with the 3x3 window gather in the operand pack, against torch's CPU convolution and
against the ZeroConv2d of the Glow oracle. 3xTF32: the fp32 bar (1e-4 relative)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import RTOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,cin,cout,hw,k,pad", [(3, 6, 20, 8, 3, 1.0), (4, 2, 512, 4, 3, 0.0), (4, 512, 512, 4, 1, 0.0),
                                                   (5, 512, 8, 6, 3, 1.0), (2, 3, 7, 5, 3, 0.0)])
def test_conv2d_feature_major_matches_torch(B, cin, cout, hw, k, pad):
    from dpi_b200.generative_model import glow_ops as G
    torch.manual_seed(B * 100 + cin)
    x = torch.randn(B, cin, hw, hw)
    wt = torch.randn(cout, cin, k, k) * 0.1
    bias = torch.randn(cout) * 0.3
    scale = torch.randn(1, cout, 1, 1) * 0.2 if pad == 1.0 else None
    if k == 3:
        ref = F.conv2d(F.pad(x, [1, 1, 1, 1], value=pad), wt, bias)
    else:
        ref = F.conv2d(x, wt, bias)
    if scale is not None:
        ref = ref * torch.exp(scale * 3)
    y = G.conv2d_fm(G.to_feature_major(x.cuda()), wt.cuda(), bias.cuda(), B, hw, hw, pad_value=pad,
                    scale=None if scale is None else scale.cuda())
    out = G.from_feature_major(y, B, hw, hw).cpu()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < RTOL, rel_err(out, ref)


def test_zero_conv_matches_glow_oracle():
    """ZeroConv2d exactly as the Glow oracle (and therefore the reference, glow_model.py:246-251) computes it."""
    from dpi_b200.generative_model import glow_ops as G
    from oracle import glow_oracle as GO
    torch.manual_seed(4)
    B, cin, cout, hw = 6, 512, 4, 8
    sd = {"conv.weight": torch.randn(cout, cin, 3, 3) * 0.05, "conv.bias": torch.randn(cout) * 0.1,
          "scale": torch.randn(1, cout, 1, 1) * 0.3}
    x = torch.randn(B, cin, hw, hw)
    ref = GO.zero_conv(sd, "", x)
    y = G.conv2d_fm(G.to_feature_major(x.cuda()), sd["conv.weight"].cuda(), sd["conv.bias"].cuda(), B, hw, hw, pad_value=1.0,
                    scale=sd["scale"].cuda())
    assert rel_err(G.from_feature_major(y, B, hw, hw).cpu(), ref) < RTOL
