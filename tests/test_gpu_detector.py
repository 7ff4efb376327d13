"""-m gpu: detector kernels (tcgen05 conv + bias + SiLU ...) against a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands.  Tolerance: bf16 output rounding (2^-8 relative) + fp32
accumulation order -> rtol 1.5e-2 / atol 1.5e-2 on O(1) activations."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ref_conv(x_nhwc, w, b, stride, act):
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = x_nhwc.float().permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(x, w.to(torch.bfloat16).float(), b, stride=stride, padding=w.shape[-1] // 2)
    if act:
        y = y * torch.sigmoid(y)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # n, h, w, cin, cout, k, s, in_pitch_extra, in_coff, out_pitch_extra, out_coff, act, f32
    (2, 32, 32, 64, 64, 1, 1, 0, 0, 0, 0, True, False),
    (2, 32, 32, 64, 64, 3, 1, 0, 0, 0, 0, True, False),
    (1, 64, 64, 64, 128, 3, 2, 0, 0, 0, 0, True, False),
    (2, 40, 40, 128, 192, 3, 1, 64, 64, 128, 64, True, False),     # channel slices of concat buffers
    (1, 64, 64, 16, 64, 3, 1, 0, 0, 0, 0, True, False),            # stem: BK = 16
    (2, 20, 20, 512, 255, 1, 1, 0, 0, 1, 0, False, True),          # detect head: linear, fp32, 255 -> 256 rows
    (2, 20, 20, 256, 256, 3, 1, 0, 0, 0, 0, True, False),          # 20x20 map: TW = 4 tiles, partial tiles
    (2, 80, 80, 256, 256, 3, 1, 0, 0, 0, 0, True, False),
    (1, 80, 80, 512, 768, 3, 2, 0, 0, 0, 0, True, False),          # stride 2, 3 N-tiles of 256
    (1, 40, 40, 1536, 384, 1, 1, 0, 0, 0, 0, True, False),         # long K (24 chunks), BN 192 x 2
]


@pytest.mark.parametrize("case", CASES)
def test_conv_bias_silu_vs_torch(case):
    from b200track.conv import ConvPlan, pack_conv_weight
    n, h, w, cin, cout, k, s, ipx, icoff, opx, ocoff, act, f32 = case
    g = torch.Generator(device="cuda").manual_seed(hash(case) % (2 ** 31))
    dev = "cuda"
    in_pitch = cin + ipx + (icoff if ipx == 0 else 0)
    xbuf = (torch.randn((n, h, w, in_pitch), device=dev, generator=g) * 1.0).to(torch.bfloat16)
    wt = torch.randn((cout, cin, k, k), device=dev, generator=g) * (1.5 / (cin * k * k) ** 0.5)
    bias = torch.randn(cout, device=dev, generator=g) * 0.5
    ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    out_pitch = (cout + 7) // 8 * 8 + opx
    ybuf = torch.full((n, ho, wo, out_pitch), -77.0, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    plan = ConvPlan(xbuf, pack_conv_weight(wt), bias.contiguous(), ybuf, n, h, w, cin, icoff, cout, k, s, ocoff, act=act, out_f32=f32)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv(xbuf[..., icoff:icoff + cin], wt, bias, s, act)
    got = ybuf[..., ocoff:ocoff + cout].float()
    err = (got - ref).abs()
    tol = 1.5e-2 + 1.5e-2 * ref.abs()
    assert bool((err <= tol).all()), "max err %.4g at %s (ref %.4g)" % (err.max().item(), np.unravel_index(int(err.argmax()), err.shape), ref.flatten()[int(err.argmax())].item())
    # untouched channels of the concat buffer stay untouched
    if ocoff > 0:
        assert bool((ybuf[..., :ocoff].float() == -77.0).all())
    if out_pitch > ocoff + cout:
        assert bool((ybuf[..., ocoff + cout:].float() == -77.0).all())
