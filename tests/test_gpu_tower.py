"""f3: the conv tower on the tensor cores (seam_tower_forward) against the PyTorch modules it replaces
(MatchPredictor.conv_seq / pool / linear in eval mode, models/match_head.py:50-62, 67-69)."""
import pytest
import torch

import seam_match_rcnn_b200 as pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# fp16 operands with fp32 accumulation through 4 convolutions of K = 2304 and a K = 1024 linear layer: the stated
# tolerance is 1e-2 of the output scale (observed ~2e-3); cuDNN's default path for the reference is TF32, the same
# 10-bit-mantissa class.  The PyTorch side of the comparison runs in strict fp32.
TOL_REL = 1e-2


def _model(seed=0):
    torch.manual_seed(seed)
    m = pkg.MatchPredictor().to(DEV).eval()
    bn = m.linear[1]
    with torch.no_grad():                       # non-trivial BatchNorm statistics and affine parameters
        bn.running_mean.uniform_(-0.5, 0.5)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
    return m


def _ref(m, x):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            return m.embed_torch(x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("K", [1, 5, 37, 300])
def test_tower_matches_pytorch(K):
    """Single ROI, tiles that straddle ROI boundaries, partial last tiles (K*H*W not a multiple of 128)."""
    m = _model()
    x = torch.randn(K, 256, 14, 14, device=DEV, generator=torch.Generator(device=DEV).manual_seed(K)).relu()   # RoIAlign of post-ReLU FPN maps is non-negative in practice; sign does not matter to the kernel
    x = x + 0.3 * torch.randn_like(x)
    ref = _ref(m, x)
    got = m.embed(x)
    assert got.shape == (K, 256) and got.dtype == torch.float32 and not got.requires_grad
    scale = ref.abs().max()
    err = (got - ref).abs().max()
    assert err <= TOL_REL * scale, f"tower error {err:.3e} vs output scale {scale:.3e}"
    # deterministic: same input, same bits; and independent of what else is in the batch
    assert torch.equal(got, m.embed(x))
    if K > 5:
        assert torch.equal(m.embed(x[3:5]), got[3:5])


def test_tower_scatters_rows():
    """dst_row sends ROI i to an arbitrary row of a larger buffer (the time-major x3_1_seq slots), other rows untouched."""
    m = _model(1)
    K = 23
    x = torch.randn(K, 256, 14, 14, device=DEV)
    plain = m.embed(x)
    buf = torch.full((64, 256), 7.0, device=DEV)
    rows = torch.randperm(64, device=DEV)[:K]
    m.embed(x, out=buf, dst_row=rows)
    assert torch.equal(buf[rows], plain)
    untouched = torch.ones(64, dtype=torch.bool, device=DEV)
    untouched[rows] = False
    assert (buf[untouched] == 7.0).all()


def test_tower_follows_weight_updates_and_training_mode():
    m = _model(2)
    x = torch.randn(4, 256, 14, 14, device=DEV)
    a = m.embed(x).clone()
    with torch.no_grad():
        m.conv_seq[0].weight.mul_(1.5)
    b = m.embed(x)
    assert not torch.equal(a, b)
    assert (b - _ref(m, x)).abs().max() <= TOL_REL * _ref(m, x).abs().max()
    m.train()                                    # training mode: the PyTorch modules (batch statistics, autograd)
    assert m.embed(x).requires_grad
    m.eval()
