"""f4: gradients of the hot path (seam_aggregate_backward, seam_score_dense_backward behind torch.autograd.Function)
against PyTorch autograd through the reference's own formulation (the oracle's forward_seq_branch, which is checked
bit-for-bit against the reference modules in tests/test_oracle_golden.py)."""
import pytest
import torch

import seam_match_rcnn_b200 as pkg
from oracle import seam_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("newnlb.theta.weight", "newnlb.theta.bias", "newnlb.phi.weight", "newnlb.phi.bias", "newnlb.g.weight",
        "newnlb.g.bias", "newnlb.W.weight", "newnlb.W.bias", "newnlb.concat_project.0.weight",
        "attention_scorer.weight", "attention_scorer.bias", "last.weight", "last.bias")


def _loss(x3_1b, x5, r1, r2):
    return (x5 * r2).sum() + (x3_1b * r1).sum() + torch.logsumexp(x5, -1).mean()


@pytest.mark.parametrize("Q,T,G,ragged", [(7, 10, 9, None), (12, 6, 5, (0, 6)), (5, 16, 3, (1, 16)), (3, 1, 4, None)])
def test_training_gradients_match_autograd(weights, Q, T, G, ragged):
    seq, mask, lens = so.synth_tracks(Q, T, seed=40 + Q, ragged=ragged)
    gal = so.synth_gallery(G, 40 + Q, None)
    r1 = torch.randn(Q, 256, generator=torch.Generator().manual_seed(1))
    r2 = torch.randn(Q, G, 2, generator=torch.Generator().manual_seed(2))
    # ---- reference: autograd through the oracle's un-folded formulation, on the CPU in fp64 for a clean gradient
    w64 = {k: v.double().clone().requires_grad_(True) for k, v in weights.items()}
    seq64, gal64 = seq.double().clone().requires_grad_(True), gal.double().clone().requires_grad_(True)
    out = so.forward_seq_branch(seq64, mask, gal64, w64)
    _loss(out[0], out[2], r1.double(), r2.double()).backward()
    # ---- ours: the drop-in module in training mode
    m = pkg.TemporalAggregationNLB().to(DEV).train()
    m.load_state_dict(weights, strict=False)
    seq_d, gal_d = seq.to(DEV).requires_grad_(True), gal.to(DEV).requires_grad_(True)
    x3_1b, _, x5, _, _, _ = m(None, None, None, x3_1_seq=seq_d, x3_1_mask=mask.to(DEV), x3_2=gal_d)
    assert x5.requires_grad and x3_1b.requires_grad
    assert (x3_1b.detach().cpu() - out[0].detach().float()).abs().max() <= 2e-5
    _loss(x3_1b, x5, r1.to(DEV), r2.to(DEV)).backward()

    def check(name, got, ref):
        ref = ref.float()
        tol = 2e-4 * max(1.0, float(ref.abs().max()))
        err = float((got.cpu() - ref).abs().max())
        assert err <= tol, f"{name}: |grad - autograd|max = {err:.3e} (scale {float(ref.abs().max()):.3e})"

    check("x3_1_seq", seq_d.grad, seq64.grad)
    check("x3_2", gal_d.grad, gal64.grad)
    sd = dict(m.named_parameters())
    for k in KEYS:
        if T == 1 and not k.startswith("last"):
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0     # the block is skipped for T == 1
            continue
        check(k, sd[k].grad, w64[k].grad)


def test_x_branch_trains_end_to_end(weights):
    """x-branch in training mode: conv tower (PyTorch, autograd) -> grouping -> aggregation kernel -> scorer kernel;
    gradients reach the tower's first convolution and the ROI features."""
    torch.manual_seed(3)
    m = pkg.TemporalAggregationNLB().to(DEV).train()
    m.load_state_dict(weights, strict=False)
    x = torch.randn(9, 256, 14, 14, device=DEV, requires_grad=True)
    types = torch.tensor([1, 0, 0, 0, 1, 0, 0, 0, 0])
    ids = torch.tensor([0, 5, 2, 5, 0, 2, 5, 9, 2])
    _, _, x5, _, _, _ = m(x, types, ids)
    gts = torch.tensor([0, 1, 0, 1, 1, 0], device=DEV)
    torch.nn.functional.cross_entropy(x5.view(-1, 2), gts).backward()
    assert x.grad is not None and float(x.grad.abs().sum()) > 0
    assert float(m.conv_seq[0].weight.grad.abs().sum()) > 0 and float(m.newnlb.theta.weight.grad.abs().sum()) > 0
    # one optimiser step changes the folded weights the forward kernels use
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    seq = torch.randn(4, 5, 256, device=DEV)
    m.eval()
    a = m.aggregate(seq).clone()
    opt.step()
    b = m.aggregate(seq)
    assert not torch.equal(a, b)
