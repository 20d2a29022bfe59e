"""Generate golden vectors for the hot path from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these goldens are the parity pin: they are
the outputs of ``/root/reference/models/match_head.py::TemporalAggregationNLB`` (seq-branch)
and ``models/nlb.py::NONLocalBlock1D`` on seeded synthetic inputs.  Inputs and weights are
regenerated from seeds by ``oracle.seam_oracle`` (numpy RandomState streams are frozen), only
the reference OUTPUTS are stored.  ``pycocotools`` is not installed; the reference imports
it only for a training-time helper, so an empty stub module is placed in sys.modules.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import seam_oracle as so  # noqa: E402


def import_reference():
    for name in ("pycocotools", "pycocotools.mask"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    sys.path.insert(0, REF)
    from models.match_head import TemporalAggregationNLB  # type: ignore
    from models.nlb import NONLocalBlock1D  # type: ignore
    return TemporalAggregationNLB, NONLocalBlock1D


# name -> (Q, Tmax, ragged, G, seed, planted)
CASES = {
    "cfg1":      dict(Q=64, Tmax=10, ragged=None,   G=1000, seed=0, planted=True),
    "ragged":    dict(Q=37, Tmax=4,  ragged=(0, 4), G=300,  seed=2, planted=True),
    "t1":        dict(Q=5,  Tmax=1,  ragged=None,   G=33,   seed=5, planted=False),
    "t64":       dict(Q=8,  Tmax=64, ragged=None,   G=513,  seed=3, planted=True),
    "tiny_g":    dict(Q=9,  Tmax=3,  ragged=(1, 3), G=7,    seed=7, planted=False),
}
K_LIST = (1, 5, 10, 20)


def make_inputs(case, w):
    seq, mask, lens = so.synth_tracks(case["Q"], case["Tmax"], case["seed"], case["ragged"])
    planted = None
    if case["planted"]:
        planted, _ = so.aggregate_tracks(seq, mask, w)
    gal = so.synth_gallery(case["G"], case["seed"], planted)
    return seq, mask, lens, gal


def main():
    TemporalAggregationNLB, NONLocalBlock1D = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(1)           # deterministic accumulation order for the pin
    w = so.random_weights(seed=0)
    ref = TemporalAggregationNLB().eval()
    missing, unexpected = ref.load_state_dict(w, strict=False)
    assert not unexpected, unexpected
    assert all(not any(m.startswith(p) for p in ("newnlb", "attention_scorer", "last")) for m in missing), missing

    out = {}
    with torch.no_grad():
        for name, case in CASES.items():
            seq, mask, lens, gal = make_inputs(case, w)
            r = ref(None, None, None, x3_1_seq=seq, x3_1_mask=mask, x3_2=gal, getatt=True)
            x3_1b, x3_2, x5, _, _, ids, att = r
            assert ids.shape == (1, 2)
            # the oracle restatement must agree with the reference
            o = so.forward_seq_branch(seq, mask, gal, w, getatt=True)
            assert torch.equal(o[0], x3_1b), name
            assert torch.equal(o[2], x5), name
            for a, b in zip(o[6], att):
                assert torch.equal(a, b), name
            attp = np.zeros((case["Q"], case["Tmax"]), np.float32)
            for i, p in enumerate(att):
                attp[i, :p.shape[0]] = p[:, 0].numpy()
            s, d, idx = so.rank_topk(x5, max(K_LIST))
            out[f"{name}.x3_1b"] = x3_1b.numpy()
            out[f"{name}.att"] = attp
            out[f"{name}.lens"] = np.asarray(lens, np.int32)
            out[f"{name}.x5_head"] = x5[:4].numpy()              # first 4 queries, all G
            out[f"{name}.score_sum"] = so.match_scores(x5).double().sum(1).numpy()
            out[f"{name}.topk_score"] = s.numpy()
            out[f"{name}.topk_margin"] = d.numpy()
            out[f"{name}.topk_idx"] = idx.numpy().astype(np.int32)
            # eval-script numpy path on the same data (fp16-rounded gallery / weights)
            g16 = gal.numpy().astype(np.float16)
            aW = w["last.weight"].numpy().astype(np.float16)
            aB = w["last.bias"].numpy().astype(np.float16)
            n_eval = min(4, case["Q"])
            ev = np.stack([so.eval_aggr_scores_np(g16, x3_1b[i].numpy(), aW, aB)[0] for i in range(n_eval)])
            out[f"{name}.eval_scores_head"] = ev.astype(np.float32)

        # NONLocalBlock1D on its own (models/nlb.py:66-101), with the same weights
        nlb = NONLocalBlock1D(in_channels=256, sub_sample=False, bn_layer=False).eval()
        nlb.load_state_dict({k[len("newnlb."):]: v for k, v in w.items() if k.startswith("newnlb.")})
        rs = np.random.RandomState(11)
        for t in (2, 7, 10):
            x = torch.from_numpy(rs.randn(3, 256, t).astype(np.float32))
            z = nlb(x)
            assert torch.equal(so.nlb_forward(x, w), z)
            out[f"nlb.t{t}.x"] = x.numpy()
            out[f"nlb.t{t}.z"] = z.numpy()

        # default (zero) W makes the block the identity: SURVEY.md section 0 item 2
        w0 = so.random_weights(seed=0, randomize_W=False)
        x = torch.from_numpy(rs.randn(1, 256, 5).astype(np.float32))
        assert torch.equal(so.nlb_forward(x, w0), x)

    path = os.path.join(HERE, "hotpath_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
    print("torch", torch.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
