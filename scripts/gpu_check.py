"""Developer check run on the GPU box: each stage of the hot path against the oracle, with the
observed error printed (the pytest suite asserts; this one reports).  Usage:
    python scripts/gpu_check.py [stage ...]      stages: agg dense cand topk misc time
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seam_match_rcnn_b200 as pkg          # noqa: E402
from oracle import seam_oracle as so        # noqa: E402

dev = torch.device("cuda:0")
W = so.random_weights(0)


def engine():
    e = pkg.SeamEngine(dev)
    e.load_weights({k: v.to(dev) for k, v in W.items()})
    return e


def stage_agg(e):
    for (Q, T, rag) in [(64, 10, None), (37, 4, (0, 4)), (5, 1, None), (8, 64, None), (301, 10, (1, 10)), (1000, 3, None),
                        (77, 40, (0, 40)), (33, 64, (1, 64)), (40, 17, None), (21, 33, (30, 33))]:
        seq, mask, lens = so.synth_tracks(Q, T, seed=Q + T, ragged=rag)
        ref, att = so.aggregate_tracks(seq, mask, W)
        out, a = e.aggregate(seq.to(dev), mask.to(dev), getatt=True)
        torch.cuda.synchronize()
        err = (out.cpu() - ref).abs().max().item()
        aref = torch.zeros(Q, T)
        for i, p in enumerate(att):
            aref[i, :p.shape[0]] = p[:, 0]
        aerr = (a.cpu() - aref).abs().max().item()
        out2 = e.aggregate(seq.to(dev), None, lens=torch.as_tensor(lens))
        e2 = (out2.cpu() - ref).abs().max().item()
        print(f"[agg] Q={Q} T={T} ragged={rag}: |out-ref|max={err:.3e} (lens path {e2:.3e}) |att-ref|max={aerr:.3e} "
              f"ref_absmax={ref.abs().max():.2f}")


def stage_dense(e):
    q = torch.from_numpy(np.random.RandomState(1).randn(70, 256).astype(np.float32))
    g = torch.from_numpy(np.random.RandomState(2).randn(301, 256).astype(np.float32))
    ref = so.pair_logits(q, g, W)
    x5 = e.score_dense(q.to(dev), g.to(dev)).cpu()
    print(f"[dense] |x5-ref|max={(x5 - ref).abs().max():.3e} ref_absmax={ref.abs().max():.2f}")


def approx_margin_ref(q, g):
    """What the tensor-core pass should produce: fp16-rounded operands, exact accumulation."""
    dw = (W["last.weight"][1] - W["last.weight"][0]).double()
    a = (-2.0 * dw.float() * q).half().double()
    gh = g.half().double()
    cg = (dw * g.double() ** 2).sum(1)
    return a @ gh.T + cg[None]


def stage_cand(e):
    """Inspect the raw per-row candidate lists of the tcgen05 kernel."""
    for (Q, G) in [(64, 1000), (130, 300), (257, 5000), (1000, 20000), (3000, 40000)]:
        rs = np.random.RandomState(Q)
        q = torch.from_numpy(rs.randn(Q, 256).astype(np.float32))
        g = torch.from_numpy(rs.randn(G, 256).astype(np.float32))
        gal = e.prepare_gallery(g.to(dev))
        sc, mg, ix, st = e.score_topk(q.to(dev), gal, 20, return_stats=True)
        torch.cuda.synchronize()
        plan = e.score_plan(Q, G)
        ws = e._ws["score"]
        P, CAP = plan["ctas_per_query_tile"], plan["list_capacity"]
        nl = P * 4
        off = plan["off_rowbuf"]
        off += (-(ws.data_ptr() + off)) % (CAP * 8)       # the library aligns the base to the sub-list size
        # a record is a quad {w0,w1,w2,w3}: w0 & 63 = quad position in the sub-list's 64-column quarter,
        # (w1 & 63) | (w2 & 63) << 6 | (w3 & 63) << 12 = gallery tile (256 rows); quarter = sub-list index % 4
        capq = CAP // 2
        buf = ws[off:off + Q * nl * CAP * 8].view(torch.int32).view(Q, nl, capq, 4).cpu()
        cnts = ws[plan["off_rowcnt"]:plan["off_rowcnt"] + Q * nl * 4].view(torch.int32).view(Q, nl).cpu()
        ref = approx_margin_ref(q, g)
        top = ref.topk(min(32, G), dim=1).indices
        miss, verr, over = 0, 0.0, 0
        cnt = cnts.sum(1)
        for i in range(Q):
            vals, idx = [], []
            for l in range(nl):
                n = int(cnts[i, l])
                over += n > capq
                n = min(n, capq)
                rec = buf[i, l, :n].long()
                tile = (rec[:, 1] & 63) | ((rec[:, 2] & 63) << 6) | ((rec[:, 3] & 63) << 12)
                col0 = tile * 256 + (l % 4) * 64 + (rec[:, 0] & 63) * 4
                cols = col0[:, None] + torch.arange(4)[None, :]
                v = buf[i, l, :n].contiguous().view(torch.float32)
                keep = cols < G                                    # padded columns hold NaN
                vals.append(v[keep])
                idx.append(cols[keep])
            vals, idx = torch.cat(vals), torch.cat(idx)
            if idx.numel():
                verr = max(verr, float((vals.double() - ref[i, idx]).abs().max()))
            sset = set(idx.tolist())
            if len(sset) != idx.numel():
                over += 1000000                            # duplicates would break the top-32 selection
            miss += sum(1 for j in top[i].tolist() if j not in sset)
        print(f"[cand] Q={Q} G={G} plan={ {k: plan[k] for k in ('query_tiles', 'gallery_tiles', 'ctas', 'ctas_per_query_tile', 'list_capacity')} } "
              f"|cand_v - ref|max={verr:.3e} missing_from_top32={miss} quads per row mean={cnt.float().mean():.1f} "
              f"max={int(cnt.max())} overflow_rows={over} fallback_rows={int(st[0])}")


def check_topk(sc, mg, ix, x5, k, tol=3e-5):
    d = so.logit_margin(x5)
    s_ref, d_ref, i_ref = so.rank_topk(x5, k)
    ix = ix.long()
    same = (ix == i_ref).all(1)
    dmine = torch.gather(d, 1, ix.clamp(min=0))
    max_merr = (mg - dmine).abs()[ix >= 0].max().item() if (ix >= 0).any() else 0.0
    bad = 0
    for r in (~same).nonzero().flatten().tolist():
        # any disagreement must be a tie inside tol
        if not torch.allclose(d_ref[r], dmine[r], atol=tol, rtol=0):
            bad += 1
    serr = (sc - torch.gather(so.match_scores(x5), 1, ix.clamp(min=0))).abs().max().item()
    return int((~same).sum()), bad, max_merr, serr


def stage_topk(e):
    for (Q, T, G, k, seed) in [(64, 10, 1000, 20, 0), (37, 4, 300, 20, 2), (9, 3, 7, 20, 7), (200, 10, 5000, 5, 9),
                               (513, 2, 2049, 1, 4), (64, 10, 1000, 32, 0)]:
        seq, mask, _ = so.synth_tracks(Q, T, seed)
        qref, _ = so.aggregate_tracks(seq, mask, W)
        g = so.synth_gallery(G, seed, qref)
        x5 = so.pair_logits(qref, g, W)
        gal = e.prepare_gallery(g.to(dev))
        # feed the ORACLE's aggregated queries so this stage isolates the scorer
        sc, mg, ix, st = e.score_topk(qref.to(dev), gal, k, return_stats=True)
        torch.cuda.synchronize()
        kk = min(k, G)
        ndiff, bad, merr, serr = check_topk(sc.cpu()[:, :kk], mg.cpu()[:, :kk], ix.cpu()[:, :kk], x5, kk)
        pad_ok = bool((ix.cpu()[:, kk:] == -1).all())
        print(f"[topk] Q={Q} G={G} k={k}: rows differing={ndiff} (non-tie: {bad}) |margin err|max={merr:.3e} "
              f"|score err|max={serr:.3e} fallback_rows={int(st[0])} pad_ok={pad_ok}")


def stage_misc(e):
    # nlb_forward
    rs = np.random.RandomState(11)
    for t in (1, 2, 7, 10, 64):
        x = torch.from_numpy(rs.randn(3, 256, t).astype(np.float32))
        z = e.nlb_forward(x.to(dev)).cpu()
        ref = so.nlb_forward(x, W)
        print(f"[nlb] T={t}: |z-ref|max={(z - ref).abs().max():.3e}")
    # rank of target
    q = torch.from_numpy(rs.randn(50, 256).astype(np.float32))
    g = torch.from_numpy(rs.randn(777, 256).astype(np.float32))
    tgt = torch.from_numpy(rs.randint(0, 777, size=50))
    r, m = e.rank_of_target(q.to(dev), g.to(dev), tgt)
    rr = so.rank_of_target(so.pair_logits(q, g, W), tgt)
    print(f"[rank] mismatches={(r.cpu().long() != rr).sum().item()} of 50")
    # merge
    x5 = so.pair_logits(q, g, W)
    parts = [(0, 300), (300, 500), (500, 777)]
    lists = [so.rank_topk(x5[:, a:b], 20) for a, b in parts]
    S = torch.stack([l[0] for l in lists]).to(dev)
    M = torch.stack([l[1] for l in lists]).to(dev)
    I = torch.stack([(l[2] + a).int() for l, (a, b) in zip(lists, parts)]).to(dev)
    s, mg, ix = e.merge_topk(S, M, I)
    s_ref, d_ref, i_ref = so.rank_topk(x5, 20)
    print(f"[merge] idx identical={bool((ix.cpu().long() == i_ref).all())} |score err|={(s.cpu() - s_ref).abs().max():.1e}")


def stage_time(e):
    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    for (Q, T, G) in [(15000, 10, 15000), (10000, 10, 125000), (100000, 64, 0)]:
        seq = torch.randn(1 + T, Q, 256, device=dev)
        t_agg = timeit(lambda: e.aggregate(seq))
        bytes_ = Q * (T + 1) * 1024
        msg = f"[time] Q={Q} T={T}: aggregate {t_agg * 1e3:.1f} us = {bytes_ / t_agg / 1e6:.0f} GB/s"
        if G:
            g = torch.randn(G, 256, device=dev)
            q = e.aggregate(seq)
            gal = e.prepare_gallery(g)
            t_prep = timeit(lambda: e.prepare_gallery(g))
            t_sc = timeit(lambda: e.score_topk(q, gal, 20))
            _, _, _, st = e.score_topk(q, gal, 20, return_stats=True)
            msg += (f" | prepare_gallery {t_prep * 1e3:.1f} us | score_topk {t_sc * 1e3:.1f} us = "
                    f"{Q * G / t_sc / 1e6:.2f} Gpairs/s = {Q * G * 512 / t_sc / 1e9:.1f} TFLOP/s; "
                    f"fallback_rows={int(st[0])} ctas_per_query_tile={e.score_plan(Q, G)['ctas_per_query_tile']}")
        print(msg)


if __name__ == "__main__":
    stages = sys.argv[1:] or ["agg", "dense", "cand", "topk", "misc", "time"]
    print(torch.cuda.get_device_name(0), "stages:", stages)
    e = engine()
    for s in stages:
        t0 = time.time()
        try:
            globals()["stage_" + s](e)
            torch.cuda.synchronize()
        except Exception as ex:   # keep going: later stages may still be informative
            print(f"[{s}] FAILED: {type(ex).__name__}: {ex}")
            try:
                torch.cuda.synchronize()
            except Exception as ex2:
                print(f"[{s}] device unusable after failure: {ex2}")
                break
        print(f"[{s}] done in {time.time() - t0:.1f}s")
