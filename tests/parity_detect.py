"""End-to-end detection parity harness (test infrastructure): the product engine (CUDA, through the C ABI) against
the float64 CPU oracle run END TO END -- network, region decode, per-class NMS -- on the same weights and images.

BASELINE.json's north_star states the bar: "decoded boxes and scores within 1e-3 relative (bf16 convs, fp32 accumulate)".
What is compared:
  * net          rel-L2 and max |diff| of the network output [N,S,S,125] (post BN + leaky, O(1) values)
  * detections   keep lists of the engine vs keep lists of the oracle (its own net -> its own decode -> its own NMS):
                 fraction of identical (image, class) lists, Jaccard index over (image, class, box) triples
  * matched      for triples kept by BOTH: max and rms of |score - score*| / score*  and  |box - box*| / |box*| per
                 coordinate (cx, cy, w, h all > 0)
  * per_layer    rel-L2 of every layer's activation against the oracle's (error attribution)
Used by tests/test_parity_gpu.py (asserts) and tools/parity_report.py (prints the table committed under profiles/).
"""
import time

import numpy as np
import torch

from oracle import yolo2_oracle as O
from tests.helpers import make_store, oracle_params, rel_l2


def oracle_end_to_end(img_u8, core_p, head_p, score_thresh, iou_thresh, chunk=8, want_inter=False):
    """float64 oracle: preprocess -> core (inference BN: per-image, run in chunks to bound memory) -> head (batch
    statistics over the WHOLE batch, the reference quirk) -> region decode -> per-class NMS."""
    x = torch.tensor(O.preprocess_u8(img_u8))
    N = x.shape[0]
    inter = None
    with torch.no_grad():
        cores, inters = [], []
        for i in range(0, N, chunk):
            xi = x[i:i + chunk].to(torch.float64)
            li_out = []
            for (k, cin, cout, pool), p in zip(O.CORE_PLAN, core_p):
                xi, _, _ = O.conv_bn_layer(xi, p, False, torch.float64)
                if pool:
                    xi = O.max_pool_2x2(xi)
                if want_inter:
                    li_out.append(xi.to(torch.float32))
            cores.append(xi)
            inters.append(li_out)
        h = torch.cat(cores, 0)
        if want_inter:
            inter = [torch.cat([c[l] for c in inters], 0).numpy() for l in range(len(core_p))]
        for p in head_p:
            h, _, _ = O.conv_bn_layer(h, p, True, torch.float64)
            if want_inter:
                inter.append(h.to(torch.float32).numpy())
    net = h.numpy()
    boxes, sthr, sraw = O.region_decode_v2(net.astype(np.float32), O.VOC_ANCHORS, 20, score_thresh)
    # decode from the float64 net (region_decode_v2 computes in float64 from float32 input; feed the rounded net so the
    # oracle's own decode input is what a float32 TF graph would hold)
    keeps = [O.nms_per_class(boxes[n], sthr[n], iou_thresh, score_thresh) for n in range(N)]
    return dict(net=net, boxes=boxes, scores=sthr, scores_raw=sraw, keeps=keeps, inter=inter, score_thresh=score_thresh,
                iou_thresh=iou_thresh)


def _explain_diff(n, k, got, want, oracle_out, edge):
    """Why do two keep lists differ?  Every box in the symmetric difference must be (a) a score within `edge` (relative)
    of the score threshold, (b) a box whose IoU with some kept box of the class is within `edge` of the IoU threshold, or
    (c) a consequence: it overlaps (IoU > threshold) another box of the difference (kept / dropped because that one
    flipped).  Returns the number of boxes with none of these explanations."""
    thr, ithr = oracle_out['score_thresh'], oracle_out['iou_thresh']
    boxes = oracle_out['boxes'][n]
    diff = sorted(set(got) ^ set(want))
    union = sorted(set(got) | set(want))
    unexplained = 0
    for b in diff:
        s0 = float(oracle_out['scores_raw'][n, b, k])
        if abs(s0 - thr) <= edge * thr:
            continue
        others = [u for u in union if u != b]
        if others:
            iou = O.get_iou(np.broadcast_to(boxes[b], (len(others), 4)), boxes[others])
            if np.any(np.abs(iou - ithr) <= edge):
                continue
            if any(o in diff and i > ithr for o, i in zip(others, iou)):
                continue
        unexplained += 1
    return unexplained


def compare(engine_out, oracle_out, num_class=20, edge=2e-3):
    """engine_out: dict(net, boxes, scores, keep_idx, keep_count) numpy; oracle_out from oracle_end_to_end."""
    net, wnet = engine_out['net'].astype(np.float64), oracle_out['net']
    N = net.shape[0]
    res = dict(net_rel_l2=rel_l2(net, wnet), net_max_abs=float(np.abs(net - wnet).max()), net_abs_max_ref=float(np.abs(wnet).max()))
    ki, kc = engine_out['keep_idx'], engine_out['keep_count']
    lists_same = lists_total = 0
    inter_n = union_n = 0
    s_err, b_err = [], []
    n_or = n_en = 0
    unexplained = 0
    for n in range(N):
        for k in range(num_class):
            got = [int(v) for v in ki[n, k, :kc[n, k]]]
            want = [int(v) for v in oracle_out['keeps'][n][k]]
            n_or += len(want)
            n_en += len(got)
            if not got and not want:
                continue
            lists_total += 1
            lists_same += int(got == want)
            sg, sw = set(got), set(want)
            if sg != sw:
                unexplained += _explain_diff(n, k, got, want, oracle_out, edge)
            inter_n += len(sg & sw)
            union_n += len(sg | sw)
            for b in sg & sw:
                s0, s1 = float(oracle_out['scores'][n, b, k]), float(engine_out['scores'][n, b, k])
                s_err.append(abs(s1 - s0) / s0)
                b0, b1 = oracle_out['boxes'][n, b].astype(np.float64), engine_out['boxes'][n, b].astype(np.float64)
                b_err.append(float(np.max(np.abs(b1 - b0) / np.abs(b0))))
    s_err, b_err = np.asarray(s_err), np.asarray(b_err)
    res.update(detections_oracle=n_or, detections_engine=n_en, nonempty_lists=lists_total,
               keep_lists_identical=(lists_same / lists_total) if lists_total else 1.0,
               detections_jaccard=(inter_n / union_n) if union_n else 1.0, matched=int(len(s_err)),
               unexplained_list_differences=int(unexplained),
               score_rel_max=float(s_err.max()) if len(s_err) else 0.0,
               score_rel_rms=float(np.sqrt(np.mean(s_err ** 2))) if len(s_err) else 0.0,
               box_rel_max=float(b_err.max()) if len(b_err) else 0.0,
               box_rel_rms=float(np.sqrt(np.mean(b_err ** 2))) if len(b_err) else 0.0)
    return res


def run_case(batch, image_size, tame, precision, score_thresh=0.3, iou_thresh=0.45, seed=0, img_seed=1234, per_layer=False,
             oracle_cache=None, engine_kwargs=None, images=None):
    """One parity case.  Weights: the reference's initialiser with seed 0 (tame=False: what bench.py times) or the
    He-scaled variant (tame=True).  Returns the compare() dict (+ per-layer rel-L2 list)."""
    from tensorflow_yolo2_b200.engine import Yolo2Engine
    st, layers = make_store(125, seed=seed, tame=tame)
    core_p, head_p = oracle_params(st, layers)
    if images is not None:               # given uint8 [batch, IS, IS, 3] images (e.g. the reference's fixture) instead of random ones
        img = np.ascontiguousarray(images, dtype=np.uint8)
        assert img.shape == (batch, image_size, image_size, 3)
        img_seed = 'given:%d' % int(img.astype(np.int64).sum())
    else:
        img = np.random.RandomState(img_seed).randint(0, 256, (batch, image_size, image_size, 3)).astype(np.uint8)
    key = (batch, image_size, tame, seed, img_seed, score_thresh, iou_thresh, per_layer)
    t0 = time.time()
    if oracle_cache is not None and key in oracle_cache:
        want = oracle_cache[key]
    else:
        want = oracle_end_to_end(img, core_p, head_p, score_thresh, iou_thresh, want_inter=per_layer)
        if oracle_cache is not None:
            oracle_cache[key] = want
    t_oracle = time.time() - t0
    kw = dict(score_thresh=score_thresh, iou_thresh=iou_thresh, use_cuda_graph=False)
    kw.update(engine_kwargs or {})
    if precision is not None:
        kw['precision'] = precision
    eng = Yolo2Engine(batch, image_size, 125, store=st, **kw)
    r = eng.infer(torch.tensor(img))
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in r.items()}
    res = compare(out, want)
    res.update(batch=batch, image_size=image_size, tame=tame, precision=precision or 'bf16', oracle_seconds=round(t_oracle, 1))
    if per_layer:
        pl = []
        for li in range(len(eng.layers)):
            a = eng.layer_activation(li) if hasattr(eng, 'layer_activation') else eng.acts[li].float()
            pl.append(rel_l2(a.cpu().numpy(), want['inter'][li]))
        res['per_layer_rel_l2'] = pl
    del eng
    torch.cuda.empty_cache()
    return res
