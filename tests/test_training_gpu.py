"""GPU parity tests of the training path (SURVEY a11 + a'): BN/leaky/pool backward, data gradient (the forward
tcgen05 kernel on transposed+flipped weights), weight gradient (tcgen05, MN-major operands), first-layer weight
gradient, region loss, and a whole training step against the oracle (torch float64 autograd over the oracle's
restatement of the reference graph)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import yolo2_oracle as O
from tests.helpers import make_store, oracle_params, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from tensorflow_yolo2_b200 import ops as _ops
    return _ops


def cu(a, dtype=None):
    t = torch.as_tensor(np.asarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def bf16r(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,H,C,pool,dy_bf16', [(2, 8, 64, False, True), (2, 8, 64, True, True), (3, 6, 125, False, False),
                                                 (2, 12, 32, True, True), (1, 4, 1024, False, True)])
def test_bn_leaky_pool_bwd_vs_autograd(ops, N, H, C, pool, dy_bf16):
    rs = np.random.RandomState(C + H)
    ldh = (C + 31) // 32 * 32
    h = rs.randn(N * H * H, ldh).astype(np.float32) * 2 + 0.5
    gamma = (rs.uniform(0.5, 1.5, C) * np.where(rs.rand(C) < 0.2, -1, 1)).astype(np.float32)
    beta = (rs.randn(C) * 0.3).astype(np.float32)
    Ho = H // 2 if pool else H
    dy = rs.randn(N, Ho, Ho, C).astype(np.float32)
    if dy_bf16:
        dy = bf16r(dy).astype(np.float32)
    # oracle: autograd through batch_norm -> leaky -> pool (oracle primitives)
    ht = torch.tensor(h[:, :C].reshape(N, H, H, C), dtype=torch.float64, requires_grad=True)
    gt_, bt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True), torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    z, _, _ = O.batch_norm(ht, gt_, bt, torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64), True)
    y = torch.maximum(O.ALPHA * z, z)
    if pool:
        y = O.max_pool_2x2(y)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    # product
    hd = cu(h)
    mean, var = ops.bn_stats(hd, C, ld=ldh)
    ld_dh = (C + 63) // 64 * 64
    dgamma, dbeta, dh = ops.bn_leaky_pool_bwd(hd, cu(dy, torch.bfloat16 if dy_bf16 else torch.float32), mean, var, cu(gamma),
                                              cu(beta), N, H, H, C, ldh=ldh, leaky=True, pool=pool, ld_dh=ld_dh)
    torch.cuda.synchronize()
    assert rel_l2(dgamma.cpu().numpy(), gt_.grad.numpy()) < 1e-4
    assert rel_l2(dbeta.cpu().numpy(), bt.grad.numpy()) < 1e-4
    got = dh.float().cpu().numpy()
    assert np.all(got[:, C:] == 0), 'padding columns of dh must be zero'
    assert rel_l2(got[:, :C], ht.grad.numpy().reshape(-1, C)) < 6e-3          # bf16 output rounding


@pytest.mark.parametrize('N,H,cin,cout,k', [(2, 13, 64, 128, 3), (2, 10, 128, 64, 3), (1, 16, 256, 128, 1),
                                             (2, 13, 1024, 125, 1), (2, 8, 32, 64, 3), (1, 13, 512, 1024, 3),
                                             (2, 40, 64, 128, 3),      # tile groups: two M tiles per unit (block_n 128), ragged last group
                                             (3, 36, 32, 64, 3),       # three M tiles per unit (block_n 64)
                                             (2, 34, 128, 64, 3)])
def test_dgrad_and_wgrad_vs_autograd(ops, N, H, cin, cout, k):
    rs = np.random.RandomState(cin + cout + k)
    x = bf16r(rs.randn(N, H, H, cin))
    w = (rs.randn(k, k, cin, cout) / np.sqrt(k * k * cin)).astype(np.float32)
    ld_dh = (cout + 63) // 64 * 64
    dh = np.zeros((N * H * H, ld_dh), dtype=np.float64)
    dh[:, :cout] = bf16r(rs.randn(N * H * H, cout))
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(bf16r(w), dtype=torch.float64, requires_grad=True)
    y = O.conv2d_same(xt, wt)
    (y * torch.tensor(dh[:, :cout].reshape(N, H, H, cout))).sum().backward()
    # weight gradient
    dw = torch.zeros((k, k, cin, cout), dtype=torch.float32, device='cuda')
    ops.conv_wgrad_bf16(cu(x, torch.bfloat16), cu(dh, torch.bfloat16), k, cin, cout, dw)
    torch.cuda.synchronize()
    err_w = rel_l2(dw.cpu().numpy(), wt.grad.numpy())
    # data gradient: forward kernel on transposed + flipped weights
    wp = ops.pack_weights_dgrad_bf16(cu(w), ld_dh)
    dx = ops.conv_fwd_bf16(cu(dh, torch.bfloat16).view(N, H, H, ld_dh), wp, k, ld_dh, cin, scale=None, shift=None, leaky=False,
                           pool=False)
    torch.cuda.synchronize()
    err_x = rel_l2(dx.float().cpu().numpy(), xt.grad.numpy())
    print('wgrad rel_l2=%.3g dgrad rel_l2=%.3g' % (err_w, err_x))
    assert err_w < 1e-4, 'weight gradient mismatch: %g' % err_w       # fp32 accumulation of exact bf16 products
    assert err_x < 5e-3, 'data gradient mismatch: %g' % err_x         # bf16 output rounding


def test_wgrad_split_k_large_map(ops):
    """K = 2*104*104 pixels with forced split-K and a ragged last stage."""
    rs = np.random.RandomState(5)
    N, H, W, cin, cout, k = 2, 50, 37, 64, 64, 3
    x = bf16r(rs.randn(N, H, W, cin))
    dh = bf16r(rs.randn(N * H * W, cout))
    xt = torch.tensor(x, dtype=torch.float64)
    wt = torch.zeros((k, k, cin, cout), dtype=torch.float64, requires_grad=True)
    y = O.conv2d_same(xt, wt)
    (y * torch.tensor(dh.reshape(N, H, W, cout))).sum().backward()
    dw = torch.zeros((k, k, cin, cout), dtype=torch.float32, device='cuda')
    ops.conv_wgrad_bf16(cu(x, torch.bfloat16), cu(dh, torch.bfloat16), k, cin, cout, dw)
    torch.cuda.synchronize()
    assert rel_l2(dw.cpu().numpy(), wt.grad.numpy()) < 1e-4


@pytest.mark.parametrize('ffma', [False, True])
@pytest.mark.parametrize('N,H,W,cout', [(2, 40, 72, 32), (1, 13, 19, 32), (3, 64, 64, 24)])
def test_wgrad_first_layer(ops, monkeypatch, ffma, N, H, W, cout):
    """First-layer weight gradient: the mma.sync kernel (default) and the FFMA kernel (Y2_BN_BWD_GENERIC=1) against float64
    autograd -- ragged tiles, fewer than 32 output channels."""
    if ffma:
        monkeypatch.setenv('Y2_BN_BWD_GENERIC', '1')
        ops.reload_env()
    rs = np.random.RandomState(9)
    x3 = bf16r(rs.uniform(-1, 1, (N, H, W, 3)))
    x8 = np.zeros((N, H, W, 8), dtype=np.float64)
    x8[..., :3] = x3
    dh = bf16r(rs.randn(N * H * W, cout))
    wt = torch.zeros((3, 3, 3, cout), dtype=torch.float64, requires_grad=True)
    y = O.conv2d_same(torch.tensor(x3, dtype=torch.float64), wt)
    (y * torch.tensor(dh.reshape(N, H, W, cout))).sum().backward()
    dw = torch.zeros((3, 3, 3, cout), dtype=torch.float32, device='cuda')
    ops.conv_wgrad_c3(cu(x8, torch.bfloat16), cu(dh, torch.bfloat16), cout, dw)
    torch.cuda.synchronize()
    assert rel_l2(dw.cpu().numpy(), wt.grad.numpy()) < 1e-4


# ---------------------------------------------------------------------------------------------
def _random_gt(rs, N, G, S):
    counts = rs.randint(0, min(G, 6) + 1, N).astype(np.int32)
    boxes = np.zeros((N, G, 4), dtype=np.float32)
    classes = np.zeros((N, G), dtype=np.int32)
    for n in range(N):
        for g in range(counts[n]):
            boxes[n, g] = [rs.uniform(0.02, 0.98), rs.uniform(0.02, 0.98), rs.uniform(0.05, 0.7), rs.uniform(0.05, 0.7)]
            classes[n, g] = rs.randint(0, 20)
        if counts[n] >= 2:          # two ground truths in the same cell with similar shapes: slot collision
            boxes[n, 1] = boxes[n, 0] + np.array([0.001, 0.001, 0.01, 0.01], dtype=np.float32)
    return boxes, classes, counts


@pytest.mark.parametrize('N,S', [(3, 13), (2, 19), (5, 7)])
def test_region_loss_vs_oracle(ops, N, S):
    rs = np.random.RandomState(S)
    net = (rs.randn(N, S, S, 125) * 1.5).astype(np.float32)
    boxes, classes, counts = _random_gt(rs, N, 8, S)
    want_loss, want_grad, want_terms = O.region_loss_torch(net, boxes, classes, counts)
    terms, dnet = ops.region_loss(cu(net), cu(O.VOC_ANCHORS), cu(boxes), cu(classes), cu(counts))
    torch.cuda.synchronize()
    t = terms.cpu().numpy()
    np.testing.assert_allclose(t[:4], want_terms, rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(t[4], want_loss, rtol=2e-4)
    assert rel_l2(dnet.cpu().numpy(), want_grad) < 1e-4


# ---------------------------------------------------------------------------------------------
def _labels_v1(rs, N, S, IS, C=20):
    lab = np.zeros((N, S, S, 5 + C), dtype=np.float32)
    for n in range(N):
        for _ in range(rs.randint(1, 4)):
            cx, cy = rs.uniform(0, IS, 2)
            w, h = rs.uniform(20, IS * 0.7, 2)
            j, i = int(cx * S / IS), int(cy * S / IS)
            if lab[n, i, j, 0] == 1:
                continue
            lab[n, i, j, 0] = 1
            lab[n, i, j, 1:5] = [cx, cy, w, h]
            lab[n, i, j, 5 + rs.randint(0, C)] = 1
    return lab


def test_training_step_vs_oracle(ops):
    """One full iteration (forward, loss, backward, Adam) at 96x96, batch 6, reference loss (B=2 -> 30 channels, the
    reference's own output_filter) on He-scaled weights.

    A 22-layer batch-statistics network amplifies bf16 rounding noise chaotically (every layer re-normalises), so the
    end-to-end gradients of two correct implementations that round differently agree only loosely.  The tight check is
    therefore layer-local: for every layer, the oracle (float64 autograd over the oracle's conv/BN/leaky/pool
    restatement) is evaluated on the product's OWN saved layer input and incoming gradient, and must reproduce the
    product's dW, dgamma, dbeta and outgoing data gradient.  The whole-chain comparison is a direction check."""
    from tensorflow_yolo2_b200.trainer import Yolo2Trainer
    N, IS, B = 6, 96, 2
    S = IS // 32
    st, layers = make_store(30, tame=True)
    core_p, head_p = oracle_params(st, layers)
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (N, IS, IS, 3)).astype(np.uint8)
    labels = _labels_v1(rs, N, S, IS)
    tr = Yolo2Trainer(N, IS, 30, store=st, loss='v1', B=B, device='cuda:0')
    p_before = [{k: v.clone() for k, v in tr.P[li].items()} for li in range(len(layers))]
    tr.set_labels(labels)
    cap = {}
    terms = tr.step(torch.tensor(img), capture=cap)
    torch.cuda.synchronize()

    # ---- whole chain vs the oracle: forward, loss, gradient direction ----
    lab_t = torch.tensor(labels, dtype=torch.float64)
    loss_fn = lambda net: O.loss_v1_graph(net, lab_t, 20, N, IS, S, B)[0]
    want_loss, grads, stats, net = O.train_step_reference(O.preprocess_u8(img), core_p, head_p, loss_fn, bf16_operands=True)
    got_net = tr.acts[-1].cpu().numpy()
    print('net rel_l2 %.3g  loss %.6g vs %.6g' % (rel_l2(got_net, net), float(terms[4]), want_loss))
    assert rel_l2(got_net, net) < 8e-2
    assert abs(float(terms[4]) - want_loss) / abs(want_loss) < 5e-2
    for li in (21, 15, 8, 0):
        a, b = tr.G[li]['W'].cpu().numpy().ravel().astype(np.float64), grads[li]['W'].ravel()
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
        print('layer %2d dW cosine vs whole-chain oracle %.3f' % (li + 1, cos))
        assert cos > 0.8
    # loss gradient on the product's own net output
    _, dnet_want = O.get_loss_torch(got_net, labels, 20, N, IS, S, B)
    assert rel_l2(cap[21].cpu().numpy(), dnet_want) < 1e-4

    # ---- layer-local backward parity on the product's own forward state ----
    nl = len(layers)
    for li in reversed(range(nl)):
        L = layers[li]
        xin = tr.x0[..., :3] if li == 0 else tr.acts[li - 1]
        x = xin.double().cpu().requires_grad_(True)
        q = {k: p_before[li][k].double().cpu().requires_grad_(True) for k in ('W', 'b', 'gamma', 'beta')}
        q['mm'] = torch.zeros(L['cout'], dtype=torch.float64)
        q['mv'] = torch.ones(L['cout'], dtype=torch.float64)
        y, _, _ = O.conv_bn_layer(x, q, True, torch.float64, bf16_operands=True)
        if L['pool']:
            y = O.max_pool_2x2(y)
        (y * cap[li].double().cpu()).sum().backward()
        eW = rel_l2(tr.G[li]['W'].cpu().numpy(), q['W'].grad.numpy())
        eg = rel_l2(tr.G[li]['gamma'].cpu().numpy(), q['gamma'].grad.numpy())
        eb = rel_l2(tr.G[li]['beta'].cpu().numpy(), q['beta'].grad.numpy())
        ex = rel_l2(cap[li - 1].float().cpu().numpy(), x.grad.numpy()) if li > 0 else 0.0
        print('layer %2d local: dW %.3g dgamma %.3g dbeta %.3g dx %.3g' % (li + 1, eW, eg, eb, ex))
        assert eW < 1e-2 and eg < 1e-2 and eb < 1e-2 and ex < 1e-2, (li, eW, eg, eb, ex)
        assert float(tr.G[li]['b'].abs().max()) == 0.0           # bias gradient: analytically zero in front of BN
        assert np.abs(q['b'].grad.numpy()).max() < 1e-6 * max(1.0, np.abs(q['W'].grad.numpy()).max()) + 1e-9
    # ---- Adam (TF1 form), first step ----
    for li in (0, 10, 21):
        for key in ('W', 'gamma', 'beta'):
            g = tr.G[li][key].cpu().numpy().astype(np.float64)
            p_new = O.adam_step(p_before[li][key].cpu().numpy().astype(np.float64), g, np.zeros(g.shape), np.zeros(g.shape), 1)[0]
            assert rel_l2(tr.P[li][key].cpu().numpy(), p_new) < 1e-5
    # moving averages were updated (UPDATE_OPS): mm = 0.99*mm + 0.01*batch_mean
    bn0 = layers[0]['bn']
    mm_want = 0.99 * np.asarray(core_p[0]['mm']) + 0.01 * stats[0][0]
    assert rel_l2(st[bn0['moving_mean']].cpu().numpy(), mm_want) < 1e-2


def test_training_loss_decreases(ops):
    """A few Adam iterations on one fixed batch reduce the reference loss (end-to-end sanity of signs/scales)."""
    from tensorflow_yolo2_b200.trainer import Yolo2Trainer
    N, IS = 4, 96
    S = IS // 32
    st, _ = make_store(45, tame=True)
    rs = np.random.RandomState(3)
    img = torch.tensor(rs.randint(0, 256, (N, IS, IS, 3)).astype(np.uint8))
    tr = Yolo2Trainer(N, IS, 45, store=st, loss='v1', B=5, device='cuda:0')
    tr.set_labels(_labels_v1(rs, N, S, IS))
    losses = [float(tr.step(img)[4]) for _ in range(12)]
    print(losses)
    assert losses[-1] < 0.7 * losses[0]


def test_training_step_cuda_graph_equals_eager(ops):
    """Yolo2Trainer(use_cuda_graph=True): the captured step (forward, loss, backward, Adam with the step size read from
    device memory) replays to the same weights as eager launches -- three iterations, so the bias correction really changes
    between replays -- up to the float atomics of the split-K weight gradient."""
    from tensorflow_yolo2_b200.trainer import Yolo2Trainer
    N, IS = 4, 96
    S = IS // 32
    rs = np.random.RandomState(4)
    img = torch.tensor(rs.randint(0, 256, (N, IS, IS, 3)).astype(np.uint8))
    lab = _labels_v1(rs, N, S, IS)
    res = []
    for graph in (False, True):
        st, _ = make_store(45, tame=True)
        tr = Yolo2Trainer(N, IS, 45, store=st, loss='v1', B=5, device='cuda:0', use_cuda_graph=graph)
        tr.set_labels(lab)
        losses = [float(tr.step(img)[4]) for _ in range(3)]
        torch.cuda.synchronize()
        assert (tr.graph is not None) == graph and tr.iteration == 3
        assert float(tr.grads.abs().max()) == 0.0                      # cleared by the update kernel
        res.append((losses, tr.params.clone(), tr.adam_v.clone(), st[tr.layers[5]['bn']['moving_mean']].clone()))
        if graph:
            assert tr.launches_per_step > 200
    (l0, p0, v0, mm0), (l1, p1, v1, mm1) = res
    print(l0, l1)
    np.testing.assert_allclose(l0, l1, rtol=2e-2)
    # Adam's first steps move every weight by ~lr whatever the gradient's size, so agreement is judged against that scale
    assert float((p0 - p1).abs().mean()) < 3e-4 and float((p0 - p1).abs().max()) < 1e-2
    np.testing.assert_allclose(mm0.cpu().numpy(), mm1.cpu().numpy(), rtol=2e-2, atol=1e-3)


def test_ddp_two_gpus_matches_mean_of_shard_gradients():
    """World-size-2 NCCL run of tools/ddp_check.py (skipped on a single-GPU box)."""
    import subprocess
    import sys
    import os
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(root, 'tools', 'ddp_check.py')],
                       capture_output=True, text=True, timeout=600, cwd=root)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and 'OK' in r.stdout


@pytest.mark.parametrize('pool,dy_bf16,C,ld', [(True, True, 64, 64), (False, False, 128, 128), (True, False, 32, 64), (False, True, 96, 128)])
def test_bn_bwd_apply_rows_equals_generic(ops, monkeypatch, pool, dy_bf16, C, ld):
    """The row-walking bn_bwd apply kernel == the generic one, bit for bit (incl. zero padding columns), and so is the
    pooled row-walking forward affine."""
    import torch
    rs = np.random.RandomState(C + ld)
    N, H, W = 2, 12, 10
    M = N * H * W
    cu = lambda a, dt=None: torch.tensor(a).cuda() if dt is None else torch.tensor(a).cuda().to(dt)
    h = cu((rs.randn(M, ld) * 2).astype(np.float32))
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    dy = cu(rs.randn(N, Ho, Wo, C).astype(np.float32), torch.bfloat16 if dy_bf16 else None)
    mean, var = ops.bn_stats(h, C, ld=ld)
    gamma, beta = cu(rs.uniform(0.5, 1.5, C).astype(np.float32)), cu(rs.randn(C).astype(np.float32))
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv('Y2_BN_BWD_GENERIC', '1')
            monkeypatch.setenv('Y2_AFFINE_GENERIC', '1')
            ops.reload_env()                     # the launchers cache the Y2_* switches
        dg, db, dh = ops.bn_leaky_pool_bwd(h, dy, mean, var, gamma, beta, N, H, W, C, ldh=ld, leaky=True, pool=pool, ld_dh=ld)
        fwd = None
        if C % 8 == 0:
            fwd = ops.affine_leaky_pool(h, N, H, W, C, ldx=ld, sub=mean, scale=gamma, shift=beta, leaky=True, pool=pool, out_bf16=True)
        outs.append((dg.clone(), db.clone(), dh.clone(), fwd))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2].float() + 0.0, outs[1][2].float() + 0.0)          # (+0.0: -0 == 0 in the padding columns)
    assert float(outs[0][2].float().abs().sum()) > 0
    if outs[0][3] is not None:
        assert torch.equal(outs[0][3], outs[1][3])


# ---- the ImageNet classifier's training graph (imagenet_train_darknet.py:46-61) ---------------------------------------------
@pytest.mark.parametrize('N,H,C', [(5, 7, 1000), (3, 1, 37), (2, 3, 40)])
def test_softmax_xent_vs_torch(ops, N, H, C):
    """y2_softmax_xent_fwd_bwd = average pool + sparse softmax cross-entropy + reduce_mean + accuracy + d loss / d (pre-pool
    map), against torch float64 autograd."""
    rs = np.random.RandomState(N * 100 + C)
    net = (rs.randn(N, H, H, C) * 3).astype(np.float32)
    labels = rs.randint(0, C, N).astype(np.int32)
    net[0, :, :, labels[0]] += 50.0                             # image 0 is classified correctly for sure
    x = torch.tensor(net, dtype=torch.float64, requires_grad=True)
    logits = x.mean(dim=(1, 2))
    losses = torch.nn.functional.cross_entropy(logits, torch.tensor(labels, dtype=torch.long), reduction='none')
    losses.mean().backward()
    arg = cu(net) if H > 1 else cu(net.reshape(N, C))
    r = ops.softmax_xent(arg, cu(labels), want_grad=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(r['logits'].cpu().numpy(), logits.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(r['losses'].cpu().numpy(), losses.detach().numpy(), rtol=1e-5, atol=1e-5)
    want_acc = float((logits.argmax(dim=1).numpy() == labels).mean())
    assert abs(float(r['terms'][0]) - float(losses.mean())) < 1e-5 * max(1.0, float(losses.mean()))
    assert abs(float(r['terms'][1]) - want_acc) < 1e-6 and float(r['correct'][0]) == 1.0
    assert rel_l2(r['dnet'].cpu().numpy().reshape(N, H, H, C), x.grad.numpy()) < 1e-5


def test_momentum_step_matches_tf_formula(ops):
    rs = np.random.RandomState(0)
    n = 4096 + 64
    p, g, a = (rs.randn(n).astype(np.float32) for _ in range(3))
    P, G, A = cu(p), cu(g), cu(a)
    ops.momentum_step(P, G, A, 0.001, 0.9, grad_scale=0.5, zero_grad=True)
    a2 = 0.9 * a.astype(np.float64) + 0.5 * g
    np.testing.assert_allclose(A.cpu().numpy(), a2, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(P.cpu().numpy(), p - 0.001 * a2, rtol=1e-6, atol=1e-7)
    assert float(G.abs().max()) == 0.0


def test_classifier_training_step_vs_oracle(ops):
    """Yolo2Trainer(loss='softmax', optimizer='momentum') -- one iteration of imagenet_train_darknet.py:110-113 -- at 96x96
    (3x3 final map), batch 4, 40 classes: loss / accuracy / logits gradient against torch on the product's own net, the whole
    chain against the oracle's autograd (direction), the logits layer's backward layer-locally, and the momentum update."""
    from tensorflow_yolo2_b200.trainer import Yolo2Trainer
    N, IS, C = 4, 96, 40
    st, layers = make_store(C, tame=True, classifier=True)
    assert len(layers) == 19 and layers[-1]['W'] == 'darknet19/Variable_36' \
        and layers[-1]['bn']['gamma'] == 'darknet19/batch_normalization_18/gamma'           # darknet.py:114 continues the core's numbering
    core_p, head_p = oracle_params(st, layers)
    rs = np.random.RandomState(1)
    img = rs.uniform(-1, 1, (N, IS, IS, 3)).astype(np.float32)          # ilsvrc_cls.get(): already x/255*2-1
    labels = rs.randint(0, C, N)
    tr = Yolo2Trainer(N, IS, store=st, loss='softmax', num_class=C, optimizer='momentum', lr=0.001, momentum=0.9, device='cuda:0')
    p_before = [{k: v.clone() for k, v in tr.P[li].items()} for li in range(len(layers))]
    tr.set_class_labels(labels)
    cap = {}
    terms = tr.step(torch.tensor(img), capture=cap)
    torch.cuda.synchronize()
    lab_t = torch.tensor(labels, dtype=torch.long)
    loss_fn = lambda net: torch.nn.functional.cross_entropy(net.mean(dim=(1, 2)), lab_t)
    want_loss, grads, stats, net = O.train_step_reference(img, core_p, head_p, loss_fn, bf16_operands=True)
    got_net = tr.acts[-1].cpu().numpy()
    print('net rel_l2 %.3g  loss %.6g vs %.6g' % (rel_l2(got_net, net), float(terms[0]), want_loss))
    assert rel_l2(got_net, net) < 8e-2 and abs(float(terms[0]) - want_loss) / abs(want_loss) < 5e-2
    for li in (18, 12, 0):
        a, b = tr.G[li]['W'].cpu().numpy().ravel().astype(np.float64), grads[li]['W'].ravel()
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
        print('layer %2d dW cosine vs whole-chain oracle %.3f' % (li + 1, cos))
        assert cos > 0.8
    # loss, accuracy and d loss / d net on the product's own net output
    x = torch.tensor(got_net, dtype=torch.float64, requires_grad=True)
    l = loss_fn(x)
    l.backward()
    assert abs(float(terms[0]) - float(l)) < 1e-5 * float(l)
    assert abs(float(terms[1]) - float((x.detach().mean(dim=(1, 2)).argmax(dim=1) == lab_t).double().mean())) < 1e-6
    assert rel_l2(cap[18].cpu().numpy(), x.grad.numpy()) < 1e-5
    # the logits layer (1x1, 1024 -> C, BN, leaky), layer-locally
    li = 18
    xin = tr.acts[li - 1].double().cpu().requires_grad_(True)
    q = {k: p_before[li][k].double().cpu().requires_grad_(True) for k in ('W', 'b', 'gamma', 'beta')}
    q['mm'], q['mv'] = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    y, _, _ = O.conv_bn_layer(xin, q, True, torch.float64, bf16_operands=True)
    (y * cap[li].double().cpu()).sum().backward()
    for key in ('W', 'gamma', 'beta'):
        assert rel_l2(tr.G[li][key].cpu().numpy(), q[key].grad.numpy()) < 1e-2, key
    assert rel_l2(cap[li - 1].float().cpu().numpy(), xin.grad.numpy()) < 1e-2
    # MomentumOptimizer, first step: accum = g, p -= lr * g
    for li in (0, 9, 18):
        for key in ('W', 'gamma', 'beta'):
            g = tr.G[li][key].cpu().numpy().astype(np.float64)
            assert rel_l2(tr.P[li][key].cpu().numpy(), p_before[li][key].cpu().numpy() - 0.001 * g) < 1e-6
    st_names = tr.optimizer_state()
    assert 'darknet19/Variable_36/Momentum' in st_names and 'beta1_power' not in st_names


def test_classifier_training_loss_decreases_graph(ops):
    """A few momentum iterations on one fixed batch, CUDA-graphed, reduce the cross-entropy."""
    from tensorflow_yolo2_b200.trainer import Yolo2Trainer
    N, IS, C = 8, 64, 16
    st, _ = make_store(C, tame=True, classifier=True)
    rs = np.random.RandomState(2)
    img = torch.tensor(rs.uniform(-1, 1, (N, IS, IS, 3)).astype(np.float32))
    tr = Yolo2Trainer(N, IS, store=st, loss='softmax', num_class=C, optimizer='momentum', lr=0.01, device='cuda:0',
                      use_cuda_graph=True)
    tr.set_class_labels(rs.randint(0, C, N))
    losses = [float(tr.step(img)[0]) for _ in range(15)]
    print(losses)
    assert tr.graph is not None and losses[-1] < 0.8 * losses[0]

