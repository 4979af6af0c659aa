"""The two drop-in entry scripts (reference: src/pascal/pascal_detect_darknet.py, pascal_train_darknet.py)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.fixture()
def fresh_env(tmp_path, monkeypatch):
    from tensorflow_yolo2_b200 import config as cfg
    from tensorflow_yolo2_b200 import variables
    monkeypatch.setattr(cfg, 'ROOT_DIR', str(tmp_path))            # ckpts/ tensorboard/ go to a scratch dir
    monkeypatch.setattr(cfg, 'PASCAL_PATH', str(tmp_path / 'data' / 'VOCdevkit'))
    monkeypatch.setattr(cfg, 'COMPUTE', 'bf16')
    variables.reset_default_store(seed=0)
    yield cfg
    variables.reset_default_store(seed=0)


def test_detect_script_runs_and_matches_oracle_decode(fresh_env, golden_dir):
    from oracle import yolo2_oracle as O
    from tensorflow_yolo2_b200.pascal import pascal_detect_darknet as script
    img = os.path.join(golden_dir, 'testImg2.jpg')
    dets = script.main(['pascal_detect_darknet.py', img, '--no-show'])
    assert isinstance(dets, list)
    # every detection is (x0, y0, w, h, class, conf) with conf > 0.5 and a valid class (random weights: any count)
    for d in dets:
        assert len(d) == 6 and d[5] > 0.5 and 0 <= d[4] < 20


def test_train_script_synthetic_three_iterations(fresh_env, capsys):
    from tensorflow_yolo2_b200.pascal import pascal_train_darknet as script
    tr = script.main(['pascal_train_darknet.py', '--synthetic', '--iters', '10'])
    torch.cuda.synchronize()
    out = capsys.readouterr().out
    assert 'iter 10/10, total loss:' in out
    assert tr.iteration == 10 and tr.N == 24 and tr.OF == 30
    assert np.isfinite(float(tr.terms[4]))


def test_detect_script_detections_equal_oracle_decode_of_its_own_net(fresh_env, golden_dir):
    """The detect script end to end (image read / resize / preprocess, builders, reshape, show_yolo_detection) against the
    oracle's restatement of net_utils.py:393-421 (decode + the reference's integer draw arithmetic) on the script's own
    network output: same boxes, classes and confidences, in the reference's loop order."""
    from PIL import Image
    from oracle import yolo2_oracle as O
    from tensorflow_yolo2_b200 import config as cfg
    from tensorflow_yolo2_b200.pascal import pascal_detect_darknet as script
    img = os.path.join(golden_dir, 'testImg2.jpg')
    dets, predicts = script.main(['pascal_detect_darknet.py', img, '--no-show'], return_predicts=True)
    assert predicts.shape == (1, cfg.S, cfg.S, 5 * cfg.B + 20)
    im_w, im_h = Image.open(img).size
    for thresh in (0.5, 0.0):
        from tensorflow_yolo2_b200.yolo2_nets.net_utils import decode_yolo_detection
        got = decode_yolo_detection(predicts, im_w, im_h, 20, cfg.S, cfg.B, object_thresh=thresh)
        want = O.draw_list_ref_v1(O.decode_ref_v1(predicts[0], cfg.S, cfg.B, 20, thresh), im_w, im_h)
        assert len(got) == len(want)
        if thresh == 0.5:
            assert [tuple(d[:5]) for d in dets] == [tuple(int(v) for v in w[:5]) for w in want]
        for g, w in zip(got, want):
            assert tuple(g[:5]) == tuple(int(v) for v in w[:5])          # integer pixel arithmetic: exact
            assert np.float32(g[5]) == np.float32(w[5])
    assert len(got) > 0                                                  # thresh 0 -> every predictor with conf > 0


def test_train_script_snapshot_and_resume_restores_adam_state(fresh_env, capsys):
    """pascal_train_darknet.py:83-86,111-114: snapshot at iteration N, then a second run resumes at N + 1 with the weights,
    BN moving statistics AND Adam's slot variables / step count the Saver would have restored."""
    from tensorflow_yolo2_b200 import variables
    from tensorflow_yolo2_b200.pascal import pascal_train_darknet as script
    tr = script.main(['pascal_train_darknet.py', '--synthetic', '--iters', '4', '--snapshot-every', '4'])
    torch.cuda.synchronize()
    assert 'Model saved in file' in capsys.readouterr().out
    params, m, v = tr.params.clone(), tr.adam_m.clone(), tr.adam_v.clone()
    mm = tr.store[tr.layers[3]['bn']['moving_mean']].clone()
    assert float(m.abs().sum()) > 0 and float(v.abs().sum()) > 0
    variables.reset_default_store(seed=123)                                # a fresh process: different initial weights
    tr2 = script.main(['pascal_train_darknet.py', '--synthetic', '--iters', '0'])
    out = capsys.readouterr().out
    assert 'Restored.' in out
    assert tr2.iteration == 4
    assert torch.equal(tr2.params, params) and torch.equal(tr2.adam_m, m) and torch.equal(tr2.adam_v, v)
    assert torch.equal(tr2.store[tr2.layers[3]['bn']['moving_mean']], mm)
    # and the next step is bit-identical to continuing the first run (same batch, same state)
    img = np.random.RandomState(5).uniform(-1, 1, (tr.N, tr.IS, tr.IS, 3))
    lab = np.zeros((tr.N, tr.S, tr.S, 25))
    lab[:, 3, 3, 0] = 1
    lab[:, 3, 3, 1:5] = [100, 100, 50, 60]
    lab[:, 3, 3, 7] = 1
    for t in (tr, tr2):
        t.set_labels(lab)
        t.step(img)
    torch.cuda.synchronize()
    assert tr.iteration == tr2.iteration == 5
    # (the weight-gradient kernel reduces its split-K partials with float atomics, so two runs agree to rounding, not bits)
    assert float((tr.params - params).abs().max()) > 5e-4                  # the step moved the weights by ~lr ...
    assert float((tr.params - tr2.params).abs().max()) < 5e-5              # ... the same way in both runs
    fresh = script.Yolo2Trainer(tr.N, tr.IS, tr.OF, store=variables.VariableStore(seed=0), loss='v1', B=tr.B)
    fresh.params.copy_(params)                                             # same weights but NO optimizer state: differs
    fresh.set_labels(lab)
    fresh.step(img)
    assert float((fresh.params - tr.params).abs().max()) > 2e-4


def test_train_script_writes_scalar_and_histogram_summaries(fresh_env):
    """pascal_train_darknet.py:47,87-91,104 + net_utils.py:361-370: five scalars and five histograms per iteration."""
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from tensorflow_yolo2_b200.pascal import pascal_train_darknet as script
    from tensorflow_yolo2_b200.yolo2_nets.net_utils import SUMMARY_HISTOGRAMS, SUMMARY_SCALARS
    script.main(['pascal_train_darknet.py', '--synthetic', '--iters', '3'])
    torch.cuda.synchronize()
    tb_dir, _ = fresh_env.get_output_tb_dir('darknet19', 'voc_2007', val=False)
    acc = EventAccumulator(tb_dir, size_guidance={'histograms': 0, 'scalars': 0})
    acc.Reload()
    tags = acc.Tags()
    assert set(SUMMARY_SCALARS) <= set(tags['scalars'])
    assert set(SUMMARY_HISTOGRAMS) <= set(tags['histograms'])
    assert [e.step for e in acc.Scalars('total_loss')] == [1, 2, 3]
    h = acc.Histograms('iou')
    assert len(h) == 3 and h[0].histogram_value.num == 24 * 7 * 7 * 2          # every cell and predictor (unmasked)
    assert 0.0 <= h[0].histogram_value.min and h[0].histogram_value.max <= 1.0


def test_imagenet_test_and_predict_scripts(fresh_env, golden_dir, capsys):
    """Drop-ins for src/imagenet/imagenet_test_darknet.py and imagenet_predict_darknet.py on a synthetic database: same
    loop, prints and return values; the logits of the predict script equal the oracle's classifier on the same weights
    (raw 0..255 pixels, like the reference feeds them)."""
    import cv2
    from oracle import yolo2_oracle as O
    from tensorflow_yolo2_b200 import variables
    from tensorflow_yolo2_b200.imagenet import imagenet_predict_darknet as predict
    from tensorflow_yolo2_b200.imagenet import imagenet_test_darknet as test
    fresh_env.COMPUTE = 'fp32'                                   # exact path: raw 0..255 inputs through 19 layers
    acc = test.main(['imagenet_test_darknet.py', '--synthetic', '100', '--batches', '2'])
    out = capsys.readouterr().out
    assert '######batch number: 2' in out and 'batch 2/2, acc:' in out and '###########validation accuracy:' in out
    assert 0.0 <= acc <= 1.0
    variables.reset_default_store(seed=0)
    img = os.path.join(golden_dir, 'testImg1.jpg')
    preds, probs = predict.main(['imagenet_predict_darknet.py', img, '--synthetic', '50'])
    assert len(preds) == 5 and np.all(np.diff(probs) <= 0)
    st = variables.default_store()
    names = st.names()
    core, cls = [], None
    for li in range(19):
        w, b = 'darknet19/Variable' + ('' if li == 0 else '_%d' % (2 * li)), 'darknet19/Variable_%d' % (2 * li + 1)
        bn = 'darknet19/batch_normalization' + ('' if li == 0 else '_%d' % li) + '/'
        g = lambda n: torch.tensor(np.asarray(st[n].cpu() if hasattr(st[n], 'cpu') else st[n]))
        p = dict(W=g(w), b=g(b), gamma=g(bn + 'gamma'), beta=g(bn + 'beta'), mm=g(bn + 'moving_mean'), mv=g(bn + 'moving_variance'))
        if li < 18:
            core.append(p)
        else:
            cls = p
    assert len(names) == 19 * 6
    x = cv2.resize(cv2.imread(img), (224, 224)).astype(np.float32)[None]
    want = O.darknet19_classifier_forward(torch.tensor(x), core, cls, training=False, dtype=torch.float64).numpy()[0]
    order = np.argsort(-want, kind='stable')[:5]
    np.testing.assert_allclose(probs, want[order], rtol=2e-3)
    assert list(preds) == list(order)


def test_imagenet_train_script_synthetic(fresh_env, capsys):
    """Drop-in for src/imagenet/imagenet_train_darknet.py on a synthetic database: the reference's log lines, a validation
    pass (is_training = 0) every --val-every iterations, the epoch snapshot with the Momentum slots, and a resumed run that
    picks the snapshot up and continues at the next epoch (:80-98)."""
    from tensorflow_yolo2_b200 import variables
    from tensorflow_yolo2_b200.imagenet import imagenet_train_darknet as script
    argv = ['imagenet_train_darknet.py', '--synthetic', '64', '--batch', '16', '--iters', '6', '--val-every', '3']
    tr, hist = script.main(argv)
    torch.cuda.synchronize()
    out = capsys.readouterr().out
    assert 'epoch 1, iter 1/4, training loss:' in out and out.count('###validation loss:') == 2
    assert 'No darknet19 snapshot' in out and 'Model saved in file:' in out
    assert len(hist) == 6 and all(np.isfinite(l) and 0.0 <= a <= 1.0 for l, a in hist)
    assert abs(hist[0][0] - np.log(1000.0)) < 3.0                # random init: cross-entropy around ln(1000)
    ck = os.path.join(fresh_env.get_ckpts_dir('darknet19', 'ilsvrc_2017_cls'), 'train_epoch_0.ckpt.npz')
    data = np.load(ck)
    assert 'darknet19/Variable_36' in data.files and 'darknet19/Variable_36/Momentum' in data.files
    w_saved = np.array(data['darknet19/Variable_36'])            # (the resumed run rewrites the file)
    data.close()
    assert tr.graph is not None and tr.optimizer == 'momentum' and tr.iteration == 6
    # resume: the snapshot is found, the epoch continues from the file name
    variables.reset_default_store(seed=0)
    tr2, hist2 = script.main(argv[:5] + ['--iters', '1'])
    out = capsys.readouterr().out
    assert 'Restorining model snapshots from' in out and 'Restored.' in out and 'epoch 1, iter 1/4' in out
    assert tr2.iteration == 2 and np.isfinite(hist2[0][0])       # y2_iteration = 1 came back with the snapshot
    # the restored weights are the snapshot's (taken after iteration 0's update) moved by one more momentum step
    assert np.abs(tr2.P[18]['W'].cpu().numpy() - w_saved).max() < 0.05

