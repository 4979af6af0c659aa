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
