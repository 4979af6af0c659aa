"""Host-side logic of the drop-in surface that needs no GPU: the label encoder + flip (a9), TF-style variable naming,
checkpoint discovery / iteration parsing / warm start (f1).  Everything here is PRODUCT code (tensorflow_yolo2_b200.*)
checked against goldens produced by the reference's own source (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest


@pytest.fixture()
def cfg_scratch(tmp_path, monkeypatch):
    from tensorflow_yolo2_b200 import config as cfg
    from tensorflow_yolo2_b200 import variables
    monkeypatch.setattr(cfg, 'ROOT_DIR', str(tmp_path))
    variables.reset_default_store(seed=0)
    yield cfg
    variables.reset_default_store(seed=0)


# ---- a9: pascal_voc.load_pascal_annotation (pascal_voc.py:125-165) and the flip of prepare() (:69-86) -------------------
@pytest.mark.parametrize('IS,S', [(224, 7), (416, 13), (608, 19)])
def test_encode_annotation_vs_reference_golden(golden_dir, IS, S):
    import cv2
    from tensorflow_yolo2_b200.img_dataset.pascal_voc import VOC_CLASSES, encode_annotation
    g = np.load(os.path.join(golden_dir, 'ref_labels.npz'))
    im = cv2.imread(os.path.join(golden_dir, 'testImg2.jpg'))
    c2i = dict(zip(VOC_CLASSES, range(20)))
    label, num = encode_annotation(os.path.join(golden_dir, 'testImg2Anno.xml'), im.shape[0], im.shape[1], IS, S, c2i)
    assert num == 2 and label.dtype == np.float64
    np.testing.assert_array_equal(label, g['label_%d_%d' % (IS, S)])        # float64 host arithmetic: bit-exact


def _voc_tree(root, golden_dir):
    import shutil
    base = os.path.join(root, 'VOCdevkit', 'VOC2007')
    for sub in ('JPEGImages', 'Annotations', os.path.join('ImageSets', 'Main')):
        os.makedirs(os.path.join(base, sub))
    shutil.copy(os.path.join(golden_dir, 'testImg2.jpg'), os.path.join(base, 'JPEGImages', '000001.jpg'))
    shutil.copy(os.path.join(golden_dir, 'testImg2Anno.xml'), os.path.join(base, 'Annotations', '000001.xml'))
    with open(os.path.join(base, 'ImageSets', 'Main', 'trainval.txt'), 'w') as f:
        f.write('000001\n')
    return os.path.join(root, 'VOCdevkit')


def test_pascal_voc_loader_get_and_flip_vs_reference_golden(golden_dir, tmp_path, monkeypatch):
    """The class itself (prepare / load_labels / get / image_read) on a one-image VOCdevkit built from the reference's
    fixtures, flipped records included -- the golden comes from the reference's class on the same tree."""
    from tensorflow_yolo2_b200 import config as cfg
    from tensorflow_yolo2_b200.img_dataset.pascal_voc import pascal_voc
    g = np.load(os.path.join(golden_dir, 'ref_labels.npz'))
    monkeypatch.setattr(cfg, 'PASCAL_PATH', _voc_tree(str(tmp_path), golden_dir))
    monkeypatch.setattr(cfg, 'CACHE_PATH', str(tmp_path / 'cache'))
    monkeypatch.setattr(cfg, 'IMAGE_SIZE', 416)
    monkeypatch.setattr(cfg, 'S', 13)
    monkeypatch.setattr(cfg, 'FLIPPED', True)
    imdb = pascal_voc('trainval', batch_size=1, rebuild=True)
    assert len(imdb.gt_labels) == 2
    plain = [r for r in imdb.gt_labels if not r['flipped']][0]
    flip = [r for r in imdb.gt_labels if r['flipped']][0]
    np.testing.assert_array_equal(plain['label'], g['label_416_13'])
    np.testing.assert_array_equal(flip['label'], g['label_416_13_flipped'])
    im = imdb.image_read(plain['imname'], False)
    imf = imdb.image_read(flip['imname'], True)
    assert im.dtype == np.float32
    np.testing.assert_allclose([im.sum(dtype=np.float64), np.abs(im).sum(dtype=np.float64), im[5, 7, 1]],
                               g['image_416_checksum'], rtol=1e-6)
    np.testing.assert_allclose([imf.sum(dtype=np.float64), np.abs(imf).sum(dtype=np.float64), imf[5, 7, 1]],
                               g['image_416_flipped_checksum'], rtol=1e-6)
    np.testing.assert_array_equal(imf, im[:, ::-1, :])
    images, labels = imdb.get()                                   # float64 batches like the reference (:43-46)
    assert images.dtype == np.float64 and labels.dtype == np.float64 and images.shape == (1, 416, 416, 3)
    # the pickle cache is read back on the next construction
    imdb2 = pascal_voc('trainval', batch_size=1, rebuild=False)
    assert len(imdb2.gt_labels) == 2


# ---- TF-style auto-generated variable names -----------------------------------------------------------
def test_variable_scope_names_like_tf(cfg_scratch):
    from tensorflow_yolo2_b200 import variables
    from tensorflow_yolo2_b200.yolo2_nets.darknet import _variable_scope
    st = variables.default_store()
    names = []
    for _ in range(2):
        with _variable_scope('darknet19'):
            names.append(st.weight_variable([1, 1, 2, 2])[0])
            names.append(st.bias_variable([2])[0])
            names.append(st.batch_norm_variables(2)['gamma'])
        with _variable_scope('darknet19_detection'):
            with _variable_scope('conv1'):
                names.append(st.weight_variable([1, 1, 2, 2])[0])
    # second entry WITHOUT reuse opens <scope>_1 (what TF's name scope does for unnamed tf.Variable's)
    assert names == ['darknet19/Variable', 'darknet19/Variable_1', 'darknet19/batch_normalization/gamma',
                     'darknet19_detection/conv1/Variable',
                     'darknet19_1/Variable', 'darknet19_1/Variable_1', 'darknet19_1/batch_normalization/gamma',
                     'darknet19_detection_1/conv1/Variable']
    with _variable_scope('darknet19', reuse=True):
        assert st.weight_variable([1, 1, 2, 2])[0] == 'darknet19/Variable'


# ---- f1: checkpoint store / discovery / resume iteration / ImageNet warm start (net_utils.py:14-110) ---------------------
class _Imdb:
    name = 'voc_2007'


def test_checkpoint_discovery_resume_and_warm_start(cfg_scratch):
    import time
    from tensorflow_yolo2_b200 import variables
    from tensorflow_yolo2_b200.engine import create_variables
    from tensorflow_yolo2_b200.yolo2_nets import net_utils as nu
    cfg = cfg_scratch
    st = variables.default_store()
    create_variables(st, 30)
    assert len(st.names()) == 132
    imdb = _Imdb()
    # no snapshot at all -> 0, variables untouched
    before = {k: np.array(st[k]) for k in st.names()}
    assert nu.restore_darknet19_variables(None, imdb, 'darknet19', save_epoch=False) == 0
    # ImageNet warm start: a snapshot of the classifier (core variables + its own 19th layer) under ilsvrc_2017_cls
    inet = variables.VariableStore(seed=7)
    create_variables(inet, 30)
    core_names = [n for n in inet.names() if n.startswith('darknet19/')]
    wdir = cfg.get_ckpts_dir('darknet19', 'ilsvrc_2017_cls')
    arrays = {n: np.asarray(inet[n]) + 1.0 for n in core_names}
    arrays['darknet19/Variable_36'] = np.zeros((1, 1, 1024, 1000), np.float32)          # not in the detection graph
    np.savez(os.path.join(wdir, cfg.TRAIN_SNAPSHOT_PREFIX + '_epoch_88.ckpt.npz'), **arrays)
    open(os.path.join(wdir, cfg.TRAIN_SNAPSHOT_PREFIX + '_epoch_88.ckpt.meta'), 'w').close()
    assert nu.restore_darknet19_variables(None, imdb, 'darknet19', save_epoch=False) == 0     # warm start returns 0 (:101)
    for n in st.names():
        if n.startswith('darknet19/'):
            np.testing.assert_array_equal(np.asarray(st[n]), arrays[n])                  # intersection restored (:85-89)
        else:
            np.testing.assert_array_equal(np.asarray(st[n]), before[n])                  # the head keeps its initial values
    # two training snapshots: the newest by mtime wins and its iteration is parsed from the file name (:104-110)
    cdir = cfg.get_ckpts_dir('darknet19', imdb.name)
    for it, bump in ((40000, 0.0), (80000, 3.0)):
        for n in st.names():
            st[n] = np.asarray(st[n]) + bump
        nu.save_checkpoint(os.path.join(cdir, cfg.TRAIN_SNAPSHOT_PREFIX + '_iter_%d.ckpt' % it), st, extra={'beta1_power': np.float32(0.5)})
        time.sleep(0.02)
    want = {k: np.array(st[k]) for k in st.names()}
    for n in st.names():
        st[n] = np.zeros_like(np.asarray(st[n]))
    assert nu.latest_checkpoint(imdb, 'darknet19', save_epoch=False).endswith('_iter_80000.ckpt')
    assert nu.restore_darknet19_variables(None, imdb, 'darknet19', save_epoch=False) == 80000
    for n in st.names():
        np.testing.assert_array_equal(np.asarray(st[n]), want[n])
    assert [os.path.basename(p) for p in nu.get_ordered_ckpts(None, imdb, 'darknet19', save_epoch=False)] == \
        [cfg.TRAIN_SNAPSHOT_PREFIX + '_iter_40000.ckpt', cfg.TRAIN_SNAPSHOT_PREFIX + '_iter_80000.ckpt']


def test_ilsvrc_cls_synthetic_loader_matches_reference_read_path():
    """ilsvrc_cls (the ImageNet scripts' database, ilsvrc2017_cls_multithread.py:95-117,320-323,408-415) on its in-memory
    synthetic set: batches are cv2.resize((IS, IS)) -> float32 -> x/255*2-1 of the records in cursor order, labels 1-D, the
    epoch counter advances and the list is reshuffled when the cursor wraps; augmentation / prefetch options raise."""
    import cv2
    from tensorflow_yolo2_b200.img_dataset.ilsvrc2017_cls import ilsvrc_cls
    db = ilsvrc_cls('val', batch_size=6, image_size=64, synthetic=10)
    assert db.name == 'ilsvrc_2017_cls' and db.num_class == 1000 and db.image_num == 10 and db.total_batch == 2 and db.epoch == 1
    recs = [dict(r) for r in db.gt_labels[:6]]
    images, labels = db.get()
    assert images.shape == (6, 64, 64, 3) and labels.shape == (6,)
    for k, r in enumerate(recs):
        want = cv2.resize(db._synthetic[r['imname']], (64, 64)).astype(np.float32) / 255.0 * 2.0 - 1.0
        np.testing.assert_array_equal(images[k].astype(np.float32), want)
        assert labels[k] == r['label']
    assert images.min() >= -1.0 and images.max() <= 1.0
    db.get()                                           # records 6..9, wraps: epoch 2, cursor back at 2
    assert db.epoch == 2 and db.cursor == 2
    with pytest.raises(NotImplementedError):
        ilsvrc_cls('train', data_aug=True, synthetic=4)
    with pytest.raises(NotImplementedError):
        ilsvrc_cls('train', multithread=True, synthetic=4)
