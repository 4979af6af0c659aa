#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE over a torch-backed TF1 shim.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/ref_*.npz

How: the reference is Python-2 / TensorFlow-1.x and cannot be imported as is.  This script
  1. reads the reference modules from /root/reference/src (never copied into the repo),
  2. applies token-level py2->py3 fixes IN MEMORY (print statements, xrange, `range(S) * S * B`,
     py2 integer `/`), and
  3. exec()s them against fake `tensorflow`, `matplotlib` modules defined below.  The fake `tf`
     implements the handful of primitives the path uses (conv2d SAME, max_pool, batch_normalization,
     stack/transpose/tile/reduce_*, maximum/minimum/clip, ...) on torch float64 tensors, so the
     reference's graph-building code runs eagerly and torch.autograd gives the gradient of the
     reference's own loss graph.

So the *structure* of every golden (slicing, offsets, masks, reductions, layer plan, variable
creation order, label encoding, draw-loop integer math) is produced by reference code; the
*primitive semantics* (SAME padding, BN eps=1e-3/momentum .99, ...) are this shim's statement of
TF behaviour.  Weights come from tensorflow_yolo2_b200.variables.truncated_normal with
RandomState(0) in creation order, which the product's VariableStore reproduces, so goldens carry
inputs and outputs only.
"""
import os
import re
import sys
import types
import shutil
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, '..', '..'))
REF = '/root/reference'
sys.path.insert(0, REPO)
from tensorflow_yolo2_b200.variables import truncated_normal  # noqa: E402

DT = torch.float64


# ------------------------------------------------------------------------------------------
# py2 -> py3 source fixes (in memory)
# ------------------------------------------------------------------------------------------
def py2to3(src):
    lines = src.split('\n')
    out = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r'^(\s*)print\s+(?!\()(.*)$', ln) or re.match(r'^(\s*)print\s+(\(.*\)\s*%.*)$', ln)
        if m:
            indent, body = m.group(1), m.group(2)
            while body.rstrip().endswith('\\'):
                i += 1
                body = body.rstrip()[:-1] + ' ' + lines[i].strip()
            out.append('%sprint(%s)' % (indent, body))
        else:
            out.append(ln)
        i += 1
    src = '\n'.join(out)
    src = src.replace('xrange(', 'range(')
    src = src.replace('np.array(range(S) * S * B)', 'np.array(list(range(S)) * S * B)')
    src = re.sub(r'(predict_[wh]) / 2\b', r'\1 // 2', src)      # py2 int / int
    return src


def load_ref_module(name, relpath, extra_globals=None):
    with open(os.path.join(REF, relpath)) as f:
        src = py2to3(f.read())
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, relpath)
    if extra_globals:
        mod.__dict__.update(extra_globals)
    sys.modules[name] = mod
    exec(compile(src, mod.__file__, 'exec'), mod.__dict__)
    return mod


# ------------------------------------------------------------------------------------------
# fake tensorflow
# ------------------------------------------------------------------------------------------
class Graph:
    def __init__(self, seed=0):
        self.rng = np.random.RandomState(seed)
        self.scope = []
        self.counters = {}
        self.variables = {}          # name -> tensor (creation order)

    def unique(self, base):
        key = ('/'.join(self.scope), base)
        k = self.counters.get(key, 0)
        self.counters[key] = k + 1
        leaf = base if k == 0 else '%s_%d' % (base, k)
        return '/'.join(self.scope + [leaf])


G = Graph()


def reset_graph(seed=0):
    global G
    G = Graph(seed)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=DT)


class _Scope:
    def __init__(self, name_or_scope, default_name=None, values=None, reuse=None):
        self.name = name_or_scope if name_or_scope is not None else default_name

    def __enter__(self):
        G.scope.append(self.name)
        return self

    def __exit__(self, *a):
        G.scope.pop()
        return False


def make_tf():
    tf = types.ModuleType('tensorflow')
    tf.float32 = DT
    tf.bool = torch.bool
    tf.variable_scope = _Scope

    def Variable(initial):
        name = G.unique('Variable')
        v = _t(initial).clone()
        G.variables[name] = v
        return v
    tf.Variable = Variable
    tf.truncated_normal = lambda shape, stddev=1.0: _t(truncated_normal(shape, stddev, G.rng))

    def constant(value, shape=None, dtype=None):
        if shape is not None:
            return torch.full(list(shape), float(np.float32(value)), dtype=DT)   # fp32 variable
        return _t(value)
    tf.constant = constant

    nn = types.ModuleType('tensorflow.nn')

    def conv2d(x, W, strides, padding):
        assert padding == 'SAME' and strides == [1, 1, 1, 1]
        k = W.shape[0]
        y = F.conv2d(x.permute(0, 3, 1, 2), W.permute(3, 2, 0, 1), stride=1, padding=k // 2)
        return y.permute(0, 2, 3, 1)
    nn.conv2d = conv2d

    def max_pool(x, ksize, strides, padding):
        assert ksize == [1, 2, 2, 1] and strides == [1, 2, 2, 1]
        assert x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0      # SAME == VALID on even maps
        return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    nn.max_pool = max_pool
    tf.nn = nn

    layers = types.ModuleType('tensorflow.layers')

    def batch_normalization(x, center=True, scale=True, training=False):
        scope = G.unique('batch_normalization')
        C = x.shape[-1]
        names = {}
        for leaf, val in (('gamma', 1.0), ('beta', 0.0), ('moving_mean', 0.0),
                          ('moving_variance', 1.0)):
            names[leaf] = scope + '/' + leaf
            if names[leaf] not in G.variables:
                G.variables[names[leaf]] = torch.full([C], val, dtype=DT)
        gamma, beta = G.variables[names['gamma']], G.variables[names['beta']]
        if training:
            mean = x.mean(dim=(0, 1, 2))
            var = x.var(dim=(0, 1, 2), unbiased=False)
        else:
            mean, var = G.variables[names['moving_mean']], G.variables[names['moving_variance']]
        return (x - mean) * torch.rsqrt(var + 1e-3) * gamma + beta
    layers.batch_normalization = batch_normalization
    tf.layers = layers

    def _bin(fn):
        return lambda a, b: fn(_t(a), _t(b))
    tf.maximum = _bin(torch.maximum)
    tf.minimum = _bin(torch.minimum)
    tf.add = _bin(torch.add)
    tf.matmul = _bin(torch.matmul)
    tf.reshape = lambda x, shape: _t(x).reshape([int(s) for s in shape])
    tf.stack = lambda xs, axis=0: torch.stack(list(xs), dim=axis)
    tf.transpose = lambda x, perm: _t(x).permute(*perm)
    tf.tile = lambda x, reps: _t(x).repeat(*reps)
    tf.square = lambda x: _t(x) ** 2
    tf.sqrt = lambda x: torch.sqrt(_t(x))
    tf.expand_dims = lambda x, axis: _t(x).unsqueeze(axis)
    tf.ones_like = lambda x, dtype=None: torch.ones_like(_t(x))
    tf.cast = lambda x, dtype: x.to(DT)
    tf.clip_by_value = lambda x, lo, hi: torch.clamp(x, lo, hi)

    def _axes(axis):
        return tuple(axis) if isinstance(axis, (list, tuple)) else axis
    tf.reduce_sum = lambda x, axis=None, name=None: x.sum() if axis is None else x.sum(dim=_axes(axis))
    tf.reduce_mean = lambda x, axis=None, name=None: x.mean() if axis is None else x.mean(dim=_axes(axis))
    tf.reduce_max = lambda x, axis, keep_dims=False: x.max(dim=axis, keepdim=keep_dims).values

    summary = types.ModuleType('tensorflow.summary')
    summary.scalar = lambda *a, **k: None
    summary.histogram = lambda *a, **k: None
    tf.summary = summary

    contrib = types.ModuleType('tensorflow.contrib')
    contrib.slim = types.ModuleType('slim')
    contrib.framework = types.ModuleType('framework')
    tf.contrib = contrib
    return tf


class FakeAx:
    def __init__(self):
        self.rects, self.texts = [], []

    def imshow(self, im):
        pass

    def add_patch(self, r):
        self.rects.append(r)

    def text(self, x, y, s, color=None):
        self.texts.append((x, y, s))


LAST_AX = None


def make_matplotlib():
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    patches = types.ModuleType('matplotlib.patches')

    def subplots(n):
        global LAST_AX
        LAST_AX = FakeAx()
        return None, LAST_AX
    plt.subplots = subplots
    plt.show = lambda: None
    patches.Rectangle = lambda xy, w, h, **kw: (xy[0], xy[1], w, h)
    mpl.pyplot, mpl.patches = plt, patches
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = plt
    sys.modules['matplotlib.patches'] = patches


def install():
    sys.modules['tensorflow'] = make_tf()
    make_matplotlib()
    cfg = load_ref_module('config', 'src/config.py')
    nu = load_ref_module('yolo2_nets_net_utils', 'src/yolo2_nets/net_utils.py')
    dk = load_ref_module('yolo2_nets_darknet', 'src/yolo2_nets/darknet.py')
    return cfg, nu, dk


def set_grid(cfg, image_size, S, B):
    """Re-evaluate config.py:34,38-42 for another grid (the reference hard-codes 224/7/2)."""
    cfg.IMAGE_SIZE, cfg.S, cfg.B = image_size, S, B
    off = np.array(list(range(S)) * S * B)
    cfg.YOLO_GRID_OFFSET = np.transpose(np.reshape(off, (B, S, S)), (1, 2, 0))


# ------------------------------------------------------------------------------------------
def gen_iou(nu, out):
    rs = np.random.RandomState(11)
    b1 = rs.uniform(0, 1, (2, 3, 3, 2, 4))
    b2 = rs.uniform(0, 1, (2, 3, 3, 2, 4))
    b2[0, 0, 0, 0] = b1[0, 0, 0, 0]                       # identical boxes -> 1.0
    b1[0, 0, 1, 0] = [.5, .5, 1, 1]; b2[0, 0, 1, 0] = [.5, .5, .5, .5]       # 0.25
    b1[0, 0, 2, 0] = [.1, .1, .1, .1]; b2[0, 0, 2, 0] = [.8, .8, .1, .1]     # disjoint
    b1[0, 1, 0, 0] = [.5, .5, .2, .2]; b2[0, 1, 0, 0] = [.6, .5, .2, .2]     # 1/3
    b1[0, 1, 1, 0] = [.5, .5, 0, 0]                                         # zero area
    iou = nu.get_iou(_t(b1), _t(b2)).numpy()
    np.savez(os.path.join(out, 'ref_iou.npz'), boxes1=b1, boxes2=b2, iou=iou)
    print('iou KATs', iou[0, 0, 0, 0], iou[0, 0, 1, 0], iou[0, 0, 2, 0], iou[0, 1, 0, 0], iou[0, 1, 1, 0])


def voc_tmpdir():
    d = tempfile.mkdtemp(prefix='voc_')
    base = os.path.join(d, 'VOCdevkit', 'VOC2007')
    for sub in ('JPEGImages', 'Annotations', os.path.join('ImageSets', 'Main')):
        os.makedirs(os.path.join(base, sub))
    shutil.copy(os.path.join(REF, 'tests', 'testImg2.jpg'), os.path.join(base, 'JPEGImages', '000001.jpg'))
    shutil.copy(os.path.join(REF, 'tests', 'testImg2Anno.xml'), os.path.join(base, 'Annotations', '000001.xml'))
    with open(os.path.join(base, 'ImageSets', 'Main', 'trainval.txt'), 'w') as f:
        f.write('000001\n')
    return d


def gen_labels(cfg, out):
    res = {}
    for (IS, S) in ((224, 7), (416, 13), (608, 19)):
        d = voc_tmpdir()
        set_grid(cfg, IS, S, 2)
        cfg.PASCAL_PATH = os.path.join(d, 'VOCdevkit')
        cfg.CACHE_PATH = os.path.join(d, 'cache')
        pv = load_ref_module('img_dataset_pascal_voc', 'src/img_dataset/pascal_voc.py')
        imdb = pv.pascal_voc('trainval', batch_size=1, rebuild=True)
        images, labels = imdb.get()
        res['label_%d_%d' % (IS, S)] = labels[0]
        res['image_%d_checksum' % IS] = np.array([images.sum(), np.abs(images).sum(), images[0, 5, 7, 1]])
        if IS == 224:
            res['image_224'] = images[0].astype(np.float32)
        if IS == 416:
            # cfg.FLIPPED = True (config.py:47 ships False): prepare() appends the mirrored records (pascal_voc.py:69-86)
            cfg.FLIPPED = True
            imdb_f = pv.pascal_voc('trainval', batch_size=1, rebuild=True)
            cfg.FLIPPED = False
            rec = [r for r in imdb_f.gt_labels if r['flipped']]
            assert len(imdb_f.gt_labels) == 2 and len(rec) == 1
            res['label_416_13_flipped'] = rec[0]['label']
            im_f = imdb_f.image_read(rec[0]['imname'], True)
            res['image_416_flipped_checksum'] = np.array([im_f.sum(), np.abs(im_f).sum(), im_f[5, 7, 1]])
        shutil.rmtree(d)
    np.savez_compressed(os.path.join(out, 'ref_labels.npz'), **res)
    set_grid(cfg, 224, 7, 2)
    return res


def gen_loss(cfg, nu, labels_res, out):
    res = {}
    rs = np.random.RandomState(5)
    cases = []
    lab7 = labels_res['label_224_7'][None]
    cases.append(('katA', np.zeros((1, 7, 7, 30)), lab7, 224, 7, 2))
    cases.append(('katB', np.full((1, 7, 7, 30), 0.5), lab7, 224, 7, 2))
    cases.append(('katC', np.random.RandomState(0).uniform(0, 1, (1, 7, 7, 30)), lab7, 224, 7, 2))

    def rand_labels(N, S, IS, C=20):
        lab = np.zeros((N, S, S, 5 + C))
        for n in range(N):
            for _ in range(rs.randint(1, 4)):
                cx, cy = rs.uniform(0, IS, 2)
                w, h = rs.uniform(20, 300, 2)
                j, i = int(cx * S / IS), int(cy * S / IS)
                if lab[n, i, j, 0] == 1:
                    continue
                lab[n, i, j, 0] = 1
                lab[n, i, j, 1:5] = [cx, cy, w, h]
                lab[n, i, j, 5 + rs.randint(0, C)] = 1
        return lab
    cases.append(('rand7', rs.uniform(-0.2, 1.0, (3, 7, 7, 30)), rand_labels(3, 7, 224), 224, 7, 2))
    lab13 = np.concatenate([labels_res['label_416_13'][None], rand_labels(3, 13, 416)], 0)
    cases.append(('rand13', rs.uniform(-0.2, 1.0, (4, 13, 13, 45)), lab13, 416, 13, 5))
    cases.append(('rand19', rs.uniform(-0.2, 1.0, (2, 19, 19, 45)), rand_labels(2, 19, 608), 608, 19, 5))
    for name, net, lab, IS, S, B in cases:
        set_grid(cfg, IS, S, B)
        x = _t(net).clone().requires_grad_(True)
        loss, ious, mask = nu.get_loss(x, _t(lab), num_class=20, batch_size=net.shape[0],
                                       image_size=IS, S=S, B=B, OFFSET=cfg.YOLO_GRID_OFFSET)
        loss.backward()
        res[name + '_net'], res[name + '_labels'] = net, lab
        res[name + '_cfg'] = np.array([IS, S, B])
        res[name + '_loss'] = loss.detach().numpy()
        res[name + '_ious'] = ious.detach().numpy()
        res[name + '_mask'] = mask.detach().numpy()
        res[name + '_dnet'] = x.grad.numpy()
        print('loss', name, float(loss), float(mask.sum()))
    set_grid(cfg, 224, 7, 2)
    np.savez_compressed(os.path.join(out, 'ref_loss.npz'), **res)


class FakeImdb:
    num_class = 20
    classes = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair',
               'cow', 'diningtable', 'dog', 'horse', 'motorbike', 'person', 'pottedplant',
               'sheep', 'sofa', 'train', 'tvmonitor')


def gen_decode(cfg, nu, out):
    res = {}
    img = os.path.join(REF, 'tests', 'testImg2.jpg')        # 353 wide x 500 high
    import io
    import contextlib
    for name, S, B, seed in (('s7', 7, 2, 3), ('s13', 13, 5, 4)):
        set_grid(cfg, 224 if S == 7 else 416, S, B)
        pred = np.random.RandomState(seed).uniform(0, 1, (1, S, S, 20 + 5 * B)).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            nu.show_yolo_detection(img, pred, FakeImdb(), object_thresh=0.5)
        rects = np.array(LAST_AX.rects, dtype=np.int64).reshape(-1, 4)
        res[name + '_pred'] = pred
        res[name + '_rects'] = rects
        res[name + '_texts'] = np.array([t[2] for t in LAST_AX.texts])
        print('decode', name, len(rects), 'boxes; first', rects[:2].tolist(), LAST_AX.texts[:1])
    res['im_wh'] = np.array([353, 500])
    set_grid(cfg, 224, 7, 2)
    np.savez_compressed(os.path.join(out, 'ref_decode.npz'), **res)


def gen_darknet(dk, out):
    res = {}
    for name, hw, nb, of in (('d64_30', 64, 2, 30), ('d96_125', 96, 2, 125)):
        reset_graph(seed=0)
        x = np.random.RandomState(21).uniform(-1, 1, (nb, hw, hw, 3)).astype(np.float32)
        with torch.no_grad():
            core = dk.darknet19_core(_t(x), is_training=False)      # pascal_detect_darknet.py:41
            outp = dk.darknet19_detection(core, of)                 # :42 (is_training default True)
        res[name + '_x'] = x
        res[name + '_core'] = core.numpy()
        res[name + '_out'] = outp.numpy()
        res[name + '_varnames'] = np.array(list(G.variables.keys()))
        w0 = G.variables['darknet19/Variable'].numpy()
        res[name + '_w0'] = w0
        print('darknet', name, core.shape, outp.shape, 'nvars', len(G.variables),
              float(core.abs().max()), float(outp.abs().max()))
    # training-mode core on a tiny map, to pin batch-stat BN through the whole stack
    reset_graph(seed=0)
    x = np.random.RandomState(22).uniform(-1, 1, (2, 32, 32, 3)).astype(np.float32)
    with torch.no_grad():
        core = dk.darknet19_core(_t(x), is_training=True)
        outp = dk.darknet19_detection(core, 30)
    res['t32_x'], res['t32_core'], res['t32_out'] = x, core.numpy(), outp.numpy()
    np.savez_compressed(os.path.join(out, 'ref_darknet.npz'), **res)


def main():
    assert os.path.isdir(REF), 'needs /root/reference (build container only)'
    out = HERE
    cfg, nu, dk = install()
    gen_iou(nu, out)
    lab = gen_labels(cfg, out)
    gen_loss(cfg, nu, lab, out)
    gen_decode(cfg, nu, out)
    gen_darknet(dk, out)
    # data fixtures (not source): config-1 input image and the only annotation pair
    for f in ('testImg1.jpg', 'testImg2.jpg', 'testImg2Anno.xml'):
        shutil.copy(os.path.join(REF, 'tests', f), os.path.join(out, f))
        os.chmod(os.path.join(out, f), 0o644)
    print('done')


if __name__ == '__main__':
    main()
