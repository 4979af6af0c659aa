"""Loss / IoU / decode / checkpoint helpers with the reference's interface
(src/yolo2_nets/net_utils.py), backed by the CUDA kernels of libyolo2_b200.so.

    get_iou(boxes1, boxes2, scope='iou')                                  net_utils.py:222
    get_loss(net, labels, num_class, batch_size, image_size, S, B, OFFSET, scope)   :263
    show_yolo_detection(image_path, predict_output, imdb, object_thresh=0.5)        :375
    get_ordered_ckpts / get_ordered_ckpts_by_dbname / restore_darknet19_variables   :14-110

plus the pieces `north_star` names that the reference does not have:
    region_decode(net, anchors, ...)   and   nms(boxes, scores, ...)  -> detections.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from .. import config as cfg
from .. import ops
from ..variables import default_store

VOC_ANCHORS = np.array([[1.3221, 1.73145], [3.19275, 4.00944], [5.05587, 8.09892],
                        [9.47112, 4.84053], [11.2364, 10.0071]], dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# checkpoints (net_utils.py:14-110).  TF Saver files cannot be read without TensorFlow; snapshots
# here are `<prefix>_{iter|epoch}_<N>.ckpt.npz` keyed by the TF variable names, with an empty
# `<...>.ckpt.meta` marker so that the reference's glob + mtime discovery rule is unchanged.
# ---------------------------------------------------------------------------------------------
def _ordered(ckpts_dir, save_epoch):
    save_interval = 'epoch' if save_epoch else 'iter'
    sfiles = glob.glob(os.path.join(ckpts_dir, cfg.TRAIN_SNAPSHOT_PREFIX + '_' + save_interval + '_*.ckpt.meta'))
    sfiles.sort(key=os.path.getmtime)
    return [ss.replace('.meta', '') for ss in sfiles]


def get_ordered_ckpts(sess, imdb, net_name, save_epoch=True):
    """net_utils.py:14-34: snapshots of `net_name` on `imdb.name`, oldest first."""
    return _ordered(cfg.get_ckpts_dir(net_name, imdb.name), save_epoch)


def get_ordered_ckpts_by_dbname(sess, imdb_name, net_name, save_epoch=True):
    """net_utils.py:41-61."""
    return _ordered(cfg.get_ckpts_dir(net_name, imdb_name), save_epoch)


def save_checkpoint(path, store=None, extra=None):
    """Counterpart of tf.train.Saver.save (pascal_train_darknet.py:111-114): writes `path`.npz and
    the `path`.meta marker."""
    store = store or default_store()
    store.save_npz(path + '.npz', extra=extra)
    open(path + '.meta', 'w').close()
    return path


def restore_checkpoint(path, store=None):
    store = store or default_store()
    return store.load_npz(path + '.npz')


def latest_checkpoint(imdb, net_name='darknet19', save_epoch=True):
    """Path (without .npz) of the newest snapshot restore_darknet19_variables would pick, or None."""
    sfiles = get_ordered_ckpts(None, imdb, net_name, save_epoch=save_epoch)
    return str(sfiles[-1]) if sfiles else None


def restore_darknet19_variables(sess, imdb, net_name='darknet19', save_epoch=True):
    """net_utils.py:64-110.  No snapshot for this dataset -> warm-start from the newest ImageNet
    snapshot by variable-name intersection (the rest keep their initial values) and return 0;
    otherwise restore the newest snapshot and return the iteration parsed from its file name."""
    sfiles = get_ordered_ckpts(sess, imdb, net_name, save_epoch=save_epoch)
    if len(sfiles) == 0:
        imagenet_sfiles = get_ordered_ckpts_by_dbname(sess, 'ilsvrc_2017_cls', net_name, save_epoch=True)
        if imagenet_sfiles:
            print('Initializing new variables to train from imagenet trained model')
            print('Restorining model snapshots from {:s}'.format(imagenet_sfiles[-1]))
            restore_checkpoint(str(imagenet_sfiles[-1]))
        else:
            print('No snapshot found: keeping freshly initialised variables')
        return 0
    print('Restorining model snapshots from {:s}'.format(sfiles[-1]))
    restore_checkpoint(str(sfiles[-1]))
    print('Restored.')
    fnames = sfiles[-1].split('_')
    return int(fnames[-1][:-5])


# ---------------------------------------------------------------------------------------------
# IoU and loss
# ---------------------------------------------------------------------------------------------
def get_iou(boxes1, boxes2, scope='iou'):
    """net_utils.py:222-260.  [BATCH,S,S,B,4] x2 (x_center, y_center, w, h) -> [BATCH,S,S,B]."""
    return ops.iou(boxes1.float().contiguous(), boxes2.float().contiguous())


class LossResult(tuple):
    """(loss, ious, object_mask) like the reference, carrying the extra kernel outputs:
    .terms (class, coord, object, noobject, total) and .dnet (d loss / d net)."""
    terms = None
    dnet = None


class _YoloLossV1(torch.autograd.Function):
    """get_loss as one autograd node: forward and the analytic gradient come out of the SAME fused kernel (loss.cu); backward
    hands out d loss / d net scaled by the upstream gradient of the scalar loss -- so the reference's
    `optimizer.minimize(loss)` pattern (pascal_train_darknet.py:44-51) composes with torch autograd.  `ious` and
    `object_mask` are non-differentiable outputs (TF: the masks come from a comparison + cast; the IoU path's gradient is
    already inside d loss / d net)."""

    @staticmethod
    def forward(ctx, net, labels, S, B, num_class, image_size, lambda_coord, lambda_noobj):
        terms, ious, mask, dnet = ops.loss_v1(net, labels, S, B, num_class, image_size, lambda_coord, lambda_noobj)
        ctx.save_for_backward(dnet)
        ctx.mark_non_differentiable(ious, mask, terms, dnet)
        return terms[4].clone(), ious, mask, terms, dnet

    @staticmethod
    def backward(ctx, g_loss, g_ious, g_mask, g_terms, g_dnet):
        (dnet,) = ctx.saved_tensors
        g = g_loss.detach().to(torch.float32).contiguous()
        return ops.scale_by_device_scalar(dnet, g), None, None, None, None, None, None, None


def get_loss(net, labels, num_class, batch_size, image_size, S, B, OFFSET=None, scope='loss_layer'):
    """net_utils.py:263-372.  Returns (loss, ious, object_mask), differentiable w.r.t. `net`: `loss.backward()` (or
    torch.autograd.grad) delivers the gradient TF's autodiff would produce -- it comes out of the same fused kernel as the
    forward (also available directly as `.dnet`, with the five loss terms as `.terms`).  OFFSET must be the standard
    config.YOLO_GRID_OFFSET (offset[y,x,b] = x); it is generated in-kernel."""
    if OFFSET is not None:
        exp = cfg._grid_offset(S, B)
        if np.asarray(OFFSET).shape != exp.shape or not np.array_equal(np.asarray(OFFSET), exp):
            raise ValueError('get_loss: only the standard YOLO_GRID_OFFSET is supported')
    net4 = net.reshape(batch_size, S, S, num_class + 5 * B)
    if net4.dtype != torch.float32:
        net4 = ops.cast(net4.detach(), torch.float32)
    net4 = net4 if net4.is_contiguous() else net4.contiguous()
    if not isinstance(labels, torch.Tensor):      # pascal_voc.get() hands out float64 numpy; the placeholder is float32 (:35)
        labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.float32))
    labels = labels.to(net.device).reshape(batch_size, S, S, 5 + num_class)
    labels = (labels if labels.dtype == torch.float32 else labels.float()).contiguous()
    loss, ious, mask, terms, dnet = _YoloLossV1.apply(net4, labels, S, B, num_class, float(image_size),
                                                      float(cfg.LAMBDA_COORD), float(cfg.LAMBDA_NOOBJ))
    res = LossResult((loss, ious, mask))
    res.terms, res.dnet = terms, dnet
    return res


# ---------------------------------------------------------------------------------------------
# summaries: tf.summary.scalar / tf.summary.histogram of get_loss (net_utils.py:361-370) + 'total_loss'
# (pascal_train_darknet.py:47), merged by tf.summary.merge_all() (:87) and written every iteration (:104)
# ---------------------------------------------------------------------------------------------
SUMMARY_SCALARS = ('class_loss', 'object_loss', 'noobject_loss', 'coord_loss', 'total_loss')
SUMMARY_HISTOGRAMS = ('boxes_delta_x', 'boxes_delta_y', 'boxes_delta_w', 'boxes_delta_h', 'iou')


def merged_summary(net, labels, terms, ious, num_class, image_size, S, B):
    """What `sess.run(merged)` evaluates in the reference's training loop: the five loss scalars and the five histogram
    tensors (unmasked box deltas of every cell / predictor, and the IoUs).  net [N,S,S,C+5B] and labels [N,S,S,5+C] are the
    CUDA tensors get_loss saw; terms = (class, coord, object, noobject, total) from the loss kernel.  Returns
    dict(scalars={name: float}, histograms={name: float32 ndarray}) -- one device->host copy of ~N*S*S*B*5 floats."""
    deltas = ops.loss_v1_box_deltas(net.float().contiguous(), labels.float().contiguous(), S, B, num_class, image_size)
    t = terms.detach().cpu().numpy()
    d = deltas.cpu().numpy()
    return dict(scalars=dict(class_loss=float(t[0]), coord_loss=float(t[1]), object_loss=float(t[2]),
                             noobject_loss=float(t[3]), total_loss=float(t[4])),
                histograms=dict(boxes_delta_x=d[..., 0], boxes_delta_y=d[..., 1], boxes_delta_w=d[..., 2],
                                boxes_delta_h=d[..., 3], iou=ious.detach().cpu().numpy()))


def add_summary(writer, summary, step):
    """train_writer.add_summary(summary, i) (pascal_train_darknet.py:104) on a torch.utils.tensorboard SummaryWriter."""
    for name, v in summary['scalars'].items():
        writer.add_scalar(name, v, step)
    for name, v in summary['histograms'].items():
        writer.add_histogram(name, v, step)


# ---------------------------------------------------------------------------------------------
# decode + display (net_utils.py:375-439)
# ---------------------------------------------------------------------------------------------
def decode_yolo_detection(predict_output, im_w, im_h, num_class, S=None, B=None, object_thresh=0.5):
    """The NumPy half of show_yolo_detection (:393-421) on the GPU: returns the list of
    (upper_left_x, upper_left_y, w, h, class_index, confidence) in the reference's loop order, with
    its integer pixel arithmetic (int() truncation, py2 floor division)."""
    S = S or cfg.S
    B = B or cfg.B
    p = torch.as_tensor(predict_output).reshape(1, S, S, num_class + B * 5).float().cuda().contiguous()
    boxes, conf, keep, cls = [t[0].cpu().numpy() for t in ops.decode_ref_v1(p, S, B, num_class, object_thresh)]
    out = []
    for c in range(S):
        for r in range(S):
            for i in range(B):
                if keep[c, r, i]:
                    px = int(boxes[c, r, i, 0] * im_w)
                    py = int(boxes[c, r, i, 1] * im_h)
                    pw = int(boxes[c, r, i, 2] * im_w)
                    ph = int(boxes[c, r, i, 3] * im_h)
                    out.append((px - pw // 2, py - ph // 2, pw, ph, int(cls[c, r]), float(conf[c, r, i])))
    return out


def show_yolo_detection(image_path, predict_output, imdb, object_thresh=0.5, show=True):
    """net_utils.py:375-439.  Prints the boxes like the reference; draws them when matplotlib is
    importable (it is optional here).  Returns the detection list."""
    from PIL import Image
    im = np.array(Image.open(image_path), dtype=np.uint8)
    im_h, im_w, _ = im.shape
    dets = decode_yolo_detection(predict_output, im_w, im_h, imdb.num_class, cfg.S, cfg.B, object_thresh)
    for (x0, y0, w, h, c, cf) in dets:
        print("predicted bounding boxes: ({:d}, {:d}), width:{:d}, height:{:d}".format(x0, y0, w, h))
    if show:
        try:
            import matplotlib.pyplot as plt
            import matplotlib.patches as patches
        except ImportError:
            return dets
        fig, ax = plt.subplots(1)
        ax.imshow(im)
        for (x0, y0, w, h, c, cf) in dets:
            ax.add_patch(patches.Rectangle((x0, y0), w, h, linewidth=1, edgecolor='r', facecolor='none'))
            ax.text(x0, y0, imdb.classes[int(c)] + ":" + str(np.float32(cf)), color='r')
        plt.show()
    return dets


# ---------------------------------------------------------------------------------------------
# region decode + NMS (absent from the reference)
# ---------------------------------------------------------------------------------------------
def region_decode(net, anchors=VOC_ANCHORS, num_class=20, thresh=0.3):
    """net [N,S,S,A*(5+C)] -> boxes [N,S*S*A,4] (cx,cy,w,h in [0,1]), scores [N,S*S*A,C]."""
    an = torch.as_tensor(np.asarray(anchors, dtype=np.float32)).to(net.device)
    return ops.decode_region(net.float().contiguous(), an, num_class, thresh)


def nms(boxes, scores, score_thresh=0.3, iou_thresh=0.45, max_keep=None):
    """Per-class greedy NMS -> (keep_idx [N,C,max_keep] int32, keep_count [N,C] int32)."""
    return ops.nms(boxes, scores, score_thresh, iou_thresh, max_keep)


def detections_from_keep(boxes, scores, keep_idx, keep_count):
    """Host-side gather: list (per image) of (class, score, cx, cy, w, h), kept boxes only."""
    b, s = boxes.cpu().numpy(), scores.cpu().numpy()
    ki, kc = keep_idx.cpu().numpy(), keep_count.cpu().numpy()
    out = []
    for n in range(b.shape[0]):
        dets = []
        for k in range(s.shape[2]):
            for t in range(min(int(kc[n, k]), ki.shape[2])):
                i = int(ki[n, k, t])
                dets.append((k, float(s[n, i, k])) + tuple(float(v) for v in b[n, i]))
        out.append(dets)
    return out
