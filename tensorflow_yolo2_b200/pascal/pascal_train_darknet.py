"""Train darknet19 on Pascal VOC -- drop-in for the reference's src/pascal/pascal_train_darknet.py.

Same hyper-parameters and loop as the reference (:23-28, :96-114): ADD_ITER = 80000, BATCH_SIZE = 24, Adam with TF
defaults, all layers in training mode, a log line every 10 iterations, a snapshot every 40000.  One iteration of
the reference's `sess.run([merged, loss, train_op, ious, object_mask], ...)` is `Yolo2Trainer.step` (forward,
get_loss, backward, BN moving-average updates, Adam) on libyolo2_b200.so.

    python tensorflow_yolo2_b200/pascal/pascal_train_darknet.py [--iters N] [--synthetic] [--no-graph]
    torchrun --nproc-per-node 8 ... pascal_train_darknet.py      # data parallel: NCCL gradient all-reduce

--synthetic feeds random images/labels when VOCdevkit is not on disk (the reference asserts in that case,
pascal_voc.py:36-39 -- and so does this script without the flag).
"""
import os
import sys

import numpy as np
import torch

FILE_DIR = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(FILE_DIR, '..', '..'))

from tensorflow_yolo2_b200 import config as cfg                                            # noqa: E402
from tensorflow_yolo2_b200.img_dataset.pascal_voc import pascal_voc                         # noqa: E402
from tensorflow_yolo2_b200.trainer import Yolo2Trainer                                      # noqa: E402
from tensorflow_yolo2_b200.utils.timer import Timer                                         # noqa: E402
from tensorflow_yolo2_b200.variables import default_store                                   # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.net_utils import (add_summary, latest_checkpoint, merged_summary,   # noqa: E402
                                                       restore_darknet19_variables, save_checkpoint)

# set hyper parameters (:23-28)
ADD_ITER = 80000
BATCH_SIZE = 24
SNAPSHOT_EVERY = 40000


class _SyntheticVoc(pascal_voc):
    """Random batches with the label layout of pascal_voc.load_pascal_annotation (pascal_voc.py:125-165)."""

    def __init__(self, batch_size, rank=0):
        pascal_voc.__init__(self, 'trainval', batch_size=batch_size, require_data=False)
        self._rs = np.random.RandomState(rank)           # every rank draws its own batches

    def get(self):
        IS, S = self.image_size, self.cell_size
        images = self._rs.uniform(-1, 1, (self.batch_size, IS, IS, 3))
        labels = np.zeros((self.batch_size, S, S, 25))
        for n in range(self.batch_size):
            for _ in range(self._rs.randint(1, 4)):
                cx, cy = self._rs.uniform(0, IS - 1, 2)
                w, h = self._rs.uniform(10, IS * 0.7, 2)
                x_ind, y_ind = int(cx * S / IS), int(cy * S / IS)
                if labels[n, y_ind, x_ind, 0] == 1:
                    continue                                   # first object wins a cell (:159-160)
                labels[n, y_ind, x_ind, 0] = 1
                labels[n, y_ind, x_ind, 1:5] = [cx, cy, w, h]
                labels[n, y_ind, x_ind, 5 + self._rs.randint(0, 20)] = 1
        return images, labels


def main(argv):
    add_iter = int(argv[argv.index('--iters') + 1]) if '--iters' in argv else ADD_ITER
    snapshot_every = int(argv[argv.index('--snapshot-every') + 1]) if '--snapshot-every' in argv else SNAPSHOT_EVERY
    summary_every = int(argv[argv.index('--summary-every') + 1]) if '--summary-every' in argv else 1     # the reference: every iteration
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    IMAGE_SIZE, S, B = cfg.IMAGE_SIZE, cfg.S, cfg.B
    # create database instance
    if '--synthetic' in argv:
        imdb = _SyntheticVoc(BATCH_SIZE, rank)
    else:
        # data parallel: one common-seed shuffle, then rank r takes records r, r + world, ... (disjoint shards; the
        # reference is single-process and reshuffles with the global, unseeded np.random -- pascal_voc.py:56,84)
        if world > 1:
            np.random.seed(cfg.__dict__.get('SHUFFLE_SEED', 0))
        imdb = pascal_voc('trainval', batch_size=BATCH_SIZE, rebuild=cfg.REBUILD)
        if world > 1:
            imdb.gt_labels = imdb.gt_labels[rank::world]
            np.random.seed(cfg.__dict__.get('SHUFFLE_SEED', 0) + 1 + rank)
    NUM_CLASS = imdb.num_class
    CKPTS_DIR = cfg.get_ckpts_dir('darknet19', imdb.name)

    # graph: darknet19_core + darknet19_detection(core, 30) + get_loss + AdamOptimizer().minimize (:39-51)
    store = default_store()
    trainer = Yolo2Trainer(BATCH_SIZE, IMAGE_SIZE, 5 * B + NUM_CLASS, store=store, loss='v1', num_class=NUM_CLASS, B=B,
                           lambda_coord=float(cfg.LAMBDA_COORD), lambda_noobj=float(cfg.LAMBDA_NOOBJ),
                           device=torch.device('cuda', local), use_cuda_graph='--no-graph' not in argv)
    last_iter_num = restore_darknet19_variables(None, imdb, net_name='darknet19', save_epoch=False)
    if last_iter_num > 0:
        # tf.train.Saver() restores every global variable (:54,83): Adam's slots and beta powers come back with the weights
        slots = trainer.load_optimizer_state(latest_checkpoint(imdb, 'darknet19', save_epoch=False) + '.npz', iteration=last_iter_num)
        if not slots and rank == 0:
            print('snapshot holds no optimizer state: Adam restarts from zero moments at t = %d' % last_iter_num)

    writer = None
    if rank == 0:
        try:                                                   # tf.summary.FileWriter (:89-91)
            from torch.utils.tensorboard import SummaryWriter
            tb_dir, _ = cfg.get_output_tb_dir('darknet19', imdb.name, val=False)
            writer = SummaryWriter(tb_dir)
        except Exception:
            writer = None

    TOTAL_ITER = add_iter + last_iter_num
    T = Timer()
    T.tic()
    for i in range(last_iter_num + 1, TOTAL_ITER + 1):
        image, gt_labels = imdb.get()
        trainer.set_labels(gt_labels)
        terms = trainer.step(image)
        if writer is not None and i % summary_every == 0:
            # summary = sess.run(merged): 5 scalars + 5 histograms (net_utils.py:361-370, :47); add_summary every iteration (:104)
            summary = merged_summary(trainer.acts[-1], trainer.labels, terms, trainer.ious, NUM_CLASS, IMAGE_SIZE, S, B)
            add_summary(writer, summary, i)
        if i % 10 == 0:
            t = terms.cpu().numpy()                            # class, coord, object, noobject, total (:361-364)
        if i % 10 == 0 and rank == 0:
            _time = T.toc(average=False)
            print('iter {:d}/{:d}, total loss: {:.3}, take {:.2}s'.format(i, TOTAL_ITER, float(t[4]), _time))
            T.tic()
        if i % snapshot_every == 0:
            trainer.sync_moving_statistics()                   # (a collective: every rank takes part)
            if rank == 0:
                save_path = save_checkpoint(os.path.join(CKPTS_DIR, cfg.TRAIN_SNAPSHOT_PREFIX + '_iter_' + str(i) + '.ckpt'),
                                            store, extra=trainer.optimizer_state())
                print("Model saved in file: %s" % save_path)
    if writer is not None:
        writer.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return trainer


if __name__ == '__main__':
    main(sys.argv)
