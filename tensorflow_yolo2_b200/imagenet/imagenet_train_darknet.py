"""Train the darknet19 classifier on ILSVRC-2017 -- drop-in for the reference's src/imagenet/imagenet_train_darknet.py.

Same graph and loop as the reference (:46-61, :106-138): `darknet19(input, is_training)` logits, sparse softmax
cross-entropy averaged over the batch, the BN moving-average UPDATE_OPS, `MomentumOptimizer(0.001, 0.9).minimize(loss)`,
arg-max accuracy; ten epochs from the restored one, a validation batch of 64 every 25 iterations (is_training = 0: moving
statistics), loss / accuracy scalars to the train and val writers, a snapshot `train_epoch_<E>.ckpt` every two epochs.
One `sess.run([train_op, loss, accuracy, merged], ...)` is `Yolo2Trainer(loss='softmax', optimizer='momentum').step` on
libyolo2_b200.so (tcgen05 convolutions forward / data gradient / weight gradient, fused BN kernels, one softmax-xent kernel
that also writes the gradient of the pre-pool map, one momentum kernel over the flat parameter arena).

    python tensorflow_yolo2_b200/imagenet/imagenet_train_darknet.py [--synthetic N] [--iters K] [--batch B] [--val-every V]
    torchrun --nproc-per-node 8 ... imagenet_train_darknet.py        # data parallel: NCCL gradient all-reduce

Not reproduced (host-side data plumbing): the random rotation / crop / colour augmentation and the
multi-process prefetcher of ilsvrc2017_cls_multithread.py, and the child process that prefetches validation batches (:24-41) --
the validation batch is read inline.  --synthetic N runs on an in-memory database of N random images (the reference asserts
when the dataset is absent, and so does this script without the flag).  With no snapshot on disk the reference fails at
`ckpts[-1]` (:81-92); this script starts from freshly initialised variables at epoch 1 instead.
"""
import os
import sys

import numpy as np
import torch

FILE_DIR = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(FILE_DIR, '..', '..'))

from tensorflow_yolo2_b200 import config as cfg                                            # noqa: E402
from tensorflow_yolo2_b200 import ops                                                      # noqa: E402
from tensorflow_yolo2_b200.img_dataset.ilsvrc2017_cls import ilsvrc_cls                    # noqa: E402
from tensorflow_yolo2_b200.trainer import Yolo2Trainer                                      # noqa: E402
from tensorflow_yolo2_b200.utils.timer import Timer                                         # noqa: E402
from tensorflow_yolo2_b200.variables import default_store                                   # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.darknet import darknet19                              # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.net_utils import get_ordered_ckpts, restore_checkpoint, save_checkpoint   # noqa: E402

LEARNING_RATE, MOMENTUM = 0.001, 0.9          # :58
VAL_BATCH = 64                                # :31
VAL_EVERY = 25                                # :118
EPOCHS = 10                                   # :106


def _arg(argv, flag, default, cast=int):
    return cast(argv[argv.index(flag) + 1]) if flag in argv else default


def evaluate(images, labels):
    """sess.run([loss, accuracy], {is_training: 0}) (:120-122): the builder with the moving statistics, then the same loss kernel
    on the ready-made logits."""
    x = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)).cuda()
    logits = darknet19(x, is_training=False, reuse=True).float().contiguous()
    lab = torch.from_numpy(np.asarray(labels).astype(np.int32)).cuda()
    t = ops.softmax_xent(logits, lab)['terms'].cpu().numpy()
    return float(t[0]), float(t[1])


def main(argv):
    synthetic = _arg(argv, '--synthetic', 0)
    batch = _arg(argv, '--batch', cfg.BATCH_SIZE)
    val_every = _arg(argv, '--val-every', VAL_EVERY)
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    if world > 1:
        np.random.seed(cfg.__dict__.get('SHUFFLE_SEED', 0))            # one common shuffle, then disjoint rank shards
    imdb = ilsvrc_cls('train', batch_size=batch, synthetic=synthetic)  # (data_aug / multithread: see the module docstring)
    val_imdb = ilsvrc_cls('val', batch_size=VAL_BATCH, synthetic=min(synthetic, 2 * VAL_BATCH) if synthetic else 0)
    if world > 1:
        imdb.gt_labels = imdb.gt_labels[rank::world]
        imdb.image_num = len(imdb.gt_labels)
        imdb.total_batch = int(np.ceil(imdb.image_num / float(imdb.batch_size)))
    CKPTS_DIR = cfg.get_ckpts_dir('darknet19', imdb.name)

    # graph (:46-61): darknet19 logits + softmax cross-entropy + MomentumOptimizer(0.001, 0.9).minimize
    store = default_store()
    trainer = Yolo2Trainer(batch, 224, store=store, loss='softmax', num_class=imdb.num_class, optimizer='momentum',
                           lr=LEARNING_RATE, momentum=MOMENTUM, device=torch.device('cuda', local),
                           use_cuda_graph='--no-graph' not in argv)
    darknet19(torch.zeros((1, 224, 224, 3), dtype=torch.float32, device='cuda'), is_training=False, reuse=True)   # (checks the names)

    # load previous models (:80-98)
    ckpts = get_ordered_ckpts(None, imdb, 'darknet19')
    if ckpts:
        print('Restorining model snapshots from {:s}'.format(ckpts[-1]))
        restore_checkpoint(str(ckpts[-1]))
        slots = trainer.load_optimizer_state(str(ckpts[-1]) + '.npz')
        print('Restored.' + ('' if slots else ' (no Momentum slots in the snapshot: accumulators start at zero)'))
        old_epoch = int(str(ckpts[-1]).split('_')[-1][:-5])
        imdb.epoch = old_epoch + 1
    elif rank == 0:
        print('No darknet19 snapshot for {:s}: training from freshly initialised variables'.format(imdb.name))

    train_writer = val_writer = None
    if rank == 0:
        try:                                                   # tf.summary.FileWriter (:72-73)
            from torch.utils.tensorboard import SummaryWriter
            tb_train, tb_val = cfg.get_output_tb_dir('darknet19', imdb.name)
            train_writer, val_writer = SummaryWriter(tb_train), SummaryWriter(tb_val)
        except Exception:
            train_writer = val_writer = None

    n_iters = _arg(argv, '--iters', imdb.total_batch * EPOCHS + 1)
    T = Timer()
    history = []
    for i in range(n_iters):
        T.tic()
        images, labels = imdb.get()
        trainer.set_class_labels(labels)
        terms = trainer.step(images).cpu().numpy()             # [loss, accuracy] of this batch
        loss_value, acc_value = float(terms[0]), float(terms[1])
        _time = T.toc(average=False)
        history.append((loss_value, acc_value))
        if rank == 0:
            print('epoch {:d}, iter {:d}/{:d}, training loss: {:.3}, training acc: {:.3}, take {:.2}s'
                  .format(imdb.epoch, (i + 1) % imdb.total_batch, imdb.total_batch, loss_value, acc_value, _time))

        if (i + 1) % val_every == 0 and rank == 0:
            T.tic()
            val_images, val_labels = val_imdb.get()
            val_loss_value, val_acc_value = evaluate(val_images, val_labels)
            _val_time = T.toc(average=False)
            print('###validation loss: {:.3}, validation acc: {:.3}, take {:.2}s'.format(val_loss_value, val_acc_value, _val_time))
            global_step = imdb.epoch * imdb.total_batch + (i % imdb.total_batch)
            if train_writer is not None:
                for w, (l, a) in ((train_writer, (loss_value, acc_value)), (val_writer, (val_loss_value, val_acc_value))):
                    w.add_scalar('loss', l, global_step)
                    w.add_scalar('accuracy', a, global_step)

        if i % (imdb.total_batch * 2) == 0:
            trainer.sync_moving_statistics()                   # (a collective: every rank takes part)
            if rank == 0:
                save_path = save_checkpoint(os.path.join(CKPTS_DIR, cfg.TRAIN_SNAPSHOT_PREFIX + '_epoch_' + str(imdb.epoch - 1)
                                                         + '.ckpt'), store, extra=trainer.optimizer_state())
                print("Model saved in file: %s" % save_path)
    for w in (train_writer, val_writer):
        if w is not None:
            w.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return trainer, history


if __name__ == '__main__':
    main(sys.argv)
