"""Validation accuracy of the darknet19 ImageNet classifier -- drop-in for the reference's
src/imagenet/imagenet_test_darknet.py (:21-67): batches of 50 validation images through `darknet19(is_training=False)`,
arg-max accuracy per batch, the running mean at the end.  `sess.run([logits, accuracy])` becomes a direct call on CUDA
tensors executed by libyolo2_b200.so (the 18 core layers + 1x1 conv 1024 -> 1000 + 7x7 average pool, darknet.py:61-123).

    python tensorflow_yolo2_b200/imagenet/imagenet_test_darknet.py [--synthetic N] [--batches K]
"""
import os
import sys

import numpy as np
import torch

FILE_DIR = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(FILE_DIR, '..', '..'))

from tensorflow_yolo2_b200.img_dataset.ilsvrc2017_cls import ilsvrc_cls                    # noqa: E402
from tensorflow_yolo2_b200.utils.timer import Timer                                         # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.darknet import darknet19                              # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.net_utils import get_ordered_ckpts, restore_checkpoint   # noqa: E402


def build_and_restore(imdb, image_size=224):
    """darknet19 graph + the newest snapshot (:27,46-50).  Variables exist after the first builder call, so build once on a
    dummy batch, restore, and let the callers run with reuse=True."""
    dummy = torch.zeros((1, image_size, image_size, 3), dtype=torch.float32, device='cuda')
    darknet19(dummy, is_training=False)
    ckpts = get_ordered_ckpts(None, imdb, 'darknet19')
    if ckpts:
        print('Restorining model snapshots from {:s}'.format(ckpts[-1]))
        restore_checkpoint(str(ckpts[-1]))
        print('Restored.')
    else:
        print('No darknet19 snapshot found for {:s}: running with freshly initialised variables'.format(imdb.name))
    return ckpts


def main(argv):
    synthetic = int(argv[argv.index('--synthetic') + 1]) if '--synthetic' in argv else 0
    imdb = ilsvrc_cls('val', batch_size=50, synthetic=synthetic)
    assert 0 == (imdb.image_num % imdb.batch_size)
    print("######image number:", imdb.image_num)
    print("######batch number:", imdb.total_batch)
    build_and_restore(imdb)
    n_batches = int(argv[argv.index('--batches') + 1]) if '--batches' in argv else imdb.total_batch

    T = Timer()
    accumulated_acc = 0.0
    accumulated_time = 0.0
    for i in range(n_batches):
        images, labels = imdb.get()
        T.tic()
        x = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)).cuda()      # the float32 placeholder (:25)
        logits = darknet19(x, is_training=False, reuse=True)
        pred = logits.argmax(dim=1).cpu().numpy().astype(np.int32)                       # tf.argmax(logits, 1) (:32)
        accuracy_value = float(np.mean(pred == labels.astype(np.int32)))
        _time = T.toc(average=False)
        print("batch {:d}/{:d}, acc: {:3f}, time: {:2f}sec".format(i + 1, n_batches, accuracy_value, _time))
        accumulated_acc += accuracy_value
        accumulated_time += _time
    print("###########validation accuracy:", (accumulated_acc / float(n_batches)))
    print("###########average time per batch:", (accumulated_time / float(n_batches)))
    return accumulated_acc / float(n_batches)


if __name__ == '__main__':
    main(sys.argv)
