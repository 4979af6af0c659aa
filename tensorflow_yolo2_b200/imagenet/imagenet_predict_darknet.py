"""Top-5 prediction for one image with the darknet19 ImageNet classifier -- drop-in for the reference's
src/imagenet/imagenet_predict_darknet.py (:19-65).  Like the reference, the image is resized to 224x224 and fed AS IS
(uint8 values 0..255 cast to float, no /255*2-1: :51-52,58,61), and the "probabilities" printed are the top-5 LOGITS
(tf.nn.top_k(logits, k=5), :26).  The reference hard-codes the image path (:48); it is argv[1] here and cv2.imshow /
waitKey (:55-56) are skipped.

    python tensorflow_yolo2_b200/imagenet/imagenet_predict_darknet.py image.jpg [--synthetic 50]
"""
import os
import sys

import cv2
import numpy as np
import torch

FILE_DIR = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(FILE_DIR, '..', '..'))

from tensorflow_yolo2_b200.img_dataset.ilsvrc2017_cls import ilsvrc_cls                    # noqa: E402
from tensorflow_yolo2_b200.imagenet.imagenet_test_darknet import build_and_restore          # noqa: E402
from tensorflow_yolo2_b200.utils.timer import Timer                                         # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.darknet import darknet19                              # noqa: E402


def main(argv):
    synthetic = int(argv[argv.index('--synthetic') + 1]) if '--synthetic' in argv else 0
    imdb = ilsvrc_cls('val', batch_size=1, synthetic=synthetic)
    build_and_restore(imdb)
    f = argv[1]
    image = cv2.imread(f)
    assert image is not None, 'cannot read {}'.format(f)
    image = cv2.resize(image, (224, 224))
    image = image.reshape((1, 224, 224, 3))

    T = Timer()
    T.tic()
    x = torch.from_numpy(image.astype(np.float32)).cuda()
    logits = darknet19(x, is_training=False, reuse=True)
    # tf.nn.top_k(logits, k=5): descending values, ties -> lower index first (a stable descending sort on the host)
    lg = logits[0].cpu().numpy()
    preds = np.argsort(-lg, kind='stable')[:5]
    probs = lg[preds]
    _time = T.toc(average=False)
    print("predictions:", [imdb.classes[i] for i in preds])
    print("probabilities:", probs)
    print("takes time:", _time)
    return preds, probs


if __name__ == '__main__':
    main(sys.argv)
