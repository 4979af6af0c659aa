"""Use trained darknet19 to detect -- drop-in for the reference's src/pascal/pascal_detect_darknet.py.

Same flow, constants and helper calls as the reference script (:21-63); what changes is what the calls run on:
`tf.placeholder` / `tf.Session.run` become direct calls on CUDA tensors, executed by libyolo2_b200.so.

    python tensorflow_yolo2_b200/pascal/pascal_detect_darknet.py [image_path]

The reference hard-codes the image path (":20 TODO: make the image path to be user input"): it is argv[1] here,
defaulting to the reference's tests/testImg2.jpg fixture.  Like the reference (:42), darknet19_detection is called
WITHOUT is_training, so the four head layers normalise with batch statistics even at inference.
"""
import os
import sys

import cv2
import numpy as np
import torch

FILE_DIR = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(FILE_DIR, '..', '..'))

from tensorflow_yolo2_b200 import config as cfg                                            # noqa: E402
from tensorflow_yolo2_b200.img_dataset.pascal_voc import pascal_voc                         # noqa: E402
from tensorflow_yolo2_b200.utils.timer import Timer                                         # noqa: E402,F401
from tensorflow_yolo2_b200.yolo2_nets.net_utils import (restore_checkpoint, restore_darknet19_variables,  # noqa: E402
                                                       show_yolo_detection)
from tensorflow_yolo2_b200.yolo2_nets.darknet import darknet19_core, darknet19_detection    # noqa: E402


def main(argv, return_predicts=False):
    """return_predicts=True also returns the [1,S,S,5B+C] network output the detections were decoded from (tests)."""
    image_path = argv[1] if len(argv) > 1 else os.path.join(cfg.ROOT_DIR, 'tests', 'golden', 'testImg2.jpg')
    IMAGE_SIZE, S, B = cfg.IMAGE_SIZE, cfg.S, cfg.B
    # create database instance (the reference needs VOCdevkit on disk even for detection, pascal_voc.py:36-39; the
    # class list is all that is used, so a missing devkit is tolerated here)
    imdb = pascal_voc('trainval', require_data=os.path.exists(cfg.PASCAL_PATH))
    NUM_CLASS = imdb.num_class

    # read in the test image (:34-38)
    image = cv2.imread(image_path)
    assert image is not None, 'cannot read {}'.format(image_path)
    image = cv2.resize(image, (IMAGE_SIZE, IMAGE_SIZE))
    image = image.astype(np.float32)
    image = (image / 255.0) * 2.0 - 1.0
    image = image.reshape((1, IMAGE_SIZE, IMAGE_SIZE, 3))
    input_data = torch.from_numpy(image).cuda()

    # Load from weight file or checkpoint (:54-60); variables are created on the first builder call, so build once
    # with `reuse=None`, restore, then run with reuse=True
    core_net = darknet19_core(input_data, is_training=False)
    final_conv_layer = darknet19_detection(core_net, 5 * B + NUM_CLASS)
    if os.path.isfile(cfg.darknet_pascal_weight_path + ".meta"):
        print('Restorining model from weight file {:s}'.format(cfg.darknet_pascal_weight_path))
        restore_checkpoint(cfg.darknet_pascal_weight_path)
        print('Restored.')
    else:
        _ = restore_darknet19_variables(None, imdb, net_name='darknet19', save_epoch=False)
    core_net = darknet19_core(input_data, is_training=False, reuse=True)
    final_conv_layer = darknet19_detection(core_net, 5 * B + NUM_CLASS, reuse=True)
    grid_net = final_conv_layer.reshape(-1, S, S, 5 * B + NUM_CLASS)

    predicts = grid_net.float().cpu().numpy()
    dets = show_yolo_detection(image_path, predicts, imdb, show='--no-show' not in argv)
    return (dets, predicts) if return_predicts else dets


if __name__ == '__main__':
    main(sys.argv)
