"""Pascal VOC 2007 image database with the interface of the reference's
src/img_dataset/pascal_voc.py (class `pascal_voc`: .name, .classes, .num_class, .get(), ...).

Host-side I/O; the part that is on the hot path is the *label layout* consumed by get_loss:
label[y, x] = [1, cx, cy, w, h, one-hot(20)] with the box in pixels of the resized image and the
first object claiming a cell (pascal_voc.py:137-163).  `encode_annotation` is that rule as a pure
function so it can be tested without a VOCdevkit on disk.
"""
from __future__ import annotations

import copy
import os
import pickle
import xml.etree.ElementTree as ET

import numpy as np

from .. import config as cfg

VOC_CLASSES = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow',
               'diningtable', 'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train',
               'tvmonitor')


def encode_annotation(xml_path, im_h, im_w, image_size, cell_size, class_to_ind):
    """pascal_voc.py:133-165.  Returns (label [S,S,25] float64, number of objects in the file)."""
    sy, sx = float(image_size) / im_h, float(image_size) / im_w
    hi = image_size - 1
    label = np.zeros((cell_size, cell_size, 5 + len(class_to_ind)))
    objs = ET.parse(xml_path).findall('object')
    for obj in objs:
        bb = obj.find('bndbox')
        # VOC pixels are 1-based; scale into the resized image and clamp to [0, IS-1]
        x1, y1, x2, y2 = [min(max((float(bb.find(tag).text) - 1) * s, 0), hi)
                          for tag, s in (('xmin', sx), ('ymin', sy), ('xmax', sx), ('ymax', sy))]
        cx, cy, w, h = (x1 + x2) / 2.0, (y1 + y2) / 2.0, x2 - x1, y2 - y1
        col, row = int(cx * cell_size / image_size), int(cy * cell_size / image_size)
        if label[row, col, 0] == 1:          # first object wins the cell
            continue
        label[row, col, 0] = 1
        label[row, col, 1:5] = (cx, cy, w, h)
        label[row, col, 5 + class_to_ind[obj.find('name').text.lower().strip()]] = 1
    return label, len(objs)


def preprocess_image(image_bgr_u8, image_size, flipped=False):
    """pascal_voc.py:60-67: cv2.resize -> float32 -> x/255*2-1 (BGR kept), optional W flip."""
    import cv2
    image = cv2.resize(image_bgr_u8, (image_size, image_size)).astype(np.float32)
    image = (image / 255.0) * 2.0 - 1.0
    return image[:, ::-1, :] if flipped else image


class pascal_voc:
    def __init__(self, image_set, batch_size=cfg.BATCH_SIZE, rebuild=False, require_data=True):
        self.name = 'voc_2007'
        self.devkit_path = cfg.PASCAL_PATH
        self.data_path = os.path.join(self.devkit_path, 'VOC2007')
        self.cache_path = cfg.CACHE_PATH
        self.batch_size = batch_size
        self.image_size = cfg.IMAGE_SIZE
        self.cell_size = cfg.S
        self.classes = VOC_CLASSES
        self.num_class = len(self.classes)
        self.class_to_ind = dict(zip(self.classes, range(self.num_class)))
        self.flipped = cfg.FLIPPED
        self.image_set = image_set
        self.rebuild = rebuild
        self.cursor = 0
        self.gt_labels = None
        if require_data:
            # same error convention as the reference (pascal_voc.py:36-39)
            assert os.path.exists(self.devkit_path), 'VOCdevkit path does not exist: {}'.format(self.devkit_path)
            assert os.path.exists(self.data_path), 'Path does not exist: {}'.format(self.data_path)
            self.prepare()

    # -- batches ------------------------------------------------------------------------------
    def get(self):
        """pascal_voc.py:42-58: next batch (float64 arrays, like the reference), reshuffling at the
        end of an epoch."""
        images = np.zeros((self.batch_size, self.image_size, self.image_size, 3))
        labels = np.zeros((self.batch_size, self.cell_size, self.cell_size, 25))
        for slot in range(self.batch_size):
            rec = self.gt_labels[self.cursor]
            images[slot] = self.image_read(rec['imname'], rec['flipped'])
            labels[slot] = rec['label']
            self.cursor += 1
            if self.cursor >= len(self.gt_labels):
                np.random.shuffle(self.gt_labels)
                self.cursor = 0
        return images, labels

    def image_read(self, imname, flipped=False):
        import cv2
        return preprocess_image(cv2.imread(imname), self.image_size, flipped)

    # -- label preparation ----------------------------------------------------------------------
    def prepare(self):
        gt_labels = self.load_labels()
        if self.flipped:
            print('Appending horizontally-flipped training examples ...')
            mirrored = copy.deepcopy(gt_labels)
            for rec in mirrored:
                rec['flipped'] = True
                rec['label'] = rec['label'][:, ::-1, :]
                hit = rec['label'][..., 0] == 1
                rec['label'][hit, 1] = self.image_size - 1 - rec['label'][hit, 1]      # pascal_voc.py:78-82
            gt_labels += mirrored
        np.random.shuffle(gt_labels)
        self.gt_labels = gt_labels
        return gt_labels

    def load_labels(self):
        cache_file = os.path.join(self.cache_path, 'pascal_' + self.image_set + '_gt_labels.pkl')
        if os.path.isfile(cache_file) and not self.rebuild:
            print('Loading gt_labels from: ' + cache_file)
            with open(cache_file, 'rb') as f:
                return pickle.load(f)
        print('Processing gt_labels from: ' + self.data_path)
        os.makedirs(self.cache_path, exist_ok=True)
        txtname = os.path.join(self.data_path, 'ImageSets', 'Main', self.image_set + '.txt')
        assert os.path.exists(txtname), 'Path does not exist: {}'.format(txtname)
        with open(txtname, 'r') as f:
            self.image_index = [x.strip() for x in f.readlines()]
        gt_labels = []
        for index in self.image_index:
            label, num = self.load_pascal_annotation(index)
            if num == 0:
                continue
            gt_labels.append({'imname': os.path.join(self.data_path, 'JPEGImages', index + '.jpg'),
                              'label': label, 'flipped': False})
        print('Saving gt_labels to: ' + cache_file)
        with open(cache_file, 'wb') as f:
            pickle.dump(gt_labels, f)
        return gt_labels

    def load_pascal_annotation(self, index):
        import cv2
        im = cv2.imread(os.path.join(self.data_path, 'JPEGImages', index + '.jpg'))
        return encode_annotation(os.path.join(self.data_path, 'Annotations', index + '.xml'), im.shape[0],
                                 im.shape[1], self.image_size, self.cell_size, self.class_to_ind)
