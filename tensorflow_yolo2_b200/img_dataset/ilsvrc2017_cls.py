"""ILSVRC-2017 classification image database with the interface the reference's ImageNet scripts use
(src/img_dataset/ilsvrc2017_cls_multithread.py: class `ilsvrc_cls` -- .name, .classes, .num_class, .image_num,
.total_batch, .epoch, .get(), .image_read()).

Scope: what imagenet_test_darknet.py / imagenet_predict_darknet.py / imagenet_train_darknet.py feed the network with: single
process, no augmentation --
`image_read` is cv2.imread -> (optional BGR->RGB) -> cv2.resize((IS, IS)) -> float32 -> x/255*2-1
(ilsvrc2017_cls_multithread.py:320-323,408-415).  The training-time augmentation (random rotation / crop / colour,
:328-407) and the 10-process prefetcher (:119-205, :221-318) are host-side data plumbing outside the hot path (SURVEY
section 2 row 11) and are not reproduced; `data_aug=True` / `multithread=True` raise.

`synthetic=N` builds an in-memory database of N random images so that the scripts run where the dataset is not on disk
(the reference asserts in that case, :36-39 -- and so does this class without the flag).
"""
from __future__ import annotations

import math
import os
import pickle
import random
import xml.etree.ElementTree as ET

import numpy as np

from .. import config as cfg


class ilsvrc_cls:
    def __init__(self, image_set, rebuild=False, data_aug=False, multithread=False, batch_size=cfg.BATCH_SIZE,
                 image_size=cfg.IMAGE_SIZE, RGB=False, synthetic=0):
        if data_aug or multithread:
            raise NotImplementedError('ilsvrc_cls: the training-time augmentation / multi-process prefetcher are not reproduced')
        self.name = 'ilsvrc_2017_cls'
        self.devkit_path = cfg.ILSVRC_PATH
        self.data_path = self.devkit_path
        self.cache_path = cfg.CACHE_PATH
        self.batch_size = batch_size
        self.image_size = image_size
        self.image_set = image_set
        self.rebuild = rebuild
        self.multithread = False
        self.data_aug = False
        self.RGB = RGB
        self.cursor = 0
        self.epoch = 1
        self.gt_labels = None
        self._synthetic = None
        if synthetic:
            rs = np.random.RandomState(0)
            self.classes = ['n%08d' % i for i in range(1000)]
            self.num_class = 1000
            self.class_to_ind = dict(zip(self.classes, range(1000)))
            self._synthetic = rs.randint(0, 256, (int(synthetic), 64, 80, 3)).astype(np.uint8)
            self.gt_labels = [{'imname': i, 'label': int(rs.randint(0, 1000))} for i in range(int(synthetic))]
            self.image_num = len(self.gt_labels)
            self.total_batch = int(math.ceil(self.image_num / float(self.batch_size)))
        else:
            assert os.path.exists(self.devkit_path), 'ILSVRC path does not exist: {}'.format(self.devkit_path)
            self.load_classes()
            self.prepare()
        self.get = self._get

    def load_classes(self):
        """:208-219 -- the class list is the list of training folders."""
        img_folder = os.path.join(self.data_path, 'Data', 'CLS-LOC', 'train')
        print('Loading class info from ' + img_folder)
        self.classes = [item for item in os.listdir(img_folder) if os.path.isdir(os.path.join(img_folder, item))]
        self.num_class = len(self.classes)
        assert self.num_class == 1000, 'number of classes is not 1000!'
        self.class_to_ind = dict(zip(self.classes, range(self.num_class)))

    def prepare(self):
        """:49-93 -- (image path, label) list, cached as a pickle, shuffled."""
        cache_file = os.path.join(self.cache_path, 'ilsvrc_cls_' + self.image_set + '_gt_labels.pkl')
        if os.path.isfile(cache_file) and not self.rebuild:
            print('Loading gt_labels from: ' + cache_file)
            with open(cache_file, 'rb') as f:
                gt_labels = pickle.load(f)
        else:
            imgset_fname = 'train_cls.txt' if self.image_set == 'train' else self.image_set + '.txt'
            imgset_file = os.path.join(self.data_path, 'ImageSets', 'CLS-LOC', imgset_fname)
            anno_dir = os.path.join(self.data_path, 'Annotations', 'CLS-LOC', self.image_set)
            print('Processing gt_labels using ' + imgset_file)
            gt_labels = []
            with open(imgset_file, 'r') as f:
                for line in f.readlines():
                    img_path = line.strip().split()[0]
                    if self.image_set == 'train':
                        label = self.class_to_ind[img_path.split('/')[0]]
                    else:
                        tree = ET.parse(os.path.join(anno_dir, img_path + '.xml'))
                        label = self.class_to_ind[tree.find('object').find('name').text]
                    gt_labels.append({'imname': os.path.join(self.data_path, 'Data', 'CLS-LOC', self.image_set, img_path + '.JPEG'),
                                      'label': label})
            os.makedirs(self.cache_path, exist_ok=True)
            print('Saving gt_labels to: ' + cache_file)
            with open(cache_file, 'wb') as f:
                pickle.dump(gt_labels, f)
        random.shuffle(gt_labels)
        self.gt_labels = gt_labels
        self.image_num = len(gt_labels)
        self.total_batch = int(math.ceil(self.image_num / float(self.batch_size)))

    def _get(self):
        """:95-117 -- next batch (float64 images like the reference's np.zeros default, 1-D labels), reshuffle per epoch."""
        images = np.zeros((self.batch_size, self.image_size, self.image_size, 3))
        labels = np.zeros(self.batch_size)
        for count in range(self.batch_size):
            rec = self.gt_labels[self.cursor]
            images[count] = self.image_read(rec['imname'])
            labels[count] = rec['label']
            self.cursor += 1
            if self.cursor >= len(self.gt_labels):
                random.shuffle(self.gt_labels)
                self.cursor = 0
                self.epoch += 1
        return images, labels

    def image_read(self, imname, data_aug=False):
        """:320-323,408-415 (no augmentation)."""
        import cv2
        image = self._synthetic[imname] if self._synthetic is not None else cv2.imread(imname)
        if self.RGB:
            image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
        image = cv2.resize(image, (self.image_size, self.image_size)).astype(np.float32)
        return (image / 255.0) * 2.0 - 1.0
