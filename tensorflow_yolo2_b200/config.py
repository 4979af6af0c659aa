"""Mirror of the reference's src/config.py (same names), with the grid made configurable.

Reference: config.py:7-24 (paths), :30-51 (hyper-parameters), :55-89 (directory helpers).
The reference hard-codes IMAGE_SIZE=224, S=7, B=2; `set_grid()` (or the env vars Y2_IMAGE_SIZE /
Y2_S / Y2_B) re-derives YOLO_GRID_OFFSET for the 416/13 and 608/19 configurations.
"""
import os
import numpy as np

##########
# Pathes #
##########
SRC_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT_DIR = os.environ.get('Y2_ROOT_DIR', os.path.abspath(os.path.join(SRC_DIR, os.pardir)))

PASCAL_PATH = os.path.join(ROOT_DIR, 'data', 'VOCdevkit')
ILSVRC_PATH = os.path.join(ROOT_DIR, 'data', 'ILSVRC')
FLOWERS_PATH = os.path.join(ROOT_DIR, 'data', 'TF_flowers')
CACHE_PATH = os.path.join(ROOT_DIR, 'cache')
WEIGHTS_PATH = os.path.join(ROOT_DIR, 'weights')
CKPTS_PATH = os.path.join(ROOT_DIR, 'ckpts')
TENSORBOARD_PATH = os.path.join(ROOT_DIR, 'tensorboard')

# weights pathes (the reference's spelling, config.py:23-24); here they are .npz stores keyed by
# the TF variable names
darknet_pascal_weight_path = os.path.join(WEIGHTS_PATH, "darknet19_pascal.ckpt")
darknet_imagenet_weight_path = os.path.join(WEIGHTS_PATH, "darkent19_imagenet.ckpt")

##########
# Hypers #
##########
TRAIN_SNAPSHOT_PREFIX = 'train'
BATCH_SIZE = 48
IMAGE_SIZE = int(os.environ.get('Y2_IMAGE_SIZE', 224))
RAND_CROP_UPBOUND = 292

# YOLO1 VOC settings
S = int(os.environ.get('Y2_S', 7))
B = int(os.environ.get('Y2_B', 2))


def _grid_offset(S_, B_):
    # config.py:40-42 (py2 `range(S) * S * B` == list repetition): offset[y, x, b] = x
    off = np.array(list(range(S_)) * S_ * B_)
    off = np.reshape(off, (B_, S_, S_))
    return np.transpose(off, (1, 2, 0))  # [Y,X,B]


YOLO_GRID_OFFSET = _grid_offset(S, B)

LAMBDA_COORD = 5
LAMBDA_NOOBJ = 0.5

FLIPPED = False
REBUILD = False
MULTITHREAD = True

# compute path of the network builders: 'bf16' (tcgen05 tensor cores) or 'fp32' (exact FFMA path)
COMPUTE = os.environ.get('Y2_COMPUTE', 'bf16')


def set_grid(image_size, s, b):
    """Switch IMAGE_SIZE / S / B (e.g. 416/13/5, 608/19/5) and rebuild YOLO_GRID_OFFSET."""
    global IMAGE_SIZE, S, B, YOLO_GRID_OFFSET
    IMAGE_SIZE, S, B = int(image_size), int(s), int(b)
    YOLO_GRID_OFFSET = _grid_offset(S, B)


###########################
# Configuration Functions #
###########################
def get_output_tb_dir(network_name, imdb_name, val=True):
    """config.py:55-75: <ROOT>/tensorboard/<net>/<imdb>/{train,val}, created on demand."""
    outdir = os.path.abspath(os.path.join(ROOT_DIR, 'tensorboard', network_name, imdb_name))
    traindir = os.path.join(outdir, 'train')
    os.makedirs(traindir, exist_ok=True)
    if val:
        valdir = os.path.join(outdir, 'val')
        os.makedirs(valdir, exist_ok=True)
    else:
        valdir = None
    return traindir, valdir


def get_ckpts_dir(network_name, imdb_name):
    """config.py:78-89: <ROOT>/ckpts/<net>/<imdb>, created on demand."""
    outdir = os.path.abspath(os.path.join(ROOT_DIR, 'ckpts', network_name, imdb_name))
    os.makedirs(outdir, exist_ok=True)
    return outdir
