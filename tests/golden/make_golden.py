"""Generate tests/golden/*.npz from the reference's MNIST files.  Run HERE
(container with /root/reference); the fixtures travel to the GPU box.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixedl_oracle as O  # noqa: E402

REF = "/root/reference/mllib/MNIST"
out = os.path.dirname(os.path.abspath(__file__))

imgs = O.read_idx(os.path.join(REF, "train-images-idx3-ubyte"))
labs = O.read_idx(os.path.join(REF, "train-labels-idx1-ubyte"))
data, labels, keep = O.select_per_label(imgs, labs, 100)     # readMNIST, Ntrain=100
raw = imgs[keep].reshape(-1, 14, 2, 14, 2).astype(np.uint16).sum(axis=(2, 4)).reshape(-1, 196)
np.savez_compressed(os.path.join(out, "mnist_100_per_label_14x14.npz"),
                    sum4=raw.astype(np.uint16), labels=labels.astype(np.int64), file_index=keep)
print("wrote mnist_100_per_label_14x14.npz", raw.shape)
