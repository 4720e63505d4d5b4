"""Fingerprint of the reference's default initialisation under torch.manual_seed(0) (build container only:
imports /root/reference/model.py).  Stored as (key, shape, sum, sum of squares) per state_dict entry."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import model as ref  # noqa: E402

torch.manual_seed(0)
m = ref.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False,
            only_pixelcnn=False, sigma_decoder=0.1, input_image_size=64)
sd = m.state_dict()
keys = np.array(list(sd.keys()))
shapes = np.zeros((len(sd), 4), dtype=np.int64)
sums = np.zeros(len(sd)); sumsq = np.zeros(len(sd))
for i, (k, v) in enumerate(sd.items()):
    shapes[i, : v.dim()] = list(v.shape)
    sums[i] = v.double().sum().item(); sumsq[i] = (v.double() ** 2).sum().item()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "init_fingerprint_seed0.npz"),
                    keys=keys, shapes=shapes, sums=sums, sumsq=sumsq)
print("ok", len(sd))
