"""Host mirrors of the reference's patch helpers, for callers that still hold numpy images
(the GPU path fuses both into the first-layer loads / last-layer stores: Codec.encode_images /
Codec.decode_images).  Same names, arguments and results as utils/utils.py."""
from __future__ import annotations

import numpy as np


def crop_image_input_patches(image, patch_size):
    """utils.crop_image_input_patches (utils/utils.py:96-133): reflect-pad bottom/right to a multiple of
    patch_size, return the row-major list of [P,P,C] crops."""
    height, width, _ = image.shape
    pad_h = (patch_size - height % patch_size) % patch_size
    pad_w = (patch_size - width % patch_size) % patch_size
    padded = np.pad(image, ((0, pad_h), (0, pad_w), (0, 0)), "reflect")
    ph, pw = padded.shape[0] // patch_size, padded.shape[1] // patch_size
    view = padded.reshape(ph, patch_size, pw, patch_size, -1).swapaxes(1, 2)
    return [view[i, j] for i in range(ph) for j in range(pw)]


def concat_patches(patches, height, width, patch_size):
    """utils.concat_patches (utils/utils.py:136-167): stitch row-major patches, crop to [height, width]."""
    hn = -(-height // patch_size)
    wn = -(-width // patch_size)
    arr = np.asarray(patches)
    if arr.shape[0] != hn * wn:
        raise ValueError(f"expected {hn * wn} patches for a {height}x{width} image, got {arr.shape[0]}")
    full = arr.reshape(hn, wn, patch_size, patch_size, -1).swapaxes(1, 2).reshape(hn * patch_size, wn * patch_size, -1)
    return full[:height, :width]


def read_image_list(path):
    """utils.read_image_list: one path per line."""
    with open(path) as f:
        return [ln.strip() for ln in f if ln.strip()]
