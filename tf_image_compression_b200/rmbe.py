"""rm_block_effect post-filter driver with the reference's surface: rmbe(image)
(submit/2/rmbe/rmbe.py:15-25; the root rm_block_effect/rmbe.py copy is broken).

The reference builds a new tf.Graph + Session and restores the params on EVERY pass
(submit/2/rmbe/rmbe.py:29-44); here the post-filter graph lives in the Codec handle and both
passes run as two kernel sequences over tiles addressed in place inside the image."""
from __future__ import annotations

import numpy as np

patch_size = 128  # submit/2/rmbe/rmbe.py:12

_codec = None


def bind(codec):
    """Select the Codec (with set_postfilter done) that rmbe() runs on."""
    global _codec
    _codec = codec


def rmbe(image):
    """De-block one [H,W,3] image (values 0..255); returns a float32 array like the reference
    (which mutates and returns the array it was given when that array is float32)."""
    if _codec is None:
        raise RuntimeError("rmbe.bind(codec) first: the post-filter runs on a configured Codec")
    if hasattr(image, "is_cuda"):
        return _codec.postfilter_images(image)
    img = np.ascontiguousarray(image, dtype=np.float32)
    _codec.postfilter_images(img)
    return img
