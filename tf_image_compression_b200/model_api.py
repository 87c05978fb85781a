"""The reference's per-model module surface — `model.encoder(input, patch_size, quan_scale)`,
`model.decoder(input, quan_scale)` (model_0/model.py:34,147; 4-argument encoder in
base_model/reduced_btn_32/model.py:34) — as eager callables over a Codec.

The reference builds a TF graph once and fetches numpy batches with sess.run (encode.py:147-165,
decode.py:167-220); these callables take the fed array directly and return what sess.run returned:
integer-valued float32 symbols [N,hb,wb,cb] / float32 reconstructions [N,P,P,3] in [0,255]."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from . import variants as V
from .codec import Codec


class ModelModule:
    """Stands where `from model_N import model` stood (encode.py:225-232)."""

    def __init__(self, variant, codec: Codec):
        self.variant = V.resolve(variant)
        self.codec = codec

    def encoder(self, input, patch_size, quan_scale, bottleneck_channel=None):
        if int(quan_scale) != self.codec.quan_scale:
            raise ValueError(f"quan_scale {quan_scale} differs from the codec's {self.codec.quan_scale}")
        if bottleneck_channel is not None and int(bottleneck_channel) != self.codec.enc_layers[-1].cout:
            raise ValueError("bottleneck_channel differs from the configured graph")
        x = np.asarray(input)
        x = x.reshape(-1, patch_size, patch_size, 3)  # tf.reshape(input, [-1, P, P, 3]) (model_0/model.py:39)
        if x.dtype != np.uint8:
            x = np.ascontiguousarray(x, dtype=np.float32)
        else:
            x = np.ascontiguousarray(x)
        return self.codec.encode_patches(x, out_dtype=np.float32)

    def decoder(self, input, quan_scale):
        if int(quan_scale) != self.codec.quan_scale:
            raise ValueError(f"quan_scale {quan_scale} differs from the codec's {self.codec.quan_scale}")
        s = np.asarray(input)
        if s.dtype != np.uint8:
            r = np.rint(s)
            if not np.array_equal(r, s) or r.min(initial=0) < 0 or r.max(initial=0) > quan_scale - 1:
                raise ValueError("decoder input must hold integer symbols in [0, quan_scale-1]")
            s = r.astype(np.uint8)
        return self.codec.decode_patches(np.ascontiguousarray(s))


def load_config(model_dir):
    """model_N/config.json (keys: name_sep, resolution, patch_size, quan_scale, [bottleneck_channel] …)."""
    with open(Path(model_dir) / "config.json") as f:
        return json.load(f)


def load_normalization(path):
    """data_info/channel_normalization_params.npz: keys 'mean', 'std', shape [3] (model_0/model.py:18,26-28)."""
    z = np.load(path)
    return np.asarray(z["mean"], dtype=np.float32), np.asarray(z["std"], dtype=np.float32)
