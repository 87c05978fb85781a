"""Shared fixtures for the -m gpu parity tests: seeded parameters, normalisation, inputs."""
import numpy as np

from oracle import codec_oracle as O
import tf_image_compression_b200 as T

MEAN = np.array([118.3, 113.9, 102.6], np.float32)
STD = np.array([61.7, 59.2, 63.8], np.float32)


def params_for(variant, scheme, seed=1234, q=2):
    ov = O.VARIANTS[variant]
    enc = O.init_params(ov["enc"], 3, seed, scheme)
    dec = O.init_params(ov["dec"], ov["bottleneck"], seed + 1, scheme)
    if scheme == "fanin":
        dec = O.condition_decoder(variant, dec, q)
    return enc, dec


def make_codec(variant, scheme="fanin", q=2, compute="fp32", seed=1234):
    enc, dec = params_for(variant, scheme, seed, q)
    c = T.Codec(variant, quan_scale=q, mean=MEAN, std=STD, enc_params=enc, dec_params=dec, compute=compute)
    return c, enc, dec


def patches_from_images(n_images, h, w, P, seed=0, kind="natural"):
    out = []
    for i in range(n_images):
        out += O.crop_image_input_patches(O.synthetic_image(h, w, seed + i, kind), P)
    return np.stack(out)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
