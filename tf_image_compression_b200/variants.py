"""Per-variant layer tables of the reference codec (table-driven replacement for the 18 distinct
model.py files).

Entries: ('c', scope, cout, stride, act) = basic_block.my_conv2d (basic_block/basic_block.py:27-47),
('d', scope, cout, act) = basic_block.my_conv2d_transpose (:50-71, stride 2), ('r', scope, cout) =
basic_block.res_block with two relu convs (:74-93).  Scopes are the TF variable scopes, so a
checkpoint's '<scope>/kernel' and '<scope>/bias' map one-to-one.
"""
from __future__ import annotations

from dataclasses import dataclass

RELU, ID = "relu", "id"


def _trunk(prefix, names, ch=64):
    return [("r", f"{prefix}_res_{n}", ch) for n in names]


VARIANTS = {
    # model_0/model.py:50-134 / :159-246 ; config.json patch_size 256, quan_scale 2
    "model_0": dict(
        patch_size=256, bottleneck_channel=64,
        encoder=[("c", "encode_0", 32, 2, RELU), ("c", "encode_1", 32, 2, RELU), ("c", "encode_2", 64, 2, RELU),
                 ("c", "encode_3", 64, 2, RELU), *_trunk("encode", (1, 2)), ("c", "encode_4", 64, 1, ID)],
        decoder=[("c", "decode_4", 64, 1, ID), *_trunk("decode", (1, 2)), ("d", "decode_3", 64, RELU),
                 ("d", "decode_2", 32, RELU), ("d", "decode_1", 32, RELU), ("d", "decode_0", 3, ID)]),
    # model_1/model.py (lines 52 and 226 differ from model_0: 16-channel outer stage)
    "model_1": dict(
        patch_size=256, bottleneck_channel=64,
        encoder=[("c", "encode_0", 16, 2, RELU), ("c", "encode_1", 32, 2, RELU), ("c", "encode_2", 64, 2, RELU),
                 ("c", "encode_3", 64, 2, RELU), *_trunk("encode", (1, 2)), ("c", "encode_4", 64, 1, ID)],
        decoder=[("c", "decode_4", 64, 1, ID), *_trunk("decode", (1, 2)), ("d", "decode_3", 64, RELU),
                 ("d", "decode_2", 32, RELU), ("d", "decode_1", 16, RELU), ("d", "decode_0", 3, ID)]),
    # model_2/model.py:50-122 / :147-222
    "model_2": dict(
        patch_size=128, bottleneck_channel=64,
        encoder=[("c", "encode_1", 32, 2, RELU), ("c", "encode_2", 64, 2, RELU), ("c", "encode_3", 64, 2, RELU),
                 *_trunk("encode", (1, 2)), ("c", "encode_4", 64, 2, ID)],
        decoder=[("d", "decode_4", 64, ID), *_trunk("decode", (1, 2)), ("d", "decode_3", 64, RELU),
                 ("d", "decode_2", 32, RELU), ("d", "decode_1", 3, ID)]),
    # model_3/model.py:50-161 / :186-300 (= base_model/fin, rm_block_effect/recons_model, submit/2, submit/3)
    "model_3": dict(
        patch_size=128, bottleneck_channel=80,
        encoder=[("c", "encode_1", 32, 2, RELU), ("c", "encode_2", 64, 2, RELU), *_trunk("encode", ("m1", 0)),
                 ("c", "encode_3", 64, 2, RELU), *_trunk("encode", (1, 2, 3)), ("c", "encode_4", 80, 2, ID)],
        decoder=[("d", "decode_4", 64, ID), *_trunk("decode", (1, 2, 3)), ("d", "decode_3", 64, RELU),
                 *_trunk("decode", (4, 5)), ("d", "decode_2", 32, RELU), ("d", "decode_1", 3, ID)]),
    # base_model/input_256/model.py:50-122 / :147-222 ; decode_4 is relu here
    "base_model/input_256": dict(
        patch_size=256, bottleneck_channel=64,
        encoder=[("c", "encode_1", 32, 2, RELU), ("c", "encode_2", 64, 2, RELU), ("c", "encode_3", 64, 2, RELU),
                 *_trunk("encode", (1, 2)), ("c", "encode_4", 64, 1, ID)],
        decoder=[("c", "decode_4", 64, 1, RELU), *_trunk("decode", (1, 2)), ("d", "decode_3", 32, RELU),
                 ("d", "decode_2", 32, RELU), ("d", "decode_1", 3, ID)]),
    # base_model/ch_128/model.py:50-110 / :135-198
    "base_model/ch_128": dict(
        patch_size=128, bottleneck_channel=64,
        encoder=[("c", "encode_1", 64, 2, RELU), ("c", "encode_2", 128, 2, RELU), *_trunk("encode", (1, 2), 128),
                 ("c", "encode_3", 64, 1, ID)],
        decoder=[("c", "decode_3", 128, 1, ID), *_trunk("decode", (1, 2), 128), ("d", "decode_2", 64, RELU),
                 ("d", "decode_1", 3, ID)]),
    # base_model/reduced_btn_32/model.py:50-110 / :136-199 ; 4-argument encoder, bottleneck_channel
    # from config.json (:276)
    "base_model/reduced_btn_32": dict(
        patch_size=128, bottleneck_channel=32,
        encoder=[("c", "encode_1", 32, 2, RELU), ("c", "encode_2", 64, 2, RELU), *_trunk("encode", (1, 2)),
                 ("c", "encode_3", "bottleneck_channel", 1, ID)],
        decoder=[("c", "decode_3", 64, 1, ID), *_trunk("decode", (1, 2)), ("d", "decode_2", 32, RELU),
                 ("d", "decode_1", 3, ID)]),
}

# aliases the reference itself uses (md5-identical model.py files)
ALIASES = {
    "base_model/fin": "model_3",
    "rm_block_effect/recons_model": "model_3",
    "submit/1": "model_2",
    "submit/2": "model_3",
    "submit/3": "model_3",
}

POSTFILTERS = {
    # submit/2/rmbe/model.py:113-197 == rm_block_effect/model_0/model.py:107-191
    "rmbe": [("c", "conv_1", 32, 2, RELU), ("c", "conv_2", 64, 2, RELU), ("c", "conv_3", 64, 1, RELU),
             ("c", "conv_4", 64, 1, RELU), ("d", "conv_5", 32, RELU), ("d", "conv6", 3, ID)],
    # rm_block_effect/model_1/model.py:107-168
    "rmbe_model_1": [("c", "conv_1", 32, 1, RELU), ("c", "conv_2", 64, 1, RELU), ("c", "conv_3", 32, 1, RELU),
                     ("c", "conv_4", 3, 1, ID)],
}


@dataclass(frozen=True)
class PrimLayer:
    kind: str        # 'c' | 'd'
    scope: str       # TF variable scope: '<scope>/kernel', '<scope>/bias'
    cin: int
    cout: int
    stride: int
    act: str
    res_begin: int
    res_end: int

    @property
    def kernel_shape(self):
        # conv HWIO (basic_block.py:30); deconv [kh, kw, filters, in] (basic_block.py:53)
        return (3, 3, self.cin, self.cout) if self.kind == "c" else (3, 3, self.cout, self.cin)


def resolve(name: str) -> str:
    name = name.strip("/")
    name = ALIASES.get(name, name)
    if name not in VARIANTS:
        raise ValueError(f"unknown model variant {name!r}; known: {sorted(VARIANTS) + sorted(ALIASES)}")
    return name


def primitive_layers(layers, cin, bottleneck_channel=None):
    """Flatten a table to primitive 3x3 layers; res_block -> '<scope>/conv_0', '<scope>/conv_1'."""
    out, c = [], cin
    for l in layers:
        if l[0] == "r":
            if l[2] != c:
                raise ValueError(f"res_block {l[1]} expects {l[2]} channels, got {c}")
            out.append(PrimLayer("c", l[1] + "/conv_0", c, c, 1, RELU, 1, 0))
            out.append(PrimLayer("c", l[1] + "/conv_1", c, c, 1, RELU, 0, 1))
            continue
        cout = l[2]
        if cout == "bottleneck_channel":
            if bottleneck_channel is None:
                raise ValueError("this variant needs bottleneck_channel (config.json)")
            cout = int(bottleneck_channel)
        if l[0] == "c":
            out.append(PrimLayer("c", l[1], c, cout, l[3], l[4], 0, 0))
        elif l[0] == "d":
            out.append(PrimLayer("d", l[1], c, cout, 2, l[3], 0, 0))
        else:
            raise ValueError(l)
        c = cout
    return out


def encoder_layers(variant, bottleneck_channel=None):
    v = VARIANTS[resolve(variant)]
    return primitive_layers(v["encoder"], 3, bottleneck_channel or v["bottleneck_channel"])


def decoder_layers(variant, bottleneck_channel=None):
    v = VARIANTS[resolve(variant)]
    return primitive_layers(v["decoder"], bottleneck_channel or v["bottleneck_channel"])


def postfilter_layers(name="rmbe"):
    return primitive_layers(POSTFILTERS[name], 3)


def flops_per_pixel(layers, patch_size):
    """Algorithmic FLOP (2*MAC, unpadded) per input-image pixel for a primitive layer list whose first
    layer sees a patch_size x patch_size map (encoder / post-filter), BASELINE.md §3."""
    h = w = patch_size
    macs = 0
    for l in layers:
        if l.kind == "c":
            h, w = -(-h // l.stride), -(-w // l.stride)
            macs += h * w * 9 * l.cin * l.cout
        else:
            macs += h * w * 9 * l.cin * l.cout  # every tap of every input pixel is used exactly once
            h, w = 2 * h, 2 * w
    return 2.0 * macs


def model_flops_per_pixel(variant, patch_size, bottleneck_channel=None):
    enc = encoder_layers(variant, bottleneck_channel)
    dec = decoder_layers(variant, bottleneck_channel)
    e = flops_per_pixel(enc, patch_size) / (patch_size * patch_size)
    # decoder starts from the bottleneck map
    hb = patch_size
    for l in enc:
        hb = -(-hb // l.stride) if l.kind == "c" else hb * 2
    d = flops_per_pixel(dec, hb) / (patch_size * patch_size)
    return e, d
