"""CPU restatement ("oracle") of the reference codec hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under tf_image_compression_b200/ may import this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

PARITY STATUS: **parity unpinned** against TensorFlow itself.  The reference
(bolin-chen/tf_image_compression) is TF-1.x graph code; TensorFlow, its weights, its
normalisation constants and any saved tensors are absent from /root/reference and from this
image, and the reference has no test or golden vector for the conv path.
This file restates the documented semantics of the TF ops the reference calls, and is
self-validated three ways (tests/test_oracle.py): (a) the torch-CPU contraction used here
against an independent direct-loop C restatement (oracle/tic_oracle.c) and a numpy einsum
restatement, (b) transposed conv == autograd gradient of the stride-2 SAME conv, (c) fp32
against an fp64 evaluation of the same graph.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# Per-variant layer lists.  ('c', scope, cout, stride, act) = basic_block.my_conv2d,
# ('d', scope, cout, act) = basic_block.my_conv2d_transpose (stride 2, output 2x),
# ('r', scope, cout) = basic_block.res_block with layer_num=2, relu (basic_block.py:74-93).
# act: 'relu' | 'id'.  All kernels 3x3, padding SAME.
# --------------------------------------------------------------------------------------------
VARIANTS = {
    # model_0/model.py:50-134 (encoder), :159-246 (decoder)
    "model_0": dict(
        patch_size=256, bottleneck=64,
        enc=[("c", "encode_0", 32, 2, "relu"), ("c", "encode_1", 32, 2, "relu"), ("c", "encode_2", 64, 2, "relu"),
             ("c", "encode_3", 64, 2, "relu"), ("r", "encode_res_1", 64), ("r", "encode_res_2", 64),
             ("c", "encode_4", 64, 1, "id")],
        dec=[("c", "decode_4", 64, 1, "id"), ("r", "decode_res_1", 64), ("r", "decode_res_2", 64),
             ("d", "decode_3", 64, "relu"), ("d", "decode_2", 32, "relu"), ("d", "decode_1", 32, "relu"),
             ("d", "decode_0", 3, "id")]),
    # model_1/model.py:52,226 — model_0 with a 16-channel first / last stage
    "model_1": dict(
        patch_size=256, bottleneck=64,
        enc=[("c", "encode_0", 16, 2, "relu"), ("c", "encode_1", 32, 2, "relu"), ("c", "encode_2", 64, 2, "relu"),
             ("c", "encode_3", 64, 2, "relu"), ("r", "encode_res_1", 64), ("r", "encode_res_2", 64),
             ("c", "encode_4", 64, 1, "id")],
        dec=[("c", "decode_4", 64, 1, "id"), ("r", "decode_res_1", 64), ("r", "decode_res_2", 64),
             ("d", "decode_3", 64, "relu"), ("d", "decode_2", 32, "relu"), ("d", "decode_1", 16, "relu"),
             ("d", "decode_0", 3, "id")]),
    # model_2/model.py:50-122, :147-222
    "model_2": dict(
        patch_size=128, bottleneck=64,
        enc=[("c", "encode_1", 32, 2, "relu"), ("c", "encode_2", 64, 2, "relu"), ("c", "encode_3", 64, 2, "relu"),
             ("r", "encode_res_1", 64), ("r", "encode_res_2", 64), ("c", "encode_4", 64, 2, "id")],
        dec=[("d", "decode_4", 64, "id"), ("r", "decode_res_1", 64), ("r", "decode_res_2", 64),
             ("d", "decode_3", 64, "relu"), ("d", "decode_2", 32, "relu"), ("d", "decode_1", 3, "id")]),
    # model_3/model.py:50-161, :186-300 (= base_model/fin = rm_block_effect/recons_model)
    "model_3": dict(
        patch_size=128, bottleneck=80,
        enc=[("c", "encode_1", 32, 2, "relu"), ("c", "encode_2", 64, 2, "relu"), ("r", "encode_res_m1", 64),
             ("r", "encode_res_0", 64), ("c", "encode_3", 64, 2, "relu"), ("r", "encode_res_1", 64),
             ("r", "encode_res_2", 64), ("r", "encode_res_3", 64), ("c", "encode_4", 80, 2, "id")],
        dec=[("d", "decode_4", 64, "id"), ("r", "decode_res_1", 64), ("r", "decode_res_2", 64),
             ("r", "decode_res_3", 64), ("d", "decode_3", 64, "relu"), ("r", "decode_res_4", 64),
             ("r", "decode_res_5", 64), ("d", "decode_2", 32, "relu"), ("d", "decode_1", 3, "id")]),
    # base_model/input_256/model.py:50-122, :147-222
    "base_model/input_256": dict(
        patch_size=256, bottleneck=64,
        enc=[("c", "encode_1", 32, 2, "relu"), ("c", "encode_2", 64, 2, "relu"), ("c", "encode_3", 64, 2, "relu"),
             ("r", "encode_res_1", 64), ("r", "encode_res_2", 64), ("c", "encode_4", 64, 1, "id")],
        dec=[("c", "decode_4", 64, 1, "relu"), ("r", "decode_res_1", 64), ("r", "decode_res_2", 64),
             ("d", "decode_3", 32, "relu"), ("d", "decode_2", 32, "relu"), ("d", "decode_1", 3, "id")]),
    # base_model/ch_128/model.py:50-110, :135-198
    "base_model/ch_128": dict(
        patch_size=128, bottleneck=64,
        enc=[("c", "encode_1", 64, 2, "relu"), ("c", "encode_2", 128, 2, "relu"), ("r", "encode_res_1", 128),
             ("r", "encode_res_2", 128), ("c", "encode_3", 64, 1, "id")],
        dec=[("c", "decode_3", 128, 1, "id"), ("r", "decode_res_1", 128), ("r", "decode_res_2", 128),
             ("d", "decode_2", 64, "relu"), ("d", "decode_1", 3, "id")]),
    # base_model/reduced_btn_32/model.py:50-110, :136-199; bottleneck_channel from config.json (:276)
    "base_model/reduced_btn_32": dict(
        patch_size=128, bottleneck=32,
        enc=[("c", "encode_1", 32, 2, "relu"), ("c", "encode_2", 64, 2, "relu"), ("r", "encode_res_1", 64),
             ("r", "encode_res_2", 64), ("c", "encode_3", 32, 1, "id")],
        dec=[("c", "decode_3", 64, 1, "id"), ("r", "decode_res_1", 64), ("r", "decode_res_2", 64),
             ("d", "decode_2", 32, "relu"), ("d", "decode_1", 3, "id")]),
}

# post-filter nets: submit/2/rmbe/model.py:113-197 (== rm_block_effect/model_0/model.py:107-191)
# and the 4-layer all-stride-1 alternative rm_block_effect/model_1/model.py:107-168
POSTFILTERS = {
    "rmbe": [("c", "conv_1", 32, 2, "relu"), ("c", "conv_2", 64, 2, "relu"), ("c", "conv_3", 64, 1, "relu"),
             ("c", "conv_4", 64, 1, "relu"), ("d", "conv_5", 32, "relu"), ("d", "conv6", 3, "id")],
    "rmbe_model_1": [("c", "conv_1", 32, 1, "relu"), ("c", "conv_2", 64, 1, "relu"), ("c", "conv_3", 32, 1, "relu"),
                     ("c", "conv_4", 3, 1, "id")],
}


def expand_layers(layers, cin):
    """Flatten a layer list to primitive layers with explicit channel counts and variable scopes.

    Returns dicts {kind, scope, cin, cout, stride, act, res_begin, res_end}.  res_block scopes are
    '<name>/conv_0', '<name>/conv_1' (basic_block.py:75,86)."""
    out = []
    c = cin
    for l in layers:
        if l[0] == "c":
            out.append(dict(kind="c", scope=l[1], cin=c, cout=l[2], stride=l[3], act=l[4], res_begin=0, res_end=0))
            c = l[2]
        elif l[0] == "d":
            out.append(dict(kind="d", scope=l[1], cin=c, cout=l[2], stride=2, act=l[3], res_begin=0, res_end=0))
            c = l[2]
        elif l[0] == "r":
            assert l[2] == c, "res_block keeps the channel count"
            out.append(dict(kind="c", scope=l[1] + "/conv_0", cin=c, cout=c, stride=1, act="relu", res_begin=1, res_end=0))
            out.append(dict(kind="c", scope=l[1] + "/conv_1", cin=c, cout=c, stride=1, act="relu", res_begin=0, res_end=1))
        else:
            raise ValueError(l)
    return out


def init_params(layers, cin, seed, scheme="reference"):
    """Seeded stand-in for the absent checkpoints.

    scheme 'reference': tf.random_normal_initializer(0, 0.01) kernels, zero bias (model_0/model.py:57-58).
    scheme 'fanin': sigma = sqrt(2 / (9*cin)) kernels and small random biases, so activations and
    bottleneck logits are O(1) (SURVEY.md §8d weight set B).
    Variable shapes: conv kernel [3,3,cin,cout] (basic_block.py:30), deconv kernel [3,3,cout,cin] (:53)."""
    rs = np.random.RandomState(seed)
    params = {}
    for l in expand_layers(layers, cin):
        shape = (3, 3, l["cin"], l["cout"]) if l["kind"] == "c" else (3, 3, l["cout"], l["cin"])
        if scheme == "reference":
            k = rs.normal(0.0, 0.01, size=shape)
            b = np.zeros(l["cout"])
        elif scheme == "fanin":
            fan = 9 * l["cin"] if l["kind"] == "c" else 9 * l["cin"] / 4.0
            k = rs.normal(0.0, np.sqrt(2.0 / fan), size=shape)
            b = rs.normal(0.0, 0.05, size=l["cout"])
        else:
            raise ValueError(scheme)
        params[l["scope"] + "/kernel"] = k.astype(np.float32)
        params[l["scope"] + "/bias"] = b.astype(np.float32)
    return params


def condition_decoder(variant, dec_params, quan_scale, target_std=0.4):
    """Fixture helper (not reference behaviour): rescale the LAST decoder kernel of a random parameter set
    so the pre-denormalisation output has std ~target_std.  The decoder input is the inverse-sigmoid table
    (-13.8 / +11.6 for q = 2), so an unscaled random decoder saturates 0..255 and the 1e-3 max-abs parity
    bound (0..255 scale) would be compared against pre-clip values of magnitude 1e2..1e4."""
    v = VARIANTS[variant]
    hb = 4
    sym = np.random.RandomState(99).randint(0, quan_scale, size=(2, hb, hb, v["bottleneck"]))
    x = inverse_sigmoid_lut(quan_scale)[sym]
    y = run_layers(x, v["dec"], v["bottleneck"], dec_params)
    last = expand_layers(v["dec"], v["bottleneck"])[-1]["scope"]
    out = dict(dec_params)
    out[last + "/kernel"] = (dec_params[last + "/kernel"] * np.float32(target_std / max(float(y.std()), 1e-12))).astype(np.float32)
    return out


# --------------------------------------------------------------------------------------------
# TF op semantics
# --------------------------------------------------------------------------------------------
def same_pad(n, stride, k=3):
    """TF 'SAME': out = ceil(n/stride); total = max((out-1)*stride + k - n, 0); before = total//2."""
    out = -(-n // stride)
    total = max((out - 1) * stride + k - n, 0)
    return out, total // 2, total - total // 2


def _t(x, dtype):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def conv2d_same(x, kernel, bias, stride, act, dtype=torch.float32):
    """basic_block.my_conv2d (basic_block/basic_block.py:27-47): tf.nn.conv2d(SAME) + bias_add + activation.
    x: [N,H,W,Cin] NHWC, kernel: HWIO [3,3,Cin,Cout].  Zero padding is asymmetric for stride 2
    on even sizes (0 before, 1 after)."""
    xt = _t(x, dtype).permute(0, 3, 1, 2)
    _, pt, pb = same_pad(x.shape[1], stride)
    _, pl, pr = same_pad(x.shape[2], stride)
    xt = F.pad(xt, (pl, pr, pt, pb))
    w = _t(kernel, dtype).permute(3, 2, 0, 1)
    y = F.conv2d(xt, w, None, stride=stride)
    y = y + _t(bias, dtype).view(1, -1, 1, 1)
    if act == "relu":
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def deconv2d(x, kernel, bias, act, dtype=torch.float32):
    """basic_block.my_conv2d_transpose (basic_block/basic_block.py:50-71): tf.nn.conv2d_transpose with
    stride 2, SAME, output_shape = 2x input (:54), filter [3,3,Cout,Cin] (:53).  It is the gradient of
    the stride-2 SAME conv:  out[2i+kh, 2j+kw, oc] += x[i,j,ic] * W[kh,kw,oc,ic], rows/cols >= 2H dropped."""
    n, h, w, _ = x.shape
    xt = _t(x, dtype).permute(0, 3, 1, 2)
    wt = _t(kernel, dtype).permute(3, 2, 0, 1)  # [Cin, Cout, kh, kw]
    y = F.conv_transpose2d(xt, wt, None, stride=2, padding=0)[:, :, : 2 * h, : 2 * w]
    y = y + _t(bias, dtype).view(1, -1, 1, 1)
    if act == "relu":
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def run_layers(x, layers, cin, params, dtype=torch.float32, taps=None):
    """Apply a layer list; res_block (basic_block.py:74-93): inputs + conv_1(conv_0(inputs)), both relu."""
    saved = None
    for l in expand_layers(layers, cin):
        k, b = params[l["scope"] + "/kernel"], params[l["scope"] + "/bias"]
        if l["res_begin"]:
            saved = x
        if l["kind"] == "c":
            x = conv2d_same(x, k, b, l["stride"], l["act"], dtype)
        else:
            x = deconv2d(x, k, b, l["act"], dtype)
        if l["res_end"]:
            x = saved + x
            saved = None
        if taps is not None:
            taps.append((l["scope"], x))
    return x


# --------------------------------------------------------------------------------------------
# shared scalar math (single source with the CUDA epilogue: include/tic_math.h via tic_oracle.c)
# --------------------------------------------------------------------------------------------
_CLIB = None


def clib():
    """The C restatement (oracle/tic_oracle.c), built by oracle/Makefile into oracle/_build/."""
    global _CLIB
    if _CLIB is None:
        here = Path(__file__).resolve().parent
        so = here / "_build" / "libtic_oracle.so"
        if not so.exists():
            import subprocess
            subprocess.check_call(["make", "-s", "-C", str(here)])
        lib = ctypes.CDLL(str(so))
        f32p = ctypes.POINTER(ctypes.c_float)
        lib.tico_sigmoid.argtypes = [f32p, f32p, ctypes.c_int64]
        lib.tico_quantize.argtypes = [f32p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int64, ctypes.c_int]
        lib.tico_conv2d_same.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, f32p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        lib.tico_deconv2d.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, f32p,
                                      ctypes.c_int, ctypes.c_int, f32p]
        lib.tico_normalize.argtypes = [f32p, f32p, ctypes.c_int64, f32p, f32p]
        lib.tico_denorm_clip.argtypes = [f32p, f32p, ctypes.c_int64, f32p, f32p]
        for fn in (lib.tico_sigmoid, lib.tico_quantize, lib.tico_conv2d_same, lib.tico_deconv2d, lib.tico_normalize,
                   lib.tico_denorm_clip):
            fn.restype = None
        _CLIB = lib
    return _CLIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def sigmoid_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    clib().tico_sigmoid(_fp(x), _fp(out), x.size)
    return out


def quantize(logits, quan_scale):
    """model_0/model.py:137-138: o = sigmoid(x)*(q-1); forward value (round(o)-o)+o == round(o)
    (tf.round = half-to-even).  Returns uint8 symbols."""
    x = np.ascontiguousarray(logits, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint8)
    clib().tico_quantize(_fp(x), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), x.size, int(quan_scale))
    return out


def quantize_reference_expr(logits, quan_scale):
    """The literal two-line expression of model_0/model.py:137-138 in numpy fp32 (uses the shared sigmoid):
    shows that (round(o) - o) + o == round(o) exactly."""
    o = sigmoid_f32(logits) * np.float32(quan_scale - 1)
    return (np.round(o) - o) + o


def inverse_sigmoid_lut(quan_scale):
    """model_0/model.py:153 + basic_block.reverse_sigmoid (basic_block.py:152-155), fp32 throughout:
    p = (s + 1e-6) / (q - 1 + 1e-5);  log(p / (1 - p)).  One entry per symbol value."""
    s = np.arange(quan_scale, dtype=np.float32)
    p = (s + np.float32(1e-6)) / np.float32(quan_scale - 1 + 1e-5)
    return np.log(p / (np.float32(1) - p)).astype(np.float32)


def normalize(x, mean, std):
    """(x - mean) / std, fp32 (model_0/model.py:44)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    m = np.ascontiguousarray(mean, dtype=np.float32)
    s = np.ascontiguousarray(std, dtype=np.float32)
    clib().tico_normalize(_fp(x), _fp(out), x.size // 3, _fp(m), _fp(s))
    return out


def denorm_clip(y, mean, std):
    """clip(y*std + mean, 0, 255), fp32, Mul then Add (model_0/model.py:251,259)."""
    y = np.ascontiguousarray(y, dtype=np.float32)
    out = np.empty_like(y)
    m = np.ascontiguousarray(mean, dtype=np.float32)
    s = np.ascontiguousarray(std, dtype=np.float32)
    clib().tico_denorm_clip(_fp(y), _fp(out), y.size // 3, _fp(m), _fp(s))
    return out


# --------------------------------------------------------------------------------------------
# model.encoder / model.decoder / rmbe_model.model
# --------------------------------------------------------------------------------------------
def encoder_logits(patches, variant, params, mean, std, dtype=torch.float32):
    """Everything of model.encoder before the quantiser (model_0/model.py:38-134)."""
    v = VARIANTS[variant]
    x = np.asarray(patches, dtype=np.float32)
    if dtype == torch.float64:
        x = (x.astype(np.float64) - np.asarray(mean, np.float32).astype(np.float64)) / np.asarray(std, np.float32).astype(np.float64)
    else:
        x = normalize(x, mean, std)
    return run_layers(x, v["enc"], 3, params, dtype)


def encoder(patches, variant, params, mean, std, quan_scale):
    """model.encoder(input, patch_size, quan_scale) (model_0/model.py:34-144): uint8 symbols [N,hb,wb,cb]."""
    return quantize(encoder_logits(patches, variant, params, mean, std), quan_scale)


def decoder(symbols, variant, params, mean, std, quan_scale, dtype=torch.float32):
    """model.decoder(input, quan_scale) (model_0/model.py:147-263): f32 [N,P,P,3] in [0,255]."""
    v = VARIANTS[variant]
    lut = inverse_sigmoid_lut(quan_scale)
    x = lut[np.asarray(symbols).astype(np.int64)]
    y = run_layers(x, v["dec"], v["bottleneck"], params, dtype)
    if dtype == torch.float64:
        m = np.asarray(mean, np.float32).astype(np.float64)
        s = np.asarray(std, np.float32).astype(np.float64)
        return np.clip(y * s + m, 0.0, 255.0)
    return denorm_clip(y, mean, std)


def postfilter(tiles, name, params, mean, std, dtype=torch.float32):
    """rmbe_model.model(input) (submit/2/rmbe/model.py:113-197): normalise -> 6 layers -> denorm, clip."""
    x = normalize(np.asarray(tiles, dtype=np.float32), mean, std)
    y = run_layers(x, POSTFILTERS[name], 3, params, dtype)
    return denorm_clip(y, mean, std)


# --------------------------------------------------------------------------------------------
# host glue restated: crop / concat / rmbe tiling / serialisation / statistics / metrics
# --------------------------------------------------------------------------------------------
def crop_image_input_patches(image, patch_size):
    """utils/utils.py:96-133: np.pad bottom/right 'reflect' to a multiple of P; row-major P x P crops."""
    h, w, _ = image.shape
    ph = (patch_size - h % patch_size) % patch_size
    pw = (patch_size - w % patch_size) % patch_size
    padded = np.pad(image, ((0, ph), (0, pw), (0, 0)), "reflect")
    H, W, _ = padded.shape
    return [padded[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size, :]
            for i in range(H // patch_size) for j in range(W // patch_size)]


def concat_patches(patches, height, width, patch_size):
    """utils/utils.py:136-167: row-major stitch, crop to [height, width]."""
    hn = -(-height // patch_size)
    wn = -(-width // patch_size)
    rows = [np.concatenate(patches[i * wn:(i + 1) * wn], axis=1) for i in range(hn)]
    return np.concatenate(rows, axis=0)[:height, :width]


def rmbe(image, run_model, patch_size=128, offset=64):
    """submit/2/rmbe/rmbe.py:15-111.  Pass 1 (rmbe_height, :70-89): tiles at rows i*128, cols 64+j*128;
    pass 2 (rmbe_width, :92-111): rows 64+i*128, cols j*128, reading pass-1 output; both write back in
    place (new_image = image[:, :, :] is a view, :81,:103).  run_model(list of tiles) -> [n,128,128,3]."""
    image = np.array(image, dtype=np.float32, copy=True)
    h, w, _ = image.shape
    P = patch_size
    for (oy, ox, hn, wn) in ((0, offset, h // P, (w - offset) // P), (offset, 0, (h - offset) // P, w // P)):
        tiles = [image[oy + i * P:oy + (i + 1) * P, ox + j * P:ox + (j + 1) * P, :].copy()
                 for i in range(hn) for j in range(wn)]
        if not tiles:
            continue
        new = run_model(np.stack(tiles))
        for i in range(hn):
            for j in range(wn):
                image[oy + i * P:oy + (i + 1) * P, ox + j * P:ox + (j + 1) * P, :] = new[i * wn + j]
    return image


def serialize_symbols(encoded_patches):
    """encode.py:171-182: concatenate -> reshape(-1, hb*wb*cb) -> flatten -> int list.  Bitstream order is
    patch-major (row-major patch grid), then h, w, c."""
    arr = np.asarray(encoded_patches)
    return arr.reshape(-1).astype(int).tolist()


def symbol_histogram(symbols, quan_scale):
    """get_encoded_distribution.py:113-134: freq += np.histogram(out, bins=[0..q]); prob = freq / sum."""
    freq = np.zeros(quan_scale)
    bins = [i for i in range(quan_scale + 1)]
    freq += np.histogram(np.asarray(symbols), bins)[0]
    return freq


def position_mean(symbol_batches):
    """cal_encoded_distribution.py:111-149, literally: the fetched batches are float32 tensors; the running mean is
    float64 (np.zeros), the per-batch term is np.sum(float32 batch, axis=0) / n evaluated in float32 (:126-128);
    one_prob = np.mean(seq_prob), prob = [1 - one_prob, one_prob] (:144-145); encoded_order = sorted(range(len), key =
    seq_prob[k]) — a stable sort (:149).  Returns (seq_prob, prob, encoded_order)."""
    n = 0
    seq_prob = None
    for b in symbol_batches:
        encoded_output = np.asarray(b, dtype=np.float32)
        batch_num = encoded_output.shape[0]
        if seq_prob is None:
            seq_prob = np.zeros(int(np.prod(encoded_output.shape[1:])))
        prev_n = n
        n += batch_num
        encoded_seq = np.reshape(encoded_output, (batch_num, -1))
        seq_prob = seq_prob * (1.0 * prev_n / n) + np.sum(encoded_seq, axis=0) / n
    one_prob = np.mean(seq_prob)
    prob = [1.0 - one_prob, one_prob]
    encoded_order = sorted(range(len(seq_prob)), key=lambda k: seq_prob[k])
    return seq_prob, np.asarray(prob), encoded_order


def coder_table(prob, resolution, prob_to_cum_freq):
    """encode.py:76-97 (= decode.py:79-101): freq' = prob*resolution + 1; renormalise; cum_freq table."""
    modified_freq = np.asarray(prob, dtype=np.float64) * resolution + 1
    modified_prob = modified_freq / np.sum(modified_freq)
    return prob_to_cum_freq(modified_prob, resolution=resolution)


def online_mean_and_std_channel(images):
    """processing_utils/get_normalization_params.py:67-111 (streaming per-channel mean/std, 0..255 scale)."""
    n = 0
    mean = 0
    square_mean = 0
    for x in images:
        x = np.asarray(x, dtype=np.float32)
        prev_n = n
        n += x.shape[0] * x.shape[1]
        x = x.reshape([-1, 3])
        square_x = np.square(x)
        square_mean = square_mean * (1.0 * prev_n / n) + np.sum(square_x, axis=0) / n
        mean = mean * (1.0 * prev_n / n) + np.sum(x, axis=0) / n
    var = square_mean - np.square(mean)
    return mean, np.sqrt(var)


def psnr(pairs):
    """processing_utils/evaluate.py:10-30: 20log10(255) - 10log10(sum SE / sum dims)."""
    num = 0
    se = 0.0
    for a, b in pairs:
        a = np.asarray(a, dtype=np.float32)
        b = np.asarray(b, dtype=np.float32)
        num += a.size
        se += float(np.sum(np.square(b - a)))
    return 20.0 * np.log10(255.0) - 10.0 * np.log10(se / num)


def bpp(code_bytes, pixel_num):
    """processing_utils/evaluate.py:44-49: 8 * code size / pixels (H*W, calc_pixel_num.py:21-23)."""
    return code_bytes * 8.0 / pixel_num


def synthetic_image(h, w, seed, kind="natural"):
    """SURVEY.md §8d synthetic inputs: 'natural' = sum of 6 random low-frequency 2-D sinusoids per channel
    + N(0, 8) noise, clipped to 0..255; 'uniform' = uniform noise."""
    rs = np.random.RandomState(seed)
    if kind == "uniform":
        return rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    img = np.zeros((h, w, 3))
    for c in range(3):
        acc = np.full((h, w), 128.0)
        for _ in range(6):
            fy, fx = rs.uniform(-0.02, 0.02, size=2)
            ph = rs.uniform(0, 2 * np.pi)
            acc += rs.uniform(10, 40) * np.sin(2 * np.pi * (fy * yy + fx * xx) + ph)
        img[:, :, c] = acc
    img += rs.normal(0, 8, size=img.shape)
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def around_u8(x):
    """decode.py:249: np.around (half-to-even) -> uint8."""
    return np.around(x).astype(np.uint8)


def codec_roundtrip(image, variant, enc_params, dec_params, mean, std, quan_scale, patch_size):
    """encode.py:153-182 + decode.py:204-249 without the entropy coder: crop -> encoder -> symbols ->
    decoder -> concat -> np.around -> uint8.  Returns (symbols [N,hb,wb,cb] uint8, recon uint8 [H,W,3])."""
    patches = np.stack(crop_image_input_patches(image, patch_size)).astype(np.float32)
    sym = encoder(patches, variant, enc_params, mean, std, quan_scale)
    rec = decoder(sym, variant, dec_params, mean, std, quan_scale)
    h, w, _ = image.shape
    img = concat_patches(list(rec), h, w, patch_size)
    return sym, np.around(img).astype(np.uint8)
