"""The reference's entry-point flows on top of the B200 path: compress (encode.py:125-212), uncompress
(decode.py:143-251, with the submit/2 post-filter call order submit/2/decoder.py:183-198), the one-graph round
trip (test.py:95-146) and the dataset-wide
symbol table (get_encoded_distribution.py:85-155).  File names, config keys, table construction and the
bitstream order (patch-major, then h, w, c: encode.py:171-182) are the reference's; images are processed in
batches through Codec.encode_images / decode_images instead of one sess.run per 64 patches."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from . import parallel, range_coder

DEFAULT_CONFIG = {"name_sep": "@_@", "resolution": 4096, "patch_size": 128, "quan_scale": 2}  # model_N/config.json keys


def cum_freq_table(prob, resolution):
    """encode.py:76-91 == decode.py:79-93: freq = prob * resolution + 1 (no zero-probability symbol), renormalise,
    range_coder.prob_to_cum_freq(modified_prob, resolution)."""
    prob = np.asarray(prob, dtype=np.float64)
    modified_freq = prob * resolution + 1
    modified_prob = modified_freq / np.sum(modified_freq)
    return range_coder.prob_to_cum_freq(modified_prob, resolution=resolution)


def encoded_name(image_name, image_shape, seq_data_len, encoded_patches_shape, config):
    """encode.py:102-122: '<name>@_@<eh>_<ew>_<ec>@_@<len>_<H>_<W>.encoded'."""
    sep = config["name_sep"]
    stem = str(image_name).split("/")[-1].replace(".png", "")
    eh, ew, ec = encoded_patches_shape
    height, width = image_shape[0], image_shape[1]
    return f"{stem}{sep}{eh}_{ew}_{ec}{sep}{seq_data_len}_{height}_{width}.encoded"


def parse_encoded_name(filename, config):
    """decode.py:104-140: -> (stem, (eh, ew, ec), seq_data_len, height, width)."""
    sep = config["name_sep"]
    parts = str(filename).replace(".encoded", "").split(sep)
    if len(parts) < 3:
        raise ValueError(f"not an encoded file name: {filename!r}")
    eh, ew, ec = (int(v) for v in parts[1].split("_"))
    n, h, w = (int(v) for v in parts[-1].split("_"))
    return parts[0], (eh, ew, ec), n, h, w


def compress(codec, images, names, config, prob, output_dir):
    """Encode a list of uint8 [H,W,3] images (any sizes) and write one range-coded file per image.
    Images of equal size go through the codec as one batch.  Returns [(path, n_bytes)] in input order."""
    P = int(config["patch_size"])
    cum = cum_freq_table(prob, int(config["resolution"]))
    out = [None] * len(images)
    os.makedirs(output_dir, exist_ok=True)
    by_shape = {}
    for i, im in enumerate(images):
        by_shape.setdefault(tuple(im.shape), []).append(i)
    hb, wb, cb = codec.bottleneck_shape(P)
    for shape, idx in by_shape.items():
        batch = np.ascontiguousarray(np.stack([images[i] for i in idx]), dtype=np.uint8)
        sym = codec.encode_images(batch, P)  # [B, gh*gw, hb, wb, cb] uint8, patch-major
        for k, i in enumerate(idx):
            seq = sym[k].reshape(-1)
            path = str(Path(output_dir) / encoded_name(names[i], shape, seq.size, (hb, wb, cb), config))
            enc = range_coder.RangeEncoder(path)
            enc.encode(seq, cum)
            enc.close()
            out[i] = (path, enc.bytes_written)
    return out


def uncompress(codec, input_dir, config, prob, postfilter=False):
    """Decode every '*.encoded' file of input_dir -> {stem: uint8 [H,W,3]} (np.around of the stitched float image,
    decode.py:249; with postfilter=True the float image goes through rmbe first, submit/2/decoder.py:183-198)."""
    P = int(config["patch_size"])
    cum = cum_freq_table(prob, int(config["resolution"]))
    files = sorted(f for f in os.listdir(input_dir) if f.endswith(".encoded"))
    groups = {}
    for f in files:
        stem, eshape, n, h, w = parse_encoded_name(f, config)
        dec = range_coder.RangeDecoder(str(Path(input_dir) / f))
        seq = dec.decode(n, cum, dtype=np.uint8)
        dec.close()
        groups.setdefault((h, w, eshape), []).append((stem, seq))
    result = {}
    for (h, w, eshape), items in groups.items():
        gh, gw = -(-h // P), -(-w // P)
        sym = np.stack([s.reshape(gh * gw, *eshape) for _, s in items])
        if postfilter:
            rec = codec.decode_images(sym, h, w, P, out_dtype=np.float32)
            codec.postfilter_images(rec)
            rec = codec.round_u8(rec)
        else:
            rec = codec.decode_images(sym, h, w, P, out_dtype=np.uint8)
        for k, (stem, _) in enumerate(items):
            result[stem] = rec[k]
    return result


def compress_and_uncompress(codec, images, config):
    """test.py:95-146: encoder and decoder in one graph, no bitstream — the reference's quality check.  Images of
    equal size go through Codec.roundtrip_images as one batch.  Returns the uint8 reconstructions in input order
    (io.imsave of the stitched float image rounds like decode.py:249)."""
    P = int(config["patch_size"])
    out = [None] * len(images)
    by_shape = {}
    for i, im in enumerate(images):
        by_shape.setdefault(tuple(im.shape), []).append(i)
    for shape, idx in by_shape.items():
        batch = np.ascontiguousarray(np.stack([images[i] for i in idx]), dtype=np.uint8)
        rec, _ = codec.roundtrip_images(batch, P, want_symbols=False)
        for k, i in enumerate(idx):
            out[i] = rec[k]
    return out


def get_distribution(codec, patches, group=None):
    """get_encoded_distribution.py:85-155 over this rank's shard of the patch list: encode (the histogram is fused
    into the bottleneck epilogue), all-reduce the q counts over the ranks, prob = freq / sum(freq)."""
    codec.hist_reset()
    if len(patches):
        codec.encode_patches(patches)
    return parallel.distribution(parallel.allreduce_histogram(codec, group))


def psnr(originals, reconstructions):
    """processing_utils/evaluate.py:10-30: 20 log10(255) - 10 log10(sum SE / sum HWC) over the whole set."""
    se, cnt = 0.0, 0
    for a, b in zip(originals, reconstructions):
        d = a.astype(np.float64) - b.astype(np.float64)
        se += float((d * d).sum())
        cnt += d.size
    return 20.0 * np.log10(255.0) - 10.0 * np.log10(se / cnt)


def bpp(n_bytes_list, images):
    """processing_utils/evaluate.py:33-49: 8 * sum(file bytes) / sum(H * W)."""
    return 8.0 * float(sum(n_bytes_list)) / float(sum(im.shape[0] * im.shape[1] for im in images))
