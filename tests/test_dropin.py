"""The drop-in claim (north_star: "encode.py / decode.py ... stay unchanged"): the reference's UNMODIFIED entry points,
imported from /root/reference with `tensorflow`, `range_coder` and `skimage.io` stood in by
tf_image_compression_b200.compat, run compress() then uncompress() on synthetic images, and their files / reconstructions
are compared byte for byte with entry.compress / entry.uncompress over the same codec.

On this CPU box the codec behind the stand-ins is the oracle adapter (tests/oracle_codec.py): that pins the whole host
side (compat, model_api, entry, range_coder, checkpoint, file naming).  tests/test_gpu_parity2.py runs the same scripts
over the real Codec where both a GPU and a reference checkout exist, and the GPU parity tests pin Codec == oracle."""
import argparse
import json
import os
from pathlib import Path

import numpy as np
import pytest

from oracle import codec_oracle as O
from tf_image_compression_b200 import checkpoint as K
from tf_image_compression_b200 import compat, entry

REF = Path(os.environ.get("TIC_REFERENCE_ROOT", "/root/reference"))
MEAN = np.array([118.3, 113.9, 102.6], np.float32)
STD = np.array([61.7, 59.2, 63.8], np.float32)

pytestmark = pytest.mark.skipif(not (REF / "encode.py").exists(), reason="no reference checkout (GPU box)")


def run_reference_roundtrip(codec, work, images, names, prob, ckpt_params, model_num=0):
    """Lay out a reference working directory (config.json copied verbatim, distribution table, image list, PNG
    inputs, model_N/params_for_test/params as a TF-V2 bundle), then call the unmodified compress() and uncompress()."""
    from PIL import Image
    work = Path(work)
    (work / f"model_{model_num}").mkdir(parents=True)
    (work / "data_info").mkdir()
    (work / "images").mkdir()
    cfg_text = (REF / f"model_{model_num}" / "config.json").read_text()
    (work / f"model_{model_num}" / "config.json").write_text(cfg_text)
    np.save(work / "data_info" / f"distribution_info_{model_num}.npy", prob)
    paths = []
    for im, nm in zip(images, names):
        p = work / "images" / (nm + ".png")
        Image.fromarray(im).save(p)
        paths.append(str(p))
    (work / "data_info" / "list.txt").write_text("\n".join(paths) + "\n")
    (work / f"model_{model_num}" / "params_for_test").mkdir()
    K.write_checkpoint(str(work / f"model_{model_num}" / "params_for_test" / "params"), ckpt_params)
    cwd = os.getcwd()
    os.chdir(work)
    try:
        with compat.installed(codec, reference_root=str(REF)):
            enc = compat.load_entry(REF / "encode.py")
            dec = compat.load_entry(REF / "decode.py")
            args = argparse.Namespace(model_num=str(model_num), gpu_num="0", debug_mode="off", params_file="",
                                      data_list="data_info/list.txt", output_dir="model_{}/encoded_data")
            enc.compress(compat.Session(), compat.model_module(model_num), args)
            dargs = argparse.Namespace(model_num=str(model_num), gpu_num="0", debug_mode="off", params_file="",
                                       input_dir="model_{}/encoded_data", output_dir="model_{}/recons_data")
            dec.uncompress(compat.Session(), compat.model_module(model_num), dargs)
    finally:
        os.chdir(cwd)
    enc_dir = work / f"model_{model_num}" / "encoded_data"
    rec_dir = work / f"model_{model_num}" / "recons_data"
    files = {f: (enc_dir / f).read_bytes() for f in sorted(os.listdir(enc_dir))}
    recs = {f[:-4]: np.asarray(Image.open(rec_dir / f)) for f in sorted(os.listdir(rec_dir))}
    return json.loads(cfg_text), files, recs


def check_against_entry_flows(codec, cfg, files, recs, images, names, prob, tmp_path, coders=("host", "serial")):
    """entry.compress / entry.uncompress over the same codec: byte-identical files, identical reconstructions."""
    out = entry.compress(codec, images, names, cfg, prob, str(tmp_path / "own"), coder=coders[0])
    assert sorted(os.path.basename(p) for p, _ in out) == sorted(files)
    for p, nbytes in out:
        assert open(p, "rb").read() == files[os.path.basename(p)] and nbytes == len(files[os.path.basename(p)])
    for c in coders[1:]:
        other = entry.compress(codec, images, names, cfg, prob, str(tmp_path / ("own_" + c)), coder=c)
        for (p, _), (q, _) in zip(out, other):
            assert open(p, "rb").read() == open(q, "rb").read(), c
    for c in coders:
        own = entry.uncompress(codec, str(tmp_path / "own"), cfg, prob, coder=c)
        assert sorted(own) == sorted(recs) == sorted(names)
        for nm, im in zip(names, images):
            assert recs[nm].shape == im.shape and np.array_equal(own[nm], recs[nm]), (c, nm)
    # file names carry the reference's metadata (encode.py:102-122)
    P = cfg["patch_size"]
    for f in files:
        stem, eshape, n, h, w = entry.parse_encoded_name(f, cfg)
        assert n == (-(-h // P)) * (-(-w // P)) * eshape[0] * eshape[1] * eshape[2]


def test_reference_entry_points_run_unmodified_on_the_oracle_codec(tmp_path):
    from oracle_codec import OracleCodec
    ov = O.VARIANTS["model_0"]
    enc = O.init_params(ov["enc"], 3, 1234, "fanin")
    dec = O.condition_decoder("model_0", O.init_params(ov["dec"], ov["bottleneck"], 1235, "fanin"), 2)
    # the codec starts from OTHER weights: the reference flow itself must restore the checkpoint (utils.restore_params)
    codec = OracleCodec("model_0", 2, MEAN, STD, O.init_params(ov["enc"], 3, 99, "fanin"),
                        O.init_params(ov["dec"], ov["bottleneck"], 98, "fanin"))
    images = [O.synthetic_image(256, 512, 3), O.synthetic_image(300, 260, 4)]  # the second needs reflect padding at P = 256
    names = ["kodim_a", "kodim_b"]
    prob = np.array([0.55, 0.45])
    cfg, files, recs = run_reference_roundtrip(codec, tmp_path / "ref", images, names, prob, {**enc, **dec})
    assert cfg["patch_size"] == 256 and cfg["quan_scale"] == 2  # the reference's own model_0/config.json
    assert all(np.array_equal(codec.enc_params[k], enc[k]) for k in enc)  # restore_params replaced the weights
    assert all(np.array_equal(codec.dec_params[k], dec[k]) for k in dec)
    # one sess.run per <= 64 patches, like the reference (2 and 4 patches here)
    assert ("encode_patches", 2) in codec.calls and ("encode_patches", 4) in codec.calls
    assert ("decode_patches", 2) in codec.calls and ("decode_patches", 4) in codec.calls
    check_against_entry_flows(codec, cfg, files, recs, images, names, prob, tmp_path)
    assert all(entry.parse_encoded_name(f, cfg)[1] == (16, 16, 64) for f in files)
