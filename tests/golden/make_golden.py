"""Generates tests/golden/cfg1_model0_768x512.npz with the CPU oracle (oracle/codec_oracle.py).

The reference itself (TensorFlow 1.x graph code, no weights, no saved tensors) cannot run here or
anywhere without TF, so this fixture is ORACLE output, not reference output: it freezes the oracle's
answer for BASELINE config 1 so the GPU test needs no CPU re-evaluation and any later drift of either
side is caught.  Run from the repo root:  python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import codec_oracle as O  # noqa: E402


def main():
    image = O.synthetic_image(512, 768, 1234)  # "768x512" (W x H)
    ov = O.VARIANTS["model_0"]
    enc = O.init_params(ov["enc"], 3, 1234, "fanin")
    dec = O.init_params(ov["dec"], ov["bottleneck"], 1235, "fanin")
    dec = O.condition_decoder("model_0", dec, 2)  # O(1) pre-denormalisation output
    mean, std = O.online_mean_and_std_channel([image])
    mean = np.asarray(mean, np.float32)
    std = np.asarray(std, np.float32)
    sym, recon = O.codec_roundtrip(image, "model_0", enc, dec, mean, std, 2, 128)
    out = dict(image=image, mean=mean, std=std, symbols_packed=np.packbits(sym.reshape(-1)), recon=recon)
    # weights are NOT stored: they are re-drawn from the same seeds (numpy RandomState is stable);
    # a checksum guards against a silent change of the generator
    out["weight_checksum"] = np.array([float(sum(np.float64(v).sum() for v in enc.values())),
                                       float(sum(np.float64(v).sum() for v in dec.values()))])
    path = Path(__file__).parent / "cfg1_model0_768x512.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes; symbols", sym.shape, "ones", int(sym.sum()))


if __name__ == "__main__":
    main()
