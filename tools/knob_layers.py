"""Per-layer times of one direction under an ablation-build knob: python tools/knob_layers.py dec TIC_DECONV_TAP_SLICES 0 1
(needs tools/build_ablate.sh; the knob is read at every launch)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tf_image_compression_b200 as T
from tf_image_compression_b200 import _lib as _L

_L.LIB_PATH = _L.LIB_PATH.with_name("libtic_ablate.so")
assert _L.LIB_PATH.exists(), "run tools/build_ablate.sh first"
which, knob, values = sys.argv[1], sys.argv[2], sys.argv[3:]
mean = np.array([118.3, 113.9, 102.6], np.float32)
std = np.array([61.7, 59.2, 63.8], np.float32)
VARIANT = os.environ.get("VARIANT", "model_0")
codec = T.Codec(VARIANT, quan_scale=2, mean=mean, std=std, compute="tensor")
codec.use_torch_stream()
n, H, W, P = int(os.environ.get("NIMG", "64")), 1536, 2048, 128
hb, wb, cb = codec.bottleneck_shape(P)
img = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
sym = torch.randint(0, 2, (n, 192, hb, wb, cb), dtype=torch.uint8, device="cuda")
rec = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
osym = torch.empty((n, 192, hb, wb, cb), dtype=torch.uint8, device="cuda")
ref = None
for v in values:
    os.environ[knob] = v
    for _ in range(2):
        (codec.decode_images(sym, H, W, P, out=rec) if which == "dec" else codec.encode_images(img, P, out=osym))
    codec.profile(True)
    for _ in range(3):
        (codec.decode_images(sym, H, W, P, out=rec) if which == "dec" else codec.encode_images(img, P, out=osym))
    torch.cuda.synchronize()
    t = [(l.scope, ms / 3, c // 3) for l, ms, c in codec.layer_times("decoder" if which == "dec" else "encoder")]
    codec.profile(False)
    out = (rec if which == "dec" else osym).clone()
    same = "-" if ref is None else str(bool(torch.equal(out, ref)))
    ref = out if ref is None else ref
    print(f"{knob}={v}: total {sum(x[1] for x in t):.3f} ms; same output as first: {same}")
    for s, ms, c in t:
        print(f"    {s:26s} {ms:7.3f} ms  {c} launches")
codec.close()
