"""Stage ablation of the fused kernels (needs a -DTIC_ABLATE build: TIC_DBG is read per launch).
python tools/fused_ablate.py [enc|dec]
bits: 1 no MMA2, 16 no MMA1, 2 no epilogue 2, 4 no region writes (epilogue 1 phase B), 8 no builders (encoder)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tf_image_compression_b200 as T
from tf_image_compression_b200 import _lib as _L

_L.LIB_PATH = _L.LIB_PATH.with_name("libtic_ablate.so")   # tools/build_ablate.sh; the package itself never loads it
assert _L.LIB_PATH.exists(), "run tools/build_ablate.sh first"

which = sys.argv[1] if len(sys.argv) > 1 else "dec"
mean = np.array([118.3, 113.9, 102.6], np.float32)
std = np.array([61.7, 59.2, 63.8], np.float32)
codec = T.Codec("model_0", quan_scale=2, mean=mean, std=std, compute="tensor")
codec.use_torch_stream()
n, H, W, P = 64, 1536, 2048, 128
img = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
sym = torch.randint(0, 2, (n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")
rec = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
osym = torch.empty((n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")


def run(dbg):
    os.environ["TIC_DBG"] = str(dbg)
    codec.set_compute("tensor")  # clears the sticky fp16-range flag an ablated (garbage-producing) run may have raised
    codec.profile(True)
    for _ in range(3):
        try:
            if which == "dec":
                codec.decode_images(sym, H, W, P, out=rec)
            else:
                codec.encode_images(img, P, out=osym)
        except T.TicError:
            codec.set_compute("tensor")
    torch.cuda.synchronize()
    t = {l.scope: ms / max(c, 1) for l, ms, c in codec.layer_times("decoder" if which == "dec" else "encoder")}
    codec.profile(False)
    return t["decode_1"] if which == "dec" else t["encode_0"]


run(0)  # warm-up: lazy weight images
cases = [(0, "full"), (2, "no epilogue 2"), (4, "no region writes (epilogue 1 phase B)"), (6, "no epilogue 2, no region writes"),
         (1, "no MMA2"), (16, "no MMA1"), (17, "no MMA at all"), (23, "no MMA, no epilogue work"),
         (19, "no MMA, no epilogue 2 (epilogue 1 alone)"), (21, "no MMA, no region writes (epilogue 2 alone)")]
if which == "enc":
    cases += [(8, "no builders"), (25, "no MMA, no builders"), (31, "skeleton: no MMA, no builders, no epilogue work"),
              (23 - 0, "builders alone (no MMA, no epilogue work)")]
for dbg, what in cases:
    print(f"TIC_DBG={dbg:3d} {what:50s} fused {which} pair {run(dbg):7.3f} ms")
codec.close()
