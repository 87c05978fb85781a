"""Stage ablation of the fused decoder kernel (needs a -DTIC_ABLATE build: TIC_DBG is read per launch).
python tools/fused_ablate.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tf_image_compression_b200 as T

mean = np.array([118.3, 113.9, 102.6], np.float32)
std = np.array([61.7, 59.2, 63.8], np.float32)
codec = T.Codec("model_0", quan_scale=2, mean=mean, std=std, compute="tensor")
codec.use_torch_stream()
n, H, W, P = 64, 1536, 2048, 128
sym = torch.randint(0, 2, (n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")
rec = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")


def run(dbg):
    os.environ["TIC_DBG"] = str(dbg)
    codec.profile(True)
    for _ in range(3):
        codec.decode_images(sym, H, W, P, out=rec)
    torch.cuda.synchronize()
    t = {l.scope: ms / max(c, 1) for l, ms, c in codec.layer_times("decoder")}
    codec.profile(False)
    return t["decode_1"], t["decode_2"]


for dbg, what in ((0, "full"), (2, "no epilogue 2"), (4, "no region writes (epilogue 1 phase B)"), (6, "no epilogue 2, no region writes"),
                  (1, "no MMA2"), (16, "no MMA1"), (17, "no MMA at all"), (23, "skeleton: no MMA, no epilogue work"),
                  (19, "no MMA, no epilogue 2 (epilogue 1 alone)"), (21, "no MMA, no region writes (epilogue 2 alone)")):
    f, d2 = run(dbg)
    print(f"TIC_DBG={dbg:3d} {what:50s} fused decode_1+decode_0 {f:7.3f} ms   (decode_2 {d2:6.3f} ms)")
codec.close()
