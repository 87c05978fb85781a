"""BASELINE config 5 timing: model_1 decode of 2048x1536 images at P = 256 + rmbe post-filter (356 tiles per image),
device-resident, CUDA events.  python tools/cfg5_timing.py [n_images]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tf_image_compression_b200 as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W, P = 1536, 2048, 256
mean = np.array([118.3, 113.9, 102.6], np.float32)
std = np.array([61.7, 59.2, 63.8], np.float32)
codec = T.Codec("model_1", quan_scale=2, mean=mean, std=std, compute="tensor")
codec.set_postfilter()
codec.use_torch_stream()
hb, wb, cb = codec.bottleneck_shape(P)
sym = torch.randint(0, 2, (n, (H // P) * (W // P), hb, wb, cb), dtype=torch.uint8, device="cuda")
rec = torch.empty((n, H, W, 3), dtype=torch.float32, device="cuda")
out = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t_dec = timed(lambda: codec.decode_images(sym, H, W, P, out=rec))
t_rmbe = timed(lambda: codec.postfilter_images(rec))
t_round = timed(lambda: codec.round_u8(rec, out=out))
px = n * H * W / 1e6
print(f"cfg5 {n} images: decode {t_dec:.2f} ms ({px / t_dec * 1e3:.0f} Mpixel/s), rmbe {t_rmbe:.2f} ms ({px / t_rmbe * 1e3:.0f} Mpixel/s), "
      f"round {t_round:.2f} ms; total {px / (t_dec + t_rmbe + t_round) * 1e3:.0f} Mpixel/s")
codec.profile(True)
codec.postfilter_images(rec)
torch.cuda.synchronize()
for l, ms, c in codec.layer_times("postfilter"):
    print(f"  rmbe {l.scope:10s} {ms:7.3f} ms in {c} launches")
codec.close()
