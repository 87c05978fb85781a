"""Workload for ncu captures of one fused kernel (shipped library): python tools/run_fused.py [enc|dec] [repeats]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tf_image_compression_b200 as T

which = sys.argv[1] if len(sys.argv) > 1 else "enc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mean = np.array([118.3, 113.9, 102.6], np.float32)
std = np.array([61.7, 59.2, 63.8], np.float32)
codec = T.Codec("model_0", quan_scale=2, mean=mean, std=std, compute="tensor")
codec.use_torch_stream()
n, H, W, P = 64, 1536, 2048, 128
img = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
sym = torch.randint(0, 2, (n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")
rec = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
osym = torch.empty((n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    if which == "dec":
        codec.decode_images(sym, H, W, P, out=rec)
    else:
        codec.encode_images(img, P, out=osym)
torch.cuda.synchronize()
codec.close()
