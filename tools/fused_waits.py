"""Where the roles of the fused kernels wait (needs a -DTIC_ABLATE build): cycles inside each barrier wait of cluster 0's
issuer, builder warp 20 (encoder), epilogue warps 4 and 8, per step.      python tools/fused_waits.py [enc|dec]"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tf_image_compression_b200 as T
from tf_image_compression_b200 import _lib as _L

_L.LIB_PATH = _L.LIB_PATH.with_name("libtic_ablate.so")   # tools/build_ablate.sh; the package itself never loads it
assert _L.LIB_PATH.exists(), "run tools/build_ablate.sh first"
from tf_image_compression_b200 import _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "enc"
lib = ctypes.CDLL(str(L.LIB_PATH))
mean = np.array([118.3, 113.9, 102.6], np.float32)
std = np.array([61.7, 59.2, 63.8], np.float32)
codec = T.Codec("model_0", quan_scale=2, mean=mean, std=std, compute="tensor")
codec.use_torch_stream()
n, H, W, P = 64, 1536, 2048, 128
img = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
sym = torch.randint(0, 2, (n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")
rec = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
osym = torch.empty((n, 192, 8, 8, 64), dtype=torch.uint8, device="cuda")


def once():
    if which == "dec":
        codec.decode_images(sym, H, W, P, out=rec)
    else:
        codec.encode_images(img, P, out=osym)


for _ in range(2):
    once()
buf = (ctypes.c_ulonglong * 64)()
assert lib.tic_debug_prof_read(buf, 1) == 0
once()
assert lib.tic_debug_prof_read(buf, 1) == 0
v = list(buf)
steps = max(1, v[5])
names = {0: "issuer  wait acc1_empty", 1: "issuer  wait op_full / in_full", 2: "issuer  wait acc2_empty", 3: "issuer  wait reg_full",
         4: "issuer  total", 8: "builder wait raw_full", 9: "builder wait op_empty", 10: "builder total",
         16: "epi w4  wait acc1_full", 17: "epi w4  wait reg_empty", 18: "epi w4  wait acc2_full (inside phase C)", 19: "epi w4  phase C total",
         20: "epi w4  bar.sync", 21: "epi w4  total",
         24: "epi w8  wait acc1_full", 25: "epi w8  wait reg_empty", 26: "epi w8  wait acc2_full (inside phase C)", 27: "epi w8  phase C total",
         28: "epi w8  bar.sync", 29: "epi w8  total",
         32: "epi w4  bar.sync 3 (acc1_full release)", 33: "epi w4  gather barrier (acc1_empty)", 34: "epi w4  phase A", 35: "epi w4  phase B",
         40: "epi w8  bar.sync 3 (acc1_full release)", 41: "epi w8  gather barrier (reg_full)", 42: "epi w8  phase A", 43: "epi w8  phase B"}
print(f"fused {which}: {steps} steps in cluster 0; cycles per step")
for k in sorted(names):
    if v[k]:
        print(f"  {names[k]:42s} {v[k] / steps:9.1f}")
codec.close()
