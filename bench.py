#!/usr/bin/env python
"""Benchmark of the codec hot path (encode + decode) on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  metric   : encode+decode Mpixel/s (BASELINE.json), image pixels H*W counted once per round trip
  workload : BASELINE config 2 — model_0, 64 synthetic 2048x1536 RGB images in 128x128 patches
  value    : device-resident throughput, WEAK arm: every rank encodes + decodes the 64 images (inputs already in HBM,
             CUDA events, max over ranks); no data-path collective
  strong   : the same 64 images in total, patch-sharded 64 / N whole images per GPU (BASELINE config 2 as written:
             "8xB200 patch-sharded"), with the path's one collective — the dataset-wide symbol-frequency table
             (get_encoded_distribution.py:113-134) all-reduced over NCCL once per step and checked against the
             single-rank table; at N = 1 it is the weak arm
  e2e      : same metric through the public API with pinned HOST buffers (H2D + kernels + D2H), plus the
             one-directional encode-only / decode-only flows of encode.py / decode.py and the encode flow with the GPU
             entropy stage (only compressed bytes come back)
  roofline : dominant kernel, algorithmic bytes (or FLOPs) per launch / CUDA-event time vs MEASURED_PEAKS.json
  range_coder : entropy-coding time of the step's symbols (excluded from the metric but printed): host thread
             pool and GPU entropy stage
  configs  : BASELINE configs 3, 4, 5 as sub-records (rank 0, N = 1 only), each with its own roofline
  cpu_baseline : the CPU oracle (torch-CPU fp32 restatement; TensorFlow is not installable) on a
             bounded sample of the same workload, timed on this box's host cores (rank 0, N=1 only)
`--impl reference` times that CPU restatement alone (rank 0) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "encode+decode Mpixel/s"
UNIT = "Mpixel/s"
VARIANT = "model_0"
P = 128
IMG_H, IMG_W = 1536, 2048
MEAN = np.array([118.3, 113.9, 102.6], np.float32)
STD = np.array([61.7, 59.2, 63.8], np.float32)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=64, help="images per step (BASELINE config 2: 64)")
    ap.add_argument("--compute", default=os.environ.get("TIC_COMPUTE", "auto"), help="fp32 | tensor | tf32 | auto")
    ap.add_argument("--cpu-images", type=int, default=2, help="images in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 3/4/5 sub-records")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("TIC_CHUNK", "0")), help="patches per launch sequence (0: library default)")
    ap.add_argument("--layers", action="store_true", help="print the per-layer time table to stderr")
    ap.add_argument("--variant", default=VARIANT, help="model variant (default: model_0 = BASELINE config 2; e.g. base_model/ch_128 for config 3)")
    ap.add_argument("--patch", type=int, default=P, help="patch size")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement (kind "port": TF itself cannot be installed here)
# ------------------------------------------------------------------------------------------------
def oracle_params():
    from oracle import codec_oracle as O
    ov = O.VARIANTS[VARIANT]
    enc = O.init_params(ov["enc"], 3, 1234, "reference")
    dec = O.init_params(ov["dec"], ov["bottleneck"], 1235, "reference")
    return enc, dec


def cpu_roundtrip_time(n_images, reps):
    """Seconds per round trip (crop -> encoder -> symbols -> decoder -> stitch -> uint8) of n_images
    2048x1536 images on the host cores; best of reps."""
    import torch
    from oracle import codec_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    enc, dec = oracle_params()
    imgs = [O.synthetic_image(IMG_H, IMG_W, 1234 + i, "uniform") for i in range(n_images)]
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        for im in imgs:
            O.codec_roundtrip(im, VARIANT, enc, dec, MEAN, STD, 2, P)
        best = min(best, time.perf_counter() - t0)
    return best, torch.get_num_threads()


def parity_sample(compute):
    """BASELINE metric's parity part on a bounded sample, next to the CPU baseline (rank 0, N = 1): one 256x384 synthetic
    image, fan-in weights, GPU path vs the oracle — symbol mismatches, reconstruction difference, bpp and PSNR deltas
    (processing_utils/evaluate.py:10-49) with the same table and range coder on both symbol streams."""
    import tempfile
    import tf_image_compression_b200 as T
    from tf_image_compression_b200 import entry, range_coder
    from oracle import codec_oracle as O
    ov = O.VARIANTS["model_0"]
    enc = O.init_params(ov["enc"], 3, 1234, "fanin")
    dec = O.condition_decoder("model_0", O.init_params(ov["dec"], ov["bottleneck"], 1235, "fanin"), 2)
    image = O.synthetic_image(256, 384, 7)
    with T.Codec("model_0", quan_scale=2, mean=MEAN, std=STD, enc_params=enc, dec_params=dec, compute=compute) as c:
        c.hist_reset()
        rec, sym = c.roundtrip_images(image[None], 128)
        counts = c.hist_read()
        cum = entry.cum_freq_table(counts / counts.sum(), 4096)
        gpu_stream, gpu_nb = c.entropy_encode(sym, cum)  # the GPU entropy stage on the same symbols
    s_ref, r_ref = O.codec_roundtrip(image, "model_0", enc, dec, MEAN, STD, 2, 128)
    sizes, blobs = [], []
    with tempfile.TemporaryDirectory() as d:
        for k, stream in enumerate((sym[0].reshape(-1), s_ref.reshape(-1).astype(np.uint8))):
            e = range_coder.RangeEncoder(os.path.join(d, f"{k}.bin"))
            e.encode(stream, cum)
            e.close()
            sizes.append(os.path.getsize(os.path.join(d, f"{k}.bin")))
            blobs.append(open(os.path.join(d, f"{k}.bin"), "rb").read())
    px = image.shape[0] * image.shape[1]
    diff = np.abs(rec[0].astype(int) - r_ref.astype(int))
    return {"sample": "one 256x384 synthetic image, model_0, fan-in weights, 6 patches",
            "symbols": int(s_ref.size), "symbol_mismatches": int((sym[0].reshape(s_ref.shape) != s_ref).sum()),
            "recon_max_abs_diff_u8": int(diff.max()), "recon_pixels_differing": int((diff != 0).sum()),
            "bpp": 8.0 * sizes[0] / px, "bpp_delta": 8.0 * (sizes[0] - sizes[1]) / px,
            "bitstream_identical_to_oracle_symbols": bool(blobs[0] == blobs[1]),
            "gpu_entropy_stage_identical_to_host_coder": bool(bytes(gpu_stream[0, :int(gpu_nb[0])]) == blobs[0]),
            "psnr_db": float(entry.psnr([image], [rec[0]])),
            "psnr_delta_db": float(entry.psnr([image], [rec[0]]) - entry.psnr([image], [r_ref]))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = 1  # bounded sample per step: one 2048x1536 image = 192 patches
    import torch
    from oracle import codec_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    enc, dec = oracle_params()
    img = O.synthetic_image(IMG_H, IMG_W, 1234, "uniform")
    for _ in range(args.warmup):
        O.codec_roundtrip(img, VARIANT, enc, dec, MEAN, STD, 2, P)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.codec_roundtrip(img, VARIANT, enc, dec, MEAN, STD, 2, P)
    dt = time.perf_counter() - t0
    mpx = n_img * IMG_H * IMG_W * args.steps / dt / 1e6
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": mpx, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{VARIANT} encode+decode, 2048x1536 RGB images in {P}x{P} patches (BASELINE config 2); "
                               f"CPU arm: bounded sample of {n_img} image (192 patches) per step", "patch_size": P},
        "cpu_baseline": {"value": mpx, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_img} image 2048x1536 per step x {args.steps} steps; torch-CPU fp32 oracle "
                                   "(TensorFlow is not installable in this image)"},
        "e2e": {"value": mpx, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (pynvml)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        self.init_error = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:
            self.ok = False
            self.init_error = repr(e)

    def sample(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception as e:  # keep the reason: "unavailable" alone does not say why
            self.error = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            self.sample()
            time.sleep(0.005)

    def result(self):
        """Call while the GPU is still busy with the timed region (before the closing synchronize): takes one more
        sample itself, so that even a timed region shorter than the thread's wake-up latency is covered."""
        if self.ok:
            self.sample()
        self.stop_flag.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "error": getattr(self, "error", self.init_error)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# per-layer roofline of one profiled pass
# ------------------------------------------------------------------------------------------------
def layer_rows(codec, graphs, npatch, patch, hb, passes):
    """[(graph, layer, ms per step, launches per step, algorithmic flops, algorithmic bytes)] from the handle's per-layer
    CUDA-event profile.  Algorithmic bytes of a layer per patch: its input and output tensors once (activations are 4 B
    per element: fp16 pair planes or fp32; u8 images / symbols / reconstructions are 1 B) + weights."""
    rows = []
    for graph, layers, h0, first_u8, last_bytes in graphs:
        hh = h0
        first, last = layers[0], layers[-1]
        for (l, tot_ms, n_launch) in codec.layer_times(graph):
            hin = hh
            if l.kind == "c":
                hh = -(-hh // l.stride)
                macs = hh * hh * 9 * l.cin * l.cout
            else:
                macs = hh * hh * 9 * l.cin * l.cout
                hh *= 2
            in_b = hin * hin * l.cin * (first_u8 if l is first else 4)
            out_b = hh * hh * l.cout * (last_bytes if l is last else 4)
            if n_launch == 0 and tot_ms == 0.0 and rows and rows[-1]["graph"] == graph:
                # this layer ran INSIDE the previous layer's launch (fused back-to-back kernel, tic_fused16.cuh): one row,
                # both layers' FLOPs, and only the pair's outer tensors as algorithmic bytes (the tensor between them never
                # leaves the SM)
                prev = rows[-1]
                prev["scope"] += "+" + l.scope
                prev["flops_step"] += 2.0 * macs * npatch
                prev["bytes_step"] += float(out_b - in_b) * npatch + 9.0 * l.cin * l.cout * 4
                prev["fused"] = True
                continue
            rows.append(dict(graph=graph, scope=l.scope, ms_step=tot_ms / passes, launches=max(1, n_launch // passes),
                             flops_step=2.0 * macs * npatch, bytes_step=float(in_b + out_b) * npatch + 9.0 * l.cin * l.cout * 4))
    return rows


def roofline_of(rows, compute, variant, patch):
    pk = peaks()
    step_ms = sum(r["ms_step"] for r in rows)
    top = max(rows, key=lambda r: r["ms_step"])
    ridge = pk["bf16_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    top_launch_s = top["ms_step"] / top["launches"] * 1e-3
    if top["flops_step"] / top["bytes_step"] >= ridge:
        achieved = top["flops_step"] / top["launches"] / top_launch_s / 1e12
        rf = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
              "frac": achieved / pk["bf16_sustained"], "peak_source": f"bf16 sustained, {pk['source']}"}
    else:
        achieved = top["bytes_step"] / top["launches"] / top_launch_s / 1e9
        rf = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
              "peak_source": f"copy bandwidth, {pk['source']}"}
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists() and compute == "tensor":
        # DRAM bytes of this kernel from the committed ncu --set full capture, scaled to this run's patches per launch
        td = json.loads(tp.read_text())
        per = td.get(variant, {}).get(f"{top['graph']}/{top['scope']}")
        if per is not None and patch == 128:
            traffic = per / td["patches_per_launch"] * top["npatch"] / top["launches"] if "npatch" in top else None
        lim = td.get("limiters", {}).get(variant, {}).get(f"{top['graph']}/{top['scope']}")
        if lim is not None and patch == 128:
            rf["ncu_limiters"] = lim   # committed ncu evidence for what bounds this kernel when it is neither HBM nor tensor math
    rf.update({"traffic": traffic, "kernel": f"{top['graph']}/{top['scope']} ({compute})",
               "kernel_ms_per_launch": top_launch_s * 1e3, "kernel_share_of_step": top["ms_step"] / max(step_ms, 1e-9),
               "kernel_algorithmic_flop_per_byte": top["flops_step"] / top["bytes_step"], "ridge_flop_per_byte": ridge,
               "step_ms_sum_of_layers": step_ms,
               "step_algorithmic_tflops": sum(r["flops_step"] for r in rows) / max(step_ms, 1e-9) / 1e9,
               "step_layer_io_gbs": sum(r["bytes_step"] for r in rows) / max(step_ms, 1e-9) / 1e6})
    return rf


def print_rows(rows, title):
    print(f"--- {title}", file=sys.stderr)
    for r in rows:
        t = max(r["ms_step"], 1e-9) * 1e-3
        print(f"  {r['graph']:10s} {r['scope']:22s} {r['ms_step']:8.3f} ms/step {r['launches']:3d} launches  "
              f"{r['flops_step'] / t / 1e12:8.2f} TFLOP/s {r['bytes_step'] / t / 1e9:8.1f} GB/s", file=sys.stderr)


def cuda_time(torch, fn, reps=3):
    """Best-of-reps device time (ms) of fn() on torch's current stream."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float("inf")
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


# ------------------------------------------------------------------------------------------------
# BASELINE configs 3, 4, 5 (rank 0, N = 1): sub-records with their own roofline
# ------------------------------------------------------------------------------------------------
def config_records(args, compute, dev, show_layers):
    import torch
    import tf_image_compression_b200 as T
    from tf_image_compression_b200 import entry, range_coder
    recs = []
    n_img = 16
    g = torch.Generator(device="cpu").manual_seed(4321)
    imgs = torch.randint(0, 256, (n_img, IMG_H, IMG_W, 3), dtype=torch.uint8, generator=g).to(dev)
    px = n_img * IMG_H * IMG_W

    def codec_for(variant):
        c = T.Codec(variant, quan_scale=2, mean=MEAN, std=STD, device=dev.index, compute=compute, seed=1234)
        c.use_torch_stream()
        return c

    # ---- config 3: base_model/ch_128 at its native 128 and at 256, base_model/input_256 at 256 --------------------
    for variant, patch in (("base_model/ch_128", 128), ("base_model/ch_128", 256), ("base_model/input_256", 256)):
        c = codec_for(variant)
        gh, gw = IMG_H // patch, IMG_W // patch
        hb, wb, cb = c.bottleneck_shape(patch)
        sym = torch.empty((n_img, gh * gw, hb, wb, cb), dtype=torch.uint8, device=dev)
        rec = torch.empty((n_img, IMG_H, IMG_W, 3), dtype=torch.uint8, device=dev)

        def step():
            c.encode_images(imgs, patch, out=sym)
            c.decode_images(sym, IMG_H, IMG_W, patch, out=rec)
        ms = cuda_time(torch, step)
        c.profile(True)
        step()
        torch.cuda.synchronize()
        rows = layer_rows(c, (("encoder", c.enc_layers, patch, 1, 1), ("decoder", c.dec_layers, hb, 1, 1)), n_img * gh * gw, patch, hb, 1)
        c.profile(False)
        if show_layers:
            print_rows(rows, f"config 3: {variant} @ {patch}")
        fl = sum(T.variants.model_flops_per_pixel(variant, patch)) * px
        recs.append({"config": 3, "workload": f"{variant} encode+decode of {n_img} synthetic 2048x1536 images in {patch}x{patch} patches "
                                              f"({n_img * gh * gw} patches), random-init weights, device-resident",
                     "metric": METRIC, "value": px / ms / 1e3, "unit": UNIT, "ms_per_step": ms,
                     "algorithmic_tflops": fl / ms / 1e9, "roofline": roofline_of(rows, compute, variant, patch)})
        c.close()

    # ---- config 4: reduced_btn_32 + dataset-wide table + range coder ----------------------------------------------
    variant, patch = "base_model/reduced_btn_32", 128
    c = codec_for(variant)
    gh, gw = IMG_H // patch, IMG_W // patch
    hb, wb, cb = c.bottleneck_shape(patch)
    sym = torch.empty((n_img, gh * gw, hb, wb, cb), dtype=torch.uint8, device=dev)

    def enc_step():
        c.hist_reset()
        c.encode_images(imgs, patch, out=sym)
    ms_enc = cuda_time(torch, enc_step)
    counts = c.hist_read()
    table_ok = bool(np.array_equal(counts.astype(np.int64), torch.bincount(sym.reshape(-1).to(torch.int64), minlength=2).cpu().numpy()))
    cum = entry.cum_freq_table(counts / counts.sum(), 4096)
    ms_gpu_coder = cuda_time(torch, lambda: c.entropy_encode(sym, cum))
    packed, nbytes = c.entropy_encode(sym, cum)
    c.check_status()
    sym_h = sym.cpu().numpy().reshape(n_img, -1)
    t0 = time.perf_counter()
    blobs = range_coder.encode_streams(sym_h, cum)
    ms_host_coder = (time.perf_counter() - t0) * 1e3
    nb = nbytes.cpu().numpy()
    ph = packed.cpu().numpy()
    identical = all(bytes(ph[i, :int(nb[i])]) == blobs[i] for i in range(n_img))
    c.profile(True)
    enc_step()
    torch.cuda.synchronize()
    rows = layer_rows(c, (("encoder", c.enc_layers, patch, 1, 1),), n_img * gh * gw, patch, hb, 1)
    c.profile(False)
    if show_layers:
        print_rows(rows, f"config 4: {variant} @ {patch} (encoder)")
    recs.append({"config": 4, "workload": f"{variant} (bottleneck_channel 32) encode of {n_img} synthetic 2048x1536 images "
                                          f"({n_img * gh * gw} patches, {sym.numel()} symbols) + fused histogram -> table -> range coder",
                 "metric": "encode Mpixel/s", "value": px / ms_enc / 1e3, "unit": UNIT, "ms_per_step": ms_enc,
                 "histogram_equals_bincount": table_ok, "cum_freq": [int(v) for v in cum],
                 "range_coder": {"symbols": int(sym.numel()), "gpu_entropy_stage_ms": ms_gpu_coder, "host_thread_pool_ms": ms_host_coder,
                                 "host_threads": os.cpu_count(), "bytes": int(nb.sum()), "bpp": 8.0 * float(nb.sum()) / px,
                                 "gpu_bytes_identical_to_host": bool(identical)},
                 "roofline": roofline_of(rows, compute, variant, patch)})
    c.close()

    # ---- config 5: model_1 decode at P = 256 + rmbe post-filter ----------------------------------------------------
    variant, patch = "model_1", 256
    c = codec_for(variant)
    c.set_postfilter()
    gh, gw = IMG_H // patch, IMG_W // patch
    hb, wb, cb = c.bottleneck_shape(patch)
    sym = torch.randint(0, 2, (n_img, gh * gw, hb, wb, cb), dtype=torch.uint8, device=dev)
    rec = torch.empty((n_img, IMG_H, IMG_W, 3), dtype=torch.float32, device=dev)
    out8 = torch.empty((n_img, IMG_H, IMG_W, 3), dtype=torch.uint8, device=dev)

    def dec_step():  # submit/2/decoder.py:183-198: decode -> stitch -> rmbe -> np.around -> uint8
        c.decode_images(sym, IMG_H, IMG_W, patch, out=rec)
        c.postfilter_images(rec)
        c.round_u8(rec, out=out8)
    ms = cuda_time(torch, dec_step)
    ms_dec = cuda_time(torch, lambda: c.decode_images(sym, IMG_H, IMG_W, patch, out=rec))
    c.profile(True)
    dec_step()
    torch.cuda.synchronize()
    tiles = (IMG_H // 128) * ((IMG_W - 64) // 128) + ((IMG_H - 64) // 128) * (IMG_W // 128)
    rows = layer_rows(c, (("decoder", c.dec_layers, hb, 1, 4),), n_img * gh * gw, patch, hb, 1)
    rows += layer_rows(c, (("postfilter", c.post_layers, 128, 4, 4),), n_img * tiles, 128, 0, 1)
    c.profile(False)
    if show_layers:
        print_rows(rows, "config 5: model_1 decode @ 256 + rmbe (356 tiles per image)")
    recs.append({"config": 5, "workload": f"model_1 decode of {n_img} synthetic 2048x1536 images at P = 256 ({n_img * gh * gw} patches) -> "
                                          f"stitch -> rmbe post-filter ({tiles} tiles of 128 per image, two passes) -> uint8",
                 "metric": "decode+postfilter Mpixel/s", "value": px / ms / 1e3, "unit": UNIT, "ms_per_step": ms,
                 "decode_only_ms": ms_dec, "roofline": roofline_of(rows, compute, variant, patch)})
    c.close()
    return recs


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import tf_image_compression_b200 as T
    from tf_image_compression_b200 import entry, parallel, range_coder
    from tf_image_compression_b200 import variants as V

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    try:  # bind this rank to the CPUs (NUMA node) next to its GPU before any pinned host buffer is touched
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # random-init weights of the named architecture (the reference ships no checkpoints)
    codec = T.Codec(VARIANT, quan_scale=2, mean=MEAN, std=STD, device=local, compute="fp32", seed=1234)
    compute = args.compute
    if compute == "auto":
        compute = os.environ.get("TIC_DEFAULT_COMPUTE", "tensor")  # fp16-pair tcgen05 path (parity-green)
    codec.set_compute(compute)
    codec.use_torch_stream()
    if args.chunk > 0:
        codec.set_chunk_patches(args.chunk)
    B = args.images
    gh, gw = IMG_H // P, IMG_W // P
    hb, wb, cb = codec.bottleneck_shape(P)
    pixels = B * IMG_H * IMG_W

    # every rank holds the SAME 64 images (seed 1234): the weak arm runs all of them on every rank, the strong arm
    # image range [lo, hi) of them on this rank
    g = torch.Generator(device="cpu").manual_seed(1234)
    host_img = torch.randint(0, 256, (B, IMG_H, IMG_W, 3), dtype=torch.uint8, generator=g).pin_memory()
    host_sym = torch.empty((B, gh * gw, hb, wb, cb), dtype=torch.uint8).pin_memory()
    host_rec = torch.empty((B, IMG_H, IMG_W, 3), dtype=torch.uint8).pin_memory()
    d_img = host_img.to(dev, non_blocking=True)
    d_sym = torch.empty((B, gh * gw, hb, wb, cb), dtype=torch.uint8, device=dev)
    d_rec = torch.empty((B, IMG_H, IMG_W, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step_device():
        codec.encode_images(d_img, P, out=d_sym)
        codec.decode_images(d_sym, IMG_H, IMG_W, P, out=d_rec)

    def step_host():
        codec.encode_images(host_img, P, out=host_sym)
        codec.decode_images(host_sym, IMG_H, IMG_W, P, out=host_rec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing, weak arm ---------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = codec.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    if sampler.ok:
        sampler.sample()  # the queue is still draining: one sample under load even if the thread has not woken up yet
    barrier()
    clocks = sampler.result()
    launches = codec.launch_count - l0
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    value = world * pixels * args.steps / (ms_max * 1e-3) / 1e6

    # ---- strong arm: the 64 images sharded over the ranks + the table all-reduce ---------------------
    strong = None
    if world > 1:
        lo, hi = parallel.shard_range(B, rank, world)
        s_img, s_sym, s_rec = d_img[lo:hi], d_sym[lo:hi], d_rec[lo:hi]
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        coll_ms = []

        def step_strong(timed=False):
            codec.hist_reset()
            codec.encode_images(s_img, P, out=s_sym)
            if timed:
                ce0.record()
            counts = parallel.allreduce_histogram(codec)  # NCCL, in place on the device histogram (syncs to read it)
            if timed:
                ce1.record()
                torch.cuda.synchronize()
                coll_ms.append(ce0.elapsed_time(ce1))
            codec.decode_images(s_sym, IMG_H, IMG_W, P, out=s_rec)
            return counts
        for _ in range(args.warmup):
            step_strong()
        barrier()
        t0 = time.perf_counter()
        se0, se1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        se0.record()
        for _ in range(args.steps):
            table = step_strong(timed=True)
        se1.record()
        barrier()
        s_ms = max_over_ranks(se0.elapsed_time(se1))
        # the single-rank table of the same 64 images (outside the timed region)
        codec.hist_reset()
        codec.encode_images(d_img, P, out=d_sym)
        single = codec.hist_read()
        strong = {"value": pixels * args.steps / (s_ms * 1e-3) / 1e6, "unit": UNIT, "scaling": "strong",
                  "images_total": B, "images_per_gpu": hi - lo, "patches_per_gpu": (hi - lo) * gh * gw, "ms_per_step": s_ms / args.steps,
                  "collective": "ncclAllReduce(sum) of the uint64[256] symbol histogram, once per step (get_encoded_distribution.py:113-134)",
                  "collective_ms": max_over_ranks(float(np.median(coll_ms))),
                  "table_equals_single_rank": bool(np.array_equal(table, single)), "table": [int(v) for v in table]}

    # ---- end to end through the public API with pinned host buffers ------------------------------
    # (a) the two reference-facing calls back to back (encode.py flow, then decode.py flow): every byte crosses
    #     PCIe in one direction at a time;  (b) the one-graph round trip (test.py:95-146, Codec.roundtrip_images):
    #     same results, images H2D while reconstructions + symbols D2H.  (b) is the e2e headline, (a) sits next to it.
    def time_host(fn, scale=1.0):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return scale * world * pixels * args.steps / dt / 1e6

    def step_roundtrip():
        codec.roundtrip_images(host_img, P, out=host_rec, out_symbols=host_sym)

    e2e_separate = time_host(step_host)
    e2e_value = time_host(step_roundtrip)
    e2e_encode = time_host(lambda: codec.encode_images(host_img, P, out=host_sym))
    e2e_decode = time_host(lambda: codec.decode_images(host_sym, IMG_H, IMG_W, P, out=host_rec))
    h2d = host_img.numel()
    d2h = host_sym.numel() + host_rec.numel()

    # ---- entropy coding of the step's symbols (excluded from the metric, printed) ----------------------
    codec.hist_reset()
    codec.encode_images(d_img, P, out=d_sym)
    counts = codec.hist_read()
    cum = entry.cum_freq_table(counts / counts.sum(), 4096)
    coder_gpu_ms = cuda_time(torch, lambda: codec.entropy_encode(d_sym, cum))
    packed, nbytes = codec.entropy_encode(d_sym, cum)
    codec.check_status()
    coder_gpu_dec_ms = cuda_time(torch, lambda: codec.entropy_decode(packed, nbytes, d_sym[0].numel(), cum))
    range_coder_rec = {"symbols_per_step": int(d_sym.numel()), "streams": B, "table": [int(v) for v in cum],
                       "gpu_entropy_stage_encode_ms": coder_gpu_ms, "gpu_entropy_stage_decode_ms": coder_gpu_dec_ms,
                       "bytes": int(nbytes.sum().item()), "bpp": 8.0 * float(nbytes.sum().item()) / pixels}
    if rank == 0:
        sym_h = d_sym.cpu().numpy().reshape(B, -1)
        t0 = time.perf_counter()
        blobs = range_coder.encode_streams(sym_h, cum)
        range_coder_rec["host_thread_pool_encode_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        range_coder.encode_streams(sym_h[:4], cum, threads=1)
        range_coder_rec["host_single_thread_encode_ms_extrapolated"] = (time.perf_counter() - t0) * 1e3 * B / 4
        t0 = time.perf_counter()
        range_coder.decode_streams(blobs, [sym_h.shape[1]] * B, cum)
        range_coder_rec["host_thread_pool_decode_ms"] = (time.perf_counter() - t0) * 1e3
        range_coder_rec["host_threads"] = os.cpu_count()
        nb = nbytes.cpu().numpy()
        ph = packed[:4].cpu().numpy()
        range_coder_rec["gpu_bytes_identical_to_host"] = bool(all(bytes(ph[i, :int(nb[i])]) == blobs[i] for i in range(4)))

    # encode flow with the GPU entropy stage: images H2D, only the compressed streams D2H
    host_packed = torch.empty((B, codec.entropy_bound(d_sym[0].numel())), dtype=torch.uint8).pin_memory()

    def step_encode_entropy():
        d_in = host_img.to(dev, non_blocking=True)
        sym = codec.encode_images(d_in, P, out=d_sym)
        pk, nb = codec.entropy_encode(sym, cum)
        nmax = int(nb.max().item())
        host_packed[:, :nmax].copy_(pk[:, :nmax], non_blocking=True)
        torch.cuda.synchronize()
    e2e_encode_entropy = time_host(step_encode_entropy)

    # ---- per-layer times (separate profiled pass, not part of the timed region) -------------------
    PASSES = 3
    codec.profile(True)
    for _ in range(PASSES):
        step_device()
    torch.cuda.synchronize()
    npatch = B * gh * gw
    rows = layer_rows(codec, (("encoder", codec.enc_layers, P, 1, 1), ("decoder", codec.dec_layers, hb, 1, 1)), npatch, P, hb, PASSES)
    for r in rows:
        r["npatch"] = npatch
    codec.profile(False)
    roofline = roofline_of(rows, compute, VARIANT, P)
    roofline.update({
        "whole_step_algorithmic_tflops": world * sum(V.model_flops_per_pixel(VARIANT, P)) * pixels * args.steps / (ms_max * 1e-3) / 1e12,
        "whole_step_layer_bytes_gbs": world * sum(r["bytes_step"] for r in rows) * args.steps / (ms_max * 1e-3) / 1e9,
    })
    if args.layers and rank == 0:
        print_rows(rows, f"config 2: {VARIANT} @ {P}")

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores = cpu_roundtrip_time(args.cpu_images, reps=3)
        cpu = {"value": args.cpu_images * IMG_H * IMG_W / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_images} of the {B} 2048x1536 images (192 patches each), best of 3; torch-CPU fp32 "
                         "oracle restatement (TensorFlow not installable)"}

    parity = None
    if cpu is not None and VARIANT == "model_0":
        parity = parity_sample(compute)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs and VARIANT == "model_0":
        codec.close()
        codec = None
        del d_img, d_sym, d_rec
        torch.cuda.empty_cache()
        configs = config_records(args, compute, dev, args.layers)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tensor": "f16x3", "f16x3": "f16x3", "3xtf32": "3xtf32", "tf32": "tf32"}.get(compute, compute),
            "data": "synthetic",
            "config": {"workload": f"{VARIANT} encode+decode of {B} synthetic 2048x1536 RGB images per GPU in {P}x{P} "
                                   f"patches ({B * gh * gw} patches, BASELINE config 2), random-init weights",
                       "patch_size": P, "images_per_gpu": B, "compute": compute,
                       "l2": "inputs (604 MB per step) exceed the 126 MB L2"},
            "clocks": clocks,
            "strong": strong if strong is not None else {"value": value, "unit": UNIT, "scaling": "strong", "images_total": B,
                                                          "images_per_gpu": B, "ms_per_step": ms_max / args.steps,
                                                          "collective": None, "note": "N = 1: identical to the weak arm"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": "Codec.roundtrip_images (tic_roundtrip_images; test.py:95-146), pinned host images in, host symbols + "
                            "uint8 reconstructions out",
                    "separate_calls": {"value": e2e_separate, "unit": UNIT,
                                       "h2d_bytes_per_step": int(host_img.numel() + host_sym.numel()),
                                       "d2h_bytes_per_step": int(d2h),
                                       "call": "Codec.encode_images then Codec.decode_images (encode.py / decode.py flows)"},
                    "encode_only": {"value": e2e_encode, "unit": UNIT, "h2d_bytes_per_step": int(host_img.numel()),
                                    "d2h_bytes_per_step": int(host_sym.numel()), "call": "Codec.encode_images (encode.py flow)"},
                    "decode_only": {"value": e2e_decode, "unit": UNIT, "h2d_bytes_per_step": int(host_sym.numel()),
                                    "d2h_bytes_per_step": int(host_rec.numel()), "call": "Codec.decode_images (decode.py flow)"},
                    "encode_with_gpu_entropy_stage": {"value": e2e_encode_entropy, "unit": UNIT, "h2d_bytes_per_step": int(host_img.numel()),
                                                      "d2h_bytes_per_step": int(range_coder_rec["bytes"]),
                                                      "call": "images H2D -> Codec.encode_images -> Codec.entropy_encode -> compressed "
                                                              "streams D2H (encode.py:153-202 end to end on the device)"}},
            "gpu_launches": int(launches) * world,
            "roofline": roofline,
            "range_coder": range_coder_rec,
            "cpu_baseline": cpu,
            "parity": parity,
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if codec is not None:
        codec.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    global VARIANT, P
    args = parse()
    VARIANT, P = args.variant, args.patch
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
