#!/usr/bin/env python
"""Benchmark of the codec hot path (encode + decode) on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  metric   : encode+decode Mpixel/s (BASELINE.json), image pixels H*W counted once per round trip
  workload : BASELINE config 2 — model_0, 64 synthetic 2048x1536 RGB images in 128x128 patches per
             GPU (weak scaling: every rank encodes + decodes its own 64 images; no data-path collective)
  value    : device-resident throughput (inputs already in HBM, CUDA events, max over ranks)
  e2e      : same metric through the public API with pinned HOST buffers (H2D + kernels + D2H)
  roofline : dominant kernel, algorithmic bytes (or FLOPs) per launch / CUDA-event time vs MEASURED_PEAKS.json
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
    ap.add_argument("--images", type=int, default=64, help="images per GPU per step (BASELINE config 2: 64)")
    ap.add_argument("--compute", default=os.environ.get("TIC_COMPUTE", "auto"), help="fp32 | tensor | tf32 | auto")
    ap.add_argument("--cpu-images", type=int, default=2, help="images in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    s_ref, r_ref = O.codec_roundtrip(image, "model_0", enc, dec, MEAN, STD, 2, 128)
    cum = entry.cum_freq_table(counts / counts.sum(), 4096)
    sizes = []
    with tempfile.TemporaryDirectory() as d:
        for k, stream in enumerate((sym[0].reshape(-1), s_ref.reshape(-1).astype(np.uint8))):
            e = range_coder.RangeEncoder(os.path.join(d, f"{k}.bin"))
            e.encode(stream, cum)
            e.close()
            sizes.append(os.path.getsize(os.path.join(d, f"{k}.bin")))
    px = image.shape[0] * image.shape[1]
    diff = np.abs(rec[0].astype(int) - r_ref.astype(int))
    return {"sample": "one 256x384 synthetic image, model_0, fan-in weights, 6 patches",
            "symbols": int(s_ref.size), "symbol_mismatches": int((sym[0].reshape(s_ref.shape) != s_ref).sum()),
            "recon_max_abs_diff_u8": int(diff.max()), "recon_pixels_differing": int((diff != 0).sum()),
            "bpp": 8.0 * sizes[0] / px, "bpp_delta": 8.0 * (sizes[0] - sizes[1]) / px,
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
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import tf_image_compression_b200 as T
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

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
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

    # ---- device-resident timing -------------------------------------------------------------------
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
    barrier()
    clocks = sampler.result()
    launches = codec.launch_count - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * pixels * args.steps / (ms_max * 1e-3) / 1e6

    # ---- end to end through the public API with pinned host buffers ------------------------------
    # (a) the two reference-facing calls back to back (encode.py flow, then decode.py flow): every byte crosses
    #     PCIe in one direction at a time;  (b) the one-graph round trip (test.py:95-146, Codec.roundtrip_images):
    #     same results, images H2D while reconstructions + symbols D2H.  (b) is the e2e headline, (a) sits next to it.
    def time_host(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * pixels * args.steps / float(t.item()) / 1e6

    def step_roundtrip():
        codec.roundtrip_images(host_img, P, out=host_rec, out_symbols=host_sym)

    e2e_separate = time_host(step_host)
    e2e_value = time_host(step_roundtrip)
    h2d = host_img.numel()
    d2h = host_sym.numel() + host_rec.numel()

    # ---- per-layer times (separate profiled pass, not part of the timed region) -------------------
    # Every layer launch is bracketed by CUDA events on the codec's stream (tic_profile_*); the dominant
    # kernel's roofline is algorithmic work per launch / its average launch duration.
    PASSES = 3
    codec.profile(True)
    for _ in range(PASSES):
        step_device()
    torch.cuda.synchronize()
    pk = peaks()
    rows = []
    npatch = B * gh * gw
    for graph, layers in (("encoder", codec.enc_layers), ("decoder", codec.dec_layers)):
        hh = P if graph == "encoder" else hb
        first, last = layers[0], layers[-1]
        for (l, tot_ms, n_launch) in codec.layer_times(graph):
            hin = hh
            if l.kind == "c":
                hh = -(-hh // l.stride)
                macs = hh * hh * 9 * l.cin * l.cout
            else:
                macs = hh * hh * 9 * l.cin * l.cout
                hh *= 2
            # algorithmic bytes of this layer per patch: its input and output tensors once (activations are
            # 4 B per element: fp16 pair planes or fp32; the u8 image / symbols / u8 reconstruction are 1 B) + weights
            in_b = hin * hin * l.cin * (1 if l is first else 4)
            out_b = hh * hh * l.cout * (1 if l is last else 4)
            rows.append(dict(graph=graph, scope=l.scope, ms_step=tot_ms / PASSES, launches=max(1, n_launch // PASSES),
                             flops_step=2.0 * macs * npatch, bytes_step=float(in_b + out_b) * npatch + 9.0 * l.cin * l.cout * 4))
    codec.profile(False)
    step_ms = sum(r["ms_step"] for r in rows)
    top = max(rows, key=lambda r: r["ms_step"])
    ridge = pk["bf16_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    top_launch_s = top["ms_step"] / top["launches"] * 1e-3
    if top["flops_step"] / top["bytes_step"] >= ridge:
        achieved = top["flops_step"] / top["launches"] / top_launch_s / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16_sustained"], "peak_source": f"bf16 sustained, {pk['source']}"}
    else:
        achieved = top["bytes_step"] / top["launches"] / top_launch_s / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                    "peak_source": f"copy bandwidth, {pk['source']}"}
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists() and compute == "tensor":
        # DRAM bytes of this kernel from the committed ncu --set full capture, scaled to this run's patches per launch
        td = json.loads(tp.read_text())
        per = td.get(VARIANT, {}).get(f"{top['graph']}/{top['scope']}")
        if per is not None and P == 128:
            traffic = per / td["patches_per_launch"] * npatch / top["launches"]
    roofline.update({
        "traffic": traffic, "kernel": f"{top['graph']}/{top['scope']} ({compute})",
        "kernel_ms_per_launch": top_launch_s * 1e3, "kernel_share_of_step": top["ms_step"] / max(step_ms, 1e-9),
        "kernel_algorithmic_flop_per_byte": top["flops_step"] / top["bytes_step"], "ridge_flop_per_byte": ridge,
        "whole_step_algorithmic_tflops": world * sum(V.model_flops_per_pixel(VARIANT, P)) * pixels * args.steps / (ms_max * 1e-3) / 1e12,
        "whole_step_layer_bytes_gbs": world * sum(r["bytes_step"] for r in rows) * args.steps / (ms_max * 1e-3) / 1e9,
    })
    if args.layers and rank == 0:
        for r in rows:
            t = r["ms_step"] * 1e-3
            print(f"  {r['graph']:8s} {r['scope']:22s} {r['ms_step']:8.3f} ms/step {r['launches']:3d} launches  "
                  f"{r['flops_step'] / t / 1e12:8.2f} TFLOP/s {r['bytes_step'] / t / 1e9:8.1f} GB/s", file=sys.stderr)

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
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": "Codec.roundtrip_images (tic_roundtrip_images; test.py:95-146), pinned host images in, host symbols + "
                            "uint8 reconstructions out",
                    "separate_calls": {"value": e2e_separate, "unit": UNIT,
                                       "h2d_bytes_per_step": int(host_img.numel() + host_sym.numel()),
                                       "d2h_bytes_per_step": int(d2h),
                                       "call": "Codec.encode_images then Codec.decode_images (encode.py / decode.py flows)"}},
            "gpu_launches": int(launches) * world,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
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
