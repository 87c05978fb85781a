"""CPU tests of the host-side product code: layer tables, patch helpers, C-ABI surface."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import tf_image_compression_b200 as T
from tf_image_compression_b200 import _lib as L
from tf_image_compression_b200 import variants as V
from oracle import codec_oracle as O

ROOT = Path(__file__).resolve().parents[1]


def test_variant_tables_match_the_oracle_restatement():
    """Two independent transcriptions of the reference model files must agree."""
    for name, ov in O.VARIANTS.items():
        pv = V.VARIANTS[name]
        assert pv["patch_size"] == ov["patch_size"] and pv["bottleneck_channel"] == ov["bottleneck"], name
        for table, okey, cin in (("encoder", "enc", 3), ("decoder", "dec", ov["bottleneck"])):
            prod = V.primitive_layers(pv[table], cin, pv["bottleneck_channel"])
            orac = O.expand_layers(ov[okey], cin)
            assert len(prod) == len(orac), (name, table)
            for p, o in zip(prod, orac):
                assert (p.kind, p.scope, p.cin, p.cout, p.stride, p.act, p.res_begin, p.res_end) == (
                    o["kind"], o["scope"], o["cin"], o["cout"], o["stride"], o["act"], o["res_begin"], o["res_end"]), (name, p.scope)
    for name, ol in O.POSTFILTERS.items():
        prod = V.postfilter_layers(name)
        orac = O.expand_layers(ol, 3)
        assert [(p.kind, p.scope, p.cout, p.stride, p.act) for p in prod] == [
            (o["kind"], o["scope"], o["cout"], o["stride"], o["act"]) for o in orac]


def test_flops_per_pixel_match_baseline_md():
    expect = {("model_0", 128): (3888, 3888), ("model_1", 256): (3096, 3096), ("base_model/input_256", 256): (9648, 7920),
              ("base_model/ch_128", 128): (93024, 93024), ("base_model/ch_128", 256): (93024, 93024),
              ("base_model/reduced_btn_32", 128): (23472, 23472), ("model_3", 128): (29592, 29592)}
    for (v, p), (e, d) in expect.items():
        assert V.model_flops_per_pixel(v, p) == (e, d)
    assert V.flops_per_pixel(V.postfilter_layers(), 128) / 128 ** 2 == 14688


def test_aliases_and_unknown_variant():
    assert V.resolve("submit/2") == "model_3" and V.resolve("base_model/fin/") == "model_3"
    with pytest.raises(ValueError):
        V.resolve("model_9")
    with pytest.raises(ValueError):
        V.primitive_layers(V.VARIANTS["base_model/reduced_btn_32"]["encoder"], 3, None)
    assert V.encoder_layers("base_model/reduced_btn_32", 16)[-1].cout == 16


def test_crop_and_concat_match_oracle():
    for (h, w, p) in [(200, 300, 128), (128, 256, 128), (130, 129, 64), (512, 768, 128)]:
        img = O.synthetic_image(h, w, h + w)
        a = T.utils.crop_image_input_patches(img, p)
        b = O.crop_image_input_patches(img, p)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(T.utils.concat_patches(a, h, w, p), img)
        assert np.array_equal(T.utils.concat_patches(a, h, w, p), O.concat_patches(b, h, w, p))
    with pytest.raises(ValueError):
        T.utils.concat_patches(a[:-1], 512, 768, 128)


def test_inverse_sigmoid_lut_matches_oracle():
    for q in (2, 3, 16, 256):
        assert np.array_equal(T.inverse_sigmoid_lut(q), O.inverse_sigmoid_lut(q))


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "tic.h").read_text()
    declared = set(re.findall(r"\b(tic_[a-z0-9_]+)\s*\(", header))
    declared -= {"tic_codec"}
    assert declared, "no declarations parsed"
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))
    lib = ctypes.CDLL(str(L.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in L.load().tic_version()


def test_shipped_library_has_no_ablation_knobs():
    """The stage-ablation / tuning environment knobs (TIC_DBG, TIC_FUSE_DEC, TIC_FUSE_ENC, TIC_MAX_NBUF, TIC_FIRST_NO_TMA ...)
    exist only in builds with -DTIC_ABLATE (tools/build_lib.sh with TIC_EXTRA_FLAGS): the shipped library does not know their names,
    so no such variable can change its results."""
    blob = Path(L.LIB_PATH).read_bytes()   # (getenv itself is linked in by the static CUDA runtime)
    for knob in (b"TIC_DBG", b"TIC_FUSE_DEC", b"TIC_FUSE_ENC", b"TIC_MAX_NBUF", b"TIC_FIRST_NO_TMA", b"TIC_L2_BUDGET_MB"):
        assert knob not in blob, knob


def test_no_gpu_fails_loudly_without_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TicError, match="no CPU fallback"):
        T.Codec("model_0")


def test_product_never_imports_the_oracle():
    for p in (ROOT / "tf_image_compression_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), p
    for p in (ROOT / "tf_image_compression_b200" / "csrc").glob("*"):
        assert "oracle/" not in p.read_text(), p


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    no GPU needed."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encode+decode Mpixel/s" and d["unit"] == "Mpixel/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
