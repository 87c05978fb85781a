"""CPU tests that pin the oracle (oracle/codec_oracle.py) — the reference has no golden tensor for the
conv path, so the restatement is validated by independent restatements and
algebraic identities."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import codec_oracle as O


def _rand(shape, seed, scale=1.0):
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


def c_conv(x, k, b, stride, relu):
    n, h, w, cin = x.shape
    cout = k.shape[3]
    ho, wo = -(-h // stride), -(-w // stride)
    out = np.empty((n, ho, wo, cout), np.float32)
    O.clib().tico_conv2d_same(O._fp(x), n, h, w, cin, O._fp(k), O._fp(b), cout, stride, int(relu), O._fp(out))
    return out


def c_deconv(x, k, b, relu):
    n, h, w, cin = x.shape
    cout = k.shape[2]
    out = np.empty((n, 2 * h, 2 * w, cout), np.float32)
    O.clib().tico_deconv2d(O._fp(x), n, h, w, cin, O._fp(k), O._fp(b), cout, int(relu), O._fp(out))
    return out


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("hw", [(8, 8), (7, 10), (16, 12)])
@pytest.mark.parametrize("cin,cout", [(3, 8), (16, 5)])
def test_conv_torch_matches_direct_c(stride, hw, cin, cout):
    x = _rand((2, hw[0], hw[1], cin), 1)
    k = _rand((3, 3, cin, cout), 2, 0.2)
    b = _rand((cout,), 3, 0.1)
    for act in ("relu", "id"):
        a = O.conv2d_same(x, k, b, stride, act)
        c = c_conv(x, k, b, stride, act == "relu")
        assert a.shape == c.shape
        np.testing.assert_allclose(a, c, rtol=0, atol=2e-5)


def test_conv_same_padding_is_asymmetric_for_stride2():
    # even input, stride 2: TF pads 0 before / 1 after
    x = np.zeros((1, 4, 4, 1), np.float32)
    x[0, 0, 0, 0] = 1.0
    k = np.zeros((3, 3, 1, 1), np.float32)
    k[0, 0, 0, 0] = 1.0  # tap (0,0) reads in[2*o + 0 - pad_before]
    y = O.conv2d_same(x, k, np.zeros(1, np.float32), 2, "id")
    assert y[0, 0, 0, 0] == 1.0  # pad_before == 0
    assert O.same_pad(4, 2) == (2, 0, 1)
    assert O.same_pad(4, 1) == (4, 1, 1)
    assert O.same_pad(5, 2) == (3, 1, 1)


@pytest.mark.parametrize("hw", [(4, 4), (5, 3)])
@pytest.mark.parametrize("cin,cout", [(8, 3), (4, 6)])
def test_deconv_torch_matches_direct_c(hw, cin, cout):
    x = _rand((2, hw[0], hw[1], cin), 4)
    k = _rand((3, 3, cout, cin), 5, 0.2)
    b = _rand((cout,), 6, 0.1)
    for act in ("relu", "id"):
        a = O.deconv2d(x, k, b, act)
        c = c_deconv(x, k, b, act == "relu")
        np.testing.assert_allclose(a, c, rtol=0, atol=2e-5)


def test_deconv_is_gradient_of_stride2_same_conv():
    """tf.nn.conv2d_transpose == gradient of conv2d wrt its input; the forward conv maps [2h,2w,cout]
    -> [h,w,cin] with the SAME filter [3,3,cout,cin] read as HWIO (in=cout_deconv, out=cin_deconv)."""
    h, w, cin, cout = 5, 6, 4, 3
    x = _rand((1, h, w, cin), 7)
    k = _rand((3, 3, cout, cin), 8, 0.3)
    big = torch.zeros(1, 2 * h, 2 * w, cout, dtype=torch.float64, requires_grad=True)
    bt = big.permute(0, 3, 1, 2)
    bt = F.pad(bt, (0, 1, 0, 1))  # SAME stride 2 on even size: 0 before, 1 after
    wt = torch.from_numpy(k).double().permute(3, 2, 0, 1)  # HWIO -> OIHW with I = cout_deconv
    y = F.conv2d(bt, wt, None, stride=2)  # [1, cin, h, w]
    y.backward(torch.from_numpy(x).double().permute(0, 3, 1, 2))
    grad = big.grad.numpy()
    ours = O.deconv2d(x, k, np.zeros(cout, np.float32), "id", dtype=torch.float64)
    np.testing.assert_allclose(ours, grad, rtol=0, atol=1e-12)


def test_deconv_padding1_idiom_is_wrong():
    # the common padding=1, output_padding=1 idiom is off by one pixel vs TF SAME
    x = _rand((1, 4, 4, 2), 9)
    k = _rand((3, 3, 2, 2), 10)
    ours = O.deconv2d(x, k, np.zeros(2, np.float32), "id")
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    wt = torch.from_numpy(k).permute(3, 2, 0, 1)
    other = F.conv_transpose2d(xt, wt, None, stride=2, padding=1, output_padding=1).permute(0, 2, 3, 1).numpy()
    assert np.abs(ours - other).max() > 1e-3


def test_einsum_restatement_of_conv():
    x = _rand((1, 6, 6, 3), 11)
    k = _rand((3, 3, 3, 4), 12)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    out = np.zeros((1, 6, 6, 4))
    for kh in range(3):
        for kw in range(3):
            out += np.einsum("nhwc,co->nhwo", xp[:, kh:kh + 6, kw:kw + 6, :].astype(np.float64), k[kh, kw].astype(np.float64))
    a = O.conv2d_same(x, k, np.zeros(4, np.float32), 1, "id", dtype=torch.float64)
    np.testing.assert_allclose(a, out, rtol=0, atol=1e-12)


def test_quantiser_expression_and_half_even():
    x = _rand((4096,), 13, 3.0)
    for q in (2, 4, 256):
        sym = O.quantize(x, q)
        lit = O.quantize_reference_expr(x, q)  # (round(o) - o) + o
        assert np.array_equal(lit, sym.astype(np.float32))
        assert sym.min() >= 0 and sym.max() <= q - 1
    # q = 2: sigmoid == 0.5 exactly rounds to 0 (half-to-even), the fp32 dead zone above zero
    dead = np.array([0.0, 1e-8, 5e-8], np.float32)
    assert np.all(O.sigmoid_f32(dead) == np.float32(0.5))
    assert np.array_equal(O.quantize(dead, 2), [0, 0, 0])
    assert O.quantize(np.array([2e-7], np.float32), 2)[0] == 1
    # q = 3: o = 1.0 at x = 0 -> 1; o = 0.5 boundary is sigmoid = 0.25
    assert O.quantize(np.array([0.0], np.float32), 3)[0] == 1


def test_sigmoid_accuracy():
    xs = np.linspace(-30, 30, 200001).astype(np.float32)
    ref = 1.0 / (1.0 + np.exp(-xs.astype(np.float64)))
    got = O.sigmoid_f32(xs).astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) < 6e-7
    assert np.all(np.diff(got) >= 0)  # monotone: the symbol decision has a single threshold


def test_inverse_sigmoid_lut_values():
    lut = O.inverse_sigmoid_lut(2)
    # fp32 values, note the cancellation in 1 - p
    assert abs(lut[0] - (-13.815519)) < 1e-5 and abs(lut[1] - 11.611643) < 1e-5
    lut256 = O.inverse_sigmoid_lut(256)
    assert np.all(np.diff(lut256) > 0) and np.all(np.isfinite(lut256))


def test_crop_concat_roundtrip_and_reflect():
    img = O.synthetic_image(200, 300, 0)
    patches = O.crop_image_input_patches(img, 128)
    assert len(patches) == 2 * 3 and patches[0].shape == (128, 128, 3)
    back = O.concat_patches(patches, 200, 300, 128)
    assert np.array_equal(back, img)
    # reflect without edge repeat: padded row 200 == row 198
    assert np.array_equal(patches[3][200 - 128, :, :], img[198, 0:128, :])
    # exact multiples are not padded
    assert len(O.crop_image_input_patches(img[:128, :256], 128)) == 2


def test_rmbe_tiling_counts_and_inplace_order():
    calls = []

    def fake_model(tiles):
        calls.append(tiles.shape[0])
        return tiles + 1.0

    img = np.zeros((384, 320, 3), np.float32)
    out = O.rmbe(img, fake_model)
    assert calls == [3 * 2, 2 * 2]  # (H//128)*((W-64)//128), ((H-64)//128)*(W//128)
    assert out[0, 0, 0] == 0 and out[0, 64, 0] == 1  # pass-1 only
    assert out[64, 64, 0] == 2  # both passes
    assert out[64, 0, 0] == 1 and out[383, 319, 0] == 1 and out[383, 0, 0] == 0
    # 2048x1536: 356 tiles per image
    calls.clear()
    O.rmbe(np.zeros((1536, 2048, 3), np.float32), lambda t: (calls.append(t.shape[0]), t)[1])
    assert sum(calls) == 356


def test_fp32_graph_close_to_fp64():
    v = "model_0"
    enc = O.init_params(O.VARIANTS[v]["enc"], 3, 1, "fanin")
    mean = np.array([120.0, 115.0, 100.0], np.float32)
    std = np.array([60.0, 58.0, 62.0], np.float32)
    patches = np.stack(O.crop_image_input_patches(O.synthetic_image(128, 256, 3), 128)).astype(np.float32)
    l32 = O.encoder_logits(patches, v, enc, mean, std)
    l64 = O.encoder_logits(patches, v, enc, mean, std, dtype=torch.float64)
    assert l32.shape == (2, 8, 8, 64)
    assert np.abs(l32 - l64).max() / np.abs(l64).max() < 1e-5


def test_flops_and_shapes_of_every_variant():
    expect = {"model_0": (128, (8, 8, 64)), "model_1": (256, (16, 16, 64)), "base_model/input_256": (256, (32, 32, 64)),
              "base_model/ch_128": (128, (32, 32, 64)), "base_model/reduced_btn_32": (128, (32, 32, 32)),
              "model_3": (128, (8, 8, 80)), "model_2": (128, (8, 8, 64))}
    for v, (P, shp) in expect.items():
        h = P
        c = 3
        for l in O.expand_layers(O.VARIANTS[v]["enc"], 3):
            h = -(-h // l["stride"]) if l["kind"] == "c" else 2 * h
            c = l["cout"]
        assert (h, h, c) == shp, v
        hd = h
        for l in O.expand_layers(O.VARIANTS[v]["dec"], c):
            hd = -(-hd // l["stride"]) if l["kind"] == "c" else 2 * hd
        assert hd == P, v


def test_histogram_position_mean_and_metrics():
    sym = np.random.RandomState(5).randint(0, 2, size=(10, 8, 8, 4)).astype(np.uint8)
    freq = O.symbol_histogram(sym, 2)
    assert freq.sum() == sym.size and freq[1] == sym.sum()
    mean, prob, order = O.position_mean([sym[:4], sym[4:]])
    np.testing.assert_allclose(mean, sym.reshape(10, -1).mean(0), atol=1e-7)  # float32 batch terms (cal_encoded_distribution.py:128)
    # [1 - p, p] of the mean over positions (:144-145) and the stable position order (:149)
    assert prob.shape == (2,) and np.isclose(prob.sum(), 1) and np.isclose(prob[1], sym.mean(), atol=1e-7)
    assert sorted(order) == list(range(256)) and all(mean[order[i]] <= mean[order[i + 1]] for i in range(255))
    ties = [k for k in range(255) if mean[order[k]] == mean[order[k + 1]]]
    assert ties and all(order[k] < order[k + 1] for k in ties)  # stable: ties keep position order
    a = np.zeros((4, 4, 3), np.uint8)
    b = np.full((4, 4, 3), 1, np.uint8)
    assert abs(O.psnr([(a, b)]) - 20 * np.log10(255.0)) < 1e-9
    assert O.bpp(100, 400) == 2.0
    m, s = O.online_mean_and_std_channel([np.full((2, 2, 3), 10, np.uint8), np.full((2, 2, 3), 20, np.uint8)])
    np.testing.assert_allclose(m, [15, 15, 15])
    np.testing.assert_allclose(s, [5, 5, 5])
