"""Pins oracle/unet_oracle.py (the CPU restatement) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

CLASSES = ["UNet", "UNetSP", "UNetSPSmall", "UNetDO", "UNet4_2IC", "recAE_v2_fixed"]


def _x(cin, size, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(1, cin, size, size, size, generator=g) > 0.7).float()


def _targets(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    sk = (torch.rand(batch, size, size, size, generator=g) > 0.6).long()
    fl = ((torch.rand(batch, size, size, size, generator=g) > 0.8) & (sk > 0)).long()
    oh = lambda t: torch.nn.functional.one_hot(t, 2).permute(0, 4, 1, 2, 3).float().contiguous()
    return oh(sk), oh(fl)


@pytest.mark.parametrize("name", CLASSES)
def test_same_seed_init_matches_reference(golden, name):
    g = golden["classes"][name]
    sd = O.build_state_dict(O.PRESETS[name], seed=0)
    assert list(sd.keys()) == g["keys"]
    assert [tuple(v.shape) for v in sd.values()] == g["shapes"]
    params = [v for k, v in sd.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    assert sum(p.numel() for p in params) == g["n_params"]
    assert abs(float(sum(p.double().abs().sum() for p in params)) - g["abs_sum"]) < 1e-9
    assert torch.equal(params[0].flatten()[:3], g["first3"])


@pytest.mark.parametrize("name", CLASSES)
def test_eval_forward_matches_reference(golden, name):
    g = golden["classes"][name]["eval32"]
    cfg = O.PRESETS[name]
    sd = O.build_state_dict(cfg, seed=0)
    x = _x(cfg.input_channels, 32, 1)
    assert float(x.sum()) == g["x_sum"]
    with torch.no_grad():
        out = O.unet_forward(sd, x, cfg, training=False)
    outs = out if isinstance(out, tuple) else (out,)
    for o, s, am, sl, hs in zip(outs, g["out_sums"], g["argmax_ones"], g["out_slices"], g["hard_segm_slice"]):
        assert torch.equal(o[:, :, 12:20, 12:20, 12:20], sl)          # bit-exact: same torch ops in the same order
        assert float(o.double().sum()) == s
        assert int(torch.argmax(o, 1).sum()) == am
        assert torch.equal(O.hard_segm_from_tensor(o)[:, 12:20, 12:20, 12:20], hs)


@pytest.mark.parametrize("name", ["UNetSP", "UNetDO", "UNetSPSmall", "UNet4_2IC", "recAE_v2_fixed"])
def test_train_step_matches_reference(golden, name):
    g = golden["train_step"][name]
    cfg = O.PRESETS[name]
    sd = O.build_state_dict(cfg, seed=0)
    pnames = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    for k in pnames:
        sd[k].requires_grad_()
    gen = torch.Generator().manual_seed(7)
    x = (torch.rand(g["batch"], cfg.input_channels, g["size"], g["size"], g["size"], generator=gen) > 0.7).float()
    x.requires_grad_()
    sk_t, fl_t = _targets(g["batch"], g["size"], 11)
    out = O.unet_forward(sd, x, cfg, training=True)
    if g["handler"] == "double":
        loss, comps = O.loss_double_output(out, (sk_t, fl_t), 1.0, 1.0)
    else:
        loss, comps = O.loss_single_output(out, sk_t, 1.0, 1.0)
    loss.backward()
    assert float(loss) == pytest.approx(g["loss"], rel=1e-6)
    for k, v in g["components"].items():
        assert float(comps[k]) == pytest.approx(v, rel=1e-6)
    none = [k for k in pnames if sd[k].grad is None]
    assert none == g["grad_none"]
    for k, v in g["grad_abs_sum"].items():
        assert float(sd[k].grad.double().abs().sum()) == pytest.approx(v, rel=2e-4, abs=1e-7), k
        assert torch.allclose(sd[k].grad.flatten()[:8], g["grad_head"][k], rtol=2e-3, atol=1e-6), k
    assert float(x.grad.double().abs().sum()) == pytest.approx(g["x_grad_abs_sum"], rel=2e-4)
    for k, v in g["bn_after"].items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        else:
            assert torch.allclose(sd[k], v, rtol=1e-5, atol=1e-7), k


def test_checkpoint_quirks_recorded(golden):
    g = golden["train_step"]["UNetSP"]
    assert set(g["grad_none"]) == {"cblock.block.0.weight", "cblock.block.1.weight", "cblock.block.1.bias",
                                   "cblock.block.3.weight", "cblock.block.4.weight", "cblock.block.4.bias"}
    assert int(g["bn_after"]["d_blocks.0.block.1.num_batches_tracked"]) == 2
    assert int(g["bn_after"]["cblock.block.1.num_batches_tracked"]) == 1


def test_dice_and_hard_segm(golden):
    d = golden["dice"]
    assert float(O.dice_loss(d["p"], d["t"])) == d["value"]
    h = golden["hard_segm"]
    assert torch.equal(O.hard_segm_from_tensor(h["x"]), h["y"])
    assert torch.equal(O.hard_segm_from_tensor(h["x"][0]), h["y4"])
    assert O.hard_segm_from_tensor(h["x"]).dtype == torch.float32


def test_shape_3d_and_blank_patch(golden):
    s = golden["shape_3d"]
    sph = O.shape_3d((8, 8, 8), 4, (16, 16, 16), "sphere")
    box = O.shape_3d((8, 8, 8), 4, (16, 16, 16), "box")
    assert int((sph == 0).sum()) == s["sphere_zeros"] == 257
    assert int((box == 0).sum()) == s["box_zeros"] == 729
    assert str(sph.dtype) == s["dtype"]
    sph2 = O.shape_3d((3, 10, 5), 6, (12, 16, 14), "sphere")
    box2 = O.shape_3d((3, 10, 5), 6, (12, 16, 14), "box")
    assert np.array_equal(np.packbits(sph2.astype(np.uint8)), s["sphere2"].numpy())
    assert np.array_equal(np.packbits(box2.astype(np.uint8)), s["box2"].numpy())
    b = golden["blank_patch"]
    img = b["img"].numpy()
    masked, extracted = O.blank_patch(img, b["center"], b["size"], "sphere")
    assert np.array_equal(masked, b["masked"].numpy()) and np.array_equal(extracted, b["extracted"].numpy())
    assert np.array_equal(masked + extracted, img)
    assert list(O.radius_bounds(img.shape)) == b["radius_bounds"]
    assert int((img > 0).sum()) == b["n_nonzero"]
    assert list(O.kth_nonzero(img, 0)) == [int(v) for v in np.argwhere(img > 0)[0]]


def test_flap_shape_matches_reference_code_over_restated_raster_geometry():
    """tests/golden/flap_shape_golden.pt comes from the reference's own shape_3d / random_blank_patch (make_golden_flap.py)
    with only raster_geometry.cylinder / cube restated: the oracle must reproduce it bit for bit."""
    import os
    import random
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "flap_shape_golden.pt"))
    for c in gold["cases"]:
        shp = O.shape_3d(c["center"], c["size"], c["image_size"], "flap", c["c_diam"])
        assert str(shp.dtype) == c["dtype"] == "uint8" and int((shp == 0).sum()) == c["zeros"]
        assert np.array_equal(np.packbits(shp), c["packed"].numpy())
        np.random.seed(c["seed"])                          # the radius draw happens inside shape_3d when not supplied
        assert np.array_equal(O.shape_3d(c["center"], c["size"], c["image_size"], "flap"), shp)
    r = gold["random_blank_patch"]
    img = r["img"].numpy()
    random.seed(r["seed"])
    np.random.seed(r["seed"])
    random.uniform(0, 1)                                   # transforms.py:243
    center = O.kth_nonzero(img, int(np.random.choice(int((img > 0).sum()))))   # :252
    lo, hi = O.radius_bounds(img.shape)
    size = np.random.randint(lo, hi)                       # :268
    masked, extracted = O.blank_patch(img, center, size, "flap")                # draws c_diam (utilities.py:146)
    assert np.array_equal(masked, r["masked"].numpy()) and np.array_equal(extracted, r["extracted"].numpy())
    assert np.array_equal(masked + extracted, img) and extracted.sum() > 0


def test_encode_flaprec_batch_is_one_hot_plus_atlas():
    g = torch.Generator().manual_seed(2)
    full = (torch.rand(2, 4, 6, 8, generator=g) > 0.5).to(torch.uint8)
    flap = full * (torch.rand(2, 4, 6, 8, generator=g) > 0.5).to(torch.uint8)
    atlas = torch.rand(4, 6, 8, generator=g)
    img, (sk, fl) = O.encode_flaprec_batch(full - flap, full, flap, atlas)
    assert img.shape == (2, 2, 4, 6, 8) and sk.shape == fl.shape == (2, 2, 4, 6, 8) and sk.dtype == torch.float32
    assert torch.equal(img[:, 0], (full - flap).float()) and torch.equal(img[1, 1], atlas)
    assert torch.equal(sk[:, 1], full.float()) and torch.equal(sk[:, 0], 1 - full.float())
    assert torch.equal(fl.argmax(1), flap.long())
    assert O.encode_flaprec_batch(full - flap, full, flap)[0].shape == (2, 1, 4, 6, 8)


def test_nearest_index_matches_torch():
    for n_in, n_out in [(512, 128), (256, 128), (100, 37), (37, 100), (128, 128), (7, 3)]:
        v = torch.arange(n_in, dtype=torch.float32)[None, None, :, None, None].expand(1, 1, n_in, 1, 1)
        ref = torch.nn.functional.interpolate(v, size=(n_out, 1, 1), mode="nearest")[0, 0, :, 0, 0].long().numpy()
        assert np.array_equal(O.nearest_src_index(n_out, n_in), ref)
