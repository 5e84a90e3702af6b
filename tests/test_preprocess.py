"""Input transform: PolypDiffusionDataset.py:52-59 restated (oracle/preprocess.py), pinned
against Pillow / torchvision THEMSELVES (third-party dependencies of the reference that are installed here), and the
device pipeline against both."""
import numpy as np
import pytest
import torch

from oracle import preprocess as opre

SHAPES = [(288, 384, 128), (500, 574, 128), (64, 64, 128), (100, 37, 64), (128, 128, 128), (300, 128, 128),
          (129, 131, 64), (576, 720, 224)]


def _frames(b, h, w, seed):
    rng = np.random.default_rng(seed)
    smooth = rng.integers(0, 256, (b, h // 8 + 2, w // 8 + 2, 3)).repeat(8, 1).repeat(8, 2)[:, :h, :w]
    noise = rng.integers(-40, 41, (b, h, w, 3))
    return np.clip(smooth + noise, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,s", SHAPES)
def test_oracle_resize_is_pillow_bit_exact(h, w, s):
    from PIL import Image
    img = _frames(1, h, w, h * 7 + w)[0]
    want = np.asarray(Image.fromarray(img).resize((s, s), Image.BILINEAR))
    assert np.array_equal(opre.resize_bilinear_u8(img, s, s), want)


def test_oracle_transform_is_torchvision_bit_exact():
    from PIL import Image
    import torchvision.transforms as T
    img = _frames(1, 288, 384, 3)[0]
    for flip in (False, True):
        tv = T.Compose([T.Resize((128, 128)), T.RandomHorizontalFlip(p=1.0 if flip else 0.0), T.ToTensor(),
                        T.Normalize([0.5], [0.5])])
        assert torch.equal(tv(Image.fromarray(img)), opre.transform(img, 128, flip))


def test_product_tables_match_oracle_tables():
    from polyp_image_generator_b200.preprocess import pillow_bilinear_tables
    for n_in, n_out in ((384, 128), (288, 128), (37, 64), (64, 128), (720, 224), (131, 64)):
        b, c, k = pillow_bilinear_tables(n_in, n_out)
        ob, oc, ok = opre.precompute_coeffs(n_in, n_out)
        assert k == ok and np.array_equal(b, ob) and np.array_equal(c, oc)
        assert (c.sum(1) - (1 << 22)).__abs__().max() <= k          # weights sum to one up to rounding
    b, c, k = pillow_bilinear_tables(128, 128)
    assert k == 1 and np.array_equal(b[:, 0], np.arange(128)) and (c == 1 << 22).all()


def test_flip_draws_follow_torchvision_rng_order():
    """RandomHorizontalFlip draws torch.rand(1) per image from the global CPU RNG (PolypDiffusionDataset.__getitem__)."""
    from PIL import Image
    import torchvision.transforms as T
    from polyp_image_generator_b200.preprocess import DevicePreprocessor
    img = Image.fromarray(np.arange(12, dtype=np.uint8).reshape(2, 2, 3))
    torch.manual_seed(123)
    tv = T.RandomHorizontalFlip()
    want = [not np.array_equal(np.asarray(tv(img)), np.asarray(img)) for _ in range(16)]
    torch.manual_seed(123)
    got = DevicePreprocessor(128).draw_flips(16).tolist()
    assert got == want and any(got) and not all(got)


@pytest.mark.parametrize("h,w,s", [(100, 37, 64), (64, 64, 128), (128, 128, 128), (300, 128, 64)])
def test_device_pipeline_host_logic_through_emulation(emu_backend, h, w, s):
    from polyp_image_generator_b200.preprocess import DevicePreprocessor
    fr = _frames(3, h, w, 5)
    flips = torch.tensor([True, False, True])
    got = DevicePreprocessor(s)(torch.from_numpy(fr), flips)
    want = torch.stack([opre.transform(fr[i], s, bool(flips[i])) for i in range(3)])
    assert got.shape == (3, 3, s, s) and torch.equal(got, want)
    with pytest.raises(TypeError):
        DevicePreprocessor(s)(torch.zeros(1, 3, 8, 8), None)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,s", SHAPES)
def test_device_pipeline_bit_exact_vs_pillow(h, w, s):
    """The CUDA transform against torchvision on PIL images: bit-identical fp32 tensors."""
    from PIL import Image
    import torchvision.transforms as T
    from polyp_image_generator_b200.preprocess import DevicePreprocessor
    dev = torch.device("cuda:0")
    fr = _frames(5, h, w, h + w)
    flips = torch.tensor([False, True, True, False, True])
    got = DevicePreprocessor(s)(torch.from_numpy(fr).to(dev), flips).cpu()
    for i in range(5):
        tv = T.Compose([T.Resize((s, s)), T.RandomHorizontalFlip(p=1.0 if flips[i] else 0.0), T.ToTensor(),
                        T.Normalize([0.5], [0.5])])
        assert torch.equal(got[i], tv(Image.fromarray(fr[i]))), i
    gray = DevicePreprocessor(s)(torch.from_numpy(fr[:2, :, :, :1].copy()).to(dev), None).cpu()
    assert torch.equal(gray[0, 0], opre.transform(fr[0, :, :, :1], s, False)[0])
    assert DevicePreprocessor(s)(torch.empty(0, h, w, 3, dtype=torch.uint8, device=dev)).shape == (0, 3, s, s)
    with pytest.raises(RuntimeError, match="CUDA"):
        DevicePreprocessor(s)(torch.from_numpy(fr), flips)


@pytest.mark.gpu
def test_pinned_prefetcher_delivers_batches_in_order():
    from polyp_image_generator_b200.preprocess import DevicePreprocessor, PinnedPrefetcher
    dev = torch.device("cuda:0")
    batches = [(torch.from_numpy(_frames(4, 96, 80, i)), i) for i in range(5)]
    pre = DevicePreprocessor(64)
    seen = []
    for frames, tag in PinnedPrefetcher(batches, dev):
        assert frames.is_cuda and torch.equal(frames.cpu(), batches[tag][0])
        seen.append(tag)
        assert pre(frames).shape == (4, 3, 64, 64)
    assert seen == [0, 1, 2, 3, 4]


@pytest.mark.gpu
def test_device_pipeline_matches_committed_pillow_golden():
    """tests/golden/resize.json (written by Pillow / torchvision here): the CUDA transform reproduces it bit for bit."""
    import importlib.util
    import json
    import os
    from polyp_image_generator_b200.preprocess import DevicePreprocessor
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    gold = json.load(open(os.path.join(gdir, "resize.json")))
    spec = importlib.util.spec_from_file_location("mk3", os.path.join(gdir, "make_golden_session3.py"))
    mk3 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk3)
    img = mk3.pattern(gold["h"], gold["w"])
    dev = torch.device("cuda:0")
    t = DevicePreprocessor(gold["size"])(torch.from_numpy(img)[None].to(dev), torch.tensor([True])).cpu()[0]
    assert t.double().sum().item() == pytest.approx(gold["transform_flipped_sum"], abs=1e-9)
    assert torch.equal(t[0, 0], torch.tensor(gold["transform_flipped_first_row"]))
    plain = DevicePreprocessor(gold["size"])(torch.from_numpy(img)[None].to(dev), None).cpu()[0]
    want = (torch.tensor(gold["resized"], dtype=torch.float32).permute(2, 0, 1) / 255 - 0.5) / 0.5
    assert torch.equal(plain, want)
