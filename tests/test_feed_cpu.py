"""f3: the encoder feed — on-disk precomputed-feature format, the frozen trunk as a producer with
an explicit BN mode, and the one-batch-ahead prefetching loader.  CPU part: formats and ordering (no streams here)."""
import numpy as np
import pytest
import torch
import torch.nn as nn


class TinyResNet(nn.Module):
    """The attribute layout of torchvision's ResNet (what models.py:13-16 relies on), three orders smaller."""

    def __init__(self, dim=2048):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 8, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(8)
        self.relu = nn.ReLU()
        self.maxpool = nn.MaxPool2d(2)
        self.layer1 = nn.Sequential(nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU())
        self.layer2 = nn.Identity()
        self.layer3 = nn.Identity()
        self.layer4 = nn.Sequential(nn.Conv2d(8, dim, 1, bias=False), nn.ReLU())
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(dim, 4)


class TinyEncoder(nn.Module):
    has_backbone = True

    def __init__(self, dim=2048):
        super().__init__()
        self.resnet = TinyResNet(dim)


def _batches(n_batches, batch, dup=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        ids = [100 + i * batch + j for j in range(batch)]
        if dup and i > 0:
            ids[0] = 100                                  # an image that appears again with another caption
        out.append((torch.randn(batch, 3, 8, 8, generator=g), torch.randint(0, 50, (batch, 7), generator=g),
                    list(range(7, 7 - batch, -1)), ids))
    return out


def test_feature_store_round_trip(tmp_path):
    from show_and_tell_b200.feed import FeatureStore
    ids = [7, 3, 11, 5, 2]
    rng = np.random.default_rng(0)
    feats = np.abs(rng.standard_normal((5, 64))).astype(np.float32)
    for dtype, tol in (("float32", 0.0), ("float16", 1e-3)):
        d = tmp_path / dtype
        st = FeatureStore.create(str(d), ids, dim=64, dtype=dtype, bn_mode="eval")
        st.put(ids[:2], feats[:2])
        st.put(ids[2:], torch.from_numpy(feats[2:]))
        st.flush()
        rd = FeatureStore.open(str(d))
        assert len(rd) == 5 and rd.meta == {"dim": 64, "dtype": dtype, "bn_mode": "eval", "format": 1}
        got = rd.gather([5, 7, 7, 2])
        assert got.dtype == torch.float32 and tuple(got.shape) == (4, 64)
        want = feats[[3, 0, 0, 4]]
        assert np.abs(got.numpy() - want).max() <= tol * np.abs(want).max()
        with pytest.raises(RuntimeError):
            rd.put([7], feats[:1])
        with pytest.raises(KeyError):
            rd.gather([12345])
        assert np.load(str(d / "pooled.npy")).shape == (5, 64)          # a plain .npy any tool can read
    with pytest.raises(ValueError):
        FeatureStore.create(str(tmp_path / "dup"), [1, 1], dim=4)
    with pytest.raises(ValueError):
        FeatureStore.create(str(tmp_path / "bad"), [1], dtype="int8")


def test_trunk_feed_bn_modes_and_precompute(tmp_path):
    from show_and_tell_b200.feed import FeatureStore, PrefetchLoader, TrunkFeed
    torch.manual_seed(0)
    enc = TinyEncoder(32)
    with torch.no_grad():                                  # non-trivial running statistics
        enc.resnet.bn1.running_mean.uniform_(-0.5, 0.5)
        enc.resnet.bn1.running_var.uniform_(0.5, 2.0)
    batches = _batches(3, 4, dup=True)
    r = enc.resnet

    def direct(x):
        with torch.no_grad():
            y = r.maxpool(r.relu(r.bn1(r.conv1(x))))
            return torch.flatten(r.avgpool(r.layer4(r.layer3(r.layer2(r.layer1(y))))), 1)

    feed = TrunkFeed(enc, dtype=torch.float32, channels_last=True, bn_mode="eval")
    enc.train()                                            # the caller's mode must not leak into the feed ...
    rm = r.bn1.running_mean.clone()
    got = feed.result(feed.submit(batches[0][0]))
    assert enc.resnet.bn1.training and torch.equal(r.bn1.running_mean, rm)     # ... nor the feed's into the caller's
    enc.eval()
    assert torch.allclose(got, direct(batches[0][0]), rtol=1e-5, atol=1e-6)
    # "train" = batch statistics, what the reference's trunk really does under model.train()
    enc.train()
    want_train = direct(batches[0][0])
    r.bn1.running_mean.copy_(rm)
    feed_t = TrunkFeed(enc, dtype=torch.float32, bn_mode="train")
    assert torch.allclose(feed_t.pooled(batches[0][0]), want_train, rtol=1e-5, atol=1e-6)
    assert not torch.allclose(want_train, got, rtol=1e-3, atol=1e-4)
    with pytest.raises(RuntimeError):
        TrunkFeed(nn.Module())
    # bf16 autocast stays close to fp32
    enc.eval()
    bf = TrunkFeed(enc, dtype=torch.bfloat16, bn_mode="eval").pooled(batches[0][0])
    assert bf.dtype == torch.float32 and torch.allclose(bf, got, rtol=5e-2, atol=5e-2)

    # precompute -> store -> loader that ignores the images and serves stored features
    all_ids = sorted({i for b in batches for i in b[3]})
    store = FeatureStore.create(str(tmp_path / "s"), all_ids, dim=32)
    assert feed.precompute(batches, store) == len(all_ids) == 10
    rd = FeatureStore.open(str(tmp_path / "s"))
    served = list(PrefetchLoader(batches, "cpu", store=rd))
    assert len(served) == 3
    for (imgs, caps, lengths, ids), (f, c, l, i) in zip(batches, served):
        assert torch.equal(c, caps) and l == lengths and i == ids and tuple(f.shape) == (4, 32)
    assert torch.allclose(served[2][0][1:], direct(batches[2][0])[1:], rtol=1e-5, atol=1e-6)
    assert torch.equal(served[1][0][0], served[0][0][0])               # the repeated image: first occurrence's row


def test_prefetch_loader_order_and_trunk_on_cpu():
    from show_and_tell_b200.feed import PrefetchLoader, TrunkFeed
    batches = _batches(4, 3)
    out = list(PrefetchLoader(batches, "cpu"))
    assert len(out) == 4 and len(PrefetchLoader(batches, "cpu")) == 4
    for a, b in zip(batches, out):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
    assert list(PrefetchLoader([], "cpu")) == []
    enc = TinyEncoder(16).eval()
    feed = TrunkFeed(enc, dtype=torch.float32)
    via = list(PrefetchLoader(batches, "cpu", trunk=feed))
    for a, b in zip(batches, via):
        assert torch.allclose(b[0], feed.pooled(a[0])) and tuple(b[0].shape) == (3, 16)
