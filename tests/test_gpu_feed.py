"""f3 on the GPU: the frozen ResNet-152 trunk as a side-stream producer (channels-last, bf16, explicit
BN mode), the one-batch-ahead loader, and a FeatureStore-fed training loop that reproduces the directly-fed one."""
import numpy as np
import pytest
import torch

from test_trainer_cpu import Vocab, make_opt

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm())


def test_trunk_feed_and_prefetch_loader():
    import show_and_tell_b200 as snt
    from show_and_tell_b200.feed import PrefetchLoader, TrunkFeed
    torch.manual_seed(0)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                 # fp32 convolutions for the comparison below
    try:
        _trunk_feed_checks(snt, PrefetchLoader, TrunkFeed)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def _trunk_feed_checks(snt, PrefetchLoader, TrunkFeed):
    enc = snt.EncoderCNN(64, backbone=True, precision="fp32").cuda().eval()
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(8, 3, 64, 64, generator=g), torch.randint(4, 100, (8, 9), generator=g),
                [9, 9, 8, 8, 7, 7, 6, 6], list(range(k * 8, k * 8 + 8))) for k in range(3)]
    r = enc.resnet
    with torch.no_grad():                                   # the trunk exactly as EncoderCNN.forward runs it
        x = batches[0][0].cuda()
        x = r.maxpool(r.relu(r.bn1(r.conv1(x))))
        ref = torch.flatten(r.avgpool(r.layer4(r.layer3(r.layer2(r.layer1(x))))), 1)
    feed32 = TrunkFeed(enc, dtype=torch.float32, bn_mode="eval")
    got32 = feed32.result(feed32.submit(batches[0][0]))
    assert tuple(got32.shape) == (8, 2048) and got32.dtype == torch.float32
    assert _rel(got32, ref) < 1e-3                           # same fp32 math, channels-last algorithms
    feed = TrunkFeed(enc, dtype=torch.bfloat16, bn_mode="eval")
    got = feed.result(feed.submit(batches[0][0].pin_memory()))
    # bf16 autocast through 152 layers: a sanity bound on the feed, not a parity bar (the trunk is torch/cuDNN's)
    assert _rel(got, ref) < 0.15 and not enc.resnet.bn1.training
    # the head consumes the fed features through the C ABI exactly like directly computed ones
    assert _rel(enc.forward_pooled(got32), enc(batches[0][0].cuda())) < 1e-2
    # the loader: same order, captions intact, features = the feed's
    seen = list(PrefetchLoader(batches, "cuda", trunk=feed))
    assert len(seen) == 3
    for (imgs, caps, lengths, ids), (f, c, l, i) in zip(batches, seen):
        assert f.is_cuda and c.is_cuda and torch.equal(c.cpu(), caps) and l == lengths and i == ids
        assert _rel(f, feed.pooled(imgs.cuda())) < 1e-3
    torch.cuda.synchronize()


def test_store_fed_training_matches_direct_feeding(tmp_path):
    import show_and_tell_b200 as snt
    from show_and_tell_b200.feed import FeatureStore, PrefetchLoader
    E, H, V = 64, 128, 500
    batches = []
    for k in range(3):
        b = snt.synthetic.make_batch(32, V, seed=60 + k, pooled_dim=2048)
        batches.append((torch.from_numpy(b["pooled"]), torch.from_numpy(b["captions"]), b["lengths"],
                        list(range(k * 32, k * 32 + 32))))
    store = FeatureStore.create(str(tmp_path / "feat"), list(range(96)), dim=2048, dtype="float32")
    for pooled, _, _, ids in batches:
        store.put(ids, pooled)
    store.flush()
    rd = FeatureStore.open(str(tmp_path / "feat"))
    blank = [(torch.zeros(1), c, l, i) for _, c, l, i in batches]     # the images are never touched
    losses = []
    for loader in (batches, PrefetchLoader(blank, "cuda", store=rd)):
        torch.manual_seed(9)
        model = snt.CaptionModel(E, H, V, 1, backbone=False, precision="bf16").cuda()
        opt = make_opt(tmp_path / "x", embed_size=E, hidden_size=H, num_gpu=1, max_epochs=2, log_step=100)
        tr = snt.Trainer(opt, loader, None, vocab=Vocab(V), model=model)
        got = []
        step = tr.train_step
        tr.train_step = lambda *a: (got.append(step(*a)), got[-1])[1]
        tr.train()
        losses.append([float(x) for x in got])
    assert len(losses[0]) == 6 and losses[0] == losses[1]             # identical inputs, deterministic kernels
