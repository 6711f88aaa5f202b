"""f1: the train.py / eval.py loops re-hosted on the B200 modules (show_and_tell_b200.trainer).
CPU part: the loop logic — lr schedule, iteration bookkeeping, evaluation cadence, best-score rule, checkpoint files,
resume, prediction de-duplication, <end> trimming — driven with the CPU torch port (oracle/torch_port.py, test
infrastructure) standing in for the CUDA modules.  The same loops over the real CUDA modules: tests/test_gpu_trainer.py.
"""
import argparse
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import snt_oracle as O
from oracle import torch_port as TP

E, H, V, L, POOLED = 16, 24, 61, 1, 32


class Vocab:
    """utils.py:22-40 shaped stand-in: idx2word / word2idx / len; ids 0..3 as preprocess.py:75-78 fixes them."""

    def __init__(self, n):
        words = ["<pad>", "<start>", "<end>", "<unk>"] + ["w%d" % i for i in range(4, n)]
        self.idx2word = dict(enumerate(words))
        self.word2idx = {w: i for i, w in self.idx2word.items()}

    def __len__(self):
        return len(self.idx2word)


class HeadCPU(TP.EncoderHeadCPU):
    def forward_pooled(self, pooled):
        return self(pooled)


class DecCPU(TP.CaptionDecoderCPU):
    def loss(self, features, captions, lengths, targets):
        return torch.nn.functional.cross_entropy(self(features, captions, lengths), targets)

    def sample(self, features, states=None):
        return super().sample(features, states).squeeze()          # models.py:66-67


class StepperCPU:
    """train.py:137-146 with torch's own ops: zero_grad, forward, CE, backward, clip_gradient, Adam."""
    world = 1

    def __init__(self, model, lr, grad_clip):
        self.model, self.lr, self.grad_clip = model, lr, grad_clip
        self.opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr)
        self.lrs, self.calls = [], 0

    def step(self, images, captions, lengths, targets, n_tokens_global=None):
        for g in self.opt.param_groups:
            g["lr"] = self.lr
        self.lrs.append(self.lr)
        self.calls += 1
        if targets is None:                                         # the Trainer leaves the pack to the stepper (eval.py:91)
            targets = torch.nn.utils.rnn.pack_padded_sequence(captions, [int(l) for l in lengths], batch_first=True)[0]
        self.model.zero_grad()
        loss = self.model.loss(images, captions, lengths, targets)
        loss.backward()
        for g in self.opt.param_groups:
            for p in g["params"]:
                p.grad.data.clamp_(-self.grad_clip, self.grad_clip)
        self.opt.step()
        return loss.detach()


def make_opt(tmp, **kw):
    d = dict(num_gpu=0, embed_size=E, hidden_size=H, num_layers=L, learning_rate=1e-3, max_epochs=2,
             learning_rate_decay_start=1, learning_rate_decay_every=3, learning_rate_decay_rate=0.8, grad_clip=0.1,
             log_step=2, language_eval=0, save_checkpoint_every=2, expr_dir=str(tmp), start_from=None,
             load_best_score=True, load_pretrained=False, load_model_path=None, vocab_path=None)
    d.update(kw)
    return argparse.Namespace(**d)


def make_loader(n_batches, batch, seed, first_imgid=0):
    import show_and_tell_b200 as snt
    out = []
    for i in range(n_batches):
        b = snt.synthetic.make_batch(batch, V, seed=seed + i, pooled_dim=POOLED)
        imgids = [first_imgid + i * batch + j for j in range(batch)]
        if i == 1:
            imgids[1] = imgids[0]                                   # one image with two captions in the same batch
        out.append((torch.from_numpy(b["pooled"]), torch.from_numpy(b["captions"]), b["lengths"], imgids))
    return out


def make_model(seed=5):
    import show_and_tell_b200 as snt
    torch.manual_seed(seed)
    return snt.CaptionModel(E, H, V, L, encoder=HeadCPU(E, POOLED), decoder=DecCPU(E, H, V, L))


def oracle_trim(ids, end_id=2, pad_id=0):
    a, l = O.trim_captions(ids.numpy(), end_id, pad_id)
    return torch.from_numpy(a), torch.from_numpy(l)


def test_learning_rate_schedule_follows_train_py():
    import show_and_tell_b200 as snt
    from show_and_tell_b200.trainer import learning_rate_for_epoch
    for start, every, rate in [(1, 3, 0.8), (-1, 3, 0.8), (0, 2, 0.5), (4, 1, 0.9)]:
        opt = make_opt(".", learning_rate_decay_start=start, learning_rate_decay_every=every,
                       learning_rate_decay_rate=rate, learning_rate=4e-4)
        for epoch in range(1, 12):
            if epoch > start and start >= 1:                         # train.py:98-105
                want = 4e-4 * rate ** ((epoch - start) // every)
            else:
                want = 4e-4
            assert learning_rate_for_epoch(opt, epoch) == pytest.approx(want, rel=1e-12)


def test_trainer_loop_checkpoints_and_matches_hand_rolled_steps(tmp_path, capsys, monkeypatch):
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import trainer as T
    monkeypatch.setattr(T.ops, "trim_captions", oracle_trim)         # the device kernel's CPU checker stands in
    opt = make_opt(tmp_path, learning_rate_decay_start=1, learning_rate_decay_every=1, learning_rate_decay_rate=0.5)
    train, valid = make_loader(3, 6, seed=10), make_loader(2, 5, seed=90, first_imgid=1000)
    model = make_model()
    stepper = StepperCPU(model, opt.learning_rate, opt.grad_clip)
    tr = snt.Trainer(opt, train, valid, vocab=Vocab(V), model=model, stepper=stepper)
    assert tr.total_train_iter == 3 and tr.total_valid_iter == 2
    infos = tr.train()
    # 2 epochs x 3 iterations; epoch 2 runs at half the learning rate (decay start 1, every 1, rate 0.5)
    assert stepper.calls == 6 and stepper.lrs == [1e-3] * 3 + [5e-4] * 3
    assert infos["total_iter"] == 6 and infos["epoch"] == 2 and infos["iter"] == 3
    assert sorted(infos["val_result_history"]) == [2, 4, 6] and sorted(infos["lr_history"]) == [2, 4, 6]
    assert infos["lr_history"] == {2: 1e-3, 4: 5e-4, 6: 5e-4}
    losses = [infos["val_result_history"][k]["loss"] for k in (2, 4, 6)]
    assert infos["best_val_score"] == pytest.approx(-min(losses))    # language_eval == 0: score = -val_loss
    for f in ("infos.pkl", "infos-best.pkl", "model-best.pth"):
        assert os.path.exists(tmp_path / f), f
    with open(tmp_path / "infos.pkl", "rb") as f:
        assert pickle.load(f)["total_iter"] == 6
    sd = torch.load(tmp_path / "model-best.pth")
    assert set(sd) == set(model.state_dict()) and "decoder.lstm.weight_hh_l0" in sd and "encoder.bn.running_mean" in sd
    # predictions: one entry per distinct image id (eval.py:112-116), words only, cut before <end>
    preds = infos["val_result_history"][6]["predictions"]
    assert len(preds) == 9 and len({p["image_id"] for p in preds}) == 9          # 10 rows, one duplicated image
    assert all("<end>" not in p["caption"].split() for p in preds)
    assert model.training                                             # evaluation restores train mode
    out = capsys.readouterr().out
    assert out.count("Epoch [") == 2 and "Step [2/3]" in out and "model saved to" in out

    # the same six steps, hand-rolled on an identically initialised model: identical loss trajectory
    ref = make_model()
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    hand = []
    for epoch in (1, 2):
        for g in ropt.param_groups:
            g["lr"] = 1e-3 if epoch == 1 else 5e-4
        for pooled, caps, lengths, _ in train:
            targets = torch.from_numpy(snt.synthetic.pack_host(caps.numpy(), lengths))
            ref.zero_grad()
            loss = torch.nn.functional.cross_entropy(ref(pooled, caps, lengths), targets)
            loss.backward()
            for p in ref.parameters():
                p.grad.clamp_(-0.1, 0.1)
            ropt.step()
            hand.append(float(loss.detach()))
    assert float(tr.last_loss) == pytest.approx(hand[-1], rel=1e-6)
    assert infos["loss_history"][2] == pytest.approx(hand[1], rel=1e-6)
    assert infos["loss_history"][4] == pytest.approx(hand[3], rel=1e-6)
    for (k, a), b in zip(sorted(model.state_dict().items()), [v for _, v in sorted(ref.state_dict().items())]):
        assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-7), k


def test_trainer_resumes_inside_the_saved_epoch(tmp_path, monkeypatch):
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import trainer as T
    monkeypatch.setattr(T.ops, "trim_captions", oracle_trim)
    train, valid = make_loader(3, 4, seed=20), make_loader(1, 4, seed=70, first_imgid=500)
    opt = make_opt(tmp_path, max_epochs=1, save_checkpoint_every=2)
    model = make_model()
    first = snt.Trainer(opt, train, valid, vocab=Vocab(V), model=model, stepper=StepperCPU(model, 1e-3, 0.1))
    infos = first.train()
    assert infos["iter"] == 2 and infos["epoch"] == 1 and infos["total_iter"] == 2   # saved after iteration 2 of 3
    opt2 = make_opt(tmp_path, max_epochs=2, save_checkpoint_every=2, start_from=str(tmp_path))
    model2 = make_model()
    st2 = StepperCPU(model2, 1e-3, 0.1)
    second = snt.Trainer(opt2, train, valid, vocab=Vocab(V), model=model2, stepper=st2)
    infos2 = second.train()
    assert st2.calls == 1 + 3                                          # iteration 3 of epoch 1, then all of epoch 2
    assert infos2["total_iter"] == 6 and sorted(infos2["val_result_history"]) == [2, 4, 6]
    assert infos2["best_val_score"] >= infos["best_val_score"]         # the saved best score is carried over


def test_evaluation_strict_and_fused_agree_and_cider_hook(tmp_path, monkeypatch):
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import trainer as T
    monkeypatch.setattr(T.ops, "trim_captions", oracle_trim)
    valid = make_loader(2, 5, seed=33)
    model = make_model()
    vocab = Vocab(V)
    opt = make_opt(tmp_path)
    crit = torch.nn.CrossEntropyLoss()
    seen = []
    l_strict, p_strict, stats = snt.evaluation(model, crit, valid, vocab, opt,
                                               language_eval=lambda preds: seen.append(len(preds)) or {"CIDEr": 0.5})
    l_fused, p_fused, none = snt.evaluation(model, None, valid, vocab, opt)
    assert l_strict == pytest.approx(l_fused, rel=1e-6) and p_strict == p_fused
    assert stats == {"CIDEr": 0.5} and none == {} and seen == [len(p_strict)]
    # the loss is the mean over batches of the per-batch token means (eval.py:95-97,119), in eval mode (running stats)
    model.eval()
    want = []
    with torch.no_grad():
        for pooled, caps, lengths, _ in valid:
            tg = torch.from_numpy(snt.synthetic.pack_host(caps.numpy(), lengths))
            want.append(float(crit(model(pooled, caps, lengths), tg)))
    assert l_strict == pytest.approx(np.mean(want), rel=1e-6)
    # captions are the words of the reference's own loop over the sampled ids (eval.py:103-110)
    with torch.no_grad():
        ids = model.sample(valid[0][0]).numpy()
    for row, pred in zip(ids, p_strict):
        words = []
        for w in row:
            if vocab.idx2word[int(w)] == "<end>":
                break
            words.append(vocab.idx2word[int(w)])
        assert pred["caption"] == " ".join(words)
    # CIDEr drives the best-score rule when language_eval == 1 (train.py:172-175)
    opt1 = make_opt(tmp_path, language_eval=1, max_epochs=1)
    m2 = make_model()
    tr = snt.Trainer(opt1, make_loader(2, 4, seed=1), valid, vocab=vocab, model=m2, stepper=StepperCPU(m2, 1e-3, 0.1),
                     language_eval=lambda preds: {"CIDEr": 0.25})
    assert tr.train()["best_val_score"] == 0.25
