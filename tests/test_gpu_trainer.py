"""f1 on the GPU: the re-hosted train.py / eval.py loops (show_and_tell_b200.trainer) driving the
real CUDA modules through the C ABI, checked against the same loops run with torch's own CPU ops (oracle/torch_port.py)
from identical weights and batches.  Tolerances: fp32 mode 5e-4 rel on every loss of the six-step trajectory, bf16 mode 5e-3 (single-step bars: 1e-5 / 1e-3;
Adam's sign-like first updates amplify gradient noise on near-zero entries a little)."""
import numpy as np
import pytest
import torch

from oracle import snt_oracle as O
from oracle import torch_port as TP
from test_trainer_cpu import Vocab, make_opt

pytestmark = pytest.mark.gpu

E, H, V, L, POOLED = 64, 128, 1000, 1, 2048


def _loader(n_batches, batch, seed, first_imgid=0):
    import show_and_tell_b200 as snt
    out = []
    for i in range(n_batches):
        b = snt.synthetic.make_batch(batch, V, seed=seed + i, pooled_dim=POOLED)
        imgids = [first_imgid + i * batch + j for j in range(batch)]
        out.append((torch.from_numpy(b["pooled"]), torch.from_numpy(b["captions"]), b["lengths"], imgids))
    return out


def _cpu_twin(model):
    """The torch-nn CPU composition carrying the same weights (state_dict names match the reference's)."""
    head, dec = TP.EncoderHeadCPU(E, POOLED), TP.CaptionDecoderCPU(E, H, V, L)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    head.fc.load_state_dict({"weight": sd["encoder.resnet.fc.weight"], "bias": sd["encoder.resnet.fc.bias"]})
    head.bn.load_state_dict({k[len("encoder.bn."):]: v for k, v in sd.items() if k.startswith("encoder.bn.")})
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    return head, dec


@pytest.mark.parametrize("prec,tol", [("fp32", 5e-4), ("bf16", 5e-3)])
def test_trainer_loss_trajectory_matches_cpu_loop(prec, tol, tmp_path):
    import show_and_tell_b200 as snt
    torch.manual_seed(3)
    model = snt.CaptionModel(E, H, V, L, backbone=False, precision=prec).cuda()
    head, dec = _cpu_twin(model)
    train, valid = _loader(3, 48, seed=40), _loader(2, 32, seed=80, first_imgid=10_000)
    opt = make_opt(tmp_path, embed_size=E, hidden_size=H, num_gpu=1, max_epochs=2, save_checkpoint_every=3,
                   learning_rate_decay_start=1, learning_rate_decay_every=1, learning_rate_decay_rate=0.5, log_step=1)
    tr = snt.Trainer(opt, train, valid, vocab=Vocab(V), model=model)
    losses = []
    step = tr.train_step
    tr.train_step = lambda *a: (losses.append(step(*a)), losses[-1])[1]
    infos = tr.train()
    got = [float(x) for x in losses]
    assert len(got) == 6 and sorted(infos["val_result_history"]) == [3, 6]

    # train.py:137-146 with torch's CPU ops on the twin
    params = list(head.parameters()) + list(dec.parameters())
    ropt = torch.optim.Adam(params, lr=1e-3)
    want = []
    for epoch in (1, 2):
        for g in ropt.param_groups:
            g["lr"] = 1e-3 if epoch == 1 else 5e-4
        for pooled, caps, lengths, _ in train:
            targets = torch.from_numpy(snt.synthetic.pack_host(caps.numpy(), lengths))
            for p in params:
                p.grad = None
            loss = torch.nn.functional.cross_entropy(dec(head(pooled), caps, lengths), targets)
            loss.backward()
            for p in params:
                p.grad.clamp_(-0.1, 0.1)
            ropt.step()
            want.append(float(loss.detach()))
    for a, b in zip(got, want):
        assert abs(a - b) / b < tol, (got, want)
    assert got[-1] < got[0]                                            # it trains

    # the validation loss of the last checkpoint: fused path vs the CPU twin in eval mode
    head.eval(), dec.eval()
    ref = []
    with torch.no_grad():
        for pooled, caps, lengths, _ in valid:
            targets = torch.from_numpy(snt.synthetic.pack_host(caps.numpy(), lengths))
            ref.append(float(torch.nn.functional.cross_entropy(dec(head(pooled), caps, lengths), targets)))
    assert abs(infos["val_result_history"][6]["loss"] - np.mean(ref)) / np.mean(ref) < 5 * tol
    assert model.training and (tmp_path / "model-best.pth").exists()


def test_evaluation_strict_fused_and_captions(tmp_path):
    import show_and_tell_b200 as snt
    torch.manual_seed(4)
    model = snt.CaptionModel(E, H, V, L, backbone=False, precision="fp32").cuda()
    vocab, valid = Vocab(V), _loader(2, 40, seed=7)
    opt = make_opt(tmp_path, embed_size=E, hidden_size=H, num_gpu=1)
    l_strict, p_strict, _ = snt.evaluation(model, torch.nn.CrossEntropyLoss(), valid, vocab, opt)
    l_fused, p_fused, _ = snt.evaluation(model, None, valid, vocab, opt)
    assert abs(l_strict - l_fused) / l_fused < 1e-5 and p_strict == p_fused and len(p_fused) == 80
    model.eval()
    with torch.no_grad():
        ids = model.sample(valid[0][0].cuda()).cpu().numpy()
    kept_ids, kept = O.trim_captions(ids)
    for b in range(40):
        assert p_fused[b]["caption"] == " ".join(vocab.idx2word[int(w)] for w in kept_ids[b, :kept[b]])
        assert p_fused[b]["image_id"] == b
