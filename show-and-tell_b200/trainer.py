"""The reference's training / evaluation loops (train.py `Trainer`, eval.py `evaluation`) on the B200 modules
(the integration shim that proves the drop-in claim end to end).

What is kept from the reference: `Trainer(opt, trainloader, validloader)` with the `opt` fields of config.py, the
epoch / iteration bookkeeping and resume rule (train.py:66-80,96,116-121), the step-wise learning-rate decay
(train.py:98-105), one teacher-forced step per batch (train.py:123-146: forward, CrossEntropyLoss, backward,
clip_gradient, Adam), the log line (train.py:151-154), the evaluate-and-checkpoint cadence with the
"save when the validation score improves" rule and the `infos.pkl` / `model-best.pth` / `infos-best.pkl` files
(train.py:157-199); `evaluation(model, crit, loader, vocab, opt)` returns `(mean loss, predictions, lang_stats)` with
one `{'image_id', 'caption'}` entry per distinct image (eval.py:58-119).

What necessarily differs (SURVEY.md Appendix B lists the reference's defects; none is replicated):
  * the model is the `models.py` pair train.py:11 imports — `EncoderCNN` + `DecoderRNN` behind one callable
    (`CaptionModel`) — not `model2.ShowAttendTellModel` (train.py:37, an unfinished attention model);
  * the call convention is eval.py:91-93's (`targets = pack(captions, lengths)`, `model(images, captions, lengths)`):
    train.py:134-139 slices `captions[:, 1:]` / `captions[:, :-1]` for model2, which is off by one for `DecoderRNN`,
    whose timestep 0 is the image feature (models.py:50);
  * torch-0.1.12 idioms (`Variable`, `volatile=True`, `loss.data[0]`) are gone; evaluation runs under `no_grad`, and the
    module's train/eval mode is restored afterwards (the reference leaves the model in eval mode for good);
  * forward + loss + backward are one native call sequence (`parallel.DataParallelStep` -> `snt_step_run`) and
    clip_gradient + Adam one fused launch over flat parameter buffers; with `world_size > 1` (one process per GPU under torchrun, instead of train.py:43-44's
    `nn.DataParallel`) gradients are all-reduced over NCCL in readiness order;
  * sampled ids are cut at `<end>` on the device (`snt_caption_trim`) and copied to the host once per batch, instead
    of a Python loop over every word id (eval.py:101-109);
  * `language_eval` (pycocoevalcap: Python 2 + a JVM, out of scope) is a caller-supplied hook; without it the
    validation score is `-val_loss`, the reference's own `language_eval == 0` branch (train.py:172-175).
"""
from __future__ import annotations

import math
import os
import pickle
import time

import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence

from . import ops
from .models import DecoderRNN, EncoderCNN


class CaptionModel(nn.Module):
    """`self.model` of train.py:37-41 as train.py:11 intends it: the models.py pair behind the call signatures the
    loops use — `model(images, captions, lengths)` (train.py:139, eval.py:93) and `model.sample(images, state)`
    (eval.py:99).  `images` is either `[B,3,224,224]` (needs `backbone=True`) or precomputed pooled ResNet features
    `[B,2048]`.  Parameter names are `encoder.*` / `decoder.*` + the reference's own names."""

    def __init__(self, embed_size, hidden_size, vocab_size, num_layers, backbone=False, precision="bf16",
                 encoder=None, decoder=None):
        super().__init__()
        self.encoder = encoder if encoder is not None else EncoderCNN(embed_size, backbone=backbone, precision=precision)
        self.decoder = decoder if decoder is not None else DecoderRNN(embed_size, hidden_size, vocab_size, num_layers,
                                                                      precision=precision)

    def features(self, images):
        return self.encoder(images) if images.dim() == 4 else self.encoder.forward_pooled(images)

    def forward(self, images, captions, lengths):
        return self.decoder(self.features(images), captions, lengths)

    def loss(self, images, captions, lengths, targets):
        return self.decoder.loss(self.features(images), captions, lengths, targets)

    def sample(self, images, states=None):
        return self.decoder.sample(self.features(images), states)


def learning_rate_for_epoch(opt, epoch):
    """train.py:98-105: from the epoch after `learning_rate_decay_start` on, multiply by `learning_rate_decay_rate`
    once every `learning_rate_decay_every` epochs (a start < 1 disables the decay)."""
    start = opt.learning_rate_decay_start
    if start >= 1 and epoch > start:
        return opt.learning_rate * opt.learning_rate_decay_rate ** ((epoch - start) // opt.learning_rate_decay_every)
    return opt.learning_rate


def _to_device(t, device):
    return t.to(device, non_blocking=True) if isinstance(t, torch.Tensor) else torch.as_tensor(t, device=device)


def evaluation(model, crit, loader, vocab, opt, language_eval=None, trim=None):
    """eval.py:58-119.  `crit` given: the strict path, `crit(model(images, captions, lengths), targets)` with the
    logits materialised exactly as eval.py:93-95; `crit=None`: the fused `model.loss`.  -> (mean loss over the
    batches as a float, predictions, lang_stats).  `trim(ids, end_id, pad_id) -> (ids, lengths)` defaults to the
    device kernel; `<end>` / `<pad>` ids come from `vocab.word2idx` when it has them."""
    trim = trim or ops.trim_captions
    was_training = model.training
    model.eval()                                                      # eval.py:65
    device = next(model.parameters()).device
    word2idx = getattr(vocab, "word2idx", None) or {}
    end_id, pad_id = word2idx.get("<end>", 2), word2idx.get("<pad>", 0)       # preprocess.py:75-78 fixes them at 2 / 0
    loss_sum, loss_evals = 0.0, 0
    predictions, seen = [], set()
    with torch.no_grad():                                             # eval.py:79-80 `volatile=True`
        for images, captions, lengths, imgids in loader:
            images, captions = _to_device(images, device), _to_device(captions, device)
            lengths = [int(l) for l in lengths]
            targets = pack_padded_sequence(captions, lengths, batch_first=True)[0]          # eval.py:91
            feats = model.features(images) if hasattr(model, "features") else None
            if crit is not None:
                dec_in = (feats, captions, lengths) if feats is not None else (images, captions, lengths)
                outputs = (model.decoder if feats is not None else model)(*dec_in)          # eval.py:93
                loss = crit(outputs, targets)                                               # eval.py:95
            elif feats is not None:
                loss = model.decoder.loss(feats, captions, lengths, targets)
            else:
                loss = model.loss(images, captions, lengths, targets)
            loss_sum += float(loss)                                                         # eval.py:96-97
            loss_evals += 1
            ids = model.decoder.sample(feats) if feats is not None else model.sample(images)  # eval.py:99
            ids, kept = trim(ids.reshape(len(lengths), -1), end_id, pad_id)                 # eval.py:103-109 on device
            ids, kept = ids.cpu().numpy(), kept.cpu().numpy()                               # eval.py:101
            for i, imgid in enumerate(imgids):
                imgid = imgid.item() if hasattr(imgid, "item") else imgid
                if imgid in seen:                                                           # eval.py:112-116
                    continue
                seen.add(imgid)
                words = [vocab.idx2word[int(w)] for w in ids[i, :int(kept[i])]]
                predictions.append({"image_id": imgid, "caption": " ".join(words)})
    model.train(was_training)
    lang_stats = language_eval(predictions) if language_eval is not None else {}            # eval.py:117
    return loss_sum / max(loss_evals, 1), predictions, lang_stats


class Trainer(object):
    """train.py:20-199 on the B200 modules.  Extra keyword arguments exist for callers that already hold the pieces
    (tests, precomputed-feature pipelines): `vocab` (else unpickled from opt.vocab_path, train.py:33-34), `model` (else
    built from opt), `stepper` (else `parallel.DataParallelStep` over the model's pair), `language_eval` (hook)."""

    def __init__(self, opt, trainloader, validloader, vocab=None, model=None, stepper=None, language_eval=None,
                 backbone=None, precision="bf16"):
        self.opt = opt
        self.total_train_iter = len(trainloader)                      # train.py:24-25
        self.total_valid_iter = len(validloader) if validloader is not None else 0
        self.trainloader, self.validloader = trainloader, validloader
        self.num_gpu = opt.num_gpu
        if vocab is None:
            with open(opt.vocab_path, "rb") as f:
                vocab = pickle.load(f)
        self.vocab = vocab
        if model is None:
            model = CaptionModel(opt.embed_size, opt.hidden_size, len(vocab), opt.num_layers,
                                 backbone=bool(backbone), precision=precision)
            if self.num_gpu > 0:
                model.cuda()                                          # train.py:40-41 (one GPU per process)
        self.model = model
        load_path = getattr(opt, "load_model_path", None)
        if getattr(opt, "load_pretrained", False) and load_path:
            self.load_model(load_path)                                # train.py:46-48
        self.criterion = nn.CrossEntropyLoss()                        # train.py:53 (strict evaluation path)
        if stepper is None:
            from .parallel import DataParallelStep
            stepper = DataParallelStep(model.encoder, model.decoder, lr=opt.learning_rate, grad_clip=opt.grad_clip)
        self.stepper = stepper                                        # train.py:55-56,88-91,144-146
        self.language_eval = language_eval
        self.fused_eval = True
        self.last_loss = None

    # -- checkpoint glue the reference leaves empty (train.py:60-64) ------------------------------------------------
    def load_model(self, path):
        self.model.load_state_dict(torch.load(path, map_location="cpu"))

    def _infos_path(self, suffix=""):
        return os.path.join(self.opt.expr_dir, "infos" + suffix + ".pkl")

    def _load_infos(self):
        if getattr(self.opt, "start_from", None) is not None and not getattr(self.opt, "load_pretrained", False):
            with open(self._infos_path(), "rb") as f:                 # train.py:70-74
                return pickle.load(f)
        return {}

    def set_lr(self, lr):                                             # train.py:93-95
        self.opt.current_lr = lr
        self.stepper.lr = lr

    def train_step(self, images, captions, lengths):
        """train.py:123-146 for one loader batch -> the (device) loss of this rank.  The targets are
        pack(captions, lengths) (eval.py:91), gathered on the device by the step executor."""
        device = next(self.model.parameters()).device
        images, captions = _to_device(images, device), _to_device(captions, device)
        world = getattr(self.stepper, "world", 1)
        # several ranks: the SUM all-reduce of gradients scaled by 1/world = the average of the ranks' mean losses
        n_global = int(sum(int(l) for l in lengths)) * world if world > 1 else None
        return self.stepper.step(images, captions, lengths, None, n_global)

    def validate(self):
        crit = None if self.fused_eval else self.criterion
        return evaluation(self.model, crit, self.validloader, self.vocab, self.opt, self.language_eval)

    def train(self):
        opt = self.opt
        infos = self._load_infos()
        total_iteration = infos.get("total_iter", 0)                  # train.py:76-81
        loaded_iteration = infos.get("iter", 0)
        loaded_epoch = infos.get("epoch", 1)
        val_result_history = infos.get("val_result_history", {})
        loss_history = infos.get("loss_history", {})
        lr_history = infos.get("lr_history", {})
        best_val_score = infos.get("best_val_score", None) if getattr(opt, "load_best_score", True) else None
        self.model.train()
        for epoch in range(max(1, loaded_epoch), 1 + opt.max_epochs):
            self.set_lr(learning_rate_for_epoch(opt, epoch))
            for it, (images, captions, lengths, imgids) in enumerate(self.trainloader, start=1):
                if epoch == loaded_epoch and it <= loaded_iteration:
                    continue                                          # resume inside the saved epoch (train.py:119-121)
                total_iteration += 1
                start = time.time()
                loss = self.train_step(images, captions, lengths)
                self.last_loss = loss
                if it % opt.log_step == 0:                            # the only host<->device sync of a plain step
                    lv = float(loss)
                    print("Epoch [%d/%d], Step [%d/%d], Loss: %.4f, Perplexity: %5.4f, %.1f ms"
                          % (epoch, opt.max_epochs, it, self.total_train_iter, lv, math.exp(min(lv, 80.0)),
                             1e3 * (time.time() - start)))
                if self.validloader is None or total_iteration % opt.save_checkpoint_every != 0:
                    continue
                val_loss, predictions, lang_stats = self.validate()   # train.py:157-160
                val_result_history[total_iteration] = {"loss": val_loss, "lang_stats": lang_stats,
                                                       "predictions": predictions}
                loss_history[total_iteration] = float(loss)
                lr_history[total_iteration] = opt.current_lr
                if getattr(opt, "language_eval", 0) == 1 and "CIDEr" in lang_stats:
                    current_score = lang_stats["CIDEr"]               # train.py:172-175
                else:
                    current_score = -val_loss
                best_flag = best_val_score is None or current_score > best_val_score
                if best_flag:
                    best_val_score = current_score
                infos.update(total_iter=total_iteration, iter=it, epoch=epoch, best_val_score=best_val_score,
                             opt=dict(vars(opt)), val_result_history=val_result_history, loss_history=loss_history,
                             lr_history=lr_history)
                if self._is_writer():
                    os.makedirs(opt.expr_dir, exist_ok=True)
                    with open(self._infos_path(), "wb") as f:         # train.py:191-192
                        pickle.dump(infos, f)
                    if best_flag:                                     # train.py:194-199
                        torch.save(self.model.state_dict(), os.path.join(opt.expr_dir, "model-best.pth"))
                        print("model saved to {}".format(opt.expr_dir))
                        with open(self._infos_path("-best"), "wb") as f:
                            pickle.dump(infos, f)
            loaded_iteration = 0
        return infos

    @staticmethod
    def _is_writer():
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
