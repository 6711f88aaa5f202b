"""Drop-in mirror of the reference's `models.py` (EncoderCNN, DecoderRNN): same constructor signatures,
submodule / parameter names, shapes and initialisation, so `state_dict()` round-trips with the reference
and `train.py` / `eval.py` can call these modules unchanged — but every forward/backward runs in the
hand-written sm_100a kernels behind the C ABI (ops.py -> _lib.py -> libsnt_b200.so), never in
torch.nn's own kernels.  The nn.Embedding / nn.LSTM / nn.Linear / nn.BatchNorm1d children are parameter
containers only.

Reference: /root/reference/models.py:9-29 (EncoderCNN), :31-67 (DecoderRNN).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _FcOnly(nn.Module):
    """Stand-in for the frozen ResNet when features are precomputed: only `.fc` (models.py:16) exists."""

    def __init__(self, in_features, embed_size):
        super().__init__()
        self.fc = nn.Linear(in_features, embed_size)


class EncoderCNN(nn.Module):
    """models.py:9-29.  `backbone=True` builds torchvision's ResNet-152 (random init: there is no network
    for the pretrained weights the reference downloads at models.py:13) frozen as at models.py:14-15 and run
    by cuDNN as the baseline feed; `backbone=False` keeps only the head for precomputed pooled features
    (`forward_pooled`).  The trainable head — resnet.fc Linear(2048,E) + BatchNorm1d(E, momentum=0.01) —
    always runs in the snt kernels."""

    def __init__(self, embed_size, backbone=True, precision="bf16"):
        super().__init__()
        if backbone:
            import torchvision.models as tvm
            self.resnet = tvm.resnet152(weights=None)
            for p in self.resnet.parameters():
                p.requires_grad = False                       # models.py:14-15
            self.resnet.fc = nn.Linear(self.resnet.fc.in_features, embed_size)  # models.py:16
        else:
            self.resnet = _FcOnly(2048, embed_size)
        self.bn = nn.BatchNorm1d(embed_size, momentum=0.01)   # models.py:17
        self.precision = precision
        self.has_backbone = bool(backbone)
        self.init_weights()

    def init_weights(self):
        self.resnet.fc.weight.data.normal_(0.0, 0.02)          # models.py:22
        self.resnet.fc.bias.data.fill_(0)                      # models.py:23

    def forward_pooled(self, pooled):
        """pooled[B,2048] (the ResNet's global-average-pooled activations) -> features[B,E]."""
        bn = self.bn
        if self.training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        return ops.head(pooled, self.resnet.fc.weight, self.resnet.fc.bias, bn.weight, bn.bias, bn.running_mean,
                        bn.running_var, self.training, bn.momentum, bn.eps, self.precision)

    def pooled(self, images):
        """images[B,3,224,224] -> the frozen trunk's pooled activations [B,2048] (models.py:27 without the fc),
        under no_grad (models.py:14-15)."""
        if not self.has_backbone:
            raise RuntimeError("EncoderCNN(backbone=False) has no CNN: call forward_pooled(pooled[B,2048])")
        r = self.resnet
        with torch.no_grad():
            x = r.maxpool(r.relu(r.bn1(r.conv1(images))))
            x = r.layer4(r.layer3(r.layer2(r.layer1(x))))
            return torch.flatten(r.avgpool(x), 1)

    def forward(self, images):
        """models.py:25-29."""
        return self.forward_pooled(self.pooled(images))


class DecoderRNN(nn.Module):
    """models.py:31-67."""

    def __init__(self, embed_size, hidden_size, vocab_size, num_layers, precision="bf16", sample_precision="fp32"):
        super().__init__()
        self.embed = nn.Embedding(vocab_size, embed_size)                            # models.py:35
        self.lstm = nn.LSTM(embed_size, hidden_size, num_layers, batch_first=True)   # models.py:36
        self.linear = nn.Linear(hidden_size, vocab_size)                             # models.py:37
        self.ss_prob = 0.0                                                           # models.py:38 (unused there too)
        self.num_layers = num_layers
        self.precision = precision                # teacher-forced path: "bf16" (tcgen05) or "fp32" (faithful)
        self.sample_precision = sample_precision  # greedy decode defaults to the token-exact fp32 mode
        self.init_weights()

    def init_weights(self):
        self.embed.weight.data.uniform_(-0.1, 0.1)       # models.py:43
        self.linear.weight.data.uniform_(-0.1, 0.1)      # models.py:44
        self.linear.bias.data.fill_(0)                   # models.py:45

    def _lstm_weights(self):
        return [(getattr(self.lstm, f"weight_ih_l{k}"), getattr(self.lstm, f"weight_hh_l{k}"),
                 getattr(self.lstm, f"bias_ih_l{k}"), getattr(self.lstm, f"bias_hh_l{k}"))
                for k in range(self.num_layers)]

    def forward(self, features, captions, lengths):
        """Decode image feature vectors and generate captions (models.py:47-54).
        -> logits[N,V] over the packed (time-major) rows, N = sum(lengths)."""
        return ops.decoder_logits(features, captions, lengths, self.embed.weight, self._lstm_weights(),
                                  self.linear.weight, self.linear.bias, self.precision)

    def loss(self, features, captions, lengths, targets, grad_scale=1.0):
        """criterion(self(features, captions, lengths), targets) of train.py:139-143 as one fused op: the vocab
        projection is fused with log-softmax + cross-entropy and the logits never reach HBM as a whole."""
        return ops.decoder_loss(features, captions, lengths, targets, self.embed.weight, self._lstm_weights(),
                                self.linear.weight, self.linear.bias, self.precision, grad_scale)

    def sample(self, features, states=None, precision=None):
        """Samples captions for given image features (Greedy search), models.py:56-67: always 20 steps,
        no <end> early stop; returns int64 [B,20] ([20] when B == 1, the reference's .squeeze())."""
        ids = ops.greedy(features, self.embed.weight, self._lstm_weights(), self.linear.weight, self.linear.bias,
                         states, ops.SAMPLE_STEPS, precision or self.sample_precision)
        return ids.squeeze()

    def sample_trimmed(self, features, states=None, precision=None, end_id=2, pad_id=0):
        """sample() followed by eval.py:101-109's "words before the first <end>" rule on the device:
        -> (ids[B,20] with everything from the first <end> on replaced by <pad>, lengths[B] int32)."""
        ids = ops.greedy(features, self.embed.weight, self._lstm_weights(), self.linear.weight, self.linear.bias,
                         states, ops.SAMPLE_STEPS, precision or self.sample_precision)
        return ops.trim_captions(ids, end_id, pad_id)
