"""The encoder feed: what hands pooled 2048-d ResNet features to the hot path.

The reference runs the frozen ResNet-152 trunk inline, once per use of a batch (models.py:25-27; eval.py:93 and :99
even run it twice), from images the DataLoader decodes per step (data_loader.py:24-43).  The trunk has no trainable
parameter (models.py:14-15), so its output for an image never changes: north_star keeps it "on cuDNN as the baseline
feed, with features precomputed".  Three pieces, all host-side plumbing around torch/cuDNN (no kernel of ours):

  FeatureStore    the on-disk precomputed-feature format: `<dir>/pooled.npy` (a plain .npy, memory-mapped,
                  [n_images, dim] float32 or float16) + `<dir>/index.json` (image ids in row order, dim, dtype, the BN
                  mode the features were produced in).  `gather(image_ids)` fills a pinned float32 host batch.
  TrunkFeed       the frozen trunk as a producer: channels-last, bf16 autocast, `no_grad`, on a side stream, so
                  the trunk of batch i+1 overlaps the decoder step of batch i.  BN mode is explicit: "eval" (running
                  statistics; deterministic per image, what a precomputed store needs; the default) or "train"
                  (batch statistics — what the reference's trunk actually does during training, because
                  `model.train()` also flips the frozen BatchNorm2d layers).
  PrefetchLoader  wraps any loader yielding `(images, captions, lengths, imgids)` host batches: stages the next batch
                  in pinned memory and copies it (and, with a TrunkFeed, runs the trunk) on the side stream while
                  the current batch trains; yields device tensors in the same tuple shape, same order.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


class FeatureStore:
    """Precomputed pooled features on disk.  Rows are addressed by image id through `index.json`."""

    DATA, INDEX = "pooled.npy", "index.json"

    def __init__(self, path, array, image_ids, meta, writable):
        self.path, self.array, self.meta, self.writable = path, array, meta, writable
        self.image_ids = list(image_ids)
        self.row_of = {i: r for r, i in enumerate(self.image_ids)}
        if len(self.row_of) != len(self.image_ids):
            raise ValueError("FeatureStore: duplicate image ids")

    @classmethod
    def create(cls, path, image_ids, dim=2048, dtype="float32", bn_mode="eval"):
        if dtype not in ("float32", "float16"):
            raise ValueError("FeatureStore dtype must be float32 or float16")
        os.makedirs(path, exist_ok=True)
        image_ids = [i.item() if hasattr(i, "item") else i for i in image_ids]
        arr = np.lib.format.open_memmap(os.path.join(path, cls.DATA), mode="w+", dtype=np.dtype(dtype),
                                        shape=(len(image_ids), dim))
        meta = {"dim": dim, "dtype": dtype, "bn_mode": bn_mode, "format": 1}
        with open(os.path.join(path, cls.INDEX), "w") as f:
            json.dump(dict(meta, image_ids=image_ids), f)
        return cls(path, arr, image_ids, meta, True)

    @classmethod
    def open(cls, path):
        with open(os.path.join(path, cls.INDEX)) as f:
            meta = json.load(f)
        ids = meta.pop("image_ids")
        arr = np.load(os.path.join(path, cls.DATA), mmap_mode="r")
        if arr.shape != (len(ids), meta["dim"]) or str(arr.dtype) != meta["dtype"]:
            raise ValueError(f"FeatureStore at {path}: {arr.shape} {arr.dtype} does not match its index")
        return cls(path, arr, ids, meta, False)

    def __len__(self):
        return len(self.image_ids)

    def put(self, image_ids, pooled):
        """pooled: [k, dim] array or tensor (any float dtype, any device) for the given ids."""
        if not self.writable:
            raise RuntimeError("FeatureStore opened read-only")
        if isinstance(pooled, torch.Tensor):
            pooled = pooled.detach().float().cpu().numpy()
        rows = [self.row_of[i.item() if hasattr(i, "item") else i] for i in image_ids]
        self.array[rows] = np.asarray(pooled).astype(self.array.dtype)

    def flush(self):
        if self.writable:
            self.array.flush()

    def gather(self, image_ids, out=None):
        """-> float32 host tensor [k, dim] (pinned when CUDA is available) holding the rows of `image_ids`."""
        rows = np.fromiter((self.row_of[i.item() if hasattr(i, "item") else i] for i in image_ids), dtype=np.int64,
                           count=len(image_ids))
        if out is None:
            out = torch.empty(len(rows), self.meta["dim"], dtype=torch.float32,
                              pin_memory=torch.cuda.is_available())
        elif tuple(out.shape) != (len(rows), self.meta["dim"]) or out.dtype != torch.float32 or out.is_cuda:
            raise ValueError("FeatureStore.gather: `out` must be a float32 host tensor of shape [len(image_ids), dim]")
        order = np.argsort(rows, kind="stable")            # read the memory map in file order
        dst = out.numpy()
        dst[order] = self.array[rows[order]]               # float16 rows widen here, on the host
        return out


class TrunkFeed:
    """The frozen ResNet trunk of an `EncoderCNN(backbone=True)` as a side-stream producer of pooled features."""

    def __init__(self, encoder, dtype=torch.bfloat16, channels_last=True, bn_mode="eval", stream=None):
        if not getattr(encoder, "has_backbone", False):
            raise RuntimeError("TrunkFeed needs an EncoderCNN built with backbone=True")
        if bn_mode not in ("eval", "train"):
            raise ValueError("bn_mode must be 'eval' or 'train'")
        self.encoder, self.dtype, self.bn_mode = encoder, dtype, bn_mode
        self.channels_last = channels_last
        r = encoder.resnet
        self.layers = [r.conv1, r.bn1, r.relu, r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4, r.avgpool]
        if channels_last:
            for m in self.layers:
                m.to(memory_format=torch.channels_last)
        self.device = next(r.conv1.parameters()).device
        self.stream = stream if stream is not None else (torch.cuda.Stream(self.device) if self.device.type == "cuda"
                                                         else None)

    def _bn_layers(self):
        for m in self.layers:
            for sub in m.modules():
                if isinstance(sub, torch.nn.modules.batchnorm._BatchNorm):
                    yield sub

    @torch.no_grad()
    def pooled(self, images):
        """images [B,3,H,W] on the trunk's device -> pooled [B,2048] float32, on the CURRENT stream."""
        saved = [(m, m.training) for m in self._bn_layers()]
        for m, _ in saved:
            m.train(self.bn_mode == "train")
        try:
            x = images.contiguous(memory_format=torch.channels_last) if self.channels_last else images
            with torch.autocast(self.device.type, dtype=self.dtype, enabled=self.dtype != torch.float32):
                for m in self.layers:
                    x = m(x)
            return torch.flatten(x, 1).float()
        finally:
            for m, was in saved:
                m.train(was)

    def submit(self, images):
        """Start H2D (if `images` is a host tensor) + trunk on the side stream -> ticket for `result`."""
        if self.stream is None:
            return (self.pooled(images.to(self.device)), None)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))   # weights / earlier work are visible
        with torch.cuda.stream(self.stream):
            dev = images.to(self.device, non_blocking=True)
            out = self.pooled(dev)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return (out, ev)

    def result(self, ticket):
        out, ev = ticket
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            out.record_stream(cur)          # allocated on the side stream, consumed on this one
        return out

    def precompute(self, loader, store, log_every=0):
        """Fill a FeatureStore from a loader yielding `(images, captions, lengths, imgids)` (each image once)."""
        done = set()
        for k, (images, _, _, imgids) in enumerate(loader):
            ids = [i.item() if hasattr(i, "item") else i for i in imgids]
            new = [j for j, i in enumerate(ids) if i not in done]
            if new:
                pooled = self.result(self.submit(images))
                store.put([ids[j] for j in new], pooled[new])
                done.update(ids[j] for j in new)
            if log_every and (k + 1) % log_every == 0:
                print("precomputed %d images" % len(done))
        store.flush()
        return len(done)


class PrefetchLoader:
    """Iterates `loader` one batch ahead: batch i+1 is staged in pinned memory and copied to `device` (and run through
    `trunk`, when given) on a side stream while batch i is being consumed.  With `store`, the images of the loader
    are ignored and the pooled features of `imgids` come from the FeatureStore instead."""

    def __init__(self, loader, device, trunk=None, store=None, pin_thread=True, depth=2):
        """pin_thread: stage the host batches in pinned memory from a background thread, `depth` batches ahead (what
        DataLoader(pin_memory=True) does with its own thread).  Pinning 8 MB of pooled features is a 0.5-0.8 ms host copy:
        on the consumer thread it made the loop host-bound at 1.7 ms per 1.05 ms GPU step (profiles/r02_bench_a.json)."""
        self.loader, self.device = loader, torch.device(device)
        self.trunk, self.store = trunk, store
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.pin_thread = bool(pin_thread) and self.cuda
        self._dev_index = (self.device.index if self.device.index is not None else torch.cuda.current_device()) \
            if self.cuda else None
        self.depth = max(1, int(depth))

    def __len__(self):
        return len(self.loader)

    def _pin(self, t):
        t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
        return t.pin_memory() if self.cuda and not t.is_pinned() else t

    def _start(self, batch):
        images, captions, lengths, imgids = batch
        if self.store is not None:
            images = self.store.gather(imgids)
        if not self.cuda:
            if self.trunk is not None and self.store is None:
                images = self.trunk.pooled(images)
            return (images, captions, lengths, imgids, None, None)
        ticket = None
        if self.trunk is not None and self.store is None:
            ticket = self.trunk.submit(self._pin(images))
            images = None
        # No wait for the consumer's stream: the destination tensors are fresh allocations of the copy stream (the caching
        # allocator orders their reuse through record_stream in _finish), so the copy may start at once instead of at the
        # end of the step being computed - i.e. beside the head of the next step (profiles/r02_e2e_copy_window.txt).
        with torch.cuda.stream(self.stream):
            if images is not None:
                images = self._pin(images).to(self.device, non_blocking=True)
            captions = self._pin(captions).to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return (images, captions, lengths, imgids, ev, ticket)

    def _finish(self, staged):
        images, captions, lengths, imgids, ev, ticket = staged
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in (images, captions):
                if t is not None:
                    t.record_stream(cur)
        if ticket is not None:
            images = self.trunk.result(ticket)
        return images, captions, lengths, imgids

    def _pinned_batches(self):
        """The loader's batches with their tensors already pinned, produced `depth` ahead by a background thread (the
        copies into pinned memory release the GIL).  Exceptions of the loader are re-raised on the consumer side."""
        import queue
        import threading
        q = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        END = object()

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                torch.cuda.set_device(self._dev_index)   # pin_memory() needs this thread's current device
                for images, captions, lengths, imgids in self.loader:
                    if self.store is None and images is not None and self.trunk is None:
                        images = self._pin(images)
                    if not put((images, self._pin(captions), lengths, imgids)):
                        return
                put(END)
            except BaseException as e:   # noqa: BLE001 - handed to the consumer
                put(e)

        th = threading.Thread(target=work, name="snt-pin", daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is END:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()

    def __iter__(self):
        it = self._pinned_batches() if self.pin_thread else iter(self.loader)
        try:
            staged = self._start(next(it))
        except StopIteration:
            return
        for nxt in it:
            ahead = self._start(nxt)          # enqueue batch i+1 before handing out batch i
            yield self._finish(staged)
            staged = ahead
        yield self._finish(staged)
