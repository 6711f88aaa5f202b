#!/usr/bin/env python
"""bench.py — the caption-decoder hot path (BASELINE.json) on N B200s; rank 0 prints ONE JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config default|scaled]
                  [--scaling weak|strong] [--prec bf16|fp32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one synthetic COCO-shaped batch: encoder head on
precomputed pooled 2048-d features -> gather/concat/pack -> LSTM -> vocab Linear fused with log-softmax + CE ->
BPTT -> (N>1: gradient all-reduce overlapped with BPTT) -> clip_gradient + Adam, through the native step executor
(parallel.DataParallelStep -> snt_step_run), eager launches, no CUDA graph.

value   : train captions/s with the inputs resident in HBM, CUDA events, max over ranks.  The timed region cycles over
          NB = 8 DISTINCT ragged batches (different lengths, tokens and features each), so nothing depends on one
          replayed batch signature; `fixed_batch` repeats batch 0 for comparison.
e2e     : the same loop with HOST (pinned) inputs: every step pays the H2D copies of its own pooled features and
          captions (three slots, on a side stream) and a D2H read of its loss inside the timed region.
roofline: the stage with the largest share of the step, timed by CUDA events inside snt_step_run (snt_step_profile);
          `traffic` is read from the ncu capture committed under profiles/ (profiles/r02_traffic.json), never a literal.
greedy  : the second half of BASELINE's metric - greedy sample() of 4096 image features per GPU x 20 tokens
          (configs[2]) - as an object with its own roofline, cpu_baseline, e2e (features from the host, ids back) and
          the torch/cuDNN arm on the same GPU.
configs0: BASELINE configs[0] - the reference's own CPU case, its whole model with the ResNet-152 trunk, batch 128: timed on
          the host cores by `--impl reference` (`configs0`) and on the GPU as `f_rows.configs0` (trunk on cuDNN + native step).
configs3: the scaled decoder (E512/H1024/L2/V32000, batch 2048) timed the same way (N=1), also as `--config scaled`.
strong_8192 (N>1): BASELINE configs[4] as written - global batch 8192 sharded over the N ranks.
cpu_baseline / --impl reference: the reference's own models.py (oracle/_ref, placed there unmodified by build(); kind
          "reference") - or, when that is absent, its pinned restatement oracle/torch_port.py (kind "port") - timed on
          this box's host cores.
gpu_torch_reference: the same torch.nn composition on THIS GPU (cuDNN LSTM, cuBLAS; fp32, TF32, bf16 autocast), timed
          in a child process after the timed regions - the bar to beat.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CONFIGS = {
    "default": dict(B=1024, E=256, H=512, V=10000, L=1, POOLED=2048, name="BASELINE configs[1]"),
    "scaled": dict(B=2048, E=512, H=1024, V=32000, L=2, POOLED=2048, name="BASELINE configs[3]"),
}
CFG = CONFIGS["default"]
CONFIGS0 = dict(B=128, E=512, H=1024, V=10000, L=1, POOLED=2048, name="BASELINE configs[0]")   # config.py:17,27-29
CONFIGS0_WHAT = ("BASELINE configs[0]: frozen ResNet-152 trunk -> Linear+BN head -> 1-layer LSTM decoder at config.py's sizes "
                 "(embed 512 / hidden 1024, 10k vocabulary), teacher-forced train step, batch 128 synthetic 224x224 images, "
                 "captions <= 20 tokens")
GREEDY_B = 4096                                                       # BASELINE.json configs[2]
STRONG_GLOBAL_B = 8192                                                # BASELINE.json configs[4]
NB = 8                                                                # distinct ragged batches in the timed region
EXTRA_WARMUP_MULTI_GPU = 60                                           # untimed steps added to --warmup when N > 1


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def m_tok(c):
    """MAC per token: sum_k 4H(in_k + H) + H V."""
    return sum(4 * c["H"] * ((c["E"] if k == 0 else c["H"]) + c["H"]) for k in range(c["L"])) + c["H"] * c["V"]


def train_flops(B, N, c=CFG):
    """Algorithmic FLOPs of one train step: 3 * 2 * (N*M_tok + B*M_head)."""
    return 6.0 * (N * m_tok(c) + B * c["POOLED"] * c["E"])


def greedy_flops_per_token(c=CFG):
    return 2.0 * m_tok(c)


class ClockSampler:
    """SM clock / power / throttle reasons of this rank's GPU, sampled DURING the timed region.

    In-process NVML (the library behind nvidia-smi) from a background thread, every 100 ms: a `nvidia-smi -lms 50`
    subprocess per rank was measured to stretch a 2-GPU step from 1.56 ms to 7.4 ms (its query loop contends with
    kernel launches), which made the sampled run unrepresentative.  Falls back to one `nvidia-smi -lms 200`
    subprocess when pynvml is missing."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.1

    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []          # (sm_mhz, max_mhz, power_w, reasons bitmask)
        self.thread = None
        self.stop_flag = False
        self.p = None
        self.f = None
        self.source = None

    def _nvml_loop(self, nv, h):
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), float(pw), int(rs)))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
        return nv, nv.nvmlDeviceGetHandleByIndex(phys)

    def open_manual(self):
        """N > 1: no polling thread (see bench main); sample_once() is called while the GPU is under load."""
        try:
            self._nv, self._h = self._nvml()
            self._mx = self._nv.nvmlDeviceGetMaxClockInfo(self._h, self._nv.NVML_CLOCK_SM)
            self.thread = "manual"
            self.source = "nvml (in-process, on-demand samples right after the timed region)"
        except Exception:
            self.thread = None

    def sample_once(self):
        if self.thread != "manual":
            return
        nv, h = self._nv, self._h
        try:
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(self._mx),
                              nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(rs)))
        except Exception:
            pass

    def start(self):
        try:
            import threading
            nv, h = self._nvml()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            self.source = "nvml (in-process, 100 ms)"
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi -lms 200"
        except OSError:
            self.p = None

    def stop(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        sm, mx, pw, reasons = [], [], [], set()
        if self.thread is not None:
            self.stop_flag = True
            if self.thread != "manual":
                self.thread.join(timeout=2)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            for a, b, c, r in self.rows:
                sm.append(a); mx.append(b); pw.append(c)
                for n in names:
                    if r & bits[n]:
                        reasons.add(n)
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
            self.f.flush()
            rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
            os.unlink(self.f.name)
            for r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(names, r[4:8]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        loaded = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw)), "source": self.source}


# ---------------------------------------------------------------------------------------------------------------------
# reference arms
# ---------------------------------------------------------------------------------------------------------------------
def cpu_arm():
    """(module, kind, description) of the CPU arm: the reference's own models.py when build() placed it under
    oracle/_ref (it travels to the GPU box with the snapshot), else the pinned restatement oracle/torch_port.py."""
    from oracle import ref_arm as RA     # test / bench infrastructure, never shipped
    if RA.available():
        return RA, "reference", RA.what()
    from oracle import torch_port as TP
    return TP, "port", (f"torch {torch.__version__} CPU, oracle/torch_port.py pinned to the reference's goldens "
                        "(oracle/_ref absent)")


def cpu_train_baseline(c, steps, warmup, threads):
    """The reference's CPU step (head + decoder fwd, CE, bwd, clip, Adam) on the host cores:
    (captions/s, s/step, loss, tokens, kind, description)."""
    import show_and_tell_b200 as snt
    arm, kind, what = cpu_arm()
    b = snt.synthetic.make_batch(c["B"], c["V"], embed=c["E"], seed=1, pooled_dim=c["POOLED"])
    b["targets"] = snt.synthetic.pack_host(b["captions"], b["lengths"])
    cps, dt, loss = arm.time_full_train(c["B"], c["E"], c["H"], c["V"], c["L"], b, steps=steps, warmup=warmup,
                                        threads=threads)
    return cps, dt, loss, int(sum(b["lengths"])), kind, what


def cpu_greedy_baseline(c, batch, threads):
    """The reference's greedy loop on a bounded sample of the decode batch: (tokens/s, s per sample, kind, description)."""
    arm, kind, what = cpu_arm()
    feats = np.random.default_rng(1).standard_normal((batch, c["E"])).astype(np.float32)
    tps, dt = arm.time_greedy(batch, c["E"], c["H"], c["V"], c["L"], feats, steps=2, warmup=1, threads=threads)
    return tps, dt, kind, what


def run_reference(args, rank):
    """The reference's own CPU composition, all host threads, same config/metric.  Rank 0 only."""
    if rank != 0:
        return
    c = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    steps, warm = max(1, args.steps), max(1, args.warmup)
    cps, dt, loss, n_tok, kind, what = cpu_train_baseline(c, steps, warm, threads)
    g_tps, g_dt, _, _ = cpu_greedy_baseline(c, 512, threads)
    configs0 = None
    if kind == "reference" and not args.no_extras and args.config == "default":
        # BASELINE configs[0], the reference's own CPU-runnable case: its whole model (ResNet-152 trunk included) on the
        # host cores; the GPU counterpart is f_rows.configs0 of the native line
        try:
            arm, _, _ = cpu_arm()
            ips, dt0, loss0 = arm.time_configs0(threads=threads, steps=1, warmup=0)
            configs0 = {"metric": "train_images_per_s", "value": ips, "unit": "images/s", "s_per_step": dt0, "loss": loss0,
                        "config": CONFIGS0_WHAT, "sample": "ONE step of batch 128 (no warm-up step: a step takes ~10-20 s)"}
        except Exception as e:   # noqa: BLE001
            configs0 = {"error": repr(e)[:300]}
    line = {"impl": "reference", "metric": "train_captions_per_s", "value": cps, "unit": "captions/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(c, c["B"], 1, "fp32", n_tok, 20, device="cpu", threads=threads),
            "cpu_baseline": {"value": cps, "unit": "captions/s", "cores": threads, "kind": kind,
                             "sample": f"{steps} full steps of B={c['B']} (N={n_tok} tokens), {what}"},
            "e2e": {"value": cps, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "greedy": {"metric": "greedy_decode_tokens_per_s", "value": g_tps, "unit": "tokens/s",
                       "sample": f"2 x sample() of 512 image features x 20 tokens ({g_dt:.2f} s each), {what}"},
            "configs0": configs0, "loss": loss}
    print(json.dumps(line), flush=True)


def workload_config(c, b_local, world, prec, n_tok, max_len, **extra):
    d = {"workload": f"decoder train step on precomputed {c['POOLED']}-d features: head(Linear+BN) + embed/pack + "
                     f"{c['L']}-layer LSTM + fused vocab-CE + BPTT + clip/Adam; E{c['E']}/H{c['H']}/V{c['V']}/L{c['L']}, "
                     f"batch {b_local} per GPU ({c['name']})",
         "global_batch": b_local * world, "tokens_per_rank": n_tok, "max_len": max_len, "parallelism": f"dp{world}",
         "precision_mode": prec}
    d.update(extra)
    return d


def gpu_torch_reference_child(device_index):
    """Child process of `gpu_torch_reference`: prints one JSON object.  Self-contained on purpose: the composition of
    torch.nn layers the reference runs (models.py:9-67, train.py:137-146) is written out here, so this leg neither
    imports the test oracle nor touches the product's kernels."""
    import torch.nn as nn
    from torch.nn.utils.rnn import pack_padded_sequence
    import show_and_tell_b200 as snt          # synthetic batch generator only (numpy, host side)
    c = CFG
    dev = torch.device("cuda", device_index)
    torch.cuda.set_device(dev)
    b = snt.synthetic.make_batch(c["B"], c["V"], embed=c["E"], seed=1, pooled_dim=c["POOLED"])
    lengths = b["lengths"]
    pooled = torch.from_numpy(b["pooled"]).to(dev)
    caps = torch.from_numpy(b["captions"]).to(dev)
    targets = torch.from_numpy(snt.synthetic.pack_host(b["captions"], lengths)).to(dev)

    class Pair(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(c["POOLED"], c["E"])                              # models.py:16
            self.bn = nn.BatchNorm1d(c["E"], momentum=0.01)                       # models.py:17
            self.embed = nn.Embedding(c["V"], c["E"])                             # models.py:35
            self.lstm = nn.LSTM(c["E"], c["H"], c["L"], batch_first=True)         # models.py:36
            self.linear = nn.Linear(c["H"], c["V"])                               # models.py:37

        def forward(self, pooled, captions, lengths):
            feats = self.bn(self.fc(pooled))                                      # models.py:27-28
            steps = torch.cat((feats.unsqueeze(1), self.embed(captions)), 1)      # models.py:49-50
            hiddens, _ = self.lstm(pack_padded_sequence(steps, lengths, batch_first=True))   # models.py:51-52
            return self.linear(hiddens[0])                                        # models.py:53

        @torch.no_grad()
        def sample(self, features):                                               # models.py:56-67
            x, states, out = features.unsqueeze(1), None, []
            for _ in range(20):
                h, states = self.lstm(x, states)
                tok = self.linear(h.squeeze(1)).max(1, keepdim=True)[1]
                out.append(tok)
                x = self.embed(tok)
            return torch.cat(out, 1)

    out = {"unit": "captions/s",
           "what": "torch.nn composition of the same train step (head, decoder, CE, backward, clip, Adam) on this GPU, "
                   f"torch {torch.__version__}: cuDNN LSTM, cuBLAS, ATen kernels; CUDA events, 10 steps after 3 warm-up; "
                   "greedy: the reference's 20-step sample() loop on 4096 features"}
    gfeat = torch.randn(GREEDY_B, c["E"], device=dev)
    for name, tf32, amp in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
            torch.manual_seed(0)
            model = Pair().to(dev)
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)
            crit = nn.CrossEntropyLoss()

            def one():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    loss = crit(model(pooled, caps, lengths), targets)            # train.py:139-143
                loss.backward()                                                   # train.py:144
                for p in model.parameters():
                    p.grad.clamp_(-0.1, 0.1)                                      # train.py:88-91,145
                opt.step()                                                        # train.py:146
                return loss

            for _ in range(3):
                one()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                loss = one()
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3 / 10
            out[name] = {"value": c["B"] / dt, "ms_per_step": dt * 1e3, "loss": float(loss.detach())}
            model.eval()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                model.sample(gfeat)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    model.sample(gfeat)
                e1.record()
            torch.cuda.synchronize()
            out[name]["greedy_tokens_per_s"] = GREEDY_B * 20 * 3 / (e0.elapsed_time(e1) * 1e-3)
            del model, opt
        except Exception as e:   # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}
    print(json.dumps(out), flush=True)


def gpu_torch_reference(device_index, timeout_s=300):
    """the reference's own composition of torch.nn layers timed on THIS GPU (no kernel of ours),
    reported next to the bench line.  Runs in a child process after the timed regions, so that neither a crash nor a
    hang of that foreign code path can cost the bench line; any failure becomes an "error" string."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "torch-gpu", "--gpus", "1",
                            "--device-index", str(device_index)], capture_output=True, text=True, timeout=timeout_s)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": f"child exited {r.returncode}: {r.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:300]}


# ---------------------------------------------------------------------------------------------------------------------
# per-stage rooflines
# ---------------------------------------------------------------------------------------------------------------------
def stage_rooflines(prof_us, n_tok, peaks, c=CFG, b_local=None):
    """prof_us: {stage: microseconds per step} (snt_step_profile events, mean over real steps).  Algorithmic work per
    stage as SURVEY.md §8(d) counts it: no credit for recomputation, dgrad + wgrad = 2x the forward contraction."""
    B = b_local or c["B"]
    E, H, V, K, L = c["E"], c["H"], c["V"], c["POOLED"], c["L"]
    N = n_tok
    lstm_mac = sum(4 * H * ((E if k == 0 else H) + H) for k in range(L))
    n_par = V * E + sum(4 * H * ((E if k == 0 else H) + H) + 8 * H for k in range(L)) + V * H + V + E * K + 3 * E
    work = {   # stage -> (bound, algorithmic FLOPs or bytes per step)
        "vocab_ce_fwd": ("tensor", 2.0 * N * V * H),
        "vocab_ce_bwd": ("tensor", 4.0 * N * V * H),
        "lstm_fwd": ("tensor", 2.0 * N * lstm_mac),
        "lstm_bwd": ("tensor", 4.0 * N * lstm_mac),
        "head_fwd": ("tensor", 2.0 * B * K * E),
        "head_bwd": ("tensor", 2.0 * B * K * E),
        "embed_pack_fwd": ("hbm", 4.0 * N * E + 2.0 * N * E),                # read fp32 rows, write bf16 rows
        "embed_pack_bwd": ("hbm", 4.0 * N * E + 4.0 * V * E),                # read dx, write dense d_w_emb
        "clamp_adam": ("hbm", 28.0 * n_par),                                 # read p,g,m,v; write p,m,v
    }
    out = []
    for name, us in sorted(prof_us.items(), key=lambda kv: -kv[1]):
        ent = {"stage": name, "us_per_step": us}
        if name in work and us > 0:
            bound, w = work[name]
            if bound == "tensor":
                ach, peak, unit = w / (us * 1e-6) / 1e12, peaks["tf_sust"], "TFLOP/s"
            else:
                ach, peak, unit = w / (us * 1e-6) / 1e9, peaks["hbm"], "GB/s"
            ent.update({"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                        "algorithmic_work": w})
        out.append(ent)
    return out


def load_traffic():
    """profiles/r02_traffic.json: DRAM bytes per stage of one training step from the committed `ncu --set full` capture of
    this tree (written by profiles/ncu_traffic.py from the capture's raw page)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except ValueError:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """Model pair + stepper + NB distinct ragged batches (host pinned and device resident) for one configuration."""

    def __init__(self, snt, c, b_local, world, rank, dev, prec, nb=NB):
        from show_and_tell_b200 import parallel
        self.c, self.b_local, self.world, self.dev = c, b_local, world, dev
        torch.manual_seed(0)
        self.enc = snt.EncoderCNN(c["E"], backbone=False, precision=prec).to(dev).train()
        self.dec = snt.DecoderRNN(c["E"], c["H"], c["V"], c["L"], precision=prec).to(dev).train()
        self.stepper = parallel.DataParallelStep(self.enc, self.dec)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        self.batches = []
        for k in range(nb):
            gb = snt.synthetic.make_batch(b_local * world, c["V"], embed=c["E"], seed=1 + k, pooled_dim=c["POOLED"])
            sh = parallel.shard_batch(gb, world, rank)
            hp, hc = pin(sh["pooled"]), pin(sh["captions"])
            self.batches.append(dict(lengths=np.asarray(sh["lengths"], dtype=np.int64), n_tok=int(sum(sh["lengths"])),
                                     n_glob=sh["n_tokens_global"], hp=hp, hc=hc, dp=hp.to(dev), dc=hc.to(dev)))
        self.i = 0

    def step_resident(self, fixed=False):
        b = self.batches[0 if fixed else self.i % len(self.batches)]
        self.i += 1
        return self.stepper.step(b["dp"], b["dc"], b["lengths"], None, b["n_glob"] if self.world > 1 else None)

    def mean_tokens(self):
        return float(np.mean([b["n_tok"] for b in self.batches]))


class E2E:
    """Host-input loop: step i's H2D copies are issued on a side stream while step i-1 computes (three slots), and
    the loss of step i-1 is read back (pinned D2H + event) while step i runs; every step still pays its own copies and
    its own loss read inside the timed region."""

    def __init__(self, wl):
        self.wl = wl
        dev = wl.dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        tc = max(b["hc"].shape[1] for b in wl.batches)
        self.bufs = [dict(p=torch.empty_like(wl.batches[0]["dp"]),
                          c=torch.empty(wl.b_local, tc, dtype=torch.int64, device=dev),
                          ev=torch.cuda.Event(), done=torch.cuda.Event()) for _ in range(3)]
        self.loss_host = torch.zeros(1).pin_memory()
        self.loss_ev = torch.cuda.Event()
        self.i, self.pending, self.last = 0, False, float("nan")
        self.h2d = int(np.mean([b["hp"].numel() * 4 + b["hc"].numel() * 8 for b in wl.batches]))

    def _issue(self, k):
        b, src = self.bufs[k % 3], self.wl.batches[k % len(self.wl.batches)]
        with torch.cuda.stream(self.copy_stream):
            # Three slots: the slot's last user is step k - 3, finished long ago, so the copy never waits and is not tied to a
            # step boundary.  With two slots it could only start when step k - 2 ended, i.e. it ran beside the head of the
            # next step, and that loop measured 40-90 us per step slower (profiles/r02_e2e_copy_window.txt).
            self.copy_stream.wait_event(b["done"])
            b["p"].copy_(src["hp"], non_blocking=True)
            b["c"][:, :src["hc"].shape[1]].copy_(src["hc"], non_blocking=True)
            b["ev"].record(self.copy_stream)

    def step(self):
        wl, i = self.wl, self.i
        b, src = self.bufs[i % 3], wl.batches[i % len(wl.batches)]
        if i == 0:
            self._issue(0)
        self._issue(i + 1)                                  # prefetch the next step's inputs
        cur = torch.cuda.current_stream()
        cur.wait_event(b["ev"])
        loss = wl.stepper.step(b["p"], b["c"][:, :src["hc"].shape[1]], src["lengths"], None,
                               src["n_glob"] if wl.world > 1 else None)
        b["done"].record(cur)
        if self.pending:
            self.loss_ev.synchronize()                      # previous step's loss has landed on the host
            self.last = float(self.loss_host[0])
        self.loss_host.copy_(loss.reshape(1), non_blocking=True)
        self.loss_ev.record(cur)
        self.pending = True
        self.i = i + 1
        return self.last

    def final_loss(self):
        self.loss_ev.synchronize()
        return float(self.loss_host[0])


def profile_stages(wl, nsteps=10):
    """Mean device time per stage over `nsteps` real steps (CUDA events inside snt_step_run) + the optimizer launch."""
    eng = wl.stepper.engine
    acc = {}
    eng.profile(True)
    for _ in range(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        opt = wl.stepper.optimizer
        wl.stepper.optimizer = False
        wl.step_resident()
        wl.stepper.optimizer = opt
        prof = eng.profile_read()
        e0.record()
        wl.stepper.update()      # clip + Adam; several ranks: including the gradient exchange
        e1.record()
        torch.cuda.synchronize()
        prof["clamp_adam"] = e0.elapsed_time(e1)
        for k, v in prof.items():
            acc[k] = acc.get(k, 0.0) + v * 1e3 / nsteps
    eng.profile(False)
    return acc


def measure_greedy(snt, dec, c, dev, world, timed, peaks, rank, with_cpu):
    """BASELINE configs[2]: greedy sample() of 4096 image features per GPU x 20 tokens, every rank its own batch (the
    batch shards with no communication), device-timed, max over ranks, tokens of all ranks counted."""
    out = {"metric": "greedy_decode_tokens_per_s", "unit": "tokens/s", "higher_is_better": True,
           "config": {"workload": f"DecoderRNN.sample(): 20 greedy steps (LSTM step -> vocab Linear -> first-index argmax "
                                  f"-> embed), E{c['E']}/H{c['H']}/V{c['V']}/L{c['L']}, batch {GREEDY_B} image features per "
                                  "GPU (BASELINE configs[2])", "batch_per_gpu": GREEDY_B, "steps": 20}}
    feats_h = torch.randn(GREEDY_B, c["E"]).pin_memory()
    feats = feats_h.to(dev)
    dec.eval()
    reps = 5
    for prec in ("bf16", "fp32"):
        dec.sample(feats, precision=prec)
        t = timed(lambda: dec.sample(feats, precision=prec), reps)
        tps = GREEDY_B * world * 20 * reps / t
        if prec == "bf16":
            out["value"], out["dtype"], out["ms_per_decode"] = tps, "bf16", t / reps * 1e3
            tf = tps / world * greedy_flops_per_token(c) / 1e12
            out["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                               "frac": tf / peaks["tf_sust"], "traffic": None,
                               "algorithmic_work_per_token": greedy_flops_per_token(c),
                               "kernel": "per step: gemm_tc<128,LstmFwdEpi> ([x|h] concatenated operand) + "
                                         "gemm_tc<256,ArgmaxEpi> (logits stay in TMEM) + argmax_finish",
                               "timing": f"CUDA events around {reps} whole sample() calls (20 steps each)"}
        else:
            out["fp32_faithful"] = {"value": tps, "ms_per_decode": t / reps * 1e3,
                                    "note": "token-exact mode: fp32 operands, each product as six exact bf16 partial products on the tensor cores, fp32 accumulation (csrc/gemm_x3.cu)"}
    # e2e: features from pinned host memory, ids back to the host, inside the timed region
    ids_h = torch.empty(GREEDY_B, 20, dtype=torch.int64).pin_memory()

    def one_e2e():
        f = feats_h.to(dev, non_blocking=True)
        ids_h.copy_(dec.sample(f, precision="bf16"), non_blocking=True)
    one_e2e()
    t = timed(one_e2e, reps)
    out["e2e"] = {"value": GREEDY_B * world * 20 * reps / t, "unit": "tokens/s",
                  "h2d_bytes_per_step": GREEDY_B * c["E"] * 4, "d2h_bytes_per_step": GREEDY_B * 20 * 8}
    dec.train()
    if with_cpu and rank == 0:
        threads = os.cpu_count() or 1
        tps, dt, kind, what = cpu_greedy_baseline(c, 512, threads)
        out["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": threads, "kind": kind,
                               "sample": f"2 x sample() of 512 image features x 20 tokens ({dt:.2f} s each) after 1 warm-up, "
                                         f"{what}"}
    return out


def measure_f_rows(snt, dev, c, peaks):
    """Timings of the rows widened after the hot path, one GPU, after the timed regions; each guarded:
    a failure here becomes an "error" string and never costs the bench line.
      caption_trim : snt_caption_trim on a batch larger than L2 (4 M captions x 20 ids), CUDA events, HBM roofline.
      trainer_loop : show_and_tell_b200.Trainer over 24 DISTINCT ragged host batches of 1024 captions fed by
                     PrefetchLoader (pinned staging + side-stream H2D): wall clock around the whole loop including a
                     final synchronize - what train.py's loop sees.
      trunk_feed   : the frozen ResNet-152 trunk (cuDNN, channels-last bf16, side stream) feeding the decoder step:
                     trunk alone, decoder alone, and both overlapped (feed.TrunkFeed), images/s."""
    out = {}
    try:
        res = {}
        for name, B in (("large", 4 << 20), ("configs2", GREEDY_B)):
            ids = torch.randint(3, c["V"], (B, 20), device=dev)
            ids[torch.rand(B, 20, device=dev) < 0.05] = 2
            lengths = torch.empty(B, dtype=torch.int32, device=dev)
            o = torch.empty_like(ids)
            L = snt._lib
            f = lambda: L.call("snt_caption_trim", L.ptr(ids), B, 20, 2, 0, L.ptr(lengths), L.ptr(o), L.stream_ptr())
            f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                f()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            nbytes = B * 20 * 16 + B * 4
            res[name] = {"captions": B, "us": us, "achieved": nbytes / (us * 1e-6) / 1e9, "unit": "GB/s",
                         "peak": peaks["hbm"], "frac": nbytes / (us * 1e-6) / 1e9 / peaks["hbm"],
                         "algorithmic_bytes": nbytes}
            del ids, o, lengths
        out["caption_trim"] = res
    except Exception as e:   # noqa: BLE001
        out["caption_trim"] = {"error": repr(e)[:300]}
    try:
        import argparse as _ap
        from show_and_tell_b200.feed import PrefetchLoader
        nb = 24
        batches = []
        for k in range(nb):
            b = snt.synthetic.make_batch(c["B"], c["V"], seed=100 + k, pooled_dim=c["POOLED"])
            batches.append((torch.from_numpy(b["pooled"]), torch.from_numpy(b["captions"]), b["lengths"],
                            list(range(k * c["B"], (k + 1) * c["B"]))))
        torch.manual_seed(0)
        model = snt.CaptionModel(c["E"], c["H"], c["V"], c["L"], backbone=False, precision="bf16").to(dev)

        class _V:                                   # len() is all Trainer needs without a validation loader
            def __len__(self):
                return c["V"]
        opt = _ap.Namespace(num_gpu=1, embed_size=c["E"], hidden_size=c["H"], num_layers=c["L"], learning_rate=1e-3,
                            max_epochs=1, learning_rate_decay_start=1, learning_rate_decay_every=3,
                            learning_rate_decay_rate=0.8, grad_clip=0.1, log_step=10 ** 9, language_eval=0,
                            save_checkpoint_every=10 ** 9, expr_dir=tempfile.gettempdir(), start_from=None,
                            load_best_score=True, load_pretrained=False, load_model_path=None, vocab_path=None)
        tr = snt.Trainer(opt, PrefetchLoader(batches, dev), None, vocab=_V(), model=model)
        tr.train()                                  # warm-up epoch (allocator, workspaces)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.train()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["trainer_loop"] = {"value": nb * c["B"] / dt, "unit": "captions/s", "ms_per_step": dt / nb * 1e3,
                               "batches": nb, "loss": float(tr.last_loss),
                               "what": "Trainer.train() over distinct ragged host batches through PrefetchLoader, "
                                       "wall clock incl. H2D"}
        del tr, model, batches
        torch.cuda.empty_cache()
    except Exception as e:   # noqa: BLE001
        out["trainer_loop"] = {"error": repr(e)[:300]}
    try:
        out["trunk_feed"] = measure_trunk_feed(snt, dev, c)
    except Exception as e:   # noqa: BLE001
        out["trunk_feed"] = {"error": repr(e)[:300]}
    # BASELINE configs[0] (the reference's CPU case) on the GPU: the same whole-model step, trunk on cuDNN, head and decoder
    # through the native step; `serial_images_per_s` is the number that compares with the reference arm's `configs0`.
    # In a child process: a size no other leg of this line runs must not be able to cost the line.
    out["configs0"] = configs0_gpu(dev.index or 0)
    return out


def configs0_gpu_child(device_index):
    import show_and_tell_b200 as snt
    dev = torch.device("cuda", device_index)
    torch.cuda.set_device(dev)
    r = measure_trunk_feed(snt, dev, CONFIGS0)
    r["config"] = CONFIGS0_WHAT
    print(json.dumps(r), flush=True)


def configs0_gpu(device_index, timeout_s=240):
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "configs0-gpu", "--gpus", "1",
                            "--device-index", str(device_index)], capture_output=True, text=True, timeout=timeout_s)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": f"child exited {r.returncode}: {r.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:300]}


def measure_trunk_feed(snt, dev, c, batch=128, iters=6):
    """f3: the frozen ResNet-152 trunk as the producer of the hot path's input (batch 128 images of 3x224x224, the
    reference's default batch, config.py:17): images/s of (a) the trunk alone, (b) the decoder step alone on its pooled
    features, (c) TrunkFeed on a side stream overlapped with the decoder step of the previous batch."""
    from show_and_tell_b200 import parallel
    from show_and_tell_b200.feed import TrunkFeed
    torch.manual_seed(0)
    enc = snt.EncoderCNN(c["E"], backbone=True, precision="bf16").to(dev).train()
    dec = snt.DecoderRNN(c["E"], c["H"], c["V"], c["L"], precision="bf16").to(dev).train()
    st = parallel.DataParallelStep(enc, dec)
    feed = TrunkFeed(enc, bn_mode="eval")
    b = snt.synthetic.make_batch(batch, c["V"], seed=7)
    caps = torch.from_numpy(b["captions"]).to(dev)
    imgs = [torch.randn(batch, 3, 224, 224, device=dev) for _ in range(2)]

    def clock(fn, n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    pooled = feed.pooled(imgs[0])
    st.step(pooled, caps, b["lengths"])
    t_trunk = clock(lambda n: [feed.pooled(imgs[i & 1]) for i in range(n)], iters)
    t_dec = clock(lambda n: [st.step(pooled, caps, b["lengths"]) for _ in range(n)], iters)

    def overlapped(n):
        ticket = feed.submit(imgs[0])
        for i in range(n):
            p = feed.result(ticket)
            ticket = feed.submit(imgs[(i + 1) & 1])   # trunk of batch i+1 on the side stream ...
            st.step(p, caps, b["lengths"])            # ... while the decoder trains on batch i
        feed.result(ticket)
    overlapped(2)
    t_ov = clock(overlapped, iters)
    return {"batch": batch, "trunk_images_per_s": batch / t_trunk, "decoder_captions_per_s": batch / t_dec,
            "overlapped_images_per_s": batch / t_ov, "serial_images_per_s": batch / (t_trunk + t_dec),
            "what": "ResNet-152 trunk (cuDNN, channels-last bf16, eval-mode BN, side stream) + decoder train step on its "
                    "pooled features; wall clock over 6 iterations"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "torch-gpu", "configs0-gpu"])
    ap.add_argument("--config", default="default", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: BASELINE configs[4] as written, global batch 8192 sharded over the ranks")
    ap.add_argument("--device-index", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--prec", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-greedy", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip f_rows / configs3 / strong_8192")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip timing torch's own cuDNN/cuBLAS composition of the same step on this GPU")
    ap.add_argument("--stages", action="store_true", help="also print per-stage GPU time (CUDA events) to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "configs0-gpu":     # internal: the child of configs0_gpu()
        configs0_gpu_child(args.device_index)
        return
    if args.impl == "torch-gpu":        # internal: the child of gpu_torch_reference()
        gpu_torch_reference_child(args.device_index)
        return
    t_start = time.time()

    def phase(name):
        if os.environ.get("SNT_BENCH_DEBUG"):
            print(f"[dbg] rank {rank} +{time.time() - t_start:6.1f}s {name}", file=sys.stderr, flush=True)

    import torch.distributed as dist
    import show_and_tell_b200 as snt

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")   # all-reduce CTAs are placed first when SMs free up
        # The cooperative recurrence kernels need 128 of the 148 SMs at once: an all-reduce in flight with more than 20
        # CTAs makes the BPTT launch wait for the whole collective (measured: BPTT stage 294 -> 406 us at 8 GPUs).
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
    c = CONFIGS[args.config]
    peaks = load_peaks()
    b_local = c["B"] if args.scaling == "weak" else max(1, STRONG_GLOBAL_B // world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, k):
        barrier()
        st = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(k):
            fn()
        e1.record(st)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    wl = Workload(snt, c, b_local, world, rank, dev, args.prec)
    phase("warm-up")
    for _ in range(args.warmup + (EXTRA_WARMUP_MULTI_GPU if world > 1 else 0)):
        wl.step_resident()
    phase("warm-up done")
    if os.environ.get("SNT_BENCH_GC", "freeze") == "freeze":
        # Everything allocated so far is long-lived: move it to the permanent generation so that the cyclic collector's
        # full passes stay short (a full collection in the middle of a multi-GPU run pauses one rank's host thread for
        # tens of ms and, through the next all-reduce, every GPU of the job).
        import gc
        gc.collect()
        gc.freeze()
    sampler = ClockSampler(local)
    L = snt._lib.lib()
    if world == 1:
        sampler.start()          # NVML polled from a background thread for the whole timed region
    else:
        sampler.open_manual()    # polling while ranks exchange gradients stretches the step: sample right after instead
    L.snt_launch_count(1)
    t_res = timed(wl.step_resident, args.steps)
    launches = int(L.snt_launch_count(0))
    if world == 1:
        n_extra = int(np.ceil(max(0.0, 1.5 - t_res) / max(t_res / args.steps, 1e-6)))
        for _ in range(n_extra):
            wl.step_resident()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks["sampled_over"] = "timed region + continuation of the same loop to >= 1.5 s"
    else:
        for _ in range(4):
            for _ in range(5):
                wl.step_resident()
            sampler.sample_once()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks["sampled_over"] = "20 more steps of the same load right after the timed region (polling inside it perturbs multi-GPU steps)"
    t_fixed = timed(lambda: wl.step_resident(fixed=True), args.steps)
    phase("timed regions done; stage profile")

    stages = None
    if rank == 0:
        prof = profile_stages(wl, 10)
        stages = stage_rooflines(prof, wl.mean_tokens(), peaks, c, b_local)
        if args.stages:
            tot = sum(e["us_per_step"] for e in stages)
            for e in stages:
                print(f"[stages] {e['stage']:16s} {e['us_per_step']:9.1f} us/step {100 * e['us_per_step'] / tot:5.1f}%  "
                      f"{e.get('achieved', 0):8.1f} {e.get('unit', '')} ({100 * e.get('frac', 0):4.1f}% of peak)",
                      file=sys.stderr)
            print(f"[stages] sum {tot:.1f} us/step", file=sys.stderr)
    else:
        for _ in range(10):      # keep every rank's step count (and its collectives) aligned with rank 0's profile steps
            opt = wl.stepper.optimizer
            wl.stepper.optimizer = False
            wl.step_resident()
            wl.stepper.optimizer = opt
            wl.stepper.update()
    phase("stage profile done; e2e")
    e2e = E2E(wl)
    for _ in range(max(args.warmup, 12)):       # first calls: pinned staging, side-stream copy buffers, allocator growth
        e2e.step()
    t_e2e = timed(e2e.step, args.steps)
    e2e.step()
    loss_val = e2e.final_loss()
    phase("e2e done")

    total_caps = b_local * world
    value = total_caps * args.steps / t_res
    n_tok = wl.mean_tokens()
    flops_step = train_flops(b_local, n_tok, c)
    step_tf = flops_step * args.steps / t_res / 1e12      # per GPU (max-over-ranks time)

    strong = None
    if world > 1 and args.scaling == "weak" and args.config == "default" and not args.no_extras:
        # BASELINE configs[4] as written: global batch 8192 over the N ranks (4096 / 2048 / 1024 captions per GPU)
        phase("strong-scaling leg")
        ws = Workload(snt, c, STRONG_GLOBAL_B // world, world, rank, dev, args.prec, nb=4)
        for _ in range(20):
            ws.step_resident()
        ks = max(10, args.steps // 2)
        t_s = timed(ws.step_resident, ks)
        strong = {"global_batch": STRONG_GLOBAL_B, "batch_per_gpu": STRONG_GLOBAL_B // world, "n_gpus": world,
                  "value": STRONG_GLOBAL_B * ks / t_s, "unit": "captions/s", "ms_per_step": t_s / ks * 1e3,
                  "steps": ks, "scaling": "strong",
                  "what": "BASELINE configs[4]: data-parallel train step, global batch 8192 sharded by strided rows"}
        ws.stepper.close()
        del ws
        torch.cuda.empty_cache()

    greedy = None
    if not args.no_greedy:
        phase("greedy")
        greedy = measure_greedy(snt, wl.dec, CFG if args.config == "default" else c, dev, world, timed, peaks, rank,
                                with_cpu=(world == 1 and not args.no_cpu_baseline)) if c["L"] >= 1 else None

    extra = {}
    if rank == 0:
        top = stages[0]
        traffic = load_traffic()
        tr_stage = (traffic or {}).get("stages", {}).get(top["stage"]) if args.config == "default" else None
        roof = {"bound": top.get("bound"), "achieved": top.get("achieved"), "peak": top.get("peak"),
                "unit": top.get("unit"), "frac": top.get("frac"),
                "traffic": tr_stage["dram_bytes"] if tr_stage else None,
                "traffic_source": (traffic or {}).get("source") if tr_stage else None,
                "kernel": top["stage"] + (" (" + tr_stage["kernels"] + ")" if tr_stage and "kernels" in tr_stage else ""),
                "us_per_step": top["us_per_step"],
                "share_of_step": top["us_per_step"] / max(sum(e["us_per_step"] for e in stages), 1e-9),
                "algorithmic_work_per_step": top.get("algorithmic_work"),
                "peak_source": f"{peaks['src']} (MEASURED_PEAKS.json: sustained bf16 for a stage inside a long step)",
                "timing": "CUDA events inside snt_step_run around the stage on its launching stream, mean of 10 real steps"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cps, dt, _, ntk, kind, what = cpu_train_baseline(c, 3, 1, threads)
            cpu = {"value": cps, "unit": "captions/s", "cores": threads, "kind": kind,
                   "sample": f"3 full train steps (head+decoder fwd, CE, bwd, clip, Adam) of B={c['B']} ({dt:.2f} s/step) "
                             f"after 1 warm-up, {what}"}
        gpu_ref = None
        if world == 1 and not args.no_gpu_reference and args.config == "default":
            gpu_ref = gpu_torch_reference(local)
            if greedy is not None and isinstance(gpu_ref, dict):
                greedy["gpu_torch_reference"] = {k: v.get("greedy_tokens_per_s") for k, v in gpu_ref.items()
                                                 if isinstance(v, dict) and "greedy_tokens_per_s" in v}
        if world == 1 and not args.no_extras:
            if args.config == "default":
                try:      # BASELINE configs[3] in the same record
                    c3 = CONFIGS["scaled"]
                    w3 = Workload(snt, c3, c3["B"], 1, 0, dev, args.prec, nb=4)
                    for _ in range(4):
                        w3.step_resident()
                    k3 = 12
                    t3 = timed(w3.step_resident, k3)
                    f3 = train_flops(c3["B"], w3.mean_tokens(), c3)
                    st3 = stage_rooflines(profile_stages(w3, 4), w3.mean_tokens(), peaks, c3, c3["B"])
                    extra["configs3"] = {"value": c3["B"] * k3 / t3, "unit": "captions/s", "ms_per_step": t3 / k3 * 1e3,
                                         "steps": k3, "config": workload_config(c3, c3["B"], 1, args.prec,
                                                                                int(w3.mean_tokens()), 20),
                                         "step_tflops": f3 * k3 / t3 / 1e12,
                                         "step_frac_of_sustained_peak": f3 * k3 / t3 / 1e12 / peaks["tf_sust"],
                                         "stages": st3}
                    w3.stepper.close()
                    del w3
                    torch.cuda.empty_cache()
                except Exception as e:   # noqa: BLE001
                    extra["configs3"] = {"error": repr(e)[:300]}
            try:          # the fp32-faithful mode (fp32 operands split into bf16 triples, fp32 accumulation) on the same step
                wf = Workload(snt, c, c["B"], 1, 0, dev, "fp32", nb=2)
                for _ in range(2):
                    wf.step_resident()
                kf = 5
                tf_ = timed(wf.step_resident, kf)
                extra["fp32_faithful_train"] = {"value": c["B"] * kf / tf_, "unit": "captions/s",
                                                "ms_per_step": tf_ / kf * 1e3, "steps": kf,
                                                "what": "the same train step in the fp32-faithful mode (no operand "
                                                        "rounding: loss <= 1e-5, gradients <= 1e-4 of the fp64 reference)"}
                wf.stepper.close()
                del wf
                torch.cuda.empty_cache()
            except Exception as e:   # noqa: BLE001
                extra["fp32_faithful_train"] = {"error": repr(e)[:300]}
            extra["f_rows"] = measure_f_rows(snt, dev, c, peaks)
        line = {
            "metric": "train_captions_per_s", "value": value, "unit": "captions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_res / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16" if args.prec == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(c, b_local, world, args.prec, int(n_tok), int(max(b["lengths"][0] for b in wl.batches)),
                                      batches=f"{NB} distinct ragged batches cycled in the timed region (resident in HBM)",
                                      extra_warmup_steps=EXTRA_WARMUP_MULTI_GPU if world > 1 else 0,
                                      launch_mode="eager launches through the native step executor (no CUDA graph)",
                                      exchange=("none (one GPU)" if world == 1 else
                                                "gradient all-reduce + clip/Adam + parameter broadcast fused in one kernel over "
                                                "NVSwitch multicast memory (snt_dp_adam_shard: multimem.ld_reduce / multimem.st, "
                                                "optimizer sharded over the ranks)" if wl.stepper.grads_are_local else
                                                "three bucketed NCCL all-reduces overlapped with backward"),
                                      l2="no explicit flush: each step streams ~0.9 GB of activations/weights (> 126 MB L2)"),
            "fixed_batch": {"value": total_caps * args.steps / t_fixed, "ms_per_step": t_fixed / args.steps * 1e3,
                            "what": "the same loop repeating batch 0"},
            "e2e": {"value": total_caps * args.steps / t_e2e, "unit": "captions/s", "ms_per_step": t_e2e / args.steps * 1e3,
                    "h2d_bytes_per_step": e2e.h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "clocks": clocks, "roofline": roof, "stages": stages, "cpu_baseline": cpu,
            "gpu_torch_reference": gpu_ref, "greedy": greedy, "strong_8192": strong,
            "step_tflops_per_gpu": step_tf, "step_frac_of_sustained_peak": step_tf / peaks["tf_sust"],
            "algorithmic_flops_per_step": flops_step, "loss": loss_val, **extra,
        }
        print(json.dumps(line), flush=True)
    if os.environ.get("SNT_DP_DEBUG") and hasattr(wl.stepper, "exchange_report"):
        print(f"[dp] rank {rank} exchange: {wl.stepper.exchange_report()}", file=sys.stderr, flush=True)
    phase("line printed; teardown")
    wl.stepper.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    phase("exit")


if __name__ == "__main__":
    main()
