"""Where does the host-input (e2e) loop of bench.py spend its time?  Wall-clock per step of the resident loop and of
the E2E loop, plus the host time of each part of E2E.step (issue of the H2D copies, enqueue of the step, wait for
the previous loss).  python tools/e2e_probe.py [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import show_and_tell_b200 as snt


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = bench.Workload(snt, bench.CFG, 1024, 1, 0, dev, "bf16")
    for _ in range(10):
        wl.step_resident()
    torch.cuda.synchronize()
    # host enqueue time of a resident step (no sync inside)
    t = []
    for _ in range(steps):
        t0 = time.perf_counter()
        wl.step_resident()
        t.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    t = np.array(t) * 1e3
    print(f"resident: host enqueue per step median {np.median(t):.3f} ms, p90 {np.percentile(t, 90):.3f}, max {t.max():.3f}")
    t0 = time.perf_counter()
    for _ in range(steps):
        wl.step_resident()
    torch.cuda.synchronize()
    print(f"resident: wall {1e3 * (time.perf_counter() - t0) / steps:.3f} ms/step")
    e = bench.E2E(wl)
    for _ in range(6):
        e.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = []
    for _ in range(steps):
        s0 = time.perf_counter()
        e.step()
        per.append(time.perf_counter() - s0)
    e.final_loss()
    wall = (time.perf_counter() - t0) / steps * 1e3
    per = np.array(per) * 1e3
    print(f"e2e: wall {wall:.3f} ms/step; per-call median {np.median(per):.3f}, p90 {np.percentile(per, 90):.3f}, max {per.max():.3f}")
    print("e2e per-call (ms):", " ".join(f"{x:.2f}" for x in per))
    # parts
    orig_issue = e._issue
    acc = {"issue": 0.0}
    def timed_issue(k):
        s = time.perf_counter(); orig_issue(k); acc["issue"] += time.perf_counter() - s
    e._issue = timed_issue
    orig_step = wl.stepper.step
    acc["step"] = 0.0
    def timed_step(*a, **k):
        s = time.perf_counter(); r = orig_step(*a, **k); acc["step"] += time.perf_counter() - s; return r
    wl.stepper.step = timed_step
    t0 = time.perf_counter()
    for _ in range(steps):
        e.step()
    e.final_loss()
    wall = (time.perf_counter() - t0) / steps * 1e3
    print(f"e2e parts per step: issue {acc['issue'] / steps * 1e3:.3f} ms, stepper.step {acc['step'] / steps * 1e3:.3f} ms, "
          f"wall {wall:.3f} ms (rest = loss wait + copies)")
    wl.stepper.step = orig_step
    # GPU-side: is the H2D copy slowing kernels?  time steps with the copies but no loss wait
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e._issue = orig_issue
    e.pending = False
    ev0.record()
    for _ in range(steps):
        wl_i = e.i
        e.pending = False
        e.step()
    ev1.record()
    torch.cuda.synchronize()
    print(f"e2e without the per-step loss wait: device {ev0.elapsed_time(ev1) / steps:.3f} ms/step")


if __name__ == "__main__":
    main()
