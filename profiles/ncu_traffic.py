"""From the raw page of an `ncu --set full` capture of ONE training step (tools/one_step.py between
cudaProfilerStart/Stop; `ncu -i step_full.ncu-rep --page raw --csv > step_full_raw.csv`), write

  * a per-launch table (stage, duration, tensor-pipe active %, DRAM bytes read / written, L2 hit rate, grid) and
  * profiles/r02_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per stage of the step, which
    bench.py reads for `roofline.traffic` (never a literal).

    python profiles/ncu_traffic.py gpurun_out/r02/step_full_raw.csv profiles/r02_traffic.json > profiles/r02_ncu_step_full.txt

Launches are mapped to the stages of snt_step_run (csrc/step.cu) by their order and kernel names."""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def col(hdr, key):
    for i, h in enumerate(hdr):
        if h == key or h.endswith("." + key):
            return i
    return None


def stage_of(names):
    """Stage of every launch of one step, in launch order (see csrc/step.cu for the sequence)."""
    out, phase = [], "head_fwd"
    seen_lstm_prep = 0
    for n in names:
        if n.startswith("void at::") or n.startswith("at::"):
            out.append("host_glue(torch)")
            continue
        if re.search(r"emb_(tok|scan|place|small|sort|chunk|final)_kernel", n):
            out.append("embed_pack_bwd")
            continue
        if "pack_targets_kernel" in n or "embed_pack_fwd_kernel" in n:
            phase = "embed_pack_fwd"
        elif "lstm_prep_kernel" in n:
            seen_lstm_prep += 1
            phase = "lstm_fwd" if phase in ("head_fwd", "embed_pack_fwd", "lstm_fwd") else "lstm_bwd"
        elif phase == "lstm_fwd" and ("cast_bf16_kernel" in n or "RowMaxEpi" in n or "CeStoreEpi" in n or "CeFwdEpi" in n):
            phase = "vocab_ce_fwd"
        elif phase == "vocab_ce_fwd" and "ce_finish" in n:
            out.append("vocab_ce_fwd")
            phase = "vocab_ce_bwd"
            continue
        elif "bn_bwd_kernel" in n:
            phase = "head_bwd"
        elif "clamp_adam" in n:
            phase = "clamp_adam"
        out.append(phase)
    return out


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    want = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
            "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l2hit": "lts__t_sector_hit_rate.pct", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "regs": "launch__registers_per_thread"}
    idx = {k: col(hdr, v) for k, v in want.items()}
    gi = hdr.index("Grid Size")

    def val(r, k):
        i = idx[k]
        if i is None or r[i] in ("", "n/a"):
            return 0.0
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    names = [r[kn] for r in data]
    stages = stage_of(names)
    per = {}
    print(f"# {sys.argv[1]}: {len(data)} launches of one training step (configs[1]: B=1024, E256/H512/V10000), ncu --set full,")
    print("# --clock-control none; durations are cold-cache and serialised (compare shares); DRAM bytes are per launch.")
    print(f"{'#':>3s} {'stage':16s} {'us':>8s} {'tensor%':>8s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'DRAM%':>6s} {'L2hit%':>7s} {'regs':>5s} {'grid':>14s}  kernel")
    for i, (r, s) in enumerate(zip(data, stages)):
        d, rd, wr = val(r, "dur"), val(r, "rd"), val(r, "wr")
        e = per.setdefault(s, {"us": 0.0, "dram_bytes": 0.0, "launches": 0, "kernels": []})
        e["us"] += d
        e["dram_bytes"] += rd + wr
        e["launches"] += 1
        short = re.sub(r"\(.*$", "", names[i]).replace("void ", "")
        if short not in e["kernels"]:
            e["kernels"].append(short)
        print(f"{i:3d} {s:16s} {d:8.1f} {val(r, 'tensor'):8.1f} {rd / 1e6:11.2f} {wr / 1e6:11.2f} {val(r, 'dram_pct'):6.1f} "
              f"{val(r, 'l2hit'):7.1f} {int(val(r, 'regs')):5d} {r[gi]:>14s}  {short[:110]}")
    tot_us = sum(e["us"] for e in per.values())
    tot_b = sum(e["dram_bytes"] for e in per.values())
    print()
    print(f"{'stage':18s} {'launches':>8s} {'us':>9s} {'share':>6s} {'DRAM MB':>9s}")
    for s, e in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{s:18s} {e['launches']:8d} {e['us']:9.1f} {100 * e['us'] / tot_us:5.1f}% {e['dram_bytes'] / 1e6:9.1f}")
    print(f"{'total':18s} {len(data):8d} {tot_us:9.1f} {'':6s} {tot_b / 1e6:9.1f}")
    if len(sys.argv) > 2:
        js = {"source": f"ncu --set full --clock-control none, one training step (tools/one_step.py), raw page {sys.argv[1].split('/')[-1]}; "
                        "summary in profiles/r02_ncu_step_full.txt",
              "total_dram_bytes": tot_b,
              "stages": {s: {"dram_bytes": e["dram_bytes"], "launches": e["launches"], "ncu_us": e["us"],
                             "share_of_step_ncu": e["us"] / tot_us, "kernels": ", ".join(k[:60] for k in e["kernels"][:6])}
                         for s, e in per.items()}}
        json.dump(js, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
