"""From the raw page of an `ncu --set full` capture of ONE training step (tools/one_step.py between
cudaProfilerStart/Stop; `ncu -i step_full.ncu-rep --page raw --csv > step_full_raw.csv`), write

  * a per-launch table (stage, duration, tensor-pipe active %, DRAM bytes read / written, L2 hit rate, grid) and
  * profiles/r02_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per stage of the step, which
    bench.py reads for `roofline.traffic` (never a literal).

    python profiles/ncu_traffic.py gpurun_out/r02/step_full_raw.csv profiles/r02_traffic.json > profiles/r02_ncu_step_full.txt

Launches are mapped to the stages of snt_step_run (csrc/step.cu) by kernel name and grid size."""
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


def stage_of(names, grids):
    """Stage of every launch of one step.  By kernel name and grid, not by position: the executor (csrc/step.cu) runs
    weight preparation, gathers, the token plan, column sums and part of the optimizer on side streams, so the launch
    order interleaves the stages.  d_b_out's column sums (a pass over the stored numerators that runs beside the BPTT
    kernel) are booked under vocab_ce_bwd, whose gradient they are."""
    out, seen_bptt, seen_bn_bwd, last_gemm, last_colsum = [], False, False, "head_fwd", "lstm_bwd"
    for n, g in zip(names, grids):
        g0 = int(re.sub(r"[^0-9,]", "", g).split(",")[0] or 0)
        if n.startswith("void at::") or n.startswith("at::"):
            st = "host_glue(torch)"
        elif re.search(r"emb_(tok|scan|place|small|sort|chunk|final)_kernel", n):
            st = "embed_pack_bwd"
        elif "pack_targets_kernel" in n or "embed_pack_fwd_kernel" in n:
            st = "embed_pack_fwd"
        elif "bn_fwd_kernel" in n:
            st = "head_fwd"
        elif "bn_bwd_kernel" in n:
            st, seen_bn_bwd = "head_bwd", True
        elif "snt::colsum_partial_kernel" in n or "snt::colsum_final_kernel" in n or "cast2d_kernel" in n:
            st = "head_bwd"
        elif "lstm_prep" in n:
            st = "lstm_bwd" if seen_bptt else "lstm_fwd"
        elif "lstm_fwd_persistent" in n or "LstmFwdEpi" in n or "PlainEpi<256, 1>" in n:
            st = "lstm_fwd"
        elif "RowMaxEpi" in n or "CeStoreEpi" in n or "CeFwdEpi" in n or "ce_finish" in n:
            st = "vocab_ce_fwd"
        elif "lstm_bwd_persistent" in n or "lstm_bwd_point" in n:
            st, seen_bptt = "lstm_bwd", True
        elif "gemm_tc_kernel<256, 0, 1" in n:
            st = last_gemm = "vocab_ce_bwd"                       # dHs = U' . W_out
        elif "gemm_tc_kernel<256, 1, 1" in n:
            st = last_gemm = "lstm_bwd" if (seen_bptt or g0 < 100) else "vocab_ce_bwd"   # dW_ih / dW_hh (48 / 96 CTAs) or dW_out
        elif "gemm_tc_kernel<128, 0, 1" in n:
            st = last_gemm = "lstm_bwd"                           # dX
        elif "gemm_tc_kernel<128, 0, 0" in n:
            st = last_gemm = "head_fwd"
        elif "gemm_tc_kernel<128, 1, 1" in n:
            st = last_gemm = "head_bwd"
        elif "splitk_reduce" in n:
            st = last_gemm
        elif "colsum_bf16_partial_kernel" in n:
            st = last_colsum = "vocab_ce_bwd" if g0 >= 1000 else "lstm_bwd"   # d_b_out (N x V pass) or the LSTM bias
        elif "colsum_bf16_final_kernel" in n or "scale_vec_kernel" in n:
            st = last_colsum
        elif "unperm_vec_kernel" in n:
            st = "lstm_bwd"
        elif "cast_bf16_kernel" in n:
            st = "vocab_ce_fwd" if g0 >= 4000 else ("head_bwd" if seen_bn_bwd else "head_fwd")
        elif "clamp_adam" in n or "dp_adam" in n:
            st = "clamp_adam"
        else:
            st = "other"
        out.append(st)
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
    stages = stage_of(names, [r[gi] for r in data])
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
