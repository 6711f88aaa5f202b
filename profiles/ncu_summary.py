"""Extract the judged metrics from `ncu --set full` reports (run here, no GPU needed):
   python profiles/ncu_summary.py gpurun_out/prof_*.ncu-rep > profiles/rNN_ncu_full_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"##### {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in WANT:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        rd = float(d.get("dram__bytes_read.sum", "0").replace(",", "") or 0)
        wr = float(d.get("dram__bytes_write.sum", "0").replace(",", "") or 0)
        print(f"  traffic (dram read+write, units as above)                                   {rd + wr:18.3f}")
        print()
