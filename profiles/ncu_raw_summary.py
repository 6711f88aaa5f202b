"""Summarise an `ncu --page raw --csv` export (made on the GPU box from an `ncu --set full` report):
   python profiles/ncu_raw_summary.py gpurun_out/r01_hot_full_raw.csv > profiles/r01_ncu_hot_kernels.txt"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"),
        ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name")
for r in rows[2:]:
    print(r[kn][:150])
    rd = wr = 0.0
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print(f"    {label:24s} {r[i]:>16s} {units[i]}")
    print()
