"""Trim an `ncu --page raw --csv` dump to the columns the roofline discussion uses: python tools/ncu_summary.py raw.csv out.csv"""
import csv, sys

COLS = [("gpu__time_duration.sum", "duration"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [(c, n) for c, n in COLS if c in ix]
with open(sys.argv[2], "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["id", "kernel"] + [f"{n} [{units[ix[c]]}]" for c, n in cols])
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        name = name.split("(")[0].replace("void ", "").replace("kg::", "")
        w.writerow([r[ix["ID"]], name] + [r[ix[c]] for c, _ in cols])
