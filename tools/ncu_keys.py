"""print a few key metrics of an .ncu-rep (first kernel): python tools/ncu_keys.py file.ncu-rep [more.ncu-rep]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_elapsed.avg.per_second", "launch__registers_per_thread"]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("==", f, vals[hdr.index("Kernel Name")][:60])
    for i, h in enumerate(hdr):
        if h in KEYS or ("issue_stalled" in h and h.endswith("_per_warp_active.pct") and "not_issued" not in h):
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if "issue_stalled" in h and v < 2.0:
                continue
            print(f"  {h:95s} {units[i]:12s} {v:,.3f}")
