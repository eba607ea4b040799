"""Writes the judged subset of an .ncu-rep raw page as metric,unit,value rows (one block per profiled launch).
usage: python tools/ncu_raw_summary.py report.ncu-rep out.csv"""
import csv, io, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
                  r"lts__t_bytes\.sum(\.per_second)?|lts__t_sector_hit_rate\.pct|l1tex__t_bytes\.sum|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_tensor.*|sm__pipe_tensor.*active.*|"
                  r"sm__inst_executed\.avg\.per_cycle_active|sm__instruction_throughput\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(xu|fma|alu|fmaheavy|shared)_cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_(xu|lsu|uniform)\.sum|"
                  r"launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit.*|waves_per_multiprocessor)|"
                  r"smsp__average_warp.*_per_issue_stalled.*|smsp__pcsamp_sample_buffer_full|l1tex__data_bank_conflicts_pipe_lsu.*sum|"
                  r"sm__cycles_elapsed\.max|sm__cycles_active\.avg)$")
ki = hdr.index("Kernel Name")
with open(out, "w") as f:
    for r in data:
        f.write("# %s\n" % r[ki])
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, r):
            if keep.match(h):
                f.write("%s,%s,%s\n" % (h, u, v))
print("wrote", out, len(data), "launches")
