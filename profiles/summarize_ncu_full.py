"""Key counters from `ncu --set full` reports (run here, no GPU needed):
   python profiles/summarize_ncu_full.py gpurun_out/prof.ncu-rep > profiles/<name>.csv"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.max", "cycles"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"), ("lts__t_bytes.sum", "l2_bytes"),
        ("smsp__inst_executed.sum", "warp_insts"),
        ("sm__inst_executed_pipe_fp64.sum", "fp64_pipe_insts"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_insts"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")]
idx = [(hdr.index(a), b, rows[1][hdr.index(a)]) for a, b in want if a in hdr]
print(",".join(f"{b}[{u}]" if u else b for _, b, u in idx))
for r in rows[2:]:
    print(",".join(r[i].split("(")[0].replace(",", "") if b == "kernel" else r[i].replace(",", "") for i, b, _ in idx))
