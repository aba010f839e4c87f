"""prints the handful of ncu counters the roofline discussion uses from `ncu -i X.ncu-rep --page raw --csv` output
usage: ncu_summary.py raw.csv [algorithmic_bytes]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    vals = dict(zip(hdr, r))
    un = dict(zip(hdr, units))
    for w in WANT:
        if w in vals:
            print(f"{w:70s} {vals[w]} {un.get(w, '')}")
    tot = None
    try:
        def tobytes(k):
            v, u = float(vals[k]), un[k].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        tot = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        t = float(vals["gpu__time_duration.sum"]) * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1}[un["gpu__time_duration.sum"].lower()]
        print(f"{'DRAM traffic (read + write)':70s} {tot / 1e9:.4f} GB  -> {tot / t / 1e9:.0f} GB/s")
        if alg:
            print(f"{'algorithmic bytes':70s} {alg / 1e9:.4f} GB  -> {alg / t / 1e9:.0f} GB/s ; traffic / algorithmic = {tot / alg:.3f}")
    except Exception as e:
        print("(could not derive traffic:", e, ")")
    stalls = sorted(((float(v), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in vals.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v.replace(".", "").isdigit()), reverse=True)
    s = sum(x for x, _ in stalls) or 1.0
    print("warp stall samples: " + ", ".join(f"{k} {100 * x / s:.1f}%" for x, k in stalls[:7]))
    print()
