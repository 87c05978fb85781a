"""One encode+decode step from an `ncu --set full --page raw --csv` dump -> markdown table (profiles/rNN_summary.md):
python tools/ncu_step_summary.py gpurun_out/r2c_full.raw.csv [gpurun_out/r2c_launches.csv]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(r, k, default=0.0):
    if k not in hdr:
        return default
    v = r[hdr.index(k)].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return default


def unit(k):
    return units[hdr.index(k)] if k in hdr else ""


def to_gb(v, u):
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)


print("| # | kernel | grid x block | ms | DRAM read GB | DRAM write GB | DRAM GB/s | LSU smem pipe % | tensor-core smem reads % | tensor pipe % | issue % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
tot_ms = tot_rd = tot_wr = 0.0
for i, r in enumerate(data):
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    ms = col(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1.0)
    rd = to_gb(col(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"))
    wr = to_gb(col(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    tot_ms += ms; tot_rd += rd; tot_wr += wr
    print(f"| {i} | `{name}` | {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''} x {r[hdr.index('Block Size')] if 'Block Size' in hdr else ''} | {ms:.3f} | {rd:.3f} | {wr:.3f} | "
          f"{(rd + wr) / ms * 1e3:.0f} | {col(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{col(r, 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{int(col(r, 'launch__registers_per_thread'))} |")
print(f"\nstep under ncu (serialised, cold caches): {tot_ms:.3f} ms, DRAM {tot_rd:.2f} GB read + {tot_wr:.2f} GB written = {tot_rd + tot_wr:.2f} GB")
