#!/usr/bin/env python
"""Turn an .ncu-rep (read here, without a GPU) into the small text summary committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# summary of {path} (ncu --set full --clock-control none; per-launch values)")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\n## {name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:95s} {r[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(starts[:-1], starts[1:]):
        hdr = rows[a + 1]
        data = [r for r in rows[a + 2:b] if len(r) == len(hdr) and r[0].startswith("0x")]
        if not data:
            continue
        cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h}
        agg = {h: sum(int(float(r[i] or 0)) for r in data) for h, i in cols.items()}
        tot = sum(agg.values()) or 1
        print(f"\n## warp-stall samples, {rows[a][1][:90]}")
        print("   " + ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        isamp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
        print(f"   SASS instructions: {len(data)}, executed (warp-level): {sum(int(r[iex]) for r in data)}")
        top = sorted(data, key=lambda r: -int(r[isamp]))[:8]
        for r in top:
            print(f"   {int(r[isamp]):8d} samples  {r[isrc].strip()[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
