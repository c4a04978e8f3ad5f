"""Selected metrics of `ncu --set full` reports -> text (profiles/rNN_ncu_full_summary.txt).

    python scripts/ncu_summary.py gpurun_out/c3/*.ncu-rep > profiles/r02_ncu_full_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "inst_executed",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__sass_inst_executed_op_utcmma.sum",
        "smsp__inst_executed_op_tma_ld.sum"]


def main(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {p}: empty report")
            continue
        hdr, units = rows[0], rows[1]
        idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
        print(f"# {p}")
        for r in rows[2:]:
            print("---")
            for w, i in idx:
                print(f"  {w} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main(sys.argv[1:])
