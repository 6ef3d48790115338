"""Summarise an Nsight Compute report (read here, on the CPU box): one block of key metrics per
kernel launch, as committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xyz.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex_lookup_miss.sum", "smsp__inst_executed_op_global_red.sum",
    "smsp__inst_executed_op_shfl.sum" if False else "smsp__inst_executed_pipe_lsu.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("-" * 100)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-75s %s %s" % (k, r[i], units[i]))
        for i, h in enumerate(hdr):
            if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 3:
                    print("%-75s %.1f %%" % ("stall: " + h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))


if __name__ == "__main__":
    main(sys.argv[1])
