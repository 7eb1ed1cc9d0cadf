"""Summarise an .ncu-rep (ncu --set full) as a small text table of the metrics the roofline uses.
usage: python scripts/ncu_summary.py gpurun_out/<tag>/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (tf32) active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (UMMA) active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall LG throttle"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe throttle"),
]


def main(paths):
    for p in paths:
        raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print("== %s" % p.split("/")[-1])
        for r in rows[2:]:
            print("-- kernel: %s" % r[idx["Kernel Name"]][:110])
            for key, label in WANT:
                if key in idx and r[idx[key]] != "":
                    print("   %-34s %s %s" % (label, r[idx[key]], units[idx[key]]))
            extra = [h for h in hdr if "tensor" in h and h not in dict(WANT) and r[idx[h]] not in ("", "0")]
            for h in extra[:8]:
                print("   %-34s %s %s" % (h[:60], r[idx[h]], units[idx[h]]))


if __name__ == "__main__":
    main(sys.argv[1:])
