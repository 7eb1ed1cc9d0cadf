"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu) of the hand-written kernels, averaged over
the launches of one profiled step and keyed by the C-ABI entry point that launches them -> profiles/roofline_traffic.json
(bench.py puts the matching entry into roofline.traffic).
usage: python scripts/traffic_from_launches.py gpurun_out/<tag>/launches.csv > profiles/roofline_traffic.json"""
import collections
import csv
import json
import sys

KERNEL_TO_ENTRY = [
    ("conv3x3_tf32_2cta", "odwscl_conv3x3_nhwc_tf32"), ("conv3x3_tf32_kernel", "odwscl_conv3x3_nhwc_tf32"),
    ("conv3x3_wgrad_tf32", "odwscl_conv3x3_wgrad_nhwc_tf32"),
    ("roi_pool_fwd_nhwc7_kernel", "odwscl_roi_pool_fwd_nhwc_aug_f32"),
    ("fc_gemm_tf32_2cta", "odwscl_fc_gemm_tf32"),
    ("roi_pool_bwd_plane_kernel", "odwscl_roi_pool_bwd_nhwc_multi_f32"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    per_launch = collections.defaultdict(float)        # (ID) -> bytes
    name_of = {}
    for row in csv.DictReader(lines):
        m = row.get("Metric Name", "")
        if not m.startswith("dram__bytes_"):
            continue
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per_launch[row["ID"]] += v
        name_of[row["ID"]] = row["Kernel Name"]
    agg = collections.defaultdict(list)
    for i, b in per_launch.items():
        for frag, entry in KERNEL_TO_ENTRY:
            if frag in name_of[i]:
                agg[entry].append(b)
                break
    out = {k: sum(v) / len(v) for k, v in agg.items()}
    out["_note"] = ("average dram__bytes_read.sum + dram__bytes_write.sum per launch over the launches of the profiled steps "
                    "(ncu replay: caches flushed before every kernel, so reads the live step serves from L2 count as DRAM)")
    out["_launches"] = {k: len(v) for k, v in agg.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
