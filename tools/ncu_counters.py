#!/usr/bin/env python
"""tools/ncu_counters.py [workload ...] — regenerates profiles/ncu_counters.json on the GPU box (run it under gpurun, ONE GPU).

For every workload: one `ncu --metrics <the eight below> --clock-control none` pass over the k_traverse* / k_packet / k_primary launches of one warm
frame of `bench.py --workload W` (RTB_LANES=1 so launches keep program order), after the same command has run without ncu.  From
the raw page it keeps, per frame: DRAM bytes read + written (`roofline.traffic`), and — weighted by each launch's duration — active
lanes per instruction, issue-slot utilisation, L1 data-pipe wavefront utilisation and the L2 hit rate.  The file is keyed to
bench.source_hash() (a hash of the kernel sources) and to the kernel-selecting environment, so bench.py can tell a stale file from
a current one and prints nulls plus a warning instead of numbers that belong to another build.

The result is written to gpurun_out/ncu_counters.json (what gpurun brings back); copy it to profiles/ncu_counters.json and commit.
"""
import csv
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KERNELS = "k_traverse|k_packet|k_primary|k_tail"
M = {
    "ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l2hit": "lts__t_sector_hit_rate.pct",
    "inst": "smsp__inst_executed.sum",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def num(x):
    return float(x.replace(",", ""))


def frame_launch_count(workload):
    """How many launches of the traversal family one frame makes: depth + 1 traversals + depth tails (bench.WORKLOADS)."""
    depth = bench.WORKLOADS[workload][5]
    return 2 * depth + 1


def capture(workload, out_dir):
    env = dict(os.environ, RTB_LANES="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--no-c5"]
    plain = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if plain.returncode != 0:
        raise SystemExit(f"{' '.join(cmd)} failed without ncu:\n{plain.stderr[-2000:]}")
    n = frame_launch_count(workload)
    rep = os.path.join(out_dir, f"counters_{workload}")
    # skip the launches of the upload frame and of one warm-up frame: capture one frame of the timed region
    # only the metrics this file reads (a handful of replay passes instead of --set full's ~40: the capture costs GPU minutes)
    ncu = ["ncu", "--metrics", ",".join(M.values()), "--clock-control", "none", "-k", f"regex:{KERNELS}", "-s", str(2 * n), "-c", str(n), "-f", "-o", rep] + cmd
    r = subprocess.run(ncu, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"ncu failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(v) for k, v in M.items()}

    def val(r, k):
        return num(r[col[k]]) * UNIT_SCALE.get(units[col[k]], 1.0)

    ms = [val(r, "ms") for r in data]
    total_ms = sum(ms)
    w = [x / total_ms for x in ms]
    rec = {
        "env": bench.kernel_env_key(), "launches": len(data), "kernels": sorted({r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1] for r in data}),
        "traverse_ms_under_ncu": total_ms,
        "dram_bytes_per_frame": sum(val(r, "rd") + val(r, "wr") for r in data),
        "lanes_per_inst": sum(wi * num(r[col["lanes"]]) for wi, r in zip(w, data)),
        "issue_active_pct": sum(wi * num(r[col["issue"]]) for wi, r in zip(w, data)),
        "l1_wavefront_pct": sum(wi * num(r[col["l1"]]) for wi, r in zip(w, data)),
        "l2_hit_pct": sum(wi * num(r[col["l2hit"]]) for wi, r in zip(w, data)),
        "warp_instructions_per_frame": sum(num(r[col["inst"]]) for r in data),
        "per_launch_ms": [round(x, 4) for x in ms],
    }
    return rec


def main():
    workloads = sys.argv[1:] or ["c4"]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    doc = {"source_hash": bench.source_hash(), "generated_by": "tools/ncu_counters.py " + " ".join(workloads),
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "how": "ncu --metrics (dram bytes, lanes per instruction, issue active, L1 wavefronts, L2 hit rate) --clock-control none, one warm frame, RTB_LANES=1",
           "workloads": {}}
    for wl in workloads:
        doc["workloads"][wl] = capture(wl, out_dir)
        print(wl, json.dumps(doc["workloads"][wl]), file=sys.stderr)
    with open(os.path.join(out_dir, "ncu_counters.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc))


if __name__ == "__main__":
    main()
