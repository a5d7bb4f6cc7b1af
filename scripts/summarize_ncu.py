"""Summarise an `ncu --set full` capture of the sweep kernel into profiles/ (markdown table + traffic json).

usage: python scripts/summarize_ncu.py <capture.ncu-rep> <round tag, e.g. r01> <trajectories per launch> "<command line captured>"

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed).  Launches whose grid covers the whole batch
(duration > 1 ms) are the per-segment sweeps; the short ones are jump-event rounds over a handful of trajectories.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value) * scale.get(unit, 1.0)


def main():
    rep, tag, traj, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    name = body[0][col["Kernel Name"]]
    out = [f"# ncu --set full capture of the sweep kernel ({tag})", "", f"Command (B200, gpurun): `{cmd}`", "",
           f"Kernel: `{name}`; 20-qubit depth-50 Heron circuit (BASELINE configs[4]), {traj} trajectories per launch "
           f"({traj} x 16 MiB state).  Launches longer than 1 ms sweep the whole batch once (one segment); the short "
           "ones are jump-event rounds over the few trajectories that had an event in the segment.", "",
           f"Algorithmic bytes of one full sweep (SURVEY 8(d)): {traj} trajectories x 2 x 16 B x 2^20 = "
           f"{traj * 32 * 2 ** 20 / 1e6:.1f} MB.", ""]
    out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(body))) + " |")
    out.append("|---|---|" + "---|" * len(body))
    for m in METRICS:
        if m in col:
            out.append(f"| {m} | {units[col[m]]} | " + " | ".join(r[col[m]] for r in body) + " |")
    full = [r for r in body if to_bytes(r[col["gpu__time_duration.sum"]], "byte") * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]]] > 1.0]
    alg = traj * 32.0 * 2 ** 20
    rd = sum(to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) for r in full) / max(len(full), 1)
    wr = sum(to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) for r in full) / max(len(full), 1)
    out += ["", f"Full sweeps in this capture: {len(full)}.  Mean DRAM traffic per full sweep: read {rd / 1e6:.1f} MB + write "
            f"{wr / 1e6:.1f} MB = {(rd + wr) / alg:.3f} x the algorithmic bytes (the 126 MB L2 keeps part of the previous "
            "sweep's tail; nothing is re-read)."]
    md = os.path.join(ROOT, "profiles", f"{tag}_tile_kernel_ncu_full.md")
    open(md, "w").write("\n".join(out) + "\n")
    json.dump({"kernel": name, "capture": os.path.relpath(md, ROOT), "trajectories_per_launch": traj, "dram_bytes_read": rd,
               "dram_bytes_write": wr, "algorithmic_bytes": alg, "traffic_over_algorithmic": (rd + wr) / alg},
              open(os.path.join(ROOT, "profiles", "tile_kernel_traffic.json"), "w"), indent=1)
    print(open(md).read())


if __name__ == "__main__":
    main()
