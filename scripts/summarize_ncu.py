"""Summarise an `ncu --set full` capture of the sweep kernel into profiles/ (markdown table + traffic json).

usage: python scripts/summarize_ncu.py <capture.ncu-rep> <round tag, e.g. r01> <kernel: direct|tile|spec> "<command line captured>"

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed).  Every launch sweeps the state of the trajectories
in its work list once; the number of tiles comes from the grid (direct kernel: 16 tiles of 2^11 amplitudes per CTA).
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
    rep, tag, kern, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    match = {"direct": "direct_kernel", "tile": "tile_kernel", "spec": "ncb_seg_"}[kern]
    body = [r for r in body if match in r[col["Kernel Name"]]]
    name = body[0][col["Kernel Name"]]
    tiles_per_cta = 16 if kern in ("direct", "spec") else None
    tile_bytes = 2 * 16 * 2 ** 11          # one read + one write of a 2^11-amplitude tile
    out = [f"# ncu --set full capture of the {kern} sweep kernel ({tag})", "", f"Command (B200, gpurun): `{cmd}`", "",
           f"Kernel: `{name}`; 20-qubit depth-50 Heron circuit (BASELINE configs[4]), engine batches of 296 trajectories "
           "(296 x 16 MiB state).  With the shared no-jump prefix a segment's launch covers the trajectories that have "
           "already had their first jump (plus the trunk), so launches grow along the circuit.", ""]
    out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(body))) + " |")
    out.append("|---|---|" + "---|" * len(body))
    for m in METRICS:
        if m in col:
            out.append(f"| {m} | {units[col[m]]} | " + " | ".join(r[col[m]] for r in body) + " |")
    rd = wr = alg = 0.0
    if tiles_per_cta:
        for r in body:
            grid = float(r[col["launch__grid_size"]].replace(",", ""))
            alg += grid * tiles_per_cta * tile_bytes
            rd += to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr += to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        out += ["", f"Algorithmic bytes of these launches (SURVEY 8(d): 2 x 16 B per amplitude per sweep; tiles = grid x "
                f"{tiles_per_cta}, exact to one CTA): {alg / 1e6:.1f} MB.  DRAM traffic (ncu): read {rd / 1e6:.1f} MB + write "
                f"{wr / 1e6:.1f} MB = {(rd + wr) / alg:.3f} x the algorithmic bytes."]
    md = os.path.join(ROOT, "profiles", f"{tag}_{kern}_kernel_ncu_full.md")
    open(md, "w").write("\n".join(out) + "\n")
    if tiles_per_cta:
        json.dump({"kernel": name, "capture": os.path.relpath(md, ROOT), "launches": len(body), "dram_bytes_read": rd,
                   "dram_bytes_write": wr, "algorithmic_bytes": alg, "traffic_over_algorithmic": (rd + wr) / alg},
                  open(os.path.join(ROOT, "profiles", "tile_kernel_traffic.json"), "w"), indent=1)
    print(open(md).read())


if __name__ == "__main__":
    main()
