#!/usr/bin/env python
"""Turns the raw files a GPU call left in gpurun_out/ into the committed summaries under profiles/ (no GPU needed).

usage: python scripts/make_profiles.py <round tag, e.g. r02> [--launches gpurun_out/x.csv] [--seg rep] [--ev rep] [--dm rep]
"""
import argparse
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def raw_page(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


def source_hotspots(rep, top=12):
    """Per-opcode share of executed instructions and of stall samples, and the top stall reasons (page source)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    if not blocks:
        return []
    b = blocks[0]
    hdr = b["rows"][0]
    body = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, byop, samples, execd = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    for r in body:
        toks = r[idx["Source"]].split()
        op = toks[0] if toks else "?"
        if op.startswith("@") and len(toks) > 1:
            op = toks[1]
        op = op.split(".")[0]
        samples[op] += int(r[idx["# Samples"]] or 0)
        execd[op] += int(r[idx["Instructions Executed"]] or 0)
        for c in stall_cols:
            v = int(r[idx[c]] or 0)
            tot[c] += v
            byop[op][c] += v
    T, E = max(sum(samples.values()), 1), max(sum(execd.values()), 1)
    out = [f"Source-level view of `{b['name']}` ({len(body)} SASS instructions, {T} stall samples).", "",
           "Stall reasons (share of all samples): " + ", ".join(f"{c[6:]} {100 * v / T:.1f} %" for c, v in tot.most_common(9)), "",
           "| opcode | executed % | stall samples % | top stall reasons of its samples |", "|---|---|---|---|"]
    for op, s in execd.most_common(top):
        out.append(f"| {op} | {100 * s / E:.1f} | {100 * samples[op] / T:.1f} | " +
                   ", ".join(f"{c[6:]} {100 * v / max(samples[op], 1):.0f} %" for c, v in byop[op].most_common(3)) + " |")
    return out


def ncu_table(rep, pattern, title, command, note, out_path, tile_bytes=None, tiles_per_cta=None):
    hdr, units, body = raw_page(rep)
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in body if re.search(pattern, r[col["Kernel Name"]])]
    out = [f"# {title}", "", f"Command (B200, gpurun): `{command}`", "", note, "",
           "Kernels: " + ", ".join(f"`{r[col['Kernel Name']].split('(')[0]}`" for r in body), "",
           "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(body))) + " |", "|---|---|" + "---|" * len(body)]
    for m in METRICS:
        if m in col:
            out.append(f"| {m} | {units[col[m]]} | " + " | ".join(r[col[m]] for r in body) + " |")
    traffic = []
    for r in body:
        rd = float(r[col["dram__bytes_read.sum"]]) * SCALE.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr = float(r[col["dram__bytes_write.sum"]]) * SCALE.get(units[col["dram__bytes_write.sum"]], 1.0)
        t = float(r[col["gpu__time_duration.sum"]]) * TIME.get(units[col["gpu__time_duration.sum"]], 1.0)
        alg = None
        if tile_bytes and tiles_per_cta:
            alg = float(r[col["launch__grid_size"]]) * tiles_per_cta * tile_bytes
        traffic.append({"kernel": r[col["Kernel Name"]].split("(")[0], "dram_bytes": rd + wr, "seconds": t, "dram_gbs": (rd + wr) / t / 1e9,
                        "algorithmic_bytes": alg})
    out += ["", "Per launch: " + "; ".join(
        f"{x['dram_bytes'] / 1e6:.1f} MB DRAM in {x['seconds'] * 1e3:.3f} ms = {x['dram_gbs']:.0f} GB/s" +
        (f" (algorithmic {x['algorithmic_bytes'] / 1e6:.1f} MB, traffic / algorithmic = {x['dram_bytes'] / x['algorithmic_bytes']:.3f})" if x["algorithmic_bytes"] else "")
        for x in traffic), ""]
    out += source_hotspots(rep)
    open(out_path, "w").write("\n".join(out) + "\n")
    return traffic


def launch_summary(src, tag, chunk):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr, body = rows[h], rows[h + 1:]
    kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    tot, cnt = collections.Counter(), collections.Counter()
    per = collections.defaultdict(lambda: [0, 0.0, 0])
    for r in body:
        name = r[kn].split("(")[0].strip()
        ns = float(r[mv].replace(",", ""))
        grid = int(r[gs].strip("()").split(",")[0])
        m = re.match(r"ncb_seg_(\d+)(_ev)?$", name)
        if m and m.group(2):
            name = "ncb_seg_<k>_ev [generated: works with jump events]"
        elif m:
            name = "ncb_seg_<k> [generated: whole-segment sweeps]"
            p = per[int(m.group(1))]
            p[0] += grid; p[1] += ns; p[2] += 1
        tot[name] += ns
        cnt[name] += 1
    total = sum(tot.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launch_list_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct", "avg_us"])
        for name, ns in tot.most_common():
            w.writerow([name, cnt[name], f"{ns / 1e6:.3f}", f"{100 * ns / total:.2f}", f"{ns / cnt[name] / 1e3:.1f}"])
    # per-segment table of the whole-segment kernels (grid x chunk tiles of 2^11 amplitudes, 2 x 16 B each)
    import bench
    from noisycircuits_b200 import _lib
    from noisycircuits_b200.program import PackedCircuit, Program
    sq, tq, _meas, _pairs, instr = bench.workload()
    prog = Program(PackedCircuit(bench.NQ, instr, sq, tq), _lib.MODE_MCWF, specialize=True)
    desc = prog.describe().splitlines()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6536.7
    with open(os.path.join(ROOT, "profiles", f"{tag}_per_segment.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["segment", "instructions", "passes", "fast_list_ops", "launches", "tiles", "ns_per_tile", "GB_per_s", "frac_of_measured_hbm_peak"])
        for k in sorted(per):
            grid, ns, n = per[k]
            tiles = grid * chunk
            d = desc[1 + k]
            m = re.search(r"instr\[(\d+),(\d+)\)", d)
            ops = re.findall(r"\[([^\]]*)\]", d.split("ops")[1]) if "ops" in d else []
            gbs = tiles * 2048 * 32 / ns
            w.writerow([k, int(m.group(2)) - int(m.group(1)), int(re.search(r"passes (\d+)", d).group(1)), sum(len(o.split()) for o in ops), n, tiles,
                        f"{ns / tiles:.2f}", f"{gbs:.0f}", f"{gbs / peak:.3f}"])
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launch_list_summary.csv")).read())


def simple_launch_summary(src, out_path, last=None):
    """kernel, launches, total, share, average of an ncu launch list (generated per-segment kernels grouped); last: only the
    last `last` launches (one call of the profiled script)."""
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr, body = rows[h], rows[h + 1:]
    if last:
        body = body[-last:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in body:
        name = re.sub(r"ncb_seg_\d+", "ncb_seg_<k>", r[kn].split("(")[0].strip())
        name = re.sub(r"<.*", "<...>", name) if name.startswith("void cub::") else name
        tot[name] += float(r[mv].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    with open(out_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
        for name, ns in tot.most_common():
            w.writerow([name, cnt[name], f"{ns / 1e3:.1f}", f"{100 * ns / total:.2f}", f"{ns / cnt[name] / 1e3:.1f}"])
    print(open(out_path).read())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--chunk", type=int, default=8, help="tiles per CTA of the whole-segment kernels in the captured run")
    ap.add_argument("--seg")
    ap.add_argument("--ev")
    ap.add_argument("--dm")
    ap.add_argument("--res", help="ncu-rep of ncb_res (generated resident kernel, scripts/prof_c3.py)")
    ap.add_argument("--c3-launches", help="ncu launch list of scripts/prof_c3.py")
    ap.add_argument("--c4-launches", help="ncu launch list of scripts/prof_c4.py")
    ap.add_argument("--command", default="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0")
    a = ap.parse_args()
    P = os.path.join(ROOT, "profiles")
    if a.launches:
        launch_summary(a.launches, a.tag, a.chunk)
    if a.seg:
        tr = ncu_table(a.seg, r"ncb_seg_\d+$", f"ncu --set full capture of the generated whole-segment sweep kernels ({a.tag})",
                       "ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+$' -s 30 -c 3 " + a.command,
                       "20-qubit depth-50 Heron circuit (BASELINE configs[4]), engine batches of 296 trajectories (296 x 16 MiB state); R = 4 (16 amplitudes "
                       "per thread, 128-thread CTAs, four per SM), tile 2^11 amplitudes, " + str(a.chunk) + " tiles per CTA.",
                       os.path.join(P, f"{a.tag}_seg_kernel_ncu_full.md"), tile_bytes=2 * 16 * 2048, tiles_per_cta=a.chunk)
        json.dump({"source": f"profiles/{a.tag}_seg_kernel_ncu_full.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the captured launches)",
                   "dram_bytes_per_launch": sum(x["dram_bytes"] for x in tr) / len(tr),
                   "algorithmic_bytes_per_launch": sum(x["algorithmic_bytes"] for x in tr) / len(tr),
                   "traffic_over_algorithmic": sum(x["dram_bytes"] for x in tr) / sum(x["algorithmic_bytes"] for x in tr), "launches": tr},
                  open(os.path.join(P, f"{a.tag}_traffic.json"), "w"), indent=1)
    if a.ev:
        ncu_table(a.ev, r"ncb_seg_\d+_ev", f"ncu --set full capture of the generated event-piece kernels ({a.tag})",
                  "ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+_ev' -s 60 -c 3 " + a.command,
                  "The works of a segment that carry a jump event (pieces [lo, e], in-line jumps, reductions): about 6 % of a segment's works.",
                  os.path.join(P, f"{a.tag}_ev_kernel_ncu_full.md"))
    if a.dm:
        ncu_table(a.dm, r"tile_kernel", f"ncu --set full capture of the density-matrix sweeps ({a.tag})",
                  "ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 31 -c 4 python scripts/prof_dm.py",
                  "BASELINE configs[1]: 12-qubit density matrix (4^12 vectorised state, 268 MB), depth-10 layered Heron circuit; "
                  "tile_kernel<4, DENSE2|DENSE4, 256, 1>.", os.path.join(P, f"{a.tag}_dm_kernel_ncu_full.md"))

    if a.res:
        ncu_table(a.res, r"ncb_res$", f"ncu --set full capture of the generated resident kernel ({a.tag})",
                  "ncu --set full --clock-control none --import-source on -k 'regex:ncb_res$' -s 3 -c 1 python scripts/prof_c3.py",
                  "BASELINE configs[2]: 12-qubit QNN circuit (284 instructions), 10^4 trajectories per call; the 19 % of the trajectories that "
                  "leave the jump-free trunk run from the trunk's checkpoint before their first flagged draw, one CTA (512 threads, 2^12 "
                  "amplitudes in shared memory) per trajectory at a time, two CTAs per SM.  The state never touches HBM: the kernel is bound by "
                  "the FP64 pipe and the shared-memory exchanges between passes.", os.path.join(P, f"{a.tag}_res_kernel_ncu_full.md"))
    if a.c3_launches:
        simple_launch_summary(a.c3_launches, os.path.join(P, f"{a.tag}_c3_launch_list.csv"), last=8)
    if a.c4_launches:
        simple_launch_summary(a.c4_launches, os.path.join(P, f"{a.tag}_c4_launch_list.csv"))


if __name__ == "__main__":
    main()
