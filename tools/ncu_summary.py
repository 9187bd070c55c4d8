"""Summarise an .ncu-rep (raw page): per-kernel duration, DRAM bytes, throughput, occupancy, issue rate,
top stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--source N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:100])
    for w in want:
        if w in d:
            print(f"  {w:62s} {d[w]:>16s} {units[hdr.index(w)]}")
    st = []
    for h in hdr:
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                st.append((float(d[h]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("  stalls: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in st[:8]))
if "--source" in sys.argv:
    topn = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    # several kernels may follow each other; split on "Kernel Name" rows
    blocks, cur = [], None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks[:1]:
        h = b["rows"][0]
        ix = {k: i for i, k in enumerate(h)}
        data = b["rows"][1:]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
        inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
        print(f"-- source {b['name'][:80]}: {tot} samples, {inst} warp-instructions, {len(data)} SASS lines")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:topn]:
            print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:90]}")
