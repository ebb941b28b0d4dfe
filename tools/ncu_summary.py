"""Text summary of one kernel of an .ncu-rep (the metrics SURVEY.md 8d names + stall reasons + where the samples fall):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "header line" > profiles/x.txt        (needs ncu on PATH, no GPU)"""
import csv, io, subprocess, sys
rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
units = dict(zip(rows[0], rows[1]))
print(header)
print("kernel:", d.get("Kernel Name"), " grid", d.get("launch__grid_size"), "x", d.get("launch__block_size"))
print()
keys = ["gpu__time_duration.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in d:
        print(f"{k:88s} {d[k]:>20s} {units.get(k, '')}")
print()
for k in sorted(d):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        print(f"{k:88s} {d[k]:>20s}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(int(r[ia], 16), int(r[ie]), int(r[isamp]), r[hdr.index("Source")].strip()) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot, tots = sum(x[1] for x in data), sum(x[2] for x in data)
dpx = sum(x[1] for x in data if x[3].startswith(("VIADDMNMX", "VIMNMX")))
print()
print(f"instructions executed (warp level) {tot}, of which VIADDMNMX / VIMNMX3 {dpx} = {dpx / tot:.3f}")
mx = sorted(x[1] for x in data)[-200]
print("share of the executed instructions / of the stall samples by code region (hot = the unrolled step body):")
prev = None; regions = []
for a, e, sm, s in data:
    lvl = "spin" if e > 3 * mx else ("hot" if e > 0.9 * mx else ("mid" if e > 0.02 * mx else "low"))
    if lvl != prev:
        if prev is not None: regions.append((prev, start, a, acc, accs))
        prev, start, acc, accs = lvl, a, 0, 0
    acc += e; accs += sm
regions.append((prev, start, a, acc, accs))
agg = {}
for lvl, s0, s1, acc, accs in regions:
    x = agg.setdefault(lvl, [0, 0]); x[0] += acc; x[1] += accs
for lvl, (acc, accs) in agg.items():
    print(f"  {lvl:5s} inst {acc / tot:.4f}  samples {accs / tots:.4f}")
