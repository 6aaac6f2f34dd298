#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + SASS-level aggregation by barrier-delimited phase.
usage: tools/ncu_summary.py <report.ncu-rep> [nsm=148] [ksteps=16] [--json out.json --envs B]
--json writes the DRAM traffic of the captured launch (what bench.py reports as roofline.traffic)."""
import collections, csv, io, json, re, subprocess, sys
argv = [a for a in sys.argv[1:]]
json_out = envs = None
if "--json" in argv:
    i = argv.index("--json"); json_out = argv[i + 1]; del argv[i:i + 2]
if "--envs" in argv:
    i = argv.index("--envs"); envs = int(argv[i + 1]); del argv[i:i + 2]
rep = argv[0]
nsm = int(argv[1]) if len(argv) > 1 else 148
ksteps = int(argv[2]) if len(argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
if json_out:
    def _bytes(name):
        i = hdr.index(name); v = float(vals[i].replace(",", "")); u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    json.dump({"kernel": vals[hdr.index("Kernel Name")], "dram_bytes_read": _bytes("dram__bytes_read.sum"),
               "dram_bytes_write": _bytes("dram__bytes_write.sum"), "grid_envs": envs, "fused_steps": ksteps,
               "gpu_time_ms": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(units[hdr.index("gpu__time_duration.sum")], 1),
               "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)"}, open(json_out, "w"), indent=1)
for k in want:
    if k in hdr:
        i = hdr.index(k); print(f"{k:72s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot_i = tot_s = 0; stall = collections.Counter(); opc = collections.Counter()
segs = []; cur = dict(n=0, inst=0, samp=0, ops=collections.Counter())
for r in rows[2:]:
    if len(r) < len(h): continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
    op = m.group(2) if m else "?"
    ie, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    tot_i += ie; tot_s += s; opc[op] += ie
    for c in stall_cols: stall[c] += int(r[ix[c]])
    cur["n"] += 1; cur["inst"] += ie; cur["samp"] += s; cur["ops"][op] += ie
    if op == "BAR":
        segs.append(cur); cur = dict(n=0, inst=0, samp=0, ops=collections.Counter())
segs.append(cur)
per = tot_i / nsm / ksteps
print(f"\nwarp-instructions: total {tot_i}  per SM per step {per:.0f}  per thread per step {per/16:.0f} (16 warps/CTA)")
print("stall samples:", ", ".join(f"{k[6:]}={v} ({100*v/tot_s:.0f}%)" for k, v in stall.most_common(9)))
print("opcode mix   :", ", ".join(f"{k}={100*v/tot_i:.1f}%" for k, v in opc.most_common(16)))
print("phases between barriers (static instrs | % executed instrs | % stall samples | top opcodes):")
for i, s in enumerate(segs):
    if s["inst"] * 200 < tot_i: continue
    print(f"  {i:2d} {s['n']:5d} | {100*s['inst']/tot_i:5.1f}% | {100*s['samp']/tot_s:5.1f}% | " + ", ".join(f"{k}:{v*100//max(1,s['inst'])}%" for k, v in s["ops"].most_common(6)))
