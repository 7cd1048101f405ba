"""Summarise an ncu report of a step kernel into (a) a text file for profiles/ and (b) a dict of the numbers
bench.py quotes.   python scratch/ncu_metrics.py rep.ncu-rep out.txt [key=value ...]  -> prints the JSON entry"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
extra = dict(kv.split("=", 1) for kv in sys.argv[3:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {"gpu__time_duration.sum": "gpu_time", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed": "sm__pipe_tc_cycles_active_pct",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "sm__pipe_tensor_cycles_active_realtime_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "smsp__issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "sm__warps_active_pct",
        "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "launch__registers_per_thread": "registers", "launch__shared_mem_per_block_dynamic": "smem_dynamic",
        "smsp__inst_executed.sum": "warp_instructions", "sm__cycles_elapsed.max": "sm_cycles"}
lines, entries = [], []
for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    e = {"kernel": d.get("Kernel Name", "?")}
    lines.append(f"== {e['kernel']}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for k, name in want.items():
        if k in d:
            lines.append(f"{k} [{u[k]}] = {d[k]}")
            try:
                e[name] = float(d[k].replace(",", ""))
            except ValueError:
                e[name] = d[k]
            e[name + "_unit"] = u[k]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if "dram_read" in e:
        e["dram_bytes_per_launch"] = e["dram_read"] * scale.get(e["dram_read_unit"], 1) + e["dram_write"] * scale.get(e["dram_write_unit"], 1)
        lines.append(f"dram bytes per launch (read + write) = {e['dram_bytes_per_launch']:.4g}")
    entries.append(e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[1]; ix = {x: i for i, x in enumerate(h)}
    stall = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    tot = {x: 0 for x in stall}; S = I = 0; ops = {}
    for r in rows[2:]:
        try:
            sm = int(r[ix["# Samples"]]); n = int(r[ix["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        S += sm; I += n
        for x in stall:
            try: tot[x] += int(r[ix[x]])
            except ValueError: pass
        t = r[ix["Source"]].strip().split()
        if t:
            name = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[name] = ops.get(name, 0) + n
    lines.append(f"== warp-state samples (last kernel): {S} samples, {I} warp-instructions")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        lines.append(f"{k:28s} {100 * v / max(S, 1):5.1f} %")
    lines.append("== opcode mix (% of executed warp-instructions)")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
        lines.append(f"{k:12s} {100 * v / max(I, 1):5.1f} %")
open(out, "w").write("\n".join(lines) + "\n")
e = entries[-1] if entries else {}
e.update(extra)
print(json.dumps(e))
