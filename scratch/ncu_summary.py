"""Summarise an ncu report: key raw metrics, stall reasons, opcode histogram, hottest source lines."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
print("== key metrics")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); S = I = 0
op = collections.Counter(); ops = collections.Counter()
for r in data:
    try:
        s = int(r[ix["# Samples"]]); n = int(r[ix["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    S += s; I += n
    for c in stall_cols:
        try: tot[c] += int(r[ix[c]])
        except ValueError: pass
    t = r[ix["Source"]].strip().split()
    name = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    op[name] += n; ops[name] += s
print(f"== stall reasons (of {S} samples; {I} warp-instructions)")
for k, v in tot.most_common(10):
    print(f"{k:26s} {100*v/S:5.1f}%")
print("== opcodes: % of executed warp-instructions / % of samples")
for k, v in op.most_common(22):
    print(f"{k:10s} {100*v/I:5.1f}% {100*ops[k]/S:5.1f}%")
