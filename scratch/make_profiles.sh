# Summaries of the round-2 ncu captures (gpurun_out/r02_tc_<prec>_c<cfg>.ncu-rep, taken on the REAL bench.py launch)
# -> profiles/r02_tc_<prec>_c<cfg>_ncu.txt and profiles/r02_ncu_metrics.json (read by bench.py)
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import json, subprocess, os
out = {}
ncurves = {"3": 8778, "5": 100000}
for cfg in ("3", "5"):
    for prec in ("f16x3f", "f16x3", "f16"):
        rep = f"gpurun_out/r02_tc_{prec}_c{cfg}.ncu-rep"
        if not os.path.exists(rep):
            continue
        txt = f"profiles/r02_tc_{prec}_c{cfg}_ncu.txt"
        r = subprocess.run(["python", "scratch/ncu_metrics.py", rep, txt, "steps_per_launch=10", f"n_curves={ncurves[cfg]}"],
                           capture_output=True, text=True)
        e = json.loads(r.stdout.strip().splitlines()[-1])
        e["steps_per_launch"] = int(e["steps_per_launch"]); e["n_curves"] = int(e["n_curves"])
        e["command"] = f"ncu --set full --clock-control none -k regex:tc_curve_kernel -s 3 -c 1 python bench.py --config {cfg} --steps 10 --warmup 3 --precision {prec} --no-cpu --no-other"
        out.setdefault(f"config{cfg}", {})[prec] = e
        print(cfg, prec, e.get("sm__pipe_tc_cycles_active_pct"), e.get("smsp__issue_active_pct"), e.get("dram_bytes_per_launch"))
json.dump(out, open("profiles/r02_ncu_metrics.json", "w"), indent=1)
PY
