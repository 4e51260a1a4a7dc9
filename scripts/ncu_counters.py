"""Extract the per-launch counters bench.py quotes (DRAM traffic, FP32 instruction mix) from a full-size ncu capture
of the solver kernel into profiles/solver_traffic.json.   python scripts/ncu_counters.py <rep> <out.json> <note>"""
import csv, io, json, subprocess, sys

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
h, u, v = raw[0], raw[1], raw[-1]


def get(k):
    i = h.index(k)
    x = float(v[i])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u[i], 1.0)
    return x * scale


cycles = get("sm__cycles_elapsed.max")
ffma, fmul, fadd = (get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") for op in ("ffma", "fmul", "fadd"))
d = {
    "source": rep, "note": note, "kernel": v[h.index("Kernel Name")], "grid": v[h.index("launch__grid_size")],
    "duration_s_under_ncu": get("gpu__time_duration.sum"),
    "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
    "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
    "fp32_thread_inst_per_cycle": {"ffma": ffma, "fmul": fmul, "fadd": fadd},
    "fp32_flop_per_launch": (2 * ffma + fmul + fadd) * cycles,
    "fp32_flop_per_cycle": 2 * ffma + fmul + fadd, "fp32_peak_flop_per_cycle": 2 * get("sm__sass_thread_inst_executed_op_ffma_pred_on.sum.peak_sustained"),
    "issue_active_per_smsp": get("smsp__issue_active.avg.per_cycle_active"), "warps_active_per_sm": get("sm__warps_active.avg.per_cycle_active"),
    "registers_per_thread": get("launch__registers_per_thread"), "warp_inst_executed": get("smsp__inst_executed.sum"),
}
d["fp32_frac_of_peak_ncu"] = d["fp32_flop_per_cycle"] / d["fp32_peak_flop_per_cycle"]
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
