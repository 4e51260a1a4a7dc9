"""Text summary of an .ncu-rep capture of the solver kernel (run in the build container: `ncu -i` needs no GPU).

    python scripts/ncu_report.py gpurun_out/prof_X.ncu-rep > profiles/rNN_X.txt
"""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main(rep):
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = raw[0], raw[1], raw[-1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__waves_per_multiprocessor", "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum"]
    print(f"# ncu summary of {rep}")
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:70s} {vals[i]:>20s} {units[i]}")
    sass = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h2 = sass[1]
    ix = {h: i for i, h in enumerate(h2)}
    data = [r for r in sass[2:] if len(r) > ix["# Samples"]]
    F = lambda r, k: float(r[ix[k]] or 0)
    ti = sum(F(r, "Instructions Executed") for r in data)
    tt = sum(F(r, "Thread Instructions Executed") for r in data)
    ts = sum(F(r, "# Samples") for r in data)
    print(f"\nSASS instructions in kernel: {len(data)}   warp-instructions executed: {ti:.4g}   avg active threads/instr: {tt / ti:.2f}")
    hist = collections.Counter()
    for r in data:
        n = F(r, "Instructions Executed")
        if n:
            hist[int(F(r, "Avg. Threads Executed") // 4) * 4] += n
    print("active-thread histogram (share of executed warp-instructions):")
    for k in sorted(hist):
        print(f"  {k:2d}-{k + 3:2d} threads: {100 * hist[k] / ti:5.1f}%")
    stalls = {s: sum(F(r, s) for r in data) for s in h2 if s.startswith("stall_") and "Not Issued" not in s}
    print("warp stall sampling (all samples):")
    for s, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]:
        print(f"  {s:26s} {100 * v / ts:5.1f}%")
    ops = collections.Counter()
    for r in data:
        o = [x for x in r[ix["Source"]].split() if not x.startswith("@")]
        ops[o[0].split(".")[0]] += F(r, "Instructions Executed")
    print("opcode mix (executed warp-instructions): " + ", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in ops.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1])
