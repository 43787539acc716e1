#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + per-opcode instruction / shared-wavefront / stall mix."""
import collections, csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout

def main(path, per=None):
    raw = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    h, r = raw[0], raw[2]
    def g(name):
        return r[h.index(name)] if name in h else "n/a"
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
    for k in keys:
        print(f"{k:72s} {g(k)} {raw[1][h.index(k)] if k in h else ''}")
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    hh = src[1]; ix = {n: i for i, n in enumerate(hh)}
    def f(row, n):
        try: return float(row[ix[n]])
        except Exception: return 0.0
    data = src[2:]
    tot = sum(f(x, "# Samples") for x in data) or 1
    st = collections.Counter()
    for x in data:
        for n in hh:
            if n.startswith("stall_") and "(Not" not in n: st[n] += f(x, n)
    print("stalls:", {k: round(v / tot, 3) for k, v in st.most_common(9)})
    op = collections.defaultdict(lambda: [0, 0, 0, 0])
    for x in data:
        s = x[ix["Source"]].split()
        o = s[1] if s[0].startswith("@") else s[0]
        a = op[o]; a[0] += f(x, "Instructions Executed"); a[1] += f(x, "L1 Wavefronts Shared"); a[2] += f(x, "L1 Wavefronts Shared Excessive"); a[3] += f(x, "# Samples")
    div = float(per) if per else 1.0
    n_inst = sum(v[0] for v in op.values())
    print(f"total warp-instr {n_inst:.0f}  per-unit {n_inst/div:.1f}")
    for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"  {k:22s} inst/unit={v[0]/div:9.1f} shared_wf/unit={v[1]/div:8.1f} excess={v[2]/div:7.1f} samples={v[3]/tot:.3f}")

if __name__ == "__main__":
    main(*sys.argv[1:])
