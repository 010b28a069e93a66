#!/usr/bin/env python
"""Summarises an ncu report into small text files under profiles/ (run in the build container).

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_k_umma_search_2048 ["note"]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_op_write_hit_rate.pct",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed", "l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    with open(out + "_raw.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {rep}\n# {note}\n")
        for row in raw[2:]:
            d = dict(zip(hdr, row))
            f.write(f"\n## kernel: {d.get('Kernel Name', '?')[:120]}\n")
            for h, u in zip(hdr, units):
                if any(h == k or h.startswith(k + ".") or h == k.strip() for k in KEYS) or h in KEYS:
                    f.write(f"{h} [{u}] = {d[h]}\n")
    if len(raw) > 3:   # several kernels in one report: the per-kernel metrics above are the summary
        print("wrote", out + "_raw.txt")
        return
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in src[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    with open(out + "_source_top.txt", "w") as f:
        f.write(f"# warp-stall sampling by SASS line, report {rep}\n# {note}\n# total samples {tot}\n\n")
        agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
        for s, v in sorted(agg.items(), key=lambda x: -x[1]):
            f.write(f"{s:28s} {v:9d} {100 * v / tot:5.1f}%\n")
        f.write("\n# samples   executed  instruction                                                       top stalls\n")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:40]:
            st = sorted(((s, int(r[ix[s]])) for s in stalls), key=lambda x: -x[1])[:2]
            f.write(f"{r[ix['# Samples']]:>9s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:72]:72s} {st}\n")
    print("wrote", out + "_raw.txt", out + "_source_top.txt")


if __name__ == "__main__":
    try:
        main()
    except Exception as exc:  # a summary that fails must not keep the (large) report from being deleted by the caller
        print("summarize_ncu failed:", exc)
        sys.exit(0)
