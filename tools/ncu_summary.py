"""Summarise an .ncu-rep (raw + source pages) into a short text report."""
import csv
import subprocess
import sys


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for k in keys:
            if k in d:
                lines.append(f"{k:90s} {d[k]:>22s} {units[hdr.index(k)]}")
        lines.append("-- warp stall reasons (per issue-active cycle)")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                v = float(d[k] or 0)
                if v >= 0.01:
                    lines.append(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr)]

        def g(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        tot = sum(g(r, "# Samples") for r in data) or 1
        lines.append(f"-- top SASS instructions by samples (total {int(tot)})")
        for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:14]:
            stalls = {k: g(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
            lines.append(f"   {r[ix['Address']][-5:]} {r[ix['Source']][:52]:52s} {100 * g(r, '# Samples') / tot:5.1f}%  " +
                         " ".join(f"{k[6:]}={int(v)}" for k, v in top))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
