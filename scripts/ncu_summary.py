"""Turn a .ncu-rep (read with `ncu -i ... --page raw --csv`) into the short text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_final.ncu-rep profiles/r1_final_grad_umma_ncu.txt "note"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
    "sm__inst_executed_pipe_tc.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_op_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none, read with `ncu -i %s --page raw --csv`\n" % rep.split("/")[-1])
        if note:
            f.write("# %s\n" % note)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\nkernel: %s\n" % d.get("Kernel Name", "?").split("(")[0])
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write("  %-78s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            # stall reasons (top 6 by value)
            st = [(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and
                  h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
            if not st:
                st = [(float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith(".ratio") and d[h] not in ("", "n/a")]
            for v, h in sorted(st, reverse=True)[:6]:
                f.write("  %-78s %.3f\n" % (h, v))


if __name__ == "__main__":
    main()
