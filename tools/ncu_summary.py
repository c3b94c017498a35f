"""Condense an ncu report (--page raw --csv) into a per-kernel table for profiles/."""
import csv, collections, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"),
        ("smsp__inst_executed.sum", "warp_insts"),
        ("l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "smem_bank_rd%"),
        ("l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smem_bank_wr%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_lsu_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("launch__grid_size", "grid")]
agg = collections.OrderedDict()
for d in data:
    key = d[idx["Kernel Name"]][:70]
    agg.setdefault(key, []).append(d)
with open(out, "w") as f:
    f.write(f"# ncu --set full summary of {rep} (averages over the captured launches; --clock-control none)\n\n")
    for key, ds in agg.items():
        f.write(f"## {key}  ({len(ds)} launches)\n")
        for c, name in cols:
            if c not in idx:
                continue
            vals = []
            for d in ds:
                try:
                    vals.append(float(d[idx[c]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                f.write(f"- {name:14s} {sum(vals)/len(vals):14.3f} {units[idx[c]]}   ({c})\n")
        f.write("\n")
print(open(out).read())
