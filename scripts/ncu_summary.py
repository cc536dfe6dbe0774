"""Summarise an .ncu-rep as markdown: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/x.md"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
cols = [("Kernel Name", "kernel", str), ("gpu__time_duration.sum", "time", float), ("launch__registers_per_thread", "regs", float),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", float),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst", float),
        ("sm__inst_executed.avg.per_cycle_active", "IPC", float),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%", float),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%", float), ("lts__t_sector_hit_rate.pct", "L2hit%", float),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1tex%", float),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%", float),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM%", float),
        ("dram__bytes_read.sum", "dramRd", float), ("dram__bytes_write.sum", "dramWr", float),
        ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM /s", float), ("dram__bytes_read.sum.per_second", "DRAM rd /s", float),
        ("SM_B.TriageCompute.l1tex__t_sectors.sum", "L1 sectors", float),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st:long_sb", float),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st:no_inst", float),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st:not_sel", float),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st:wait", float),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st:math", float),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st:branch", float)]
idx = [(h.index(c), n, f) for c, n, f in cols if c in h]
units = rows[1]
print(f"# {title}\n\nSource: `{rep}` (`ncu --set full --clock-control none`, one B200). Units: time in {units[h.index('gpu__time_duration.sum')]}; dram bytes in {units[h.index('dram__bytes_read.sum')]}; `L2->SM /s` (crossbar bytes delivered to the SMs' L1) in {units[h.index('l1tex__m_xbar2l1tex_read_bytes.sum.per_second')] if 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second' in h else '-'}, `DRAM rd /s` in {units[h.index('dram__bytes_read.sum.per_second')] if 'dram__bytes_read.sum.per_second' in h else '-'}; `L1 sectors` = 32-byte sectors looked up in L1 (x 32 / time = bytes/s served through L1).\n")
print("| # | " + " | ".join(n for _, n, _ in idx) + " |")
print("|---|" + "|".join("---" for _ in idx) + "|")
for k, r in enumerate(rows[2:]):
    vals = []
    for i, n, f in idx:
        v = r[i]
        if f is float:
            try: v = f"{float(v.replace(',', '')):.2f}"
            except ValueError: pass
        else:
            v = v.split("(")[0].replace("void ", "").replace("jpbrt::", "")
        vals.append(v)
    print(f"| {k} | " + " | ".join(vals) + " |")
