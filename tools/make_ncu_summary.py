"""Writes profiles/ncu_summary.json (what bench.py's roofline.traffic and roofline_int_alu_pipe read) and a per-kernel
digest from an ncu --set full report.   python tools/make_ncu_summary.py gpurun_out/r2n_prof.ncu-rep profiles/r2/r2n"""
import csv, json, os, re, subprocess, sys
rep, outbase = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(outbase + "_ncu_full_raw.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
summ = {"note": "per launch, from the ncu --set full capture " + os.path.basename(rep) + " (n = 2^23, PRP chain steady state); ncu reports MB = 1e6 bytes",
        "source": outbase + "_ncu_full_raw.csv", "dram_bytes_per_launch": {}, "alu_pipe_warp_instructions_per_launch": {},
        "warp_instructions_per_launch": {}, "kernel_us_under_ncu": {}, "issue_active_pct": {}, "alu_pipe_pct": {}, "fma_pipe_pct": {}, "fmaheavy_pipe_busy_pct": {}, "alu_pipe_busy_pct": {}, "stalls_pct": {}}
seen = set()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    key = "k_mid_fast" if "k_mid_fast" in name else ("k_inv_fast_T4" if "k_inv_fast" in name else name)
    if key in seen:
        continue
    seen.add(key)
    f = lambda k: float(d[k].replace(",", ""))
    unit = dict(zip(hdr, rows[1]))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    dram = f("dram__bytes_read.sum") * scale.get(unit["dram__bytes_read.sum"], 1.0) + f("dram__bytes_write.sum") * scale.get(unit["dram__bytes_write.sum"], 1.0)
    summ["dram_bytes_per_launch"][key] = dram
    summ["warp_instructions_per_launch"][key] = f("smsp__inst_executed.sum")
    alu = f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0
    summ["alu_pipe_warp_instructions_per_launch"][key] = alu * 2.0 * f("sm__cycles_active.avg") * 148
    summ["kernel_us_under_ncu"][key] = f("gpu__time_duration.sum")
    summ["issue_active_pct"][key] = f("smsp__issue_active.avg.pct_of_peak_sustained_active")
    summ["alu_pipe_pct"][key] = alu * 100
    summ["fma_pipe_pct"][key] = f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")
    # every IMAD* issues to the FMA-heavy half; the combined FMA figure averages it with the idle FMA-lite half
    summ["fmaheavy_pipe_busy_pct"][key] = f("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")
    summ["alu_pipe_busy_pct"][key] = f("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed")
    st = []
    for h in hdr:
        if re.match(r"smsp__pcsamp_warps_issue_stalled_\w+$", h) and "not_issued" not in h:
            try: st.append((float(d[h] or 0), h[33:]))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1
    summ["stalls_pct"][key] = {k: round(100 * v / tot, 1) for v, k in sorted(st, reverse=True)[:8]}
summ["alu_pipe_note"] = "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active x 2 warp-inst/clk/SM x sm__cycles_active.avg x 148 SMs; the ALU pipe (IADD3/SHF/LOP3/ISETP/SEL) issues one warp instruction every 2 clocks per SM sub-partition"
json.dump(summ, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_summary.json"), "w"), indent=1)
print(json.dumps(summ, indent=1))
