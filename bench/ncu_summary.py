"""Summarise one kernel of an .ncu-rep (ncu -i ... --page raw --csv) into the few counters DESIGN.md argues with.

    python bench/ncu_summary.py gpurun_out/prof.ncu-rep "comment line" > profiles/rNN_kernel_ncu.csv
"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|l1tex__t_sector_hit_rate\.pct|"
    r"launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_static)|lts__t_sector_hit_rate\.pct|"
    r"lts__t_sectors_srcunit_tex_op_(read|write)\.sum|sm__cycles_elapsed\.avg|"
    r"sm__pipe_(fp64|alu|fma)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio)$")


def main():
    rep, comment = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units = rows[0], rows[1]
    for vals in rows[2:]:
        kname = vals[names.index("Kernel Name")] if "Kernel Name" in names else ""
        print(f"# {comment} [{kname}]")
        for n, u, v in zip(names, units, vals):
            if not KEEP.match(n):
                continue
            try:
                if "stalled" in n and float(v.replace(",", "")) < 0.2:
                    continue
            except ValueError:
                pass
            print(f"{n},{u},{v}")


if __name__ == "__main__":
    main()
