"""Turns an .ncu-rep (ncu --set full) into the short per-kernel summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1->XBAR request cycles"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{path.split('/')[-1]}` (ncu --set full --clock-control none)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        # warp-state samples (smsp__pcsamp_warps_issue_stalled_*): where the resident warps spend their cycles
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        total = sum(v for v, _ in stalls) or 1.0
        stalls.sort(reverse=True)
        print("| warp states (% of pc samples; `selected` = issuing) | " +
              ", ".join(f"{n} {100 * v / total:.0f}" for v, n in stalls[:7]) + " |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
