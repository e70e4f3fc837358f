"""Markdown table of the figures DESIGN.md quotes from an `ncu --set full` capture.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep            # one row per profiled launch

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", 1e-3, "{:.1f}"),            # ns -> us
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("L1tex/LSU %", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("dram rd MB", "dram__bytes_read.sum", None, "{:.1f}"),
    ("dram wr MB", "dram__bytes_write.sum", None, "{:.1f}"),
    ("regs", "launch__registers_per_thread", 1, "{:.0f}"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
]
STALLS = ["long_scoreboard", "wait", "barrier", "lg_throttle", "math_pipe_throttle", "mio_throttle", "short_scoreboard",
          "not_selected", "no_instruction", "branch_resolving", "membar", "sleeping"]
UNIT_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
UNIT_TO_NS = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " | top stalls (warps per issue) |")
    print("|---|" + "---:|" * len(COLS) + "---|")
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for _label, key, scale, fmt in COLS:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[ix[key]].replace(",", ""))
            u = units[ix[key]]
            if key == "gpu__time_duration.sum":
                v = v * UNIT_TO_NS.get(u, 1.0) * 1e-3
            elif scale is None:
                v = v * UNIT_TO_MB.get(u, 1e-6)
            else:
                v = v * scale
            cells.append(fmt.format(v))
        st = []
        for sname in STALLS:
            key = f"smsp__average_warps_issue_stalled_{sname}_per_issue_active.ratio"
            if key in ix and r[ix[key]] not in ("", "n/a"):
                st.append((float(r[ix[key]].replace(",", "")), sname))
        st.sort(reverse=True)
        print(f"| `{name}` | " + " | ".join(cells) + " | " + ", ".join(f"{n} {v:.1f}" for v, n in st[:3]) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
