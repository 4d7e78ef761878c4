"""Summarises ncu captures brought back in gpurun_out/ into small tracked text files under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/launches_X.md
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep profiles/prof_X.md
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("sgqn::", "")
    return name[:110]


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            if "spin_kernel" in r["Kernel Name"]:      # torch.cuda._sleep: bench.py parks the stream behind it in its per-launch timing pass
                continue
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for n, us, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` - per-launch times are cold-cache and "
                f"serialised: compare SHARES.\n\n{len(rows)} launches captured, {tot / 1e3:.3f} ms total.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |\n")
    print("wrote", dst)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__shared_mem_per_block_static", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__t_bytes.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        units = rows[1]
        for r in rows[2:]:
            f.write("## " + short(r[hdr.index("Kernel Name")]) + "\n\n")
            for i in idx[1:]:
                f.write(f"- {hdr[i]} [{units[i]}]: {r[i]}\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
