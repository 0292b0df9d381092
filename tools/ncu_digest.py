"""Digest of an ncu report: headline metrics + top stall sites (needs `ncu` on PATH)."""
import csv
import io
import subprocess
import sys


def raw(path):
    """Metrics of the LONGEST captured launch (a report may hold several, e.g. seed pass + main pass)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ix = rows[0].index("gpu__time_duration.sum")
    best = max(rows[2:], key=lambda r: float(r[ix].replace(",", "")))
    print(f"launches in report: {len(rows) - 2}; digest of launch ID {best[0]} ({best[ix]} {rows[1][ix]})")
    return dict(zip(rows[0], zip(rows[1], best))), best[0]


def source(path, top=18, launch_id="0"):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]      # one section per captured launch
    k = min(int(launch_id), len(starts) - 1)
    sect = rows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(rows))]
    hi = next(i for i, r in enumerate(sect) if "Source" in r and "# Samples" in r)
    hdr = sect[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in sect[hi + 1:] if len(r) >= len(hdr) - 1]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    res = [f"total samples {tot}"]
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
        st = {k.replace("stall_", ""): int(r[ix[k]]) for k in hdr
              if k.startswith("stall_") and "Not" not in k and r[ix[k]] not in ("", "0")}
        main = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
        res.append(f"{int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  "
                   f"exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:72]:72s} {main}")
    return "\n".join(res)


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__cluster_size", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]

if __name__ == "__main__":
    m, lid = raw(sys.argv[1])
    for k in KEYS:
        if k in m:
            print(f"{k:78s} {m[k][1]:>16s} {m[k][0]}")
    print(source(sys.argv[1], launch_id=lid))
