"""Summarise ncu CSV exports (launch lists and --page raw dumps) into small text files for profiles/."""
import csv
import sys
from collections import OrderedDict


def launch_list(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v * 1000.0 if r[ui] in ("ms", "msecond") else v
        agg.setdefault(r[ki].split("(")[0][-70:], []).append(v)
    total = sum(sum(v) for v in agg.values())
    out = [f"{'kernel':72s} {'launches':>8s} {'avg_us':>12s} {'total_us':>12s} {'share':>7s}"]
    for k, v in agg.items():
        out.append(f"{k:72s} {len(v):8d} {sum(v) / len(v):12.2f} {sum(v):12.1f} {sum(v) / total:7.2%}")
    out.append(f"{'TOTAL':72s} {sum(len(v) for v in agg.values()):8d} {'':12s} {total:12.1f}")
    return "\n".join(out)


KEEP = ["gpu__time_duration.sum", "gpc__cycles_elapsed.max.per_second", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def raw_page(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"--- {name[:100]}")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                out.append(f"{h:88s} {v} {u}")
    return "\n".join(out)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    print(launch_list(path) if mode == "launches" else raw_page(path))
