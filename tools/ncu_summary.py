#!/usr/bin/env python
"""Summarise ncu exports for profiles/: a launch list (gpu__time_duration pass) and a --set full raw page (CSV).

  python tools/ncu_summary.py launches <launches.csv>
  python tools/ncu_summary.py raw <raw.csv>
"""
import collections
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots busy / cycle / SMSP"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe busy % (IMAD.WIDE)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "L1 hit % local loads"),
]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("qb::", "")


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, gi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Unit")
    t = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        t.setdefault(short(r[ki]), []).append((v, r[gi]))
    tot = sum(x for v in t.values() for x, _ in v)
    print(f"{'kernel':22s} {'launches':>8s} {'mean ms':>9s} {'total ms':>9s} {'share':>6s}  grid")
    for k, v in t.items():
        s = sum(x for x, _ in v)
        print(f"{k:22s} {len(v):8d} {s / len(v):9.3f} {s:9.2f} {100 * s / tot:5.1f}%  {v[0][1]}")
    print(f"{'total':22s} {sum(len(v) for v in t.values()):8d} {'':9s} {tot:9.2f}")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    seen = collections.OrderedDict()
    gi = hdr.index("launch__grid_size")
    for r in data:  # the captured launch with the largest grid of every kernel
        k = short(r[ki])
        if k not in seen or float(r[gi].replace(",", "")) > float(seen[k][gi].replace(",", "")):
            seen[k] = r
    order = ["k_rec_coeffs", "k_coef_fx", "k_series", "k_pow", "k_scale", "k_horner", "k_rho_to_z", "k_offdiag_val", "k_offdiag_close"]
    names = sorted(seen, key=lambda n: next((i for i, o in enumerate(order) if n.startswith(o)), 99))
    print("metric".ljust(46) + "".join(n[:17].rjust(18) for n in names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        cells = []
        for n in names:
            v = seen[n][i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            cells.append((v + " " + units[i].replace("register/thread", "").replace("inst", ""))[:17].rjust(18))
        print(label.ljust(46) + "".join(cells))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
