"""Text summary of an `ncu --set full --import-source on` report: headline metrics per captured launch and the source lines
that collect the most warp-stall samples.  Used for profiles/r02_ncu_*_summary.txt.

    python tools/ncu_summary.py gpurun_out/r02_ncu_conv_tc.ncu-rep [--top 16] > profiles/r02_ncu_conv_tc_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep.split('/')[-1]}: ncu --set full --clock-control none --import-source on (one replayed launch each: cold cache, serialised)")
    for li, v in enumerate(launches):
        print(f"\n## launch {li}: {v[ix['Kernel Name']][:150]}")
        for k, label in KEYS:
            if k in ix and v[ix[k]] != "":
                print(f"  {label:45s} {v[ix[k]]} {units[ix[k]]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur = fn = None
    hdr = None
    per = {}
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            fn = r[1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
            hx = {}
            for i, h in enumerate(r):
                hx.setdefault(h, i)
        elif hdr and len(r) == len(hdr) and r[2] == "-" and cur and cur.startswith("y3_"):
            try:
                s = float(r[hx["# Samples"]])
            except ValueError:
                continue
            if s <= 0:
                continue
            st = {h[6:]: float(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            per.setdefault(fn, []).append((s, cur, r[0], r[1].strip()[:96], sorted(st.items(), key=lambda kv: -kv[1])[:2]))
    for fn, rows in per.items():
        tot = sum(x[0] for x in rows)
        print(f"\n## warp-stall samples by source line (own sources; inlined helpers are counted at every level): {fn[:110]}")
        for s, f, l, text, st in sorted(rows, key=lambda x: -x[0])[:top]:
            print(f"  {s:6.0f} {f}:{l:>4s}  {text:96s} {st[0][0]}:{st[0][1]:.0f} {st[1][0]}:{st[1][1]:.0f}")


if __name__ == "__main__":
    main()
