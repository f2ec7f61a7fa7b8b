"""Headline metrics of an .ncu-rep, machine readable.

    python tools/ncu_metrics.py REPORT.ncu-rep [--all] [--kernel REGEX] [--csv OUT.csv] [--json OUT.json]

--csv   the report's raw page (every metric ncu collected, one row per kernel instance) as CSV: the committed,
        machine-readable form of a capture (the .ncu-rep itself is scratch under gpurun_out/)
--json  the metrics bench.py quotes in roofline.binding / roofline.traffic for the first kernel matching --kernel
"""
import argparse
import csv
import json
import re
import subprocess

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
              "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--csv", default="")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    out = subprocess.run(['ncu', '-i', args.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    if args.csv:
        open(args.csv, "w").write(out)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if not args.kernel or re.search(args.kernel, dict(zip(hdr, r)).get('Kernel Name', ''))]
    for r in body if args.all else body[:1]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '')[:90])
        for k in WANT:
            if k in d:
                print(f'  {k:90s} {d[k]} {units[hdr.index(k)]}')
    if args.json and body:
        d = dict(zip(hdr, body[0]))
        u = dict(zip(hdr, units))

        def scaled(k):
            v = num(d.get(k, ""))
            return None if v is None else v * UNIT_SCALE.get(u.get(k, ""), 1.0)
        rd, wr = scaled('dram__bytes_read.sum'), scaled('dram__bytes_write.sum')
        j = {
            "source": args.report.split("/")[-1], "raw_csv": args.csv.split("/")[-1] if args.csv else None,
            "kernel": d.get('Kernel Name', ''),
            "duration_us": scaled('gpu__time_duration.sum'),
            "l1tex_data_pipe_wavefronts_pct": num(d.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', '')),
            "issue_active_pct": num(d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', '')),
            "lanes_per_inst": num(d.get('smsp__thread_inst_executed_per_inst_executed.ratio', '')),
            "warp_instructions": num(d.get('smsp__inst_executed.sum', '')),
            "l1_hit_pct": num(d.get('l1tex__t_sector_hit_rate.pct', '')), "l2_hit_pct": num(d.get('lts__t_sector_hit_rate.pct', '')),
            "l2_throughput_pct": num(d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', '')),
            "dram_throughput_pct": num(d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '')),
            "long_scoreboard_stalls_per_issue": num(d.get('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', '')),
            "registers": num(d.get('launch__registers_per_thread', '')),
            "dram_bytes_read": rd, "dram_bytes_written": wr,
            "bytes_per_launch": None if rd is None or wr is None else rd + wr,
        }
        json.dump(j, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
