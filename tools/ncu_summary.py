#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep from `ncu --set full --import-source on`) into text that can be committed
under profiles/: headline counters (duration, DRAM bytes, tensor-pipe activity, occupancy inputs), warp-stall
totals, and the hottest SASS instructions with their stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_hot] > profiles/rNN_<what>.txt
"""
import csv
import io
import subprocess
import sys

STALLS = ['stall_barrier', 'stall_branch_resolving', 'stall_long_sb', 'stall_math', 'stall_membar', 'stall_short_sb',
          'stall_wait', 'stall_not_selected', 'stall_selected', 'stall_no_inst', 'stall_mio', 'stall_lg', 'stall_sleep',
          'stall_dispatch']
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for k, vals in enumerate(raw[2:]):
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== launch {k}: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                print(f"  {h:84s} {u:16s} {v}")
    src = page(rep, "source")
    # the source page holds one block per profiled kernel: a title row, a header row, then one row per instruction
    blocks, cur, prev = [], None, []
    for r in src:
        if '# Samples' in r:
            cur = {"title": blocks_title, "hdr": r, "rows": []} if (blocks_title := (prev[1] if len(prev) > 1 else "")) is not None else None
            blocks.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
        prev = r
    seen = set()
    for b in blocks:
        key = (b["title"], len(b["rows"]))
        if key in seen:  # ncu repeats the block when the same kernel source serves several results
            continue
        seen.add(key)
        print("==", b["title"])
        ix = {x: i for i, x in enumerate(b["hdr"])}
        data = b["rows"]
        tot = sum(int(r[ix['# Samples']]) for r in data)
        print(f"warp-stall samples: {tot}")
        agg = {s: sum(int(r[ix[s]]) for r in data) for s in STALLS if s in ix}
        for s, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            if v:
                print(f"  {s:26s} {v:8d}  {100.0 * v / max(tot, 1):5.1f}%")
        print(f"hottest {nhot} SASS instructions (samples, executed, instruction, stall reasons):")
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:nhot]:
            st = {s.replace('stall_', ''): int(r[ix[s]]) for s in STALLS if s in ix and int(r[ix[s]]) > 0}
            print(f"  {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']][:64]:64s} {st}")


if __name__ == "__main__":
    main()
