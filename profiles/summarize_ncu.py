#!/usr/bin/env python
"""Condense an Nsight Compute report (gpurun_out/*.ncu-rep) into the few numbers DESIGN.md / bench.py cite.

usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
    for r in rows[2:]:
        print('kernel:', r[name_col] if name_col is not None else '?')
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.startswith(k) for k in KEYS):
                print(f'  {h} [{u}] = {v}')
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # the source page repeats its header for every kernel of the report: aggregate per kernel section
    sections, cur, ix = [], None, None
    for r in rows:
        if 'Source' in r and '# Samples' in r:
            ix = {k: i for i, k in enumerate(r)}
            cur = dict(ix=ix, hdr=r, data=[])
            sections.append(cur)
        elif cur is not None and len(r) == len(cur['hdr']):
            try:
                int(r[ix['# Samples']])
            except ValueError:
                continue
            cur['data'].append(r)
    for si, sec in enumerate(sections):
        ix, data = sec['ix'], sec['data']
        if not data:
            continue
        stall_cols = [c for c in sec['hdr'] if c.startswith('stall_') and 'Not Issued' not in c]
        tot = {}
        for c in stall_cols:
            try:
                tot[c] = sum(int(r[ix[c]]) for r in data)
            except ValueError:
                pass
        print(f'[kernel #{si}] warp-stall samples by reason (all warps):')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
            print(f'  {k}: {v}')
        print(f'[kernel #{si}] hottest SASS instructions (samples, executed, text):')
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:14]:
            print(f"  {r[ix['# Samples']]:>7} {r[ix['Instructions Executed']]:>10}  {r[ix['Source']].strip()[:100]}")
    if not sections:
        print('no source page in this report')


if __name__ == '__main__':
    main(sys.argv[1])
