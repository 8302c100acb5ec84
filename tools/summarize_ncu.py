"""Turn ncu outputs into the small text/JSON summaries committed under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md>      # per-kernel totals/shares of the last full step
  python tools/summarize_ncu.py full <report.ncu-rep> <out.json>      # key metrics per captured kernel
"""
import csv
import json
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6}.get(r[ui], 1e-6)
        seq.append((r[ki], v * scale))
    setups = [i for i, (n, _) in enumerate(seq) if 'k_setup' in n]
    # a step = from one k_setup launch to the next; take the last complete one that contains k_backward
    steps = [(a, b) for a, b in zip(setups, setups[1:] + [len(seq)])
             if any('k_backward(' in n for n, _ in seq[a:b]) and any(('k_intersect<' in n or 'k_filter_const<' in n) for n, _ in seq[a:b])]
    a, b = steps[-2] if len(steps) > 1 else steps[-1]   # a full fwd+bwd step of the main timed loop (default kernel)
    agg = {}
    for n, ms in seq[a:b]:
        key = n.split('(')[0][:80]
        cnt, tot = agg.get(key, (0, 0.0))
        agg[key] = (cnt + 1, tot + ms)
    total = sum(t for _, t in agg.values())
    lines = ['# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)', '',
             'source: `%s` - one fwd+bwd step of `bench.py` (cold-cache, serialised launches: compare shares, not absolutes)' % path, '',
             '| kernel | launches | total ms | share |', '|---|---|---|---|']
    for key, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('| `%s` | %d | %.4f | %.2f%% |' % (key, cnt, tot, 100 * tot / total))
    lines.append('| **sum** | %d | %.4f | 100%% |' % (sum(c for c, _ in agg.values()), total))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'idc__request_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def full(path, out):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for d in data:
        rec = {}
        for h, u, v in zip(hdr, units, d):
            if h == 'Kernel Name' or h in WANT:
                rec[h + (' [%s]' % u if u else '')] = v
        res.append(rec)
    json.dump(res, open(out, 'w'), indent=1)
    for r in res:
        print(json.dumps(r, indent=1))


def traffic(path, out, workload='config_e', n_gpus='1'):
    """profiles/ncu_traffic.json for bench.py's roofline.traffic: DRAM bytes (read + write) per launch of each captured
    kernel, stamped with the hash of the library sources the capture was taken on (bench.py only reports it when the
    hash matches the build it is running)."""
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
    import bench
    recs = json.load(open(path))
    kernels = {}
    for r in recs:
        name = r['Kernel Name'].split('(')[0].split('<')[0].replace('void ', '').replace('surf::', '').strip()
        rd = wr = None
        for k, v in r.items():
            if k.startswith('dram__bytes_read.sum'):
                rd = float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[k.split('[')[1].rstrip(']')]
            if k.startswith('dram__bytes_write.sum'):
                wr = float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[k.split('[')[1].rstrip(']')]
        if rd is not None and wr is not None:
            kernels[name] = int(rd + wr)
    json.dump({'source_hash': bench.source_hash(), 'workload': workload, 'n_gpus': int(n_gpus), 'kernels': kernels,
               'from': os.path.basename(path), 'what': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full'},
              open(out, 'w'), indent=1)
    print(open(out).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](*sys.argv[2:])
