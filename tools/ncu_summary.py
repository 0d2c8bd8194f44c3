#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics (raw page) and, optionally, the hottest source lines.
usage: tools/ncu_summary.py report.ncu-rep [--source KERNEL_REGEX] [--top N]"""
import csv, subprocess, sys, io, re, collections

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']

def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout

def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        print('===', r[hdr.index('Kernel Name')][:80], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('   %-80s %s %s' % (w, r[i], units[i]))
    if '--source' in sys.argv:
        pat = sys.argv[sys.argv.index('--source') + 1]
        top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
        out = run(['-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat, '--print-source', 'cuda,sass'])
        rows = list(csv.reader(io.StringIO(out)))
        hdr = None; fname = ''
        acc = collections.OrderedDict()
        for r in rows:
            if r and r[0] in ('File Name', 'File Path'):
                fname = r[1].split('/')[-1]; continue
            if r and r[0] == 'Line No':
                hdr = r; continue
            if hdr is None or len(r) != len(hdr): continue
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            try: samp = int(d.get('# Samples', '0') or 0)
            except ValueError: samp = 0
            try: inst = int(d.get('Instructions Executed', '0') or 0)
            except ValueError: inst = 0
            if samp == 0 and inst == 0: continue
            if not d.get('Line No', '').strip(): continue          # SASS rows under a source line: already counted in it
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and '(' not in k and v.isdigit() and int(v) > 0}
            key = (fname, d['Line No'], d['Source'].strip()[:100])
            a = acc.setdefault(key, [0, 0, collections.Counter()])
            a[0] += samp; a[1] += inst; a[2].update(stalls)
        tot = sum(a[0] for a in acc.values()) or 1
        toti = sum(a[1] for a in acc.values()) or 1
        print('--- hottest source lines of', pat, ': samples, % samples, % instructions, top stalls, source')
        for (fn, ln, src), (samp, inst, st) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
            ts = ','.join('%s:%d' % kv for kv in st.most_common(2))
            print('%-16s %5s %7d %5.1f%% %5.1f%%  %-28s %s' % (fn[:16], ln, samp, 100.0 * samp / tot, 100.0 * inst / toti, ts, src))

if __name__ == '__main__':
    main()
