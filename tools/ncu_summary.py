"""Summarise ncu outputs (launch list csv, .ncu-rep raw page) into text for profiles/."""
import collections, csv, re, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__waves_per_multiprocessor',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum', 'smsp__inst_executed.sum',
        'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block']

def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict(); tot = 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum': continue
        k = re.sub(r'\(.*', '', row['Kernel Name'])[:72]
        v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
        v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u, 1.0)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print(f"# launch list {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} ms (cold-cache, serialised: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:72s} n={n:4d} total={t:10.3f} ms share={t/tot*100:5.1f}% avg={t/n:9.3f} ms")

def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, units = r[0], r[1]
    print(f"# {path}")
    for row in r[2:]:
        d = dict(zip(h, row)); un = dict(zip(h, units))
        print(f"## kernel: {d.get('Kernel Name','')[:100]}  grid={d.get('Grid Size','')} block={d.get('Block Size','')}")
        for k in h:
            if any(k.startswith(p) for p in KEYS):
                print(f"  {k} = {d[k]} {un[k]}")
        st = [(k, d[k]) for k in h if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k]
        f = lambda v: float(v.replace(',', '') or 0)
        tot = sum(f(v) for _, v in st) or 1
        print("  stall samples: " + ", ".join(f"{k.split('stalled_')[1]}={f(v)/tot*100:.1f}%" for k, v in sorted(st, key=lambda kv: -f(kv[1]))[:8]))

for a in sys.argv[1:]:
    (launches if a.endswith('.csv') else rep)(a)
