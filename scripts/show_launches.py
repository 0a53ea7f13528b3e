import csv, sys
lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
by = {}
for r in rows:
    by.setdefault(r['ID'], {'k': r['Kernel Name'].replace('void ', '')[:16], 'g': r['Grid Size']})[r['Metric Name']] = r['Metric Value']
ids = sorted(by, key=int)
short = {'gpu__time_duration.sum': 'ns', 'launch__occupancy_limit_shared_mem': 'occ_smem', 'launch__occupancy_limit_registers': 'occ_reg',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps%', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu%',
         'smsp__inst_executed.sum': 'inst', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma%',
         'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
         'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active': 'fmaheavy%', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active': 'fmalite%'}
tot = 0
for i in ids[len(ids) // 2 + 1:]:
    d = by[i]
    tot += int(d.get('gpu__time_duration.sum', '0').replace(',', ''))
    print(i, d['k'], d['g'], ' '.join(f"{short.get(k, k)}={v}" for k, v in d.items() if k not in ('k', 'g')))
print('total ms', tot / 1e6)
