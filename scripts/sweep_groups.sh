#!/bin/bash
# development aid: config-2 throughput against the library's knobs; every argument is one environment ("A=1 B=2")
for spec in "$@"; do
  env $spec python bench.py --only config2 --steps 20 --warmup 5 --cpu-seconds 0.3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$spec:', round(d['reads_per_s']), 'reads/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['reads_per_s']), 'jobs/batch', round(d['queue']['jobs_per_batch'],2), 'alone ms', d['batch_latency_alone_ms'])"
done
