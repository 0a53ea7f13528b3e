#!/bin/bash
# development aid: config-2 throughput against the queue's knobs (batches in flight, merge window, callers)
for spec in "6 300 32" "4 300 32" "10 300 32" "6 1000 32" "3 300 32" "6 300 64" "12 300 64"; do
  set -- $spec
  FXG_GROUPS=$1 FXG_MERGE_WAIT_US=$2 python bench.py --only config2 --lanes $3 --steps 10 --warmup 3 --cpu-seconds 0.3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('groups $1 wait $2 lanes $3:', round(d['reads_per_s']), 'reads/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['reads_per_s']), 'jobs/batch', round(d['queue']['jobs_per_batch'],2), 'launch ms', d['roofline']['launch_event_ms_per_batch'])"
done
FXG_TRACE_BATCHES=1 python bench.py --only config2 --steps 6 --warmup 3 --cpu-seconds 0.3 2> gpurun_out/r02_trace_p.err > /dev/null
grep "batch at" gpurun_out/r02_trace_p.err | tail -40
