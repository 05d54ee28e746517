mkdir -p gpurun_out/$1
shift_dir=$1; shift
for v in "$@"; do n=${v%%:*}; e=${v#*:}; env $e python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/$shift_dir/bench_$n.json 2> gpurun_out/$shift_dir/bench_$n.err; python -c "
import json,sys
d=json.load(open('gpurun_out/$shift_dir/bench_$n.json'))
s=d['config']['stage_ms_per_step']
print('$n', round(d['value']), round(d['ms_per_step'],4), {k:round(v,3) for k,v in s.items() if v}, round(d['config'].get('ms_per_step_between_flushes',0),4), d['config'].get('p50_ms_per_scan'))
"; done
