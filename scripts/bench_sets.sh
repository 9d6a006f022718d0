for sets in 1 2 4 8; do
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --sets $sets 2>&1 | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sets', $sets, round(d['value'],1), round(d['ms_per_step']*1000,1), {k:round(v['ms']*1000,1) for k,v in d['kernels'].items()})"
done
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --sets 4 --no-pdl 2>&1 | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nopdl sets 4', round(d['value'],1), round(d['ms_per_step']*1000,1))"
