nvidia-smi --query-gpu=clocks.sm,clocks.mem,pstate,power.draw,temperature.gpu --format=csv,noheader
for pw in 5 1 1; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --prewarm-s $pw 2>&1 | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('prewarm', $pw, round(d['value'],1), round(d['ms_per_step']*1000,1), d['clocks'])"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,pstate,power.draw,temperature.gpu --format=csv,noheader
done
