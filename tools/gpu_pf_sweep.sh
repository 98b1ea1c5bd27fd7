for pf in 148 222 296 370 444; do TXG_PF=$pf python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('pf', $pf, round(d['value']), round(d['ms_per_step'],3), round(k['k_step_fused']['ms']/k['k_step_fused']['launches'],3), d['clocks']['sm_mhz'])"; done
