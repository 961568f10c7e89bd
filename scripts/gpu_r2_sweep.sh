set -x
for w in 0 6 7 8 9 10; do
ICEM_B200_CHAIN_WARPS=$w python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('warps',$w,'value',round(d['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms_avg'],3))"
done
