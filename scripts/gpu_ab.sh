for v in "" _nolock _2cta; do
ICEM_B200_LIB=icem_b200/lib/libicem_b200$v.so python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench$v', d['value'], d['ms_per_step'])"
ICEM_B200_LIB=icem_b200/lib/libicem_b200$v.so python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cheetah$v', d['value'], d['ms_per_step'])"
done
