# A/B: library variants on the two GT workloads
for v in c2_l0 c2_l1 c3_l0 c3_l1; do
  for wl in humanoid_standup_gt_n16384 halfcheetah_gt_n4096; do
    ICEM_B200_LIB=$PWD/icem_b200/lib/libicem_$v.so python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$wl', round(d['value']), 'traj/s', round(d['ms_per_step'],2),'ms', 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"
  done
done
