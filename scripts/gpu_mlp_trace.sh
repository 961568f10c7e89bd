for v in exp3; do echo "== $v"; ICEM_B200_LIB=icem_b200/lib/libicem_b200_$v.so python scripts/mlp_trace.py > gpurun_out/trace_$v.txt 2>&1; cat gpurun_out/trace_$v.txt; done
