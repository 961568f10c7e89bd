ICEM_B200_LIB=icem_b200/lib/libicem_b200_trace.so timeout 120 python scripts/mlp_trace.py > gpurun_out/trace2.txt 2>&1; cat gpurun_out/trace2.txt
