set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; tail -3 gpurun_out/r02_gputest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cox_thread -c 1 -o /tmp/r02d_ncu_full_cox python scripts/cox_probe.py 37888 > gpurun_out/ncu_r02d_cox.log 2>&1
python profiles/summarise_ncu.py rep /tmp/r02d_ncu_full_cox.ncu-rep gpurun_out/r02d_ncu_full_cox.md
python profiles/ncu_stalls.py /tmp/r02d_ncu_full_cox.ncu-rep cox_thread 25 > gpurun_out/r02d_stalls_cox.txt
grep -E "time_duration|dram__bytes|pipe_fp64" gpurun_out/r02d_ncu_full_cox.md
