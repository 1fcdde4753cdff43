# Round-2 evidence run on the GPU box: tests, bench, launch list, ncu --set full captures
# (summaries only come back: the .ncu-rep files exceed the transfer limit).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; tail -3 gpurun_out/r02_gputest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_final.err
# launch list of the bench command (cold-cache, serialised kernels: shares, not absolute times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_linear.csv python bench.py --steps 2 --warmup 3 --no-side --execute-snps 0 > gpurun_out/ncu_list.log 2>&1
python profiles/summarise_ncu.py list gpurun_out/r02_launches_linear.csv gpurun_out/r02_launches_linear_summary.csv
cap() {  # name, kernel regex, count, command...
  name=$1; regex=$2; cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -c $cnt -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/summarise_ncu.py rep /tmp/$name.ncu-rep gpurun_out/${name}.md
}
cap r02d_ncu_full_linear "contract_tc2|missing_corr_stream|linear_solve_thread" 3 python scripts/tc_variants.py 151552 0
python profiles/ncu_stalls.py /tmp/r02d_ncu_full_linear.ncu-rep contract_tc2 25 > gpurun_out/r02d_stalls_contract.txt
python profiles/ncu_stalls.py /tmp/r02d_ncu_full_linear.ncu-rep missing_corr_stream 25 > gpurun_out/r02d_stalls_corrections.txt
cap r02d_ncu_full_logistic "logistic_thread" 1 python scripts/logit_probe.py 37888
python profiles/ncu_stalls.py /tmp/r02d_ncu_full_logistic.ncu-rep logistic_thread 25 > gpurun_out/r02d_stalls_logistic.txt
cap r02d_ncu_full_cox "cox_thread" 1 python scripts/cox_probe.py 37888
python profiles/ncu_stalls.py /tmp/r02d_ncu_full_cox.ncu-rep cox_thread 25 > gpurun_out/r02d_stalls_cox.txt
cap r02d_ncu_full_gxe "contract_tc2" 1 python bench.py --model linear_gxe --steps 1 --warmup 1 --no-side
python profiles/ncu_stalls.py /tmp/r02d_ncu_full_gxe.ncu-rep contract_tc2 25 > gpurun_out/r02d_stalls_gxe.txt
ls -la gpurun_out | tail -20
