# round-end evidence on ONE B200: tests, the default bench line, the reference arm, ncu launch list + full captures
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest_final.log 2>&1; tail -1 gpurun_out/r2_pytest_final.log
( time timeout 900 python bench.py > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.err ) 2>&1 | grep real
head -c 400 gpurun_out/r2_bench_c2_final.json; echo
timeout 600 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; head -c 400 gpurun_out/r2_bench_reference_arm.json; echo
timeout 300 python bench.py --config C3 --no-sweep > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r2_bench_c3_final.err; head -c 300 gpurun_out/r2_bench_c3_final.json; echo
python scripts/step_timeline.py C2 2>&1 | grep -v arn > gpurun_out/r2_step_timeline_c2.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/r2_bench_short.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/ncu_launches.log 2>&1
timeout 120 python scripts/ncu_targets.py step > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc05|cgpl_pgls" -s 14 -c 7 -o gpurun_out/r2_step python scripts/ncu_targets.py step > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
timeout 120 python scripts/ncu_targets.py rows > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cgpl_pgls|masked_softce" -s 2 -c 2 -o gpurun_out/r2_rows_after python scripts/ncu_targets.py rows > gpurun_out/ncu_rows.log 2>&1; tail -1 gpurun_out/ncu_rows.log
ls -la gpurun_out/*.ncu-rep | tail -4
