mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r21_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r21_tests.log
tail -3 gpurun_out/r21_tests.log
timeout 900 python bench.py > gpurun_out/r1h_bench_exome.json 2> gpurun_out/r1h_bench_exome.err; echo bench rc=$?
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1h_bench_reference.json 2> gpurun_out/r1h_bench_reference.err; echo ref rc=$?
timeout 900 python bench.py --workload chr22 > gpurun_out/r1h_bench_chr22.json 2> gpurun_out/r1h_bench_chr22.err; echo chr22 rc=$?
timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1h_launches_bench_exome.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo launches rc=$?
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_window_hist|k_assemble|k_allele_call|k_replay|k_scatter" -s 18 -c 6 -o gpurun_out/prof_r1h python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo full rc=$?
ls -la gpurun_out/ | tail -12
