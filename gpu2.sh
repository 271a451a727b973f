set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_exome.json 2> gpurun_out/bench_exome.err; echo rc=$? ; tail -c 3000 gpurun_out/bench_exome.json; tail -5 gpurun_out/bench_exome.err
python bench.py --workload chr22 --cpu-sample-transcripts 100 > gpurun_out/bench_chr22.json 2> gpurun_out/bench_chr22.err; echo rc=$?; tail -c 2500 gpurun_out/bench_chr22.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo rc=$?; cat gpurun_out/bench_reference.json
nproc; free -g | head -2
python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_window_hist|k_allele_call|k_assemble" -s 9 -c 3 -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo rc=$?
ls -la gpurun_out
