# round-end measurements on one B200 (outputs under gpurun_out/, summaries are copied into profiles/ afterwards)
set -x
python bench.py --steps 20 > gpurun_out/fin_bench_8k.json 2> gpurun_out/fin_bench_8k.err
python bench.py --steps 20 --workload batch2048 > gpurun_out/fin_bench_batch2048.json 2> gpurun_out/fin_bench_batch.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/fin_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k2_exact" -s 1 -c 1 -o gpurun_out/fin_k2 -f python tools/run_once.py > gpurun_out/run_once.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_big|k1_small|k1_medium" -s 10 -c 10 -o gpurun_out/fin_k1 -f python tools/run_once.py > gpurun_out/run_once2.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_vardct_gpu.py tests/test_slab_gpu.py -m gpu -x -q > gpurun_out/fin_memcheck.log 2>&1; tail -3 gpurun_out/fin_memcheck.log
compute-sanitizer --tool racecheck python -m pytest tests/test_vardct_gpu.py tests/test_slab_gpu.py -m gpu -x -q -k "mixed_partition or batch or (full_reconstruction and 512)" > gpurun_out/fin_racecheck.log 2>&1; tail -3 gpurun_out/fin_racecheck.log
ls -la gpurun_out/fin_*
