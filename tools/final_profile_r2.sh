# round-2 measurements on one B200 (outputs under gpurun_out/, summaries are copied into profiles/ afterwards)
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/fin2_bench_8k.json 2> gpurun_out/fin2_bench_8k.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin2_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/fin2_launches.csv python tools/run_once.py > gpurun_out/fin2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k2_stream" -s 1 -c 1 -o gpurun_out/fin2_k2 -f python tools/run_once.py > gpurun_out/fin2_run_once.log 2>&1
ls -la gpurun_out/fin2_*
