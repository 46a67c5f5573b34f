# bench.py at N GPUs for the three workloads: bash tools/scale_bench.sh N
N=$1
for w in 8k batch2048 split16k; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --workload $w > gpurun_out/scale_${w}_$N.json 2> gpurun_out/scale_${w}_$N.err
  python -c "
import json,sys; j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(j['value'],1), round(j['ms_per_step'],3), j['n_gpus'])" gpurun_out/scale_${w}_$N.json || tail -5 gpurun_out/scale_${w}_$N.err
done
