python -m pytest tests/test_vardct_gpu.py tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -3
show() { python -c "
import json,sys; j=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], j['ms_per_step'], j['roofline']['stage_ms'])" $1; }
python bench.py --steps 10 > gpurun_out/b_pair160.json 2>/dev/null; show gpurun_out/b_pair160.json
JXLB200_STAGE2=3 python bench.py --steps 10 > gpurun_out/b_scalar.json 2>/dev/null; show gpurun_out/b_scalar.json
JXLB200_LIB=$PWD/jxlatte_b200/libjxlb200_t128.so python bench.py --steps 10 > gpurun_out/b_pair128.json 2>/dev/null; show gpurun_out/b_pair128.json
