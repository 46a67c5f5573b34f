# kernel-variant experiments: bench.py with alternative builds of the library (JXLB200_LIB) / stage-2 options (JXLB200_STAGE2)
show() { python -c "
import json,sys; j=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], round(j['ms_per_step'],3), j['roofline']['stage_ms'], 'e2e ms', round(j['e2e']['ms_per_step'],3))" $1; }
python bench.py --steps 10 > gpurun_out/b_default.json 2>/dev/null; show gpurun_out/b_default.json
for lib in jxlatte_b200/libjxlb200_*.so; do
  n=$(basename $lib .so)
  JXLB200_LIB=$PWD/$lib python bench.py --steps 10 > gpurun_out/b_$n.json 2>/dev/null; show gpurun_out/b_$n.json
done
