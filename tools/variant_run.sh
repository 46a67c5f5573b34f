# Kernel / schedule variant experiments in ONE gpurun call.  Build the variants here first, e.g.
#   JXLB_SO=$PWD/jxlatte_b200/libjxlb200_th96_512.so JXLB_EXTRA_FLAGS="-DKX_TH=96 -DKX_THREADS=512 -DKX_MINB=1" python -m jxlatte_b200.build --force
#   JXLB_SO=$PWD/jxlatte_b200/libjxlb200_pipe256.so  JXLB_EXTRA_FLAGS="-DJXLB200_PIPE_ROWS=256" python -m jxlatte_b200.build --force
# then:  gpurun --timeout 200 -- 'bash tools/variant_run.sh > gpurun_out/variants.txt 2>&1'
# Every line carries a CRC of the planes: a variant that changes a bit shows up at once.  SKIP_HOST=1 leaves the host-entry timing out.
L=$PWD/jxlatte_b200
python tools/variant_time.py
for lib in $L/libjxlb200_*.so; do
  [ -e "$lib" ] || continue
  JXLB200_LIB=$lib python tools/variant_time.py
done
