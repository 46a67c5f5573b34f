"""TEST TOOL: decodes a .jxl file with the C++ front end + the CPU ORACLE (not the product path) and writes a PNG, for looking
at what the checker thinks a file should decode to.    python tests/tools/decode_with_oracle.py input.jxl output.png [--bits 16]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_engine import OracleEngine  # noqa: E402
from jxlatte_b200.decoder import JXLDecoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("input")
ap.add_argument("output")
ap.add_argument("--bits", type=int, default=8)
a = ap.parse_args()
t0 = time.perf_counter()
d = JXLDecoder(a.input, engine=OracleEngine())
img = d.decode()
print("%s: %dx%d in %.2f s (front end %.0f ms)" % (os.path.basename(a.input), img.width, img.height, time.perf_counter() - t0, d.timings["front_end_s"] * 1e3))
img.write_png(a.output, a.bits)
