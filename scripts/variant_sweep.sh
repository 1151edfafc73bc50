#!/bin/bash
# Kernel tuning sweep: bench line of every library build under teloscope_b200/variants/ (built with
# make -C teloscope_b200/csrc OUT=.../variants/lib_X.so BUILD=.../build_X EXTRA=-D... <OUT>) and of the in-tree build.
for lib in "" teloscope_b200/variants/lib_*.so; do
  name=${lib:-default}
  TLS_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), round(d['kernel_only']['scan_kernels_ms'],4))"
done
