#!/bin/bash
# Multi-GPU bench lines on one box: scripts/gpu_scale.sh <N> <tag> [nopipe]   (run under gpurun --gpus N)
n=${1:-2}; tag=${2:-scale}; out=gpurun_out/$tag; mkdir -p $out
for mode in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 --scaling $mode --no-extra --no-cpu-baseline > $out/bench_n${n}_$mode.json 2> $out/bench_n${n}_$mode.err
  tail -c 1500 $out/bench_n${n}_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['kernel_only']['Gbp_per_s_all_ranks'], d['e2e']['value'])" 2>/dev/null || tail -3 $out/bench_n${n}_$mode.err
done
# the product binary over N devices: FASTA file -> files
[ "$3" = nopipe ] || timeout 600 python scripts/pipeline_bench.py 1.0 --gpus $n --repeat 3 > $out/pipeline_n$n.log 2>&1; [ "$3" = nopipe ] || tail -3 $out/pipeline_n$n.log
