#!/bin/bash
# multi-GPU session: tools/multi_session.sh <tag> <ngpus> [tests] [bench]
TAG=$1; N=$2; shift; shift
STEPS="$*"
O=gpurun_out
mkdir -p $O
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpus.txt 2>&1
if has tests; then
  timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "multi or sharded" > $O/${TAG}_tests.log 2>&1; echo "tests exit $?" >> $O/${TAG}_tests.log
  tail -4 $O/${TAG}_tests.log
fi
if has bench; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 3 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "bench exit $?"
  tail -c 700 $O/${TAG}_bench_${N}gpu.json; tail -5 $O/${TAG}_bench_${N}gpu.err
fi
if has refarm; then
  timeout 600 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; echo "ref exit $?"
  tail -c 400 $O/${TAG}_ref.json
fi
