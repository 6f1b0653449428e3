#!/bin/bash
# One gpurun session: GPU tests, bench, ncu captures.  Usage: tools/gpu_session.sh <tag> [steps...]
# steps: tests tests_sel(TESTS=...) fuzz bench benchq ab_lib(AB_CONFIGS=, AB_REPS=) ab_persist cliff
#        ncu_models ncu_cold ncu_c2 ncu_wide ncu_c4 ncu_c4s ncu_slice ncu_scada dram launches
TAG=$1; shift
STEPS="$*"
O=gpurun_out
mkdir -p $O
has() { [[ " $STEPS " == *" $1 "* ]]; }
NCU="ncu --set full --clock-control none --import-source on -k regex:mjp_line_kernel --launch-skip 1 --launch-count 1"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests exit $?" >> $O/${TAG}_tests.log
  tail -5 $O/${TAG}_tests.log
fi
if has tests_sel; then   # a subset: TESTS="tests/test_x.py -k expr"
  timeout 1200 python -m pytest $TESTS -m gpu -x -q > $O/${TAG}_tests_sel.log 2>&1; echo "tests_sel exit $?" >> $O/${TAG}_tests_sel.log
  tail -5 $O/${TAG}_tests_sel.log
fi
if has fuzz; then       # differential fuzz through the C ABI against the oracle: default layouts, forced cold layouts, wild profile
  { timeout 300 python tests/fuzz/gpu_fuzz.py ${FUZZ_S:-60} ${FUZZ_FIRST:-3000000} 2>&1 | tail -2
    MJP_COLD=1 timeout 300 python tests/fuzz/gpu_fuzz.py ${FUZZ_S:-60} $(( ${FUZZ_FIRST:-3000000} + 100000 )) 2>&1 | tail -2
    MJP_COLD=1 timeout 300 python tests/fuzz/gpu_fuzz.py ${FUZZ_S:-60} $(( ${FUZZ_FIRST:-3000000} + 200000 )) wild 2>&1 | tail -2; } > $O/${TAG}_fuzz.log 2>&1
  cat $O/${TAG}_fuzz.log
fi
if has bench; then
  timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
  tail -c 600 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
fi
if has benchq; then
  timeout 600 python bench.py --skip-configs > $O/${TAG}_benchq.json 2> $O/${TAG}_benchq.err; echo "benchq exit $?"
  tail -c 400 $O/${TAG}_benchq.json; tail -3 $O/${TAG}_benchq.err
fi
prof() {   # name
  timeout 600 python tools/profile_config.py $1 --out $O/${TAG}_$1_events.json > $O/${TAG}_$1_plain.log 2>&1 || { echo "$1 plain run failed"; tail -3 $O/${TAG}_$1_plain.log; return; }
  cat $O/${TAG}_$1_events.json; echo
  timeout 900 $NCU -f -o $O/${TAG}_$1 python tools/profile_config.py $1 > $O/${TAG}_$1_ncu.log 2>&1; echo "ncu $1 exit $?"
}
if has ab_persist; then
  for c in C3_sweep10 C4_mixed20_stats_many C5_long50 C5_long50_slice C4_mixed20_scada example_scada; do
    for v in 0 1; do
      echo -n "persist=$v $c: "; MJP_L2_PERSIST=$v timeout 600 python tools/profile_config.py $c 2>&1 | tail -1
    done
  done > $O/${TAG}_ab_persist.log 2>&1
  cat $O/${TAG}_ab_persist.log
fi
if has ab_lib; then     # the shipped library against other builds of it (mjpdes_b200/csrc/ab*/, see the Makefile);
  # AB_REPS interleaved repetitions per (config, library): boxes differ and the first launch of a process varies
  for c in ${AB_CONFIGS:-C5_long50_slice C5_long50 C4_mixed20_stats_many C3_sweep10 C4_mixed20_scada_many example_scada C2_example}; do
    for rep in $(seq 1 ${AB_REPS:-3}); do
      for lib in "" mjpdes_b200/csrc/ab*/libmjpdes_b200.so; do
        echo -n "lib=${lib:-shipped} $c: "; MJP_LIB=$lib timeout 600 python tools/profile_config.py $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],2), 'ms', d['events'], d.get('records',''), d['status_counts'])"
      done
    done
  done > $O/${TAG}_ab_lib.log 2>&1
  cat $O/${TAG}_ab_lib.log
fi
if has cliff; then
  MJP_ROTATE=0 timeout 600 python tools/cliff_probe.py > $O/${TAG}_cliff_norotate.json 2> $O/${TAG}_cliff_norotate.err; echo "cliff0 exit $?"
  timeout 600 python tools/cliff_probe.py > $O/${TAG}_cliff_rotate.json 2> $O/${TAG}_cliff_rotate.err; echo "cliff1 exit $?"
  cat $O/${TAG}_cliff_norotate.json $O/${TAG}_cliff_rotate.json
fi
if has ncu_models; then   # one capture per issue model bench.py reads (profiles/issue_model*.json)
  for c in C2_example C3_sweep10 C4_mixed20_stats C4_mixed20_stats_many C4_mixed20_scada C4_mixed20_scada_many C5_long50_slice example_scada; do prof $c; done
fi
if has ncu_cold; then for c in C3_sweep10 C4_mixed20_stats_many C5_long50_slice C4_mixed20_scada_many; do prof $c; done; fi
if has ncu_c2; then prof C2_example; fi
if has ncu_wide; then prof C3_sweep10; prof C4_mixed20_stats_many; prof C5_long50; fi
if has ncu_c4; then prof C4_mixed20_stats; prof C4_mixed20_scada; prof C4_mixed20_scada_many; fi
if has ncu_slice; then prof C5_long50_slice; fi
if has ncu_c4s; then prof C4_mixed20_stats_many; fi
if has ncu_scada; then prof C4_mixed20_scada_many; prof example_scada; fi
if has dram; then
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:mjp_line_kernel --launch-skip 1 --launch-count 1 --csv --log-file $O/${TAG}_dram_c2full.csv \
    python tools/profile_config.py C2_example_full > $O/${TAG}_dram.log 2>&1; echo "dram exit $?"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-configs --no-cpu-baseline > $O/${TAG}_launches.log 2>&1; echo "launches exit $?"
fi
ls -la $O | tail -30
