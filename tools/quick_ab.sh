# quick A/B: ms of one launch of each profiling-sized workload (tools/profile_config.py); extra env via $1
for c in ${CONFIGS:-C3_sweep10 C4_mixed20_stats_many C5_long50_slice C2_example}; do echo -n "$c: "; python tools/profile_config.py $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],2), 'ms', d['events'], d['status_counts'])"; done
