#!/usr/bin/env bash
# Time the REAL reference (Clojure on a JVM) beside bench.py, when one is available (CPU path item 3).
# Nothing here is used by the product or the tests.  The image this repository is built in has no JVM, so this
# script normally reports "unavailable"; a maintainer with `java` and an uberjar can drop the jar at
# baseline/_ref/MJPdes-standalone.jar (git-ignored) and get replications/s of the stock code path.
#
#   baseline/run_reference.sh MODEL.clj [N_PROCESSES]
#
# MODEL.clj: a model file for the reference CLI (`-i`, main.clj:7-18), e.g. the reference's resources/example.clj.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
jar="${MJPDES_JAR:-$here/_ref/MJPdes-standalone.jar}"
model="${1:-}"
nproc_="${2:-$(nproc)}"
if ! command -v java >/dev/null 2>&1; then echo '{"impl": "jvm-reference", "unavailable": "no java on PATH"}'; exit 0; fi
if [ ! -f "$jar" ]; then echo "{\"impl\": \"jvm-reference\", \"unavailable\": \"no jar at $jar\"}"; exit 0; fi
if [ -z "$model" ] || [ ! -f "$model" ]; then echo '{"impl": "jvm-reference", "unavailable": "pass a model file"}'; exit 0; fi
t0=$(date +%s.%N)
pids=()
for i in $(seq 1 "$nproc_"); do java -jar "$jar" -i "$model" -o "/tmp/mjpdes_ref_$i.clj" >/dev/null 2>&1 & pids+=($!); done
for p in "${pids[@]}"; do wait "$p"; done
t1=$(date +%s.%N)
python3 - "$t0" "$t1" "$nproc_" <<'PY'
import json, sys
t0, t1, n = float(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
print(json.dumps({"impl": "jvm-reference", "processes": n, "seconds": t1 - t0, "replications_per_sec": n / (t1 - t0),
                  "note": "one replication per JVM, JVM start-up included"}))
PY
