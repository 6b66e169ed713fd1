#!/bin/bash
# wall time of an eval sweep with 1 and 2 lock-step groups in flight per GPU
for p in 1 2; do
  python - <<PY
import subprocess, time
t0 = time.time()
subprocess.run(["python", "eval.py", "--cfg", "cfg/eval/synthetic_pie.yaml", "--limit", "${LIMIT:-48}", "--override", "--pipes", "$p"], check=True,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
print("pipes=$p: ${LIMIT:-48} edits in %.1f s wall (model load and graph capture included)" % (time.time() - t0))
PY
done
