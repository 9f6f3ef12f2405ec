#!/bin/bash
# Runs ON a 1-GPU box: times the "defaults" configuration of tools/variant_sweep.py for the production library and
# for every compile-time variant in libfluid_b200/_lib/variants/ (tools/build_variant.py).  Usage: gpu_variants.sh <tag> [names]
set -u; shopt -s nullglob
tag=$1; names=${2:-}
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("   wall %.2f  pcg/it %.4f  " % (d["ms_per_step_wall"], d["pcg_ms_per_iter"]), {k: round(v, 2) for k, v in d["phase_ms"].items() if v > 1.0})
PY
}
echo "== production"; timeout -k 5 200 python tools/variant_sweep.py --grid 256 --tag ${tag}_prod --only defaults > gpurun_out/${tag}_prod.log 2>&1; show gpurun_out/${tag}_prod.log
for lib in libfluid_b200/_lib/variants/liblfk_*.so; do
  n=$(basename $lib .so); n=${n#liblfk_}
  if [ -n "$names" ] && [[ ",$names," != *",$n,"* ]]; then continue; fi
  echo "== $n"; timeout -k 5 200 python tools/variant_sweep.py --grid 256 --tag ${tag}_$n --only defaults --lib $lib > gpurun_out/${tag}_$n.log 2>&1 || tail -3 gpurun_out/${tag}_$n.log; show gpurun_out/${tag}_$n.log
done
