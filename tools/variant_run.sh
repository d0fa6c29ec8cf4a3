#!/bin/bash
# usage (on the GPU box): tools/variant_run.sh TAG name1 name2 ...   ("main" = pst/libpst_b200.so)
tag=$1; shift
for v in "$@"; do
  lib=$PWD/protein-structure-tokenizer_b200/pst/libpst_b200_$v.so
  [ "$v" = "main" ] && lib=$PWD/protein-structure-tokenizer_b200/pst/libpst_b200.so
  PST_LIB_PATH=$lib python bench.py --no-cpu-baseline --sustained-steps 0 --steps 10 > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err || tail -5 gpurun_out/${tag}_$v.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_$v.json") if l.startswith("{")][-1])
    print("$v", "ms/step %.3f" % d["ms_per_step"], {k: (round(x, 3) if x else x) for k, x in d["roofline"]["kernel_ms_per_step"].items()}, (d.get("token_agreement") or {}).get("vs_gpu_fp32_mode_pct"))
except Exception as e:
    print("$v", "failed", e)
PY
done
