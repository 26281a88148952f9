#!/bin/bash
# round-2 GPU job 11: programmatic dependent launch A/B (bench headline + train + nms legs), then the GPU test suite with it on
set -u
mkdir -p gpurun_out
O=gpurun_out/r2j11
for cfg in "0 0" "1 0" "1 72"; do set -- $cfg; pdl=$1; export Y3_CHUNK_MB=$2
  echo "== Y3_PDL=$pdl Y3_CHUNK_MB=$Y3_CHUNK_MB"
  Y3_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --only train,nms,spp_nms > ${O}_bench_pdl${pdl}_chunk${Y3_CHUNK_MB}.json 2> ${O}_bench_pdl${pdl}_chunk${Y3_CHUNK_MB}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("${O}_bench_pdl${pdl}_chunk${Y3_CHUNK_MB}.json") if l.startswith("{")][-1])
    t = d.get("train", {})
    print("pdl=$pdl value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "parity", d.get("parity_rel_l2"))
    print("   train", {k: t.get(k) for k in ("value", "ms_per_step", "split_ms")})
    n = d.get("nms", {})
    print("   nms", {k: round(v["ms_per_batch"], 4) for k, v in n.items() if isinstance(v, dict)})
    s = d.get("spp_nms", {})
    print("   spp_nms", s.get("value"))
except Exception as e:
    print("pdl=$pdl parse failed", e)
PY
  tail -2 ${O}_bench_pdl${pdl}_chunk${Y3_CHUNK_MB}.err
done
unset Y3_CHUNK_MB
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -8 > ${O}_pytest.log; tail -4 ${O}_pytest.log
tools/gpu_sanity.sh end
