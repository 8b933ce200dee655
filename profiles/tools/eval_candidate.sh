#!/usr/bin/env bash
# One GPU call to judge a kernel candidate (run from the repo root on the GPU box, e.g.
#   gpurun --timeout 600 -- 'bash profiles/tools/eval_candidate.sh NAME'
# after `git checkout <branch> && python -c "import __graft_entry__ as g; g.build()"` in the container):
# the parity tests that exercise the grouping / big-bucket paths, then the cfg2 line and the config-5-shaped line.
# Outputs: gpurun_out/cand_<NAME>_{tests.log,cfg2.json,cfg5.json}.
set -u
name=${1:-candidate}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "table_matches or large_buckets or deep_coverage or medium_synthetic or streams or key_width or full_size or degenerate" \
    2>&1 | tail -5 > gpurun_out/cand_${name}_tests.log
timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/cand_${name}_cfg2.json 2> gpurun_out/cand_${name}_cfg2.err
timeout 200 python bench.py --workload cfg5 --reads-per-gpu 3000000 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline \
    > gpurun_out/cand_${name}_cfg5.json 2> gpurun_out/cand_${name}_cfg5.err
tail -2 gpurun_out/cand_${name}_tests.log
python - "$name" <<'PY'
import json, sys
for w in ("cfg2", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/cand_{sys.argv[1]}_{w}.json").read().strip().splitlines()[-1])
        print(w, "%.2f G k-mers/s  %.3f ms" % (d["value"] / 1e9, d["ms_per_step"]),
              {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items() if v}, d["roofline"]["pipeline_info"])
    except Exception as e:  # noqa: BLE001
        print(w, "failed:", e)
PY
