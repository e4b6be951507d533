#!/bin/bash
# Round 2, session G: full GPU suite on the consolidated build, RK64 FMA audit, small-box path, dodecane knobs
set -u
TAG=r02g
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python tools/rk64_audit.py > gpurun_out/${TAG}_rk64_audit_fma.json 2>gpurun_out/${TAG}_audit.err
PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_nofma.so timeout 600 python tools/rk64_audit.py > gpurun_out/${TAG}_rk64_audit_nofma.json 2>>gpurun_out/${TAG}_audit.err
timeout 600 python tools/bench_boxes.py > gpurun_out/${TAG}_boxes_lidryer_cvode.jsonl 2>gpurun_out/${TAG}_boxes.err
timeout 600 python tools/bench_boxes.py --reactor ReactorRK64 > gpurun_out/${TAG}_boxes_lidryer_rk64.jsonl 2>>gpurun_out/${TAG}_boxes.err
cat gpurun_out/${TAG}_boxes_*.jsonl | cut -c1-400
timeout 300 python bench.py --no-cpu-baseline --strong none 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128.json
timeout 300 python bench.py --no-cpu-baseline --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w1.json
PRX_BLOCK_WARPS=4 timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w4.json
PRX_MAX_WARPS_PER_SM=16 timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w1_mw16.json
PRX_MAX_WARPS_PER_SM=12 timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w1_mw12.json
PRX_BLOCK_WARPS=4 PRX_MAX_WARPS_PER_SM=16 timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w4_mw16.json
PRX_BLOCK_WARPS=4 PRX_COOP_A_GLOBAL=0 timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w4_asmem.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed_cells"))
    except Exception as e:
        print(f, "unreadable", e)
for f in ("fma", "nofma"):
    try:
        d = json.load(open("gpurun_out/${TAG}_rk64_audit_%s.json" % f))
        for k, v in d.items():
            if isinstance(v, dict):
                print(f, k, "pure-rel %.2e (%s, Y=%.1e) floor %.2e T %.2e" % (v["max_rel_species"], v["worst_species"], v["worst_species_Y"], v["max_rel_species_floor_1e-2"], v["max_rel_T"]))
    except Exception as e:
        print("audit", f, e)
PY
tail -3 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_audit.err gpurun_out/${TAG}_boxes.err
