#!/bin/bash
# round-2 bench line (1 GPU) + the reference arm, as the driver runs them
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
tail -c 1500 gpurun_out/r02_bench2.json; tail -3 gpurun_out/r02_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench2.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","unet_step_ms","unet_step_tflops","tail_ms_per_batch","tail","e2e","gpu_launches","launches_per_denoise_step","clocks"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac","launches_per_step","ms_per_step")})
print(d["kernel_breakdown_ms"])
PY
