"""Run bench.py under several library knobs and print one line each.
Tuning aid for the GPU box:  python tools/sweep.py 'LAYOUT=csr' 'LAYOUT=sell' 'PIPE=0' ...
(the VARIANT / CTAS / SIGMA knobs of earlier sweeps -- profiles/r1_sweep*.txt -- selected kernel
variants that have since been removed; they are ignored by the library now)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split(","):
        k, v = kv.split("=")
        env[{"LAYOUT": "EXAGS_LAYOUT", "PIPE": "EXAGS_HOST_PIPELINE", "CHUNK": "EXAGS_PIPE_CHUNK_BYTES",
             "VARIANT": "EXAGS_SELL_VARIANT", "SIGMA": "EXAGS_SELL_SIGMA", "CTAS": "EXAGS_SELL_CTAS"}[k]] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3", "--no-cpu"],
                       env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        rf = j["roofline"]
        print(f"{spec:28s} ms/step {j['ms_per_step']:.4f}  kernel min {rf['kernel_ms_min']:.4f} mean {rf['kernel_ms_mean']:.4f} "
              f"frac {rf['frac']:.3f}  GDOF/s {j['value']:.1f}  setup {j['config']['setup_s']:.1f}s", flush=True)
    except Exception as e:
        print(spec, "FAILED", e, r.stdout[-500:], r.stderr[-1500:], flush=True)
