"""Spawn np ranks of tests/mp_worker.py on this node and wait for them."""
import os
import random
import subprocess
import sys
import uuid

HERE = os.path.dirname(os.path.abspath(__file__))


def run_ranks(np_, args, timeout=300, extra_env=None):
    port = random.randint(20000, 60000)
    key = uuid.uuid4().hex[:16]
    procs = []
    for r in range(np_):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(np_), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EXAGS_JOB_KEY=key,
                   OMP_NUM_THREADS="2")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, "-X", "faulthandler", os.path.join(HERE, "mp_worker.py")] + [str(a) for a in args],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs, codes = [], []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            out = "TIMEOUT"
        outs.append(out)
        codes.append(p.returncode)
    return codes, outs
