#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import os, sys, subprocess, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import bench
w = dict(bench.WORKLOADS["c2"]); w["reads"] = 500000
idx = bench.ensure_dataset("c2")
s1, o1, _, _ = bench.make_reads(w, 7)
bench.write_fastq_sample(s1, o1, 500000, "/tmp/r.fq", "")
for env in ({}, {"CFR_B200_DUST_SCREEN": "0"}):
    e = dict(os.environ); e.update(env); e["CFR_B200_TRACE"] = "1"
    t = time.perf_counter()
    p = subprocess.run(["centrifuger_b200/centrifuger-b200", "-x", idx, "-u", "/tmp/r.fq"], stdout=open("/tmp/o.tsv", "wb"), stderr=subprocess.PIPE, env=e)
    print(env, "%.2f s" % (time.perf_counter() - t)); print(p.stderr.decode()[-1500:])
PY
