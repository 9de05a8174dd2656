#!/usr/bin/env python
"""The drop-in binary on the default bench workload's reads (FASTQ files -> TSV to /dev/null) with --gpus 1, 2, ... G:
wall seconds, pairs/s and the binary's own stage report per G, and the md5 of the TSV (must not depend on G).
usage: cli_scale.py [workload] [million pairs] [G ...]"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    mp = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    gs = [int(x) for x in sys.argv[3:]] or [1, 2]
    w = dict(bench.WORKLOADS[wl])
    idx = bench.ensure_dataset(w["dataset"])
    src = bench.ReadSource(w)
    d = tempfile.mkdtemp(prefix="cfr_cli_scale_")
    f1p, f2p = os.path.join(d, "r_1.fq"), os.path.join(d, "r_2.fq")
    n_total = 0
    with open(f1p, "wb") as f1, open(f2p, "wb") as f2:
        for j in range(mp):
            seq1, off1, seq2, off2 = src.batch(1_000_000, 7 + 100003 * j)
            bench.write_fastq_fixed(f1, seq1, 1_000_000, w["rlen"], n_total, "/1")
            bench.write_fastq_fixed(f2, seq2, 1_000_000, w["rlen"], n_total, "/2")
            n_total += 1_000_000
    exe = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")
    out = {"workload": w["desc"], "pairs": n_total, "fastq_bytes": os.path.getsize(f1p) + os.path.getsize(f2p), "runs": []}
    for g in gs:
        tsv = os.path.join(d, "out_%d.tsv" % g)
        cmd = [exe, "-x", idx, "-k", str(w["k"]), "-1", f1p, "-2", f2p] + (["--gpus", str(g)] if g > 1 else [])
        t0 = time.perf_counter()
        with open(tsv, "wb") as fo:
            r = subprocess.run(cmd, stdout=fo, stderr=subprocess.PIPE, env=dict(os.environ, CFR_B200_STAGE_REPORT="1"))
        wall = time.perf_counter() - t0
        stages = {}
        for ln in r.stderr.decode().splitlines():
            if ln.startswith("[cfr-stages]"):
                stages = json.loads(ln[len("[cfr-stages]"):])
        h = hashlib.md5()
        with open(tsv, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        out["runs"].append({"gpus": g, "rc": r.returncode, "wall_s": wall, "pairs_per_s": n_total / wall,
                            "pairs_per_s_after_load": n_total / stages.get("pipeline_s", wall), "stages": stages,
                            "tsv_md5": h.hexdigest(), "tsv_bytes": os.path.getsize(tsv),
                            "log_tail": r.stderr.decode().splitlines()[-3:]})
        os.remove(tsv)
    out["identical_tsv"] = len({r["tsv_md5"] for r in out["runs"]}) == 1
    print(json.dumps(out))
    import shutil
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
