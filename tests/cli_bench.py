#!/usr/bin/env python
"""End-to-end check of the drop-in CLI on the GPU box: FASTQ -> TSV wall time of
centrifuger-b200 vs the reference binary on the same file, and byte equality of
the two TSVs (parity at scale).  usage: cli_bench.py [workload] [n_reads]"""
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
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
    w = dict(bench.WORKLOADS[wl])
    w["reads"] = n
    idx = bench.ensure_dataset(w["dataset"])
    seq1, off1, seq2, off2 = bench.ReadSource(w).batch(n, 7)
    d = tempfile.mkdtemp(prefix="cfr_cli_")
    f1 = os.path.join(d, "r_1.fq")
    bench.write_fastq_sample(seq1, off1, n, f1, "/1" if seq2 is not None else "")
    files = ["-u", f1]
    if seq2 is not None:
        f2 = os.path.join(d, "r_2.fq")
        bench.write_fastq_sample(seq2, off2, n, f2, "/2")
        files = ["-1", f1, "-2", f2]
    ours = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
    cores = os.cpu_count() or 1
    out = {}
    for name, cmd in (("ours", [ours, "-x", idx, "-k", str(w["k"])] + files),
                      ("ours_again", [ours, "-x", idx, "-k", str(w["k"])] + files),
                      ("reference", [ref, "-x", idx, "-k", str(w["k"]), "-t", str(cores)] + files)):
        o = os.path.join(d, name + ".tsv")
        t = time.perf_counter()
        with open(o, "wb") as fo:
            subprocess.run(cmd, check=True, stdout=fo, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t
        out[name] = {"seconds": dt, "reads_per_s": n / dt, "md5": hashlib.md5(open(o, "rb").read()).hexdigest()}
    out["identical_tsv"] = out["ours"]["md5"] == out["reference"]["md5"]
    out["workload"], out["n_reads"], out["cores"] = wl, n, cores
    print(json.dumps(out))
    import shutil
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
