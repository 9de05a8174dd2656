#!/usr/bin/env python
"""bench.py -- read pairs/s of the classification hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c5|c3|c2|...]

One "step" = one pass of the hot path over the workload's reads.  Default workload (every N): BASELINE.json
configs[3] -- a synthetic 20 Gbp / 5000-sequence collection generated AND indexed on the GPU by this repo's
builder (tools/make_data.py: the files are ordinary *.cfr, the reference binary loads them too), HBM-resident
(10 GB of occ sectors + 40 GB of pair lines, far larger than the 126 MB L2), 10 M x 2x150 bp read pairs per
step per GPU as ten distinct 1 M-pair device batches, -k 5.  `--workload c5` is configs[4] (140 Gbp, one replica
per GPU), `c3` configs[2] (2 Gbp, index from the unmodified reference builder), `c2` configs[1].  Reads come from
seeded generators.

Numbers on the JSON line:
  value     pairs/s with the batches resident in HBM (cfr_classify_resident), CUDA-event time
  e2e       pairs/s through cfr_submit_batch / cfr_wait_batch with pinned HOST buffers: H2D of the
            reads, all kernels, D2H of the results, every step
  cli_e2e   the drop-in binary, FASTQ files -> TSV, on the step's reads (20 M reads at the default workload): process
            start to exit, with the seconds each pipeline stage (ingest / device / output) was busy
  roofline  dominant kernel (k_search).  `achieved` = its ALGORITHMIC bytes in this library's HBM layout
            (32 B per rank -- the pair lines serve the four ranks of two BackwardExtend steps with one 128-byte
            line --, 16 B per lookup probe, 2.25 bits per read base; the operations are counted in-kernel) / its
            CUDA-event time.  `achieved_dram` = the DRAM bytes ncu measured for the same launch
            (profiles/traffic.json) / the same time.  `bound` says "hbm" only when the sector array is larger than L2.
  cpu_baseline  the reference binary (oracle/_ref/centrifuger -t <cores>) on a bounded sample of the same reads, same box

Multi-GPU (torchrun, one rank per GPU): reads shard across ranks, the index is replicated per GPU, no data-path
collective; NCCL all-reduces the per-taxon counters ONCE, after the last step (inside the timed region); weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: dataset, reads per step per GPU, reads per device batch, read length, paired, -k
    "c2": dict(dataset="c2", reads=1_000_000, batch=1_000_000, rlen=100, paired=False, k=1,
               desc="synthetic 100 Mbp / 50-taxa index, 1M x 100 bp single-end reads (BASELINE configs[1])"),
    "small": dict(dataset="small", reads=200_000, batch=200_000, rlen=100, paired=False, k=1,
                  desc="synthetic 10 Mbp / 100-sequence index, 200k x 100 bp single-end reads"),
    "m700": dict(dataset="m700", reads=1_000_000, batch=1_000_000, rlen=100, paired=False, k=1,
                 desc="synthetic 700 Mbp / 175-taxa index (occ sectors 350 MB > L2), 1M x 100 bp single-end reads"),
    "m700pe": dict(dataset="m700", reads=1_000_000, batch=500_000, rlen=150, paired=True, k=5,
                   desc="synthetic 700 Mbp / 175-taxa index, 1M x 2x150 bp pairs per step, -k 5"),
    "c3": dict(dataset="c3", reads=10_000_000, batch=1_000_000, rlen=150, paired=True, k=5,
               desc="synthetic 2 Gbp / 500-taxa index, 10M x 2x150 bp read pairs, -k 5 (BASELINE configs[2])"),
    "c4": dict(dataset="c4", reads=10_000_000, batch=1_000_000, rlen=150, paired=True, k=5,
               desc="synthetic 20 Gbp / 5000-sequence index (built on the GPU), 10M x 2x150 bp read pairs per GPU, -k 5 "
                    "(BASELINE configs[3])"),
    # five resident batches: the index takes 140 of the 180 GB (70 GB of sectors, 17 GB lookup table, 17 GB sampled SA, 35 GB dense table)
    "c5": dict(dataset="c5", reads=5_000_000, batch=1_000_000, rlen=150, paired=True, k=5,
               desc="synthetic 140 Gbp / 35000-sequence index (built on the GPU), 5M x 2x150 bp read pairs per GPU and step, -k 5 "
                    "(BASELINE configs[4])"),
    "s2g": dict(dataset="s2g", reads=1_000_000, batch=1_000_000, rlen=150, paired=True, k=5,
                desc="synthetic 2 Gbp / 500-sequence index (built on the GPU), 1M x 2x150 bp read pairs, -k 5"),
    "s400": dict(dataset="s400", reads=2_000_000, batch=1_000_000, rlen=150, paired=True, k=5,
                 desc="synthetic 400 Mbp / 100-sequence index (built on the GPU), 2M x 2x150 bp read pairs, -k 5"),
    "c3s": dict(dataset="c3", reads=1_000_000, batch=1_000_000, rlen=150, paired=True, k=5,
                desc="synthetic 2 Gbp / 500-taxa index, 1M x 2x150 bp read pairs, -k 5 (one batch of BASELINE configs[2])"),
}
DEFAULT_WORKLOAD = "c4"
L2_BYTES = 126 << 20


def algorithmic_bytes(c, n_reads, bases):
    """useful index bytes in the reference layout + read bytes + result bytes."""
    return (120 * c["n_rank"] + 72 * c["n_access"] + 16 * c["n_search"] + 8 * c["n_locate"]
            + bases + 64 * n_reads)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled in-process
    through NVML every 10 ms (nvidia-smi as a fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self.h)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40),
                        ("sw_thermal_slowdown", 0x20)):
            if r & bit:
                self.reasons.add(nm)

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
        s = [x.strip() for x in out.split(",")]
        self.sm.append(float(s[0]))
        self.mx.append(float(s[1]))
        for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[2:6]):
            if v.lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._nvml_sample()
                else:
                    self._smi_sample()
            except Exception:
                pass
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def ensure_dataset(name):
    import make_data
    d = make_data.ensure(name, log=log)
    if d is None:
        raise SystemExit("dataset %s is missing and cannot be built (needs oracle/_ref/centrifuger-build)" % name)
    return os.path.join(d, "idx")


class ReadSource:
    """Seeded synthetic reads of one workload: batch j of rank r is always the same reads."""

    def __init__(self, w):
        import gen_data
        import make_data
        self.w, self.gd = w, gen_data
        self.syn = None
        if w["dataset"] in make_data.SYNTHETIC:  # drawn from the device generator's definition, no text on the host
            from centrifuger_b200 import builder
            spec = make_data.SYNTHETIC[w["dataset"]]
            self.syn = builder.SyntheticReads(spec["species"], spec["strains"], spec["genome_len"])
            return
        self.genomes = make_data.genomes_of(w["dataset"])
        self.cat = gen_data.concat_genomes(self.genomes)

    def batch(self, n, seed):
        """-> (seq1 bytes, off1, seq2 bytes | None, off2 | None), numpy arrays"""
        w, rl = self.w, self.w["rlen"]
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(rl))
        if self.syn is not None:
            r1, r2, _ = self.syn.pairs(n, rl, seed)
            return np.ascontiguousarray(r1).reshape(-1), off, np.ascontiguousarray(r2).reshape(-1), off.copy()
        if w["paired"]:
            r1, r2 = self.gd.make_reads_pe_fast(self.genomes, n, rl, seed=seed, cat=self.cat)
            return np.ascontiguousarray(r1).reshape(-1), off, np.ascontiguousarray(r2).reshape(-1), off.copy()
        r1 = self.gd.make_reads_se_fast(self.genomes, n, rl, seed=seed, cat=self.cat)
        return np.ascontiguousarray(r1).reshape(-1), off, None, None


def write_fastq_sample(seq, off, n, path, suffix=""):
    rl = int(off[1] - off[0]) if n > 0 else 0
    same = n > 0 and int(off[n]) == n * rl
    with open(path, "wb") as f:
        if same:  # one read length: build the file with numpy instead of a Python loop
            ids = np.char.add(np.char.add("@r", np.arange(n).astype(str)), suffix).astype("S")
            body = np.asarray(seq[:n * rl]).reshape(n, rl)
            qual = b"I" * rl
            out = bytearray()
            for i in range(n):
                out += ids[i] + b"\n" + body[i].tobytes() + b"\n+\n" + qual + b"\n"
            f.write(out)
        else:
            for i in range(n):
                s = seq[int(off[i]):int(off[i + 1])].tobytes()
                f.write(b"@r%d%s\n%s\n+\n%s\n" % (i, suffix.encode(), s, b"I" * len(s)))


def write_fastq_fixed(f, seq, n, rl, first_id, suffix):
    """n reads of one length appended to the open file f as 4-line FASTQ records with fixed-width ids
    (@r%09d<suffix>), built with numpy (20 M reads in seconds instead of a Python loop)."""
    sfx = np.frombuffer(suffix.encode(), dtype=np.uint8)
    head = 2 + 9 + len(sfx) + 1
    rec = head + rl + 3 + rl + 1
    a = np.empty((n, rec), dtype=np.uint8)
    a[:, 0], a[:, 1] = ord("@"), ord("r")
    ids = np.arange(first_id, first_id + n, dtype=np.int64)
    for d in range(9):
        a[:, 2 + d] = (ids // 10 ** (8 - d)) % 10 + 48
    a[:, 11:11 + len(sfx)] = sfx
    a[:, head - 1] = 10
    a[:, head:head + rl] = np.asarray(seq[:n * rl]).reshape(n, rl)
    a[:, head + rl:head + rl + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    a[:, head + rl + 3:head + 2 * rl + 3] = ord("I")
    a[:, rec - 1] = 10
    f.write(a.data)


def run_cli_e2e(idx, w, batches_np, n_gpus=1):
    """The drop-in binary end to end: FASTQ files -> centrifuger-b200 -> TSV (to /dev/null, like the reference arm), process
    start to exit, index load included; the stage seconds come from the binary itself (CFR_B200_STAGE_REPORT)."""
    exe = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")
    if not os.path.exists(exe):
        return None
    d = tempfile.mkdtemp(prefix="cfr_cli_e2e_")
    try:
        rl = w["rlen"]
        t0 = time.perf_counter()
        f1p, f2p = os.path.join(d, "r_1.fq"), os.path.join(d, "r_2.fq")
        n_total = 0
        with open(f1p, "wb") as f1, open(f2p, "wb") as f2:
            for seq1, off1, seq2, off2 in batches_np:
                n = len(off1) - 1
                write_fastq_fixed(f1, seq1, n, rl, n_total, "/1" if seq2 is not None else "")
                if seq2 is not None:
                    write_fastq_fixed(f2, seq2, n, rl, n_total, "/2")
                n_total += n
        t_write = time.perf_counter() - t0
        files = ["-1", f1p, "-2", f2p] if w["paired"] else ["-u", f1p]
        cmd = [exe, "-x", idx, "-k", str(w["k"])] + files + (["--gpus", str(n_gpus)] if n_gpus > 1 else [])
        env = dict(os.environ, CFR_B200_STAGE_REPORT="1")
        t0 = time.perf_counter()
        with open(os.devnull, "wb") as dn:
            r = subprocess.run(cmd, stdout=dn, stderr=subprocess.PIPE, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": r.stderr.decode()[-300:]}
        stages = {}
        for ln in r.stderr.decode().splitlines():
            if ln.startswith("[cfr-stages]"):
                stages = json.loads(ln[len("[cfr-stages]"):])
        pipe = stages.get("pipeline_s", wall)
        fq_bytes = os.path.getsize(f1p) + (os.path.getsize(f2p) if w["paired"] else 0)
        return {"value": n_total / wall, "unit": "pairs/s" if w["paired"] else "reads/s", "value_after_index_load": n_total / pipe,
                "reads": n_total * (2 if w["paired"] else 1), "fastq_bytes": fq_bytes, "wall_s": wall, "index_load_s": wall - pipe,
                "fastq_gb_per_s_after_load": fq_bytes / pipe / 1e9, "stages": stages, "fastq_write_s": t_write,
                "command": "centrifuger-b200 -x IDX -k %d %s > /dev/null" % (w["k"], "-1 r_1.fq -2 r_2.fq" if w["paired"] else "-u r.fq")}
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)


def run_reference_cpu(idx, w, seq1, off1, seq2, off2, n_sample, threads, repeat=1, t_load=None):
    """Time oracle/_ref/centrifuger (the unmodified reference) on the first n_sample reads of a
    batch, taken `repeat` times over (one process, `repeat` x n_sample reads).
    Returns (reads/s, seconds, cores, load seconds).  Index load time is measured with a 1-read run
    and subtracted."""
    exe = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
    if not os.path.exists(exe):
        return None
    d = tempfile.mkdtemp(prefix="cfr_bench_")
    try:
        f1 = os.path.join(d, "s_1.fq")
        write_fastq_sample(seq1, off1, n_sample, f1, "/1" if seq2 is not None else "")
        t1 = os.path.join(d, "t_1.fq")
        write_fastq_sample(seq1, off1, 1, t1, "/1" if seq2 is not None else "")
        if seq2 is not None:
            f2 = os.path.join(d, "s_2.fq")
            write_fastq_sample(seq2, off2, n_sample, f2, "/2")
            t2 = os.path.join(d, "t_2.fq")
            write_fastq_sample(seq2, off2, 1, t2, "/2")
            files, tiny = ["-1", f1, "-2", f2] * repeat, ["-1", t1, "-2", t2]  # one option pair per repetition
        else:
            files, tiny = ["-u", f1] * repeat, ["-u", t1]
        base = [exe, "-x", idx, "-t", str(threads), "-k", str(w["k"])]

        def timed(args):
            t = time.perf_counter()
            with open(os.devnull, "wb") as dn:
                subprocess.run(base + args, check=True, stdout=dn, stderr=dn)
            return time.perf_counter() - t

        if t_load is None:
            timed(tiny)  # page the index in
            t_load = min(timed(tiny), timed(tiny))
        t_run = timed(files)
        secs = max(t_run - t_load, 1e-6)
        return n_sample * repeat / secs, secs, threads, t_load
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)


def base_config(w, a, world):
    """The part of `config` both arms print identically."""
    return {"workload": w["desc"], "index": "data/%s/idx" % w["dataset"], "reads_per_step_per_gpu": w["reads"],
            "read_length": w["rlen"], "paired": w["paired"], "k": w["k"], "dust": not a.no_dust,
            "parallelism": "reads sharded over %d GPU(s), index replicated" % max(world, a.gpus)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)  # x 76 ms per step: a timed region of more than two seconds
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CFR_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS))
    ap.add_argument("--layout", type=int, default=0, help="0 auto, 1 run-block arrays, 2 occ sectors")
    ap.add_argument("--reads", type=int, default=0, help="override reads per step per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the drop-in binary's FASTQ -> TSV run (cli_e2e)")
    ap.add_argument("--no-dust", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="reads per device chunk in the end-to-end path (0 = auto)")
    a = ap.parse_args()

    w = dict(WORKLOADS[a.workload])
    if a.reads:
        w["reads"] = a.reads
        w["batch"] = min(w["batch"], a.reads)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "read pairs classified per second" if w["paired"] else "reads classified per second"
    unit = "pairs/s" if w["paired"] else "reads/s"
    config = base_config(w, a, world)
    nb = max(1, w["reads"] // w["batch"])      # device batches per step
    bn = w["batch"]
    n = nb * bn                                # reads per step per GPU

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        idx = ensure_dataset(w["dataset"])
        src = ReadSource(w)
        cores = os.cpu_count() or 1
        seq1, off1, seq2, off2 = src.batch(min(bn, 400_000), 7)
        # The unmodified reference binary, all host threads.  Each of the W + K steps classifies the same bounded
        # sample (the first n_sample reads of the step's first batch); all steps run in ONE process (the read
        # files are passed W + K times), so the index is loaded once -- its load time, measured with a
        # one-read run, is subtracted.  The sample is sized from a pilot run so the whole arm takes ~2 minutes.
        pilot_n = min(len(off1) - 1, 20_000)
        pilot = run_reference_cpu(idx, w, seq1, off1, seq2, off2, pilot_n, cores)
        if pilot is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/centrifuger not built"}))
            return 0
        t_load = pilot[3]
        total_steps = a.warmup + a.steps
        budget_s = float(os.environ.get("CFR_BENCH_REFERENCE_SECONDS", "90"))
        n_sample = a.cpu_sample or int(max(1000, min(len(off1) - 1, pilot[0] * budget_s / total_steps)))
        r = run_reference_cpu(idx, w, seq1, off1, seq2, off2, n_sample, cores, repeat=total_steps, t_load=t_load)
        value = r[0]
        secs_per_step = r[1] / total_steps
        line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * secs_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference",
                                 "sample": "first %d reads of the step's first batch per step, %d steps in one process, "
                                           "centrifuger -t %d, FASTQ in / TSV to /dev/null, index-load time (%.2f s) subtracted"
                                           % (n_sample, total_steps, cores, t_load)},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    import torch
    import centrifuger_b200 as cb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    idx = ensure_dataset(w["dataset"]) if rank == 0 or world == 1 else None
    if world > 1:
        dist.barrier()
        if idx is None:
            idx = ensure_dataset(w["dataset"])
    t0 = time.perf_counter()
    src = ReadSource(w)
    # the batches are independent seeded draws: a few of them are generated side by side (numpy releases the GIL in its kernels)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(4, nb, (os.cpu_count() or 4) // max(1, world)))) as ex:
        host = list(ex.map(lambda j: src.batch(bn, 7 + 1000 * rank + 100003 * j), range(nb)))
    log("rank %d: %d batches of %d reads generated in %.1f s" % (rank, nb, bn, time.perf_counter() - t0))
    bases = int(sum(b[0].size + (b[2].size if b[2] is not None else 0) for b in host))

    clf = cb.Classifier(idx, k=w["k"], dust=not a.no_dust, layout=a.layout, device=local_rank,
                        max_batch_reads=a.chunk)
    log("rank %d: index open in %.2f s, %.2f GB of HBM" % (rank, clf.info(18) / 1e6, clf.hbm_bytes / 1e9))
    # a dedicated (non-default) torch stream: its handle is passed through the C ABI so the
    # library's kernels and torch's CUDA events are on the same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    # pinned host staging (torch supplies pinned memory and events; the work is in libcfrb200.so)
    pin = lambda arr: torch.from_numpy(arr).pin_memory() if arr is not None else None
    pinned = [tuple(pin(x) for x in b) for b in host]
    del host
    NOUT = 3  # result buffers of the streaming loop (three batches in flight)
    outs = [(torch.empty(bn * 32, dtype=torch.uint8).pin_memory(),
             torch.empty(max(1, bn * w["k"]), dtype=torch.int64).pin_memory()) for _ in range(NOUT)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    from centrifuger_b200 import distributed as cdist

    batches = [clf.upload(*p, stream=sptr) for p in pinned]
    reduced = [None]

    def final_reduce():
        # the ONE collective of the path: per-taxon assignment counters, after the last step
        if world > 1:
            reduced[0] = cdist.final_counts(clf)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Resident stepping: the step's device batches are independent units of work, so they go round-robin over
    # three streams (what cfr_submit_batch does internally for host batches): the latency-bound stages of one batch
    # (SDUST, scoring) fill the SMs next to the memory-bound search of another.  One extra single-stream pass with
    # per-kernel CUDA events gives the stage times and the search kernel's own duration for the roofline.
    NSTREAM = 3
    streams = [stream] + [torch.cuda.Stream() for _ in range(NSTREAM - 1)]

    def step_resident(multi):
        for j, b in enumerate(batches):
            clf.classify_resident(b, stream=streams[j % NSTREAM].cuda_stream if multi else sptr)

    def fork():
        ev = torch.cuda.Event()
        ev.record(stream)
        for s2 in streams[1:]:
            s2.wait_event(ev)

    def join():
        for s2 in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s2)
            stream.wait_event(ev)

    # ---- profiling pass (single stream, every kernel bracketed by events) ----
    for _ in range(max(1, a.warmup)):
        flush.zero_()
        step_resident(False)
    sync_all()
    clf.reset_counters()
    clf.stage_times(reset=True)
    clf.set_profiling(True)
    prof_steps = max(1, min(a.steps, 3))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(prof_steps):
        step_resident(False)
    p1.record(stream)
    sync_all()
    single_stream_ms = p0.elapsed_time(p1) / prof_steps
    stage = clf.stage_times(reset=True)
    clf.set_profiling(False)
    counters = clf.counters()
    search_c = clf.stage_counters("search")
    launches_prof = counters["n_launches"]

    # ---- kernel-resident timing ----
    for _ in range(a.warmup):
        fork()
        step_resident(True)
        join()
    sync_all()
    clf.reset_counters()
    clf.taxon_counts_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fork()
    for _ in range(a.steps):
        step_resident(True)  # no L2 flush needed: index and the step's distinct batches are each larger than L2
    join()
    final_reduce()
    e1.record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = e0.elapsed_time(e1)
    # results must be complete (no reads left deferred) -- fetch also checks device error flags
    clf.fetch(batches[0], stream=sptr, out=outs[0])
    launches_resident = clf.counters()["n_launches"] + launches_prof
    tax_total = int(clf.taxon_counts()[clf.node_cnt + 1]) if world == 1 else int(reduced[0][clf.node_cnt + 1])

    # ---- end-to-end timing (pinned host buffers in, host results out) ----
    # The streaming form of the C ABI (cfr_submit_batch / cfr_wait_batch, what the CLI uses): every step
    # uploads its reads from pinned host memory, runs all kernels and downloads its results; three
    # batches are in flight so the copies of one overlap the kernels of the others.  All K steps' copies
    # and kernels are inside the timed region.
    def run_e2e(k_steps):
        inflight = []
        i = 0
        for _ in range(k_steps):
            for p in pinned:
                inflight.append(clf.submit(*p, stream=sptr, out=outs[i % NOUT]))
                i += 1
                if len(inflight) == NOUT:
                    clf.wait(inflight.pop(0)[0])
        while inflight:
            clf.wait(inflight.pop(0)[0])
        final_reduce()

    run_e2e(min(a.warmup, 2))
    sync_all()
    clf.reset_counters()
    link0 = (clf.info(13), clf.info(14))  # bytes the library has moved over the host link so far
    t0 = time.perf_counter()
    run_e2e(a.steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    link1 = (clf.info(13), clf.info(14))
    # The same loop with the bases packed by the producer (cfr_pack_reads -> cfr_submit_packed): 2-bit codes + N bits, 12 bytes
    # per 32 bases over the host link instead of 32.  Packing happens where the reads are produced (the CLI's ingest stage
    # does it while parsing), here before the timed region; what is timed is H2D of the packed words, kernels, D2H.
    t0 = time.perf_counter()
    packed = [cb.pack_batch(*p, threads=min(16, os.cpu_count() or 1), pinned=True) for p in pinned]
    pack_s = time.perf_counter() - t0

    def run_e2e_packed(k_steps):
        inflight = []
        i = 0
        for _ in range(k_steps):
            for pk, _keep in packed:
                inflight.append(clf.submit_packed(pk, stream=sptr, out=outs[i % NOUT]))
                i += 1
                if len(inflight) == NOUT:
                    clf.wait(inflight.pop(0)[0])
        while inflight:
            clf.wait(inflight.pop(0)[0])
        final_reduce()

    run_e2e_packed(min(a.warmup, 2))
    sync_all()
    plink0 = (clf.info(13), clf.info(14))
    t0 = time.perf_counter()
    run_e2e_packed(a.steps)
    sync_all()
    e2e_packed_s = time.perf_counter() - t0
    plink1 = (clf.info(13), clf.info(14))
    # single-call latency form (cfr_classify_batch: chunked copy/compute overlap inside one call)
    clf.classify_packed(*pinned[0], stream=sptr, out=outs[0])
    sync_all()
    t1 = time.perf_counter()
    for _ in range(3):
        clf.classify_packed(*pinned[0], stream=sptr, out=outs[0])
    sync_all()
    e2e_single_s = (time.perf_counter() - t1) / 3
    # diagnostic: what the host link delivers for a plain pinned copy of one batch's read bytes
    p_seq1 = pinned[0][0]
    hb = torch.empty(int(p_seq1.numel()), dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hb.copy_(p_seq1, non_blocking=True)
    ev0.record(stream)
    for _ in range(3):
        hb.copy_(p_seq1, non_blocking=True)
    ev1.record(stream)
    torch.cuda.synchronize()
    h2d_gbs = 3 * p_seq1.numel() / (ev0.elapsed_time(ev1) / 1000.0) / 1e9
    del hb
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches_e2e = clf.counters()["n_launches"]

    # ---- reduce over ranks (max time) ----
    t = torch.tensor([dev_ms, e2e_s * 1000.0, t_wall * 1000.0, e2e_packed_s * 1000.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, wall_ms_max, e2e_packed_ms_max = [float(x) for x in t.tolist()]
    total_reads = n * a.steps * world
    value = total_reads / (dev_ms_max / 1000.0)
    e2e_value = total_reads / (e2e_ms_max / 1000.0)

    if rank == 0:
        peak, peak_kind = measured_peak()
        s_ms, s_launch = stage["search"]
        # Algorithmic bytes of the search kernel = what its counted operations must read in THIS
        # library's HBM layout: one 32-byte occ sector per rank (the sp==ep symbol test reuses the
        # sp sector), 16 B per lookup-table probe, 2.25 bits per read base (2-bit code + N bit).
        # DESIGN.md section 3 states both this figure and the reference-layout one (120 B per
        # Sequence_RunBlock::Rank, 72 B per Access: SURVEY.md 8(d)), reported below as well.
        occ = clf.layout == 2
        per_rank = 32 if occ else 120
        per_access = 0 if occ else 72
        s_bytes = (per_rank * search_c["n_rank"] + per_access * search_c["n_access"] + 16 * search_c["n_search"]
                   + (bases * prof_steps * 9) // 32)
        s_bytes_ref = 120 * search_c["n_rank"] + 72 * search_c["n_access"] + 16 * search_c["n_search"] + bases * prof_steps
        per_launch_s = (s_ms / max(s_launch, 1)) / 1000.0
        achieved = (s_bytes / max(s_launch, 1)) / per_launch_s / 1e9 if s_ms > 0 else 0.0
        achieved_ref = (s_bytes_ref / max(s_launch, 1)) / per_launch_s / 1e9 if s_ms > 0 else 0.0
        # DRAM bytes of one k_search launch over one device batch, from the committed ncu --set full
        # capture of this workload (profiles/traffic.json; null if this workload has none)
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp)).get(w["dataset"] + ("pe" if w["paired"] and w["dataset"] == "m700" else ""), {})
                if tj.get("reads_per_launch") == bn:
                    traffic, traffic_src = tj.get("k_search_dram_bytes_per_launch"), tj.get("source")
            except Exception:
                traffic = None
        index_bytes = clf.info(15) if occ else clf.info(19) or clf.hbm_bytes
        hbm_resident = index_bytes > L2_BYTES
        achieved_dram = (traffic / per_launch_s / 1e9) if (traffic and s_ms > 0) else None
        total_alg = algorithmic_bytes(counters, n * prof_steps, bases * prof_steps) * (a.steps / prof_steps)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
            "details": {
                "layout": {1: "run-block arrays as stored", 2: "32-byte occ sectors (transcoded on the GPU at load)"}[clf.layout]
                          + ("; k_search walks 128-byte pair lines (two extends per line, four lanes per strand)" if clf.info(23) else ""),
                "l2": "no flush inside the timed region: the occ sectors (%.0f MB) and each step's %d distinct device batches "
                      "(%.0f MB of reads) are larger than the 126 MB L2" % (clf.info(15) / 1e6, nb, bases / 1e6)
                      if hbm_resident and bases > L2_BYTES else
                      "index or step input smaller than L2: L2-resident numbers (256 MiB device write before each warm-up step only)",
                "device_batches_per_step": nb, "reads_per_device_batch": bn,
                "index_hbm_bytes": clf.hbm_bytes, "occ_sector_bytes": clf.info(15), "wide_lookup_bytes": clf.info(16),
                "dense_locate_bytes": clf.info(17), "pair_line_bytes": clf.info(23), "runblock_bytes_released": clf.info(19),
                "dense_locate_shift": clf.info(20), "wide_lookup_width": clf.info(21), "position_bits": clf.info(22),
                "cfr_open_seconds": clf.info(18) / 1e6, "min_hit_len": clf.min_hit_len, "index_rows": clf.n,
                "reads_counted_by_final_reduce": tax_total},
            "e2e": {"value": e2e_value, "unit": unit,
                    # counted by the library from the copies it issues (reads of one length: the offsets
                    # are generated on the device and do not cross the link)
                    "h2d_bytes_per_step": int((link1[0] - link0[0]) // a.steps),
                    "d2h_bytes_per_step": int((link1[1] - link0[1]) // a.steps),
                    "api": "cfr_submit_batch / cfr_wait_batch, three batches in flight, pinned host buffers",
                    "single_call_value": bn / e2e_single_s,
                    "host_link_h2d_gbs": h2d_gbs},
            "e2e_packed": {"value": total_reads / (e2e_packed_ms_max / 1000.0), "unit": unit,
                           "h2d_bytes_per_step": int((plink1[0] - plink0[0]) // a.steps),
                           "d2h_bytes_per_step": int((plink1[1] - plink0[1]) // a.steps),
                           "api": "cfr_submit_packed / cfr_wait_batch: the producer hands over 2-bit codes + N bits (cfr_pack_reads, "
                                  "%.2f s for the step's reads with %d host threads, outside the timed region)"
                                  % (pack_s, min(16, os.cpu_count() or 1))},
            "gpu_launches": int(launches_resident + launches_e2e),
            "roofline": {"bound": "hbm" if hbm_resident else "l2/issue", "kernel": "k_search",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": peak_kind,
                         "achieved_useful": achieved, "achieved_dram": achieved_dram,
                         "frac_dram": (achieved_dram / peak) if achieved_dram else None,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": s_bytes / max(s_launch, 1),
                         "kernel_ms_per_launch": s_ms / max(s_launch, 1),
                         "kernel_share_of_step": s_ms / max(sum(v[0] for v in stage.values()), 1e-9),
                         "reference_layout_gbs": achieved_ref,
                         "note": ("occ sectors (%.0f MB) exceed the 126 MB L2: every rank is a DRAM line fill; `achieved` counts the "
                                  "32 useful bytes of it, `achieved_dram` the bytes DRAM really moved (128-byte fills)" % (clf.info(15) / 1e6))
                                 if hbm_resident else
                                 ("occ sectors (%.0f MB) fit the 126 MB L2: `achieved` is L2 bandwidth, not HBM; DRAM traffic is in "
                                  "`traffic`" % (clf.info(15) / 1e6)),
                         "pipeline_algorithmic_gbs": total_alg / (dev_ms_max / 1000.0) / 1e9 / world},
            "stage_ms_per_step": {k: v[0] / prof_steps for k, v in stage.items()},
            "single_stream_ms_per_step": single_stream_ms,
            "resident_streams": NSTREAM,
            "ops_per_read": {k: counters[k] / (n * prof_steps) for k in
                             ("n_rank", "n_access", "n_search", "n_locate", "n_lf", "n_extend")},
            "clocks": sampler.summary(),
            "wall_ms_per_step": wall_ms_max / a.steps,
        }
        if not a.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            seq1, off1, seq2, off2 = [x.numpy() if x is not None else None for x in pinned[0]]
            n_sample = a.cpu_sample or min(bn, (250_000 if not w["paired"] else 100_000) * max(1, cores // 8))
            r = run_reference_cpu(idx, w, seq1, off1, seq2, off2, n_sample, cores)
            rep = 1
            if r is not None and r[1] < 10.0 and not a.cpu_sample:
                # bounded sample of about 15 s of CPU work: the same reads taken several times over
                rep = int(min(64, max(2, round(15.0 / max(r[1], 0.05)))))
                r = run_reference_cpu(idx, w, seq1, off1, seq2, off2, n_sample, cores, repeat=rep, t_load=r[3])
            if r is not None:
                line["cpu_baseline"] = {"value": r[0], "unit": unit, "cores": cores, "kind": "reference",
                                        "sample": "first %d reads of the step's first batch x %d, centrifuger -t %d, %.1f s, "
                                                  "FASTQ in / TSV to /dev/null, index-load time subtracted" % (n_sample, rep, cores, r[1])}
    for b in batches:
        b.free()
    clf.close()
    if rank == 0:
        if not a.no_cli and world == 1:
            # the drop-in binary on the step's reads as FASTQ files (20 M reads at the default workload): a separate
            # process, so this one's handle is closed first (two replicas of a 20 Gbp index do not fit one GPU)
            torch.cuda.empty_cache()
            line["cli_e2e"] = run_cli_e2e(idx, w, [[x.numpy() if x is not None else None for x in p] for p in pinned])
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
