#!/usr/bin/env python
"""One process, many kernel variants: opens the workload's index once per setting (the CFR_B200_* switches
are read by cfr_open), classifies the same resident batches and prints the stage times.

    python tools/sweep.py c4 3 "" CFR_B200_PAIR_FETCH=1 CFR_B200_DENSE_LOCATE=2,CFR_B200_PAIR_SEARCH_BLOCKS=6

argv: workload, device batches per step, then one setting per argument ("" = defaults; several switches joined by commas).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

import bench  # noqa: E402
import centrifuger_b200 as cb  # noqa: E402


def main():
    wname, nb = sys.argv[1], int(sys.argv[2])
    settings = sys.argv[3:] or [""]
    w = bench.WORKLOADS[wname]
    idx = bench.ensure_dataset(w["dataset"])
    src = bench.ReadSource(w)
    bn = w["batch"]
    host = [src.batch(bn, 7 + 100003 * j) for j in range(nb)]
    pinned = [tuple(torch.from_numpy(x).pin_memory() if x is not None else None for x in b) for b in host]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    nstream = int(os.environ.get("SWEEP_STREAMS", "3"))
    streams = [stream] + [torch.cuda.Stream() for _ in range(nstream - 1)]
    steps = int(os.environ.get("SWEEP_STEPS", "4"))
    ref_sig = None
    for setting in settings:
        keys = []
        if setting in ("default", '""'):
            setting = ""
        for kv in filter(None, setting.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
            keys.append(k)
        t0 = time.perf_counter()
        clf = cb.Classifier(idx, k=w["k"], device=0)
        t_open = time.perf_counter() - t0
        batches = [clf.upload(*p, stream=sptr) for p in pinned]

        def step(multi):
            for j, b in enumerate(batches):
                clf.classify_resident(b, stream=streams[j % nstream].cuda_stream if multi else sptr)

        for _ in range(2):
            step(False)
        torch.cuda.synchronize()
        clf.reset_counters()
        clf.stage_times(reset=True)
        clf.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step(False)
        e1.record(stream)
        torch.cuda.synchronize()
        single = e0.elapsed_time(e1) / steps
        stage = clf.stage_times(reset=True)
        clf.set_profiling(False)
        # three streams
        step(True)
        torch.cuda.synchronize()
        ev = torch.cuda.Event()
        e0.record(stream)
        ev.record(stream)
        for s2 in streams[1:]:
            s2.wait_event(ev)
        for _ in range(steps):
            step(True)
        for s2 in streams[1:]:
            e = torch.cuda.Event()
            e.record(s2)
            stream.wait_event(e)
        e1.record(stream)
        torch.cuda.synchronize()
        multi = e0.elapsed_time(e1) / steps
        # the same three-stream step with every stage bracketed by events: how long each stage takes next to the others
        clf.stage_times(reset=True)
        clf.set_profiling(True)
        for _ in range(2):
            step(True)
        torch.cuda.synchronize()
        stage3 = clf.stage_times(reset=True)
        clf.set_profiling(False)
        res, ids = clf.fetch(batches[0], stream=sptr)
        sig = (int(res["score"].astype("uint64").sum()), int(ids.sum()), int(res["n_assign"].sum()))
        if ref_sig is None:
            ref_sig = sig
        out = {"setting": setting or "default", "open_s": round(t_open, 2), "hbm_gb": round(clf.hbm_bytes / 1e9, 1),
               "dense_shift": clf.info(20), "ms_per_batch_1stream": round(single / nb, 3),
               "ms_per_batch_3streams": round(multi / nb, 3),
               "Mpairs_s_1stream": round(bn * nb / single / 1e3, 2), "Mpairs_s_3streams": round(bn * nb / multi / 1e3, 2),
               "stage_ms_per_batch": {k: round(v[0] / steps / nb, 3) for k, v in stage.items()},
               "stage_ms_per_batch_3streams": {k: round(v[0] / 2 / nb, 3) for k, v in stage3.items()},
               "same_results_as_first": sig == ref_sig}
        print(json.dumps(out), flush=True)
        for b in batches:
            b.free()
        clf.close()
        for k in keys:
            del os.environ[k]
    return 0


if __name__ == "__main__":
    sys.exit(main())
