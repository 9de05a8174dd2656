#!/usr/bin/env python
"""Condense the source page of one kernel of an .ncu-rep into runs of SASS instructions with equal
execution count: where the warp instructions and the stall samples of a state-machine kernel go.
usage: ncu_blocks.py REPORT.ncu-rep KERNEL_REGEX [min_pct]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern if kern.startswith("=") is False and "|" not in kern and "\\" not in kern and "(" not in kern else "regex:" + kern],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
    # a report may hold several launches of the kernel: keep the first
    first = data[0][0]
    for j in range(1, len(data)):
        if data[j][0] == first:
            data = data[:j]
            break
    base = int(data[0][0], 16)
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    print("%s: %d SASS instructions, %d warp instructions executed, %d samples" % (rows[0][1][:60], len(data), tot_inst, tot))
    cur = None
    acc = None

    def flush():
        if cur is not None and acc["samp"] >= tot * min_pct / 100.0:
            print("%04x..%04x %3d instrs x %9s exec, thr %5s | warp-inst %5.1f%% samples %5.1f%% (long_sb %4.1f%% short_sb %4.1f%%) [%s]" % (
                acc["start"], acc["last"], acc["cnt"], cur[0], cur[1], 100.0 * acc["cnt"] * int(cur[0]) / max(tot_inst, 1),
                100.0 * acc["samp"] / max(tot, 1), 100.0 * acc["lsb"] / max(tot, 1), 100.0 * acc["ssb"] / max(tot, 1), " ".join(acc["ops"][:9])))

    for r in data:
        a = int(r[0], 16) - base
        key = (r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]])
        if key != cur:
            flush()
            cur = key
            acc = dict(start=a, last=a, cnt=0, samp=0, lsb=0, ssb=0, ops=[])
        acc["cnt"] += 1
        acc["samp"] += int(r[ix["# Samples"]])
        acc["lsb"] += int(r[ix["stall_long_sb"]])
        acc["ssb"] += int(r[ix["stall_short_sb"]])
        acc["last"] = a
        ins = r[1].strip().split()
        acc["ops"].append(ins[1] if ins[0].startswith("@") else ins[0])
    flush()


if __name__ == "__main__":
    main()
