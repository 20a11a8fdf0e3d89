#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of the config-2 training step, from CUPTI activity records
(torch.profiler; nsys is not in the image).  Unlike the ncu launch list (serialised, cold cache) this keeps the real
start / end of every kernel on every stream, so it shows what the step actually waits on:

    python tools/step_timeline.py [--relation-mode index_select|banked] [--out profiles/r02_timeline]

writes <out>.csv (kernel, stream, start_us, dur_us) and <out>.md: busy time (union over streams), idle gaps between
kernels, time with >= 2 kernels running, and per-kernel totals with their share of the step's wall time."""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--relation-mode", default="index_select")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline"))
    a = ap.parse_args()
    from gtos_b200 import hotpath
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    w = bench.WORKLOADS[a.workload]
    cfg = bench.make_cfg(w, 0.2)
    torch.manual_seed(19940117)
    model = hotpath.HotPath(cfg).to(dev).train()
    run = bench.StepRunner(a.workload, model, 0.2, dev, 0, 1, relation_mode=a.relation_mode)
    run.prepare()
    for _ in range(5):
        run.step_device()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            run.step_device()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    rows = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)) for e in evs))
    if not rows:
        print("no CUDA activity records")
        return
    # split into the three replays by the largest two gaps
    gaps = sorted(((rows[i + 1][0] - max(r[1] for r in rows[:i + 1]), i) for i in range(len(rows) - 1)), reverse=True)[:2]
    cuts = sorted(i for _, i in gaps)
    last = rows[cuts[-1] + 1:]
    t0 = last[0][0]
    t1 = max(r[1] for r in last)
    wall = t1 - t0
    ev = []
    for s, e, n, _ in last:
        ev.append((s, 1))
        ev.append((e, -1))
    ev.sort()
    busy = over = 0.0
    depth, prev = 0, t0
    for t, d in ev:
        if depth >= 1:
            busy += t - prev
        if depth >= 2:
            over += t - prev
        depth += d
        prev = t
    per = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n, _ in last:
        k = n.split("(")[0][:90]
        per[k][0] += 1
        per[k][1] += e - s
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out + ".csv", "w") as f:
        f.write("kernel,start_us,dur_us\n")
        for s, e, n, _ in last:
            f.write('"%s",%.2f,%.2f\n' % (n.split("(")[0][:120], s - t0, e - s))
    tot = sum(v[1] for v in per.values())
    with open(a.out + ".md", "w") as f:
        f.write(f"one CUDA-graph replay of the {a.workload} step ({a.relation_mode}): {len(last)} kernels / memcpy nodes, wall "
                f"{wall:.0f} us, busy (>= 1 kernel running) {busy:.0f} us = {100 * busy / wall:.1f} %, idle {wall - busy:.0f} us, "
                f">= 2 kernels running {over:.0f} us, sum of kernel durations {tot:.0f} us\n\n")
        f.write("| share of wall | sum us | launches | avg us | kernel |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f"| {100 * t / wall:.1f}% | {t:.1f} | {c} | {t / c:.1f} | `{k}` |\n")
    print(open(a.out + ".md").read())


if __name__ == "__main__":
    main()
