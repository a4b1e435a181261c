"""In-situ timeline of the pipeline (torch.profiler / CUPTI): per-kernel durations as they occur
inside a step -- warm caches, two streams -- and the GPU idle time between kernels.

    python tools/timeline.py [--steps 5] > gpurun_out/timeline.md
"""
import argparse
import collections
import json
import os
import re
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.profiler import ProfilerActivity, profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--atoms", type=int, default=30000)
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--delta", type=float, default=0.5)
    ap.add_argument("--dump-step", action="store_true", help="also list every kernel of the last step with its stream and start time")
    args = ap.parse_args()
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    p = SynthParams(n_atoms=args.atoms, n_frames=args.frames)
    edges = grid_edges_from_frame0(generate(p, 0, 1)[0], delta=args.delta)
    xyz = K.synth_generate(p)
    pipe = HopPipeline(edges, args.frames, PipelineConfig())
    for _ in range(3):
        pipe.run(xyz)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            pipe.run(xyz)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    with open(path) as f:
        trace = json.load(f)
    ev = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    # union of busy intervals over all streams
    busy, cur_s, cur_e = 0.0, None, None
    gaps = []
    for e in ev:
        s, t = e["ts"], e["ts"] + e["dur"]
        if cur_e is None:
            cur_s, cur_e = s, t
        elif s <= cur_e:
            cur_e = max(cur_e, t)
        else:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, e["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:60]))
            cur_s, cur_e = s, t
    busy += cur_e - cur_s
    per = collections.OrderedDict()
    for e in ev:
        name = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"^at::native::", "", name)
        name = re.split(r"[<(]", name)[0][:48] or e["name"][:48]
        a = per.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
    n = args.steps
    print("# In-situ timeline, %d steps of %d atoms x %d frames (torch.profiler / CUPTI)\n" % (n, args.atoms, args.frames))
    print("wall per step %.1f us, GPU busy (union over streams) %.1f us, idle %.1f us (%.1f %%)\n"
          % ((t1 - t0) / n, busy / n, (t1 - t0 - busy) / n, 100 * (t1 - t0 - busy) / (t1 - t0)))
    print("| kernel | launches/step | us/step | mean us |")
    print("|---|---:|---:|---:|")
    for k, (c, d) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %.1f | %.1f | %.1f |" % (k, c / n, d / n, d / c))
    if args.dump_step:
        hists = [e for e in ev if "hist_kernel" in e["name"]]
        start = hists[-1]["ts"]
        print("\nLast step, kernel by kernel (start us after the histogram kernel's start, duration, stream):\n")
        for e in ev:
            if e["ts"] >= start:
                nm = re.split(r"[<(]", e["name"].replace("(anonymous namespace)::", "").replace("void ", ""))[0][:40]
                print("%9.1f %8.1f  s%-3s %s" % (e["ts"] - start, e["dur"], e.get("args", {}).get("stream", "?"), nm))
    print("\nLargest idle gaps (us, kernel that follows):")
    for g, name in sorted(gaps, reverse=True)[:12]:
        print("* %.1f before %s" % (g, name))


if __name__ == "__main__":
    main()
