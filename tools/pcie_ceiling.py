#!/usr/bin/env python
"""What the box can feed: concurrent pinned host<->device copy bandwidth over N = 1, 2, 4, 8 GPUs (one process, one copy
stream pair per device, H2D and D2H alone and together).  The end-to-end frames/s of bench.py cannot exceed
    ceiling_fps = N * 256 / max(t_h2d(236 MB), t_d2h(97 MB))  per step,
which is what `e2e.pcie` reports in the same run; this tool gives the raw GB/s per N for DESIGN.md.

usage (GPU box): python tools/pcie_ceiling.py [--mb 256] [--reps 10] > gpurun_out/pcie_ceiling.json"""
import argparse
import json
import time

import torch


def measure(devs, nbytes, reps, up, down):
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            bufs.append((torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8, device="cuda"),
                         torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8, device="cuda"),
                         torch.cuda.Stream(), torch.cuda.Stream()))

    def once():
        for d, (hi, di, ho, do, s1, s2) in zip(devs, bufs):
            with torch.cuda.device(d):
                if up:
                    with torch.cuda.stream(s1):
                        di.copy_(hi, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        ho.copy_(do, non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)
    once()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = (time.perf_counter() - t0) / reps
    return len(devs) * nbytes / dt / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    ndev = torch.cuda.device_count()
    out = {"bytes_per_copy": a.mb << 20, "devices_visible": ndev, "runs": []}
    for n in (1, 2, 4, 8):
        if n > ndev:
            break
        devs = list(range(n))
        out["runs"].append({"n_gpus": n, "h2d_gbs": measure(devs, a.mb << 20, a.reps, True, False),
                            "d2h_gbs": measure(devs, a.mb << 20, a.reps, False, True),
                            "both_gbs_each_way": measure(devs, a.mb << 20, a.reps, True, True)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
