#!/usr/bin/env python
"""What the host link of this box delivers, for the end-to-end figures (spmvb200_spmv_host is bound by it).

    python tools/pcie_yardstick.py [--mb 1024] [--reps 5]

Pinned host buffers, plain cudaMemcpyAsync through torch (no kernel of ours): H2D alone, D2H alone, both directions at
once on two streams, and the traffic pattern of one end-to-end config-5 step (2 x mb up, 1 x mb down, concurrently).
Prints one JSON line.
"""
import argparse
import json

import torch


def timed(fn, reps):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.mb * (1 << 20) // 8
    h_up = [torch.ones(n, dtype=torch.float64).pin_memory() for _ in range(2)]
    h_dn = torch.empty(n, dtype=torch.float64).pin_memory()
    d_up = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    d_dn = torch.ones(n, dtype=torch.float64, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()
    gb = n * 8 / 1e9

    def h2d():
        d_up[0].copy_(h_up[0], non_blocking=True)

    def d2h():
        h_dn.copy_(d_dn, non_blocking=True)

    def both(ups):
        def run():
            s_up.wait_stream(cur)
            s_dn.wait_stream(cur)
            with torch.cuda.stream(s_up):
                for k in range(ups):
                    d_up[k].copy_(h_up[k], non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_dn, non_blocking=True)
            cur.wait_stream(s_up)
            cur.wait_stream(s_dn)
        return run

    t_up, t_dn = timed(h2d, args.reps), timed(d2h, args.reps)
    t_both, t_step = timed(both(1), args.reps), timed(both(2), args.reps)
    print(json.dumps({
        "mb_per_buffer": args.mb,
        "h2d_gbs": gb / t_up * 1e3, "d2h_gbs": gb / t_dn * 1e3,
        "bidirectional_ms": t_both, "bidirectional_h2d_gbs": gb / t_both * 1e3,
        "config5_step_pattern_ms": t_step, "config5_step_pattern_h2d_gbs": 2 * gb / t_step * 1e3,
        "what": "pinned cudaMemcpyAsync, best of %d; config5_step_pattern = 2 buffers up + 1 down concurrently "
                "(x and y_old up, y_new down: the bytes of one spmvb200_spmv_host step on config 5 when mb = 1024)" % args.reps,
    }))


if __name__ == "__main__":
    main()
