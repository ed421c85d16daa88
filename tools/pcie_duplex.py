"""Host<->device copy ceilings of this box (pinned memory, CUDA events): H2D alone, D2H alone, both at once on two
streams.  The end-to-end SpMV (`e2e` of bench.py: 80 MB of x in + 80 MB of y out per product) is bound by the
last figure; run it next to the bench so the e2e number can be read against the platform, not against a nominal
PCIe figure.  Prints one JSON line."""
import json

import torch

n = 256 << 20  # 256 MiB per direction
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
                e1.record()
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
                e2.record()
        torch.cuda.synchronize()
        t = max(e0.elapsed_time(e1) if h2d else 0.0, e0.elapsed_time(e2) if d2h else 0.0)
        best = min(best, t)
    return best * 1e-3


t_h2d, t_d2h, t_both = run(True, False), run(False, True), run(True, True)
out = {"bytes_per_direction": n, "h2d_GBs": n / t_h2d / 1e9, "d2h_GBs": n / t_d2h / 1e9,
       "duplex_aggregate_GBs": 2 * n / t_both / 1e9, "duplex_ms": t_both * 1e3,
       "l5_e2e_floor_ms": 160e6 / (2 * n / t_both), "l5_e2e_ceiling_gflops": 2 * 49978572 / (160e6 / (2 * n / t_both)) / 1e9}
print(json.dumps(out))
