"""Duplex PCIe rate for the e2e product's volumes (80 MB in + 80 MB out) when each direction is issued as c back-to-back
copies (no dependencies between the directions): the ceiling of ANY staged pipeline with c stages on this box.
Also the same with the D2H side lagging the H2D side by one chunk (what a pipeline does).  Prints one line per c."""
import torch

N = 80_000_000
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(c, lag, reps=7):
    ts = []
    step = N // c
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        evs = [torch.cuda.Event() for _ in range(c)]
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        with torch.cuda.stream(s1):
            for i in range(c):
                d_in[i * step:(i + 1) * step].copy_(h_in[i * step:(i + 1) * step], non_blocking=True)
                evs[i].record()
            e1.record()
        with torch.cuda.stream(s2):
            for i in range(c):
                if lag:
                    s2.wait_event(evs[i])
                h_out[i * step:(i + 1) * step].copy_(d_out[i * step:(i + 1) * step], non_blocking=True)
            e2.record()
        torch.cuda.synchronize()
        ts.append(max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for c in (1, 2, 4, 6, 8, 16, 32):
    for lag in (False, True):
        med, best = run(c, lag)
        print(f"chunks {c:2d} lag={int(lag)}: median {med:.3f} ms  best {best:.3f} ms  -> {2 * N / med / 1e6:.1f} GB/s aggregate", flush=True)
