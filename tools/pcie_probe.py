"""PCIe probe: pinned H2D / D2H bandwidth by copy size and stream count (events on the copy streams)."""
import torch
dev = torch.device("cuda:0")
def bw(nbytes, direction, streams, reps=30):
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go():
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                if direction in ("h2d", "both"): d.copy_(h, non_blocking=True)
                if direction == "d2h": h.copy_(d, non_blocking=True)
    for _ in range(5): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss: s.wait_event(e0)
    for _ in range(reps): go()
    for s in ss: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return reps * streams * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
for rnd in range(2):
    for mb in (1, 3, 13, 64):
        for st in (1, 2):
            print(f"round {rnd}: {mb:3d} MB x {st} stream(s): H2D {bw(mb << 20, 'h2d', st):5.1f} GB/s   D2H {bw(mb << 20, 'd2h', st):5.1f} GB/s", flush=True)
