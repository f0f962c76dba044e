"""Link speed of chunked copies against one large copy, both directions, pinned host memory (no host threads
involved): what the ring's geometry can cost by itself.  GPU box only."""
import torch

n = 2 * 7938000
dev = torch.empty(n, dtype=torch.float32, device="cuda")
host = torch.empty(n, dtype=torch.float32).pin_memory()
ring = torch.empty(8 * (1 << 19), dtype=torch.float32).pin_memory()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for chunk in (n, 1 << 21, 1 << 20, 1 << 19, 1 << 18, 1 << 16):
    events = [torch.cuda.Event() for _ in range(64)]

    def d2h(into_ring):
        k = 0
        for base in range(0, n, chunk):
            m = min(chunk, n - base)
            if into_ring and chunk <= (1 << 19):
                dst = ring[(k % 8) * (1 << 19):(k % 8) * (1 << 19) + m]
            else:
                dst = host[base:base + m]
            dst.copy_(dev[base:base + m], non_blocking=True)
            events[k % 64].record()
            k += 1

    def h2d():
        k = 0
        for base in range(0, n, chunk):
            m = min(chunk, n - base)
            dev[base:base + m].copy_(host[base:base + m], non_blocking=True)
            events[k % 64].record()
            k += 1

    t_d2h, t_ring, t_h2d = timed(lambda: d2h(False)), timed(lambda: d2h(True)), timed(h2d)
    gb = n * 4 / 1e6
    print(f"chunk {chunk * 4 / 1e6:8.2f} MB: D2H {t_d2h:6.2f} ms ({gb / t_d2h:5.1f} GB/s), D2H into an 8-slot ring {t_ring:6.2f} ms "
          f"({gb / t_ring:5.1f} GB/s), H2D {t_h2d:6.2f} ms ({gb / t_h2d:5.1f} GB/s)")
