"""H2D of one step's points (25.2 MB, pinned) as 1 copy vs k concurrent copies on k streams (tools only)."""
import torch
dev = torch.device('cuda:0')
n = 4096 * 512 * 3
host = torch.empty(n, dtype=torch.float32).pin_memory()
dst = torch.empty(n, dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
for k in (1, 2, 3, 4):
    per = -(-n // k)
    def run(reps):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in streams[:k]:
            s.wait_event(t0)
        for _ in range(reps):
            for i, s in enumerate(streams[:k]):
                with torch.cuda.stream(s):
                    dst[i * per:(i + 1) * per].copy_(host[i * per:(i + 1) * per], non_blocking=True)
        for s in streams[:k]:
            torch.cuda.current_stream().wait_stream(s)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps
    run(5)
    ms = min(run(40) for _ in range(3))
    print('%d stream(s): %.4f ms per 25.2 MB = %.1f GB/s' % (k, ms, n * 4 / ms / 1e6))
