"""One-off probe: what the host<->device link of this box can move (pinned memory, large copies), the ceiling of the
host-buffer (e2e) path.  Prints GB/s for H2D alone, D2H alone and both directions at once."""
import json, time, torch

def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        return reps * n / dt / 1e9
    run(True, True, 1)
    res = {"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True), "both_each_gbs": run(True, True)}
    # e2e ceiling for 24 B in / 16 B out per segment, both directions overlapped
    hb, db = 24, 16
    s = 32 << 20          # 768 MB in, 512 MB out: both fit the 1 GiB buffers
    hi = h_in[: s * hb]; ho = h_out[: s * db]; di = d_in[: s * hb]; do = d_out[: s * db]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res["copy_only_segments_per_s_24in_16out"] = 5 * s / dt
    print(json.dumps(res))

if __name__ == "__main__":
    main()
