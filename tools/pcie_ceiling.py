"""Host<->device copy ceilings for the end-to-end path (run on the GPU box): one direction alone and both at once."""
import torch
dev = "cuda:0"
nbytes = 4096 * 100 * 100 * 4
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20, chunks=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    n = nbytes // chunks
    for _ in range(reps):
        for c in range(chunks):
            if h2d:
                with torch.cuda.stream(s1): d_in[c*n:(c+1)*n].copy_(h_in[c*n:(c+1)*n], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[c*n:(c+1)*n].copy_(d_out[c*n:(c+1)*n], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for chunks in (1, 16):
    a, b, c = run(True, False, chunks=chunks), run(False, True, chunks=chunks), run(True, True, chunks=chunks)
    print(f"chunks={chunks}: H2D {a:.3f} ms ({nbytes/a/1e6:.1f} GB/s)  D2H {b:.3f} ms ({nbytes/b/1e6:.1f} GB/s)  both {c:.3f} ms ({nbytes/c/1e6:.1f} GB/s each way)")
