#!/usr/bin/env python
"""Phase timeline of k_rcs2_enc3 (globaltimer stamps per warp): where the encode kernel's time goes.  GPU only."""
import ctypes, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trc = importlib.import_module("turbo-range-coder_b200")
dg = importlib.import_module("turbo-range-coder_b200.datagen")
size, chunk = 100_000_000, int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1760
data = dg.zipf(size)
d_in = torch.from_numpy(data).cuda()
b = trc.DeviceBatch(trc.RCS2, size, chunk, cdfnum=256)
b.cdf, _ = trc.cdfini_dev(d_in, size, size)
for _ in range(3):
    b.encode(d_in)
torch.cuda.synchronize()
DEC = "--dec" in sys.argv
buf = torch.zeros(4096 * 32 * 8, dtype=torch.int64, device="cuda")
trc.lib.trc_debug_enc_times(ctypes.c_void_p(buf.data_ptr()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
if DEC:
    b.decode()
else:
    b.encode(d_in)
torch.cuda.synchronize()
trc.lib.trc_debug_enc_times(ctypes.c_void_p(0))
t = buf.cpu().numpy().reshape(-1, 32, 8).astype(np.float64)
used = t[:, :, 0] > 0
t0 = t[:, :, 0][used].min()
names = ["start", "tables", "loop end", "flush done", "barrier passed", "look-back done", "copy done"]
print(f"chunk {chunk}: {used.sum()} warps in {used.any(1).sum()} CTAs; times in us relative to the first warp's start")
if DEC:
    names = ["start", None, "loop end"]
for k, nm in enumerate(names):
    if nm is None:
        continue
    v = (t[:, :, k][used] - t0) / 1e3
    print(f"  {nm:16s} min {v.min():8.1f}  p10 {np.percentile(v, 10):8.1f}  median {np.median(v):8.1f}  p90 {np.percentile(v, 90):8.1f}  max {v.max():8.1f}")
d = (t[:, :, 2] - t[:, :, 0 if DEC else 1])[used] / 1e3
print(f"  main loop per warp: min {d.min():.1f} median {np.median(d):.1f} max {d.max():.1f} us")
if not DEC:
    d = (t[:, :, 6] - t[:, :, 4])[used] / 1e3
    print(f"  epilogue after the barrier per warp: min {d.min():.1f} median {np.median(d):.1f} max {d.max():.1f} us")
