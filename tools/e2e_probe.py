import sys, os, importlib, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trc = importlib.import_module("turbo-range-coder_b200"); dg = importlib.import_module("turbo-range-coder_b200.datagen")
n=100_000_000; chunk=int(sys.argv[1]) if len(sys.argv) > 1 else 1760
data = dg.zipf(n)
h_in = torch.from_numpy(data).pin_memory().numpy()
h_out = torch.empty(int(trc.lib.trc_enc_bound(n, chunk)), dtype=torch.uint8).pin_memory().numpy()
h_off = torch.empty(trc.num_chunks(n,chunk)+1, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
h_back = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
cdf = trc.cdfini(data)
for k in range(4):
    a=time.perf_counter(); s_out,s_off = trc.enc_batch_host(5, h_in, chunk, cdf=cdf, cdfnum=256, out=h_out, off=h_off); b=time.perf_counter()
    trc.dec_batch_host(5, s_out, s_off, n, chunk, cdf=cdf, cdfnum=256, out=h_back); c=time.perf_counter()
    print(f"iter {k}: enc {1e3*(b-a):.3f} ms dec {1e3*(c-b):.3f} ms", file=sys.stderr)
