#!/usr/bin/env python
"""bench.py -- encode+decode GB/s of the B200 CDF entropy path on BASELINE.json's workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--codec rcs2|ans4s|rcs|ans|rc|...] [--chunk BYTES] [--size BYTES]

One "step" = one encode pass + one decode pass of the hot path over one batch: SIZE bytes (default
100 000 000) of synthetic Zipf(1.1) bytes cut into CHUNK-byte chunks (default 4096), every chunk coded exactly
as one call of the reference function (default codec rcs2 = rccdfs2enc / rccdfsb2dec, what `turborc -e45` runs,
the id BASELINE.json's CPU config names).  The static table comes from cdfini on the whole buffer, computed
outside the timed region exactly like the reference harness does (turborc.c:429-433).

value   = SIZE / (t_encode + t_decode), GB = 1e9, buffers resident in HBM, CUDA-event time on the launching
          stream, L2 flushed (256 MiB write) before every timed encode and decode.
e2e     = the same through the C-ABI host entry points (trc_enc_batch_host / trc_dec_batch_host) on pinned host
          buffers: H2D of the input, D2H of the packed stream and offsets, H2D of the stream, D2H of the decoded
          bytes are all inside the timed region (wall clock; the calls are synchronous).
roofline / cpu_baseline: see DESIGN.md section "Measurement".
With --gpus N > 1 (torchrun): every rank codes its own SIZE-byte shard (weak scaling), the packed streams are
gathered on rank 0 over NCCL inside the step, time = max over ranks.
--impl reference: the reference's own CPU implementation (oracle/_ref, else the oracle port) on all host
threads, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODECS = {"ans4s": 0, "ans4": 1, "ans": 2, "ans1": 3, "rcs": 4, "rcs2": 5, "rc": 6, "rci": 7, "rc4": 8, "rc4i": 9, "answ": 10, "rc8": 11, "rci8": 12,
          "ansu16": 13, "ansuz16": 14, "ansv16": 15, "ansvz16": 16, "ansv32": 17, "ansvz32": 18, "rcv16": 19, "rcvz16": 20, "rcv32": 21, "rcvz32": 22, "rcu16": 23, "rcu32": 24}
REF_FN = {0: ("anscdf4senc", "anscdf4sdec"), 1: ("anscdf4enc", "anscdf4dec"), 2: ("anscdfenc", "anscdfdec"),
          3: ("anscdf1enc", "anscdf1dec"), 4: ("rccdfsenc", "rccdfsbdec"), 5: ("rccdfs2enc", "rccdfsb2dec"),
          6: ("rccdfenc", "rccdfdec"), 7: ("rccdfienc", "rccdfidec"), 8: ("rccdf4enc", "rccdf4dec"),
          9: ("rccdf4ienc", "rccdf4idec"),
          11: ("rccdfenc8", "rccdfdec8"), 12: ("rccdfienc8", "rccdfidec8"),
          13: ("anscdfuenc16", "anscdfudec16"), 14: ("anscdfuzenc16", "anscdfuzdec16"), 15: ("anscdfvenc16", "anscdfvdec16"),
          16: ("anscdfvzenc16", "anscdfvzdec16"), 17: ("anscdfvenc32", "anscdfvdec32"), 18: ("anscdfvzenc32", "anscdfvzdec32"),
          19: ("rccdfvenc16", "rccdfvdec16"), 20: ("rccdfvzenc16", "rccdfvzdec16"), 21: ("rccdfvenc32", "rccdfvdec32"),
          22: ("rccdfvzenc32", "rccdfvzdec32"), 23: ("rccdfuenc16", "rccdfudec16"), 24: ("rccdfuenc32", "rccdfudec32"),
          10: ("answenc", "answdec")}       # this repository's 32-way interleaved static rANS: no reference function, CPU leg = oracle port
CALLS_PER_SM = 384     # resident reference calls per SM in one wave of the lane-per-coder rcs2 kernels (2 CTAs x 192 calls x 2 lanes)
B200_SMS = 148         # the default chunk is sized for this part; both arms derive it from the same constants (no device query)
METRIC = "encode+decode GB/s on 100MB order-0 byte stream; bitstream bit-exact vs ref"
SRC_NAME = {"zipf-dev": "Zipf(1.1) (generated on the device)", "zipf": "Zipf(1.1)", "bwt": "BWT-shaped", "o1": "order-1 Markov", "uniform": "uniform random"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, k), int)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bit and (r & bit) and "None" not in name and "All" not in name:
                        self.reasons.add(name.replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(1.0)
        s = sorted(self.samples)
        norm = {"GpuIdle": "gpu_idle", "ApplicationsClocksSetting": "applications_clocks_setting", "SwPowerCap": "sw_power_cap",
                "HwSlowdown": "hw_slowdown", "SyncBoost": "sync_boost", "SwThermalSlowdown": "sw_thermal_slowdown",
                "HwThermalSlowdown": "hw_thermal_slowdown", "HwPowerBrakeSlowdown": "hw_power_brake_slowdown",
                "DisplayClockSetting": "display_clock_setting", "UserDefinedClocks": "applications_clocks_setting"}
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted({norm.get(r, r) for r in self.reasons} - {"gpu_idle"})}


def datagen():
    """datagen.py loaded by path: importing the package would dlopen libtrc_b200.so, which the reference arm must not do."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("trc_datagen", os.path.join(ROOT, "turbo-range-coder_b200", "datagen.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def default_chunk(args):
    """--chunk 0: one balanced wave of coder chains on a B200 (CALLS_PER_SM reference calls per SM), a multiple of 16 bytes."""
    if args.chunk:
        return args.chunk
    chunk = -(-args.size // (B200_SMS * CALLS_PER_SM)) if args.codec == "rcs2" else 4096
    return min(65536, max(256, (chunk + 15) & ~15))


def workload_config(args, chunk, world):
    """`config` of the JSON line -- built by ONE function so that both arms (ours / --impl reference) name the same workload."""
    codec = CODECS[args.codec]
    static = codec in (0, 4, 5, 10)
    n_chunks = -(-args.size // chunk)
    tab = ("static CDF (cdfini per " + (str(args.cdf_block) + "-byte block" if args.cdf_block else "whole buffer") + "), ") if static else "adaptive model, "
    return {"workload": f"{args.size} B {SRC_NAME[args.src]} bytes per GPU, {tab}batch of {chunk}-byte chunks, "
                        f"each chunk == one reference call ({REF_FN[codec][0]}/{REF_FN[codec][1]})",
            "codec": args.codec, "chunk_bytes": chunk, "n_chunks": n_chunks,
            "chunk_choice": ("one wave on %d SMs: %d calls per SM" % (B200_SMS, CALLS_PER_SM)) if not args.chunk else "--chunk",
            "l2": "GPU arm: inputs larger than L2 -- three rotating buffer sets, a kernel's input was last touched ~0.8 GB of traffic ago (no flush, no untimed gaps); CPU arm: 100 MB working set >> host LLC",
            "multi_gpu": (f"independent {args.size}-byte shard per rank, packed streams gathered on rank 0 inside the step" if world > 1 else "single GPU")}


def sha16(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def zipf_dev(torch, n, dev, seed):
    """Zipf(1.1) bytes generated on the device (numpy needs ~85 s per GB): rank ~ Zipf over 256 symbols through a fixed random
    permutation, like datagen.zipf, from torch's generator.  Deterministic for a given torch build."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    k = torch.arange(1, 257, dtype=torch.float64, device=dev)
    cdf = torch.cumsum(k ** -1.1, 0); cdf = (cdf / cdf[-1]).to(torch.float32)
    perm = torch.randperm(256, generator=g, device=dev).to(torch.uint8)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    step = 1 << 26
    for o in range(0, n, step):
        m = min(step, n - o)
        out[o:o + m] = perm[torch.searchsorted(cdf, torch.rand(m, generator=g, device=dev)).clamp_(max=255)]
    return out


def make_data(size, rank=0, src="zipf"):
    dg = datagen()
    if src == "bwt":
        return dg.bwt_shaped(size, seed=dg.BWT_SEED + rank)
    if src == "o1":
        return dg.markov1(size, seed=dg.O1_SEED + rank)
    if src == "uniform":
        return dg.uniform(size, seed=1 + rank)
    return dg.zipf(size, seed=dg.ZIPF_SEED + rank)


def cpu_reference_run(codec, data, cdf, threads, reps, sample_bytes, chunk, cpc=0):
    """Time the reference encoder + decoder (oracle/_ref, else the port) called once per `chunk`-byte chunk on `threads` pthreads."""
    from oracle import cpu
    enc, dec = REF_FN[codec]
    sample = data[:sample_bytes]
    r = cpu.cpu_bench(codec != 10, enc, dec, sample, chunk, cdf if codec in (0, 4, 5, 10) else None, 256 if codec in (4, 5, 10) else 0,
                      threads=threads, reps=reps)
    r["sample_bytes"] = int(sample.size)
    return r


def oracle_stream(codec, data, chunk, cdf, cpc=0):
    """The packed stream the CHECKER produces for this batch (compiled reference when present, else the port)."""
    from oracle import cpu
    lib = (cpu.ref() if codec != 10 else None) or cpu.port()
    static = codec in (0, 4, 5, 10)
    return cpu.batch_enc(lib, REF_FN[codec][0], data, chunk, cdf if static else None, 256 if codec in (4, 5, 10) else (16 if codec == 0 else 0), cpc)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host threads, on the SAME workload as our
    arm (same bytes, same table, same chunking => the same packed stream: size and hash are printed by both arms)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    from oracle import cpu
    codec = CODECS[args.codec]
    chunk = default_chunk(args)
    static = codec in (0, 4, 5, 10)
    threads = os.cpu_count() or 1
    data = make_data(args.size, 0, args.src)                     # rank 0's shard: CPU throughput does not depend on the shard count
    if codec in (0, 1, 8, 9) and not (codec == 0 and args.bytes_alphabet):
        data = data & 15
    lib = cpu.ref() or cpu.port()
    cdf = None
    if static:
        blk = args.cdf_block if args.cdf_block else args.size
        cdf = np.concatenate([lib.cdfini(data[o:o + blk]) for o in range(0, args.size, blk)])
    if args.cdf_block:
        raise SystemExit("--impl reference: per-block tables are not wired into the timing driver")
    for _ in range(max(min(args.warmup, 3), 1)):
        r = cpu_reference_run(codec, data, cdf, threads, 1, args.size, chunk)
    es = ds = 0.0
    for _ in range(args.steps):                                  # one step = the whole workload: ~0.3 s (rcs2, 16 threads)
        r = cpu_reference_run(codec, data, cdf, threads, 1, args.size, chunk)
        assert r["ok"], "reference round trip failed"
        es += r["enc_s"]; ds += r["dec_s"]
    es /= args.steps; ds /= args.steps
    stream, off = oracle_stream(codec, data, chunk, cdf)
    assert stream.size == r["clen"]
    val = args.size / (es + ds) / 1e9
    sample = (f"the whole workload ({args.size} B), {REF_FN[codec][0]}+{REF_FN[codec][1]} called once per {chunk}-byte chunk, "
              f"{threads} pthreads, mean of {args.steps} passes")
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round((es + ds) * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, chunk, world),
            "enc_gbs": round(args.size / es / 1e9, 4), "dec_gbs": round(args.size / ds / 1e9, 4),
            "ratio": round(r["clen"] / args.size, 5), "compressed_bytes": int(r["clen"]), "stream_sha16": sha16(stream),
            "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": threads, "kind": r["kind"], "sample": sample},
            "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def markov1_dev(torch, n, dev, seed=20261019, lanes=1 << 18):
    """BASELINE config 4 source generated ON THE DEVICE (1 GB from numpy takes ~100 s): the same order-1 process as
    datagen.markov1 -- rank ~ Zipf(1.1), byte = perm[prev][rank], 256 fixed permutations -- run as `lanes` independent chains
    laid end to end.  Deterministic for a given torch build; the oracle checks the very bytes that are coded."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    k = torch.arange(1, 257, dtype=torch.float64, device=dev)
    cdf = torch.cumsum(k ** -1.1, 0); cdf = (cdf / cdf[-1]).to(torch.float32)
    perms = torch.stack([torch.randperm(256, generator=g, device=dev) for _ in range(256)]).to(torch.uint8)
    per = -(-n // lanes)
    out = torch.empty(lanes * per, dtype=torch.uint8, device=dev).view(lanes, per)
    prev = torch.zeros(lanes, dtype=torch.long, device=dev)
    step = 4096
    for j0 in range(0, per, step):
        w = min(step, per - j0)
        ranks = torch.searchsorted(cdf, torch.rand(lanes, w, generator=g, device=dev)).clamp_(max=255)
        for j in range(w):
            v = perms[prev, ranks[:, j]]
            out[:, j0 + j] = v
            prev = v.long()
    return out.view(-1)[:n].contiguous()


def time_batch(trc, torch, batch, d_in, flush, steps, warmup=2):
    """-> (encode ms, decode ms) per pass: CUDA events on the launching stream, L2 flushed before each timed pass."""
    for _ in range(warmup):
        batch.encode(d_in); batch.decode()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    for k in range(steps):
        flush.zero_(); ev[k][0].record(); batch.encode(d_in); ev[k][1].record()
        flush.zero_(); ev[k][2].record(); batch.decode(); ev[k][3].record()
    torch.cuda.synchronize()
    return sum(e[0].elapsed_time(e[1]) for e in ev) / steps, sum(e[2].elapsed_time(e[3]) for e in ev) / steps


def run_extras(trc, torch, dev, d_zipf, cdf_dev, flush, main_chunk):
    """Extra keys of the JSON line (not the headline): the chunk-size curve of the headline codec and the adaptive / order-1
    configurations of BASELINE.json (configs[2], configs[3]) at their full sizes, each with a device round trip and the packed
    stream compared with the oracle's."""
    out = {}
    size = d_zipf.numel()
    sweep = []
    for chunk in sorted({main_chunk, 4096, 65536, 1 << 20, 4 << 20}):           # {4 KiB, 64 KiB, 1 MiB, 4 MiB}
        b = trc.DeviceBatch(CODECS["rcs2"], size, chunk, cdfnum=256, device=dev)
        b.cdf = cdf_dev
        b.prebuild_tables()
        b.encode(d_zipf); back = b.decode(); torch.cuda.synchronize()
        assert torch.equal(back, d_zipf)
        e, d = time_batch(trc, torch, b, d_zipf, flush, 1 if chunk >= (4 << 20) else 3 if chunk >= 65536 else 10)
        sweep.append({"chunk_bytes": chunk, "n_chunks": b.n, "enc_gbs": round(size / e / 1e6, 2), "dec_gbs": round(size / d / 1e6, 2),
                      "value": round(size / (e + d) / 1e6, 2), "ratio": round(b.compressed_len() / size, 5)})
        del b
    out["chunk_sweep"] = {"codec": "rcs2", "workload": f"{size} B Zipf(1.1), one table", "unit": "GB/s", "points": sweep}

    def adaptive(name, d_in, chunk, steps, src):
        n = d_in.numel()
        codec = CODECS[name]
        b = trc.DeviceBatch(codec, n, chunk, device=dev)
        b.encode(d_in); back = b.decode(); torch.cuda.synchronize()
        assert torch.equal(back, d_in), f"{name}: device round trip failed"
        clen = b.compressed_len()
        want, woff = oracle_stream(codec, d_in.cpu().numpy(), chunk, None)
        ok = bool(np.array_equal(b.off.cpu().numpy().view(np.uint64), woff) and np.array_equal(b.out[:clen].cpu().numpy(), want))
        assert ok, f"{name}: packed stream differs from the oracle's"
        e, d = time_batch(trc, torch, b, d_in, flush, steps, warmup=1)
        return {"workload": f"{n} B {src}, {REF_FN[codec][0]}/{REF_FN[codec][1]} once per {chunk}-byte chunk", "codec": name, "chunk_bytes": chunk,
                "value": round(n / (e + d) / 1e6, 3), "unit": "GB/s", "enc_gbs": round(n / e / 1e6, 3), "dec_gbs": round(n / d / 1e6, 3),
                "ratio": round(clen / n, 5), "steps": steps, "stream_equals_oracle": ok,
                "roofline_frac": round((n + clen) / (max(e, d) * 1e-3) / 1e9 / peaks()[0], 5)}

    bwt = torch.from_numpy(make_data(size, 0, "bwt")).to(dev)
    # 64 KiB = SURVEY's default batch chunk; the smaller chunks show what more concurrent calls buy (and cost in ratio)
    out["config3_adaptive_bwt_100mb"] = [adaptive(c, bwt, ck, 3, "BWT-shaped bytes") for ck in (65536, 16384, 4096) for c in ("rc", "ans")]
    del bwt
    o1 = markov1_dev(torch, 1_000_000_000, dev)
    # 4 MiB = the reference's own block size (one call per SM: 239 calls in two waves); 1 MiB chunks (954 calls) take the
    # half-warp-per-call decoder with the low-nibble tables in global memory
    out["config4_order1_1gb"] = [adaptive("ans1", o1, 4 << 20, 2, "order-1 Markov bytes (generated on the device, bench.py markov1_dev)"),
                                 adaptive("ans1", o1, 1 << 20, 2, "order-1 Markov bytes (generated on the device, bench.py markov1_dev)")]
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):     # these two levels print NCCL's version banner (claim_stdout() keeps it off stdout anyway)
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=dev)
    trc = importlib.import_module("turbo-range-coder_b200")
    shard = importlib.import_module("turbo-range-coder_b200.shard")
    trc.lib.trc_set_device(local)
    codec = CODECS[args.codec]
    size, chunk = args.size, default_chunk(args)
    static = codec in (0, 4, 5, 10)

    if args.src == "zipf-dev":
        d_in = zipf_dev(torch, size, dev, 20261017 + rank)
        data = d_in.cpu().numpy()
    else:
        data = make_data(size, rank, args.src)
    if codec in (0, 1, 8, 9) and not (codec == 0 and args.bytes_alphabet):   # 16-symbol codecs get the low nibbles
        data = data & 15
    d_in = torch.from_numpy(data).to(dev)
    batch = trc.DeviceBatch(codec, size, chunk, cdfnum=(256 if static else 0), device=dev)
    if static:                                    # cdfini on the device, outside the timed region (turborc.c:429-433)
        blk = args.cdf_block if args.cdf_block else size      # one table per cdf-block bytes (BASELINE config 5: 64 MB blocks)
        assert not args.cdf_block or blk % chunk == 0, "--cdf-block must be a multiple of --chunk"
        cdf_dev, status = trc.cdfini_dev(d_in, size, blk)
        assert int(status.abs().sum().item()) == 0
        batch.cdf = cdf_dev
        batch.cpc = blk // chunk if args.cdf_block else 0
        if not args.no_tables and chunk % 16 == 0 and (batch.cpc == 0 or batch.cpc % 128 == 0):
            batch.prebuild_tables()               # coding tables once per cdf, outside the timed region like the cdf itself (turborc.c:432)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # (extras only)

    # ---- correctness gate (untimed): device round trip, and the WHOLE packed stream + offsets byte-compared with what the
    # checker (compiled reference, else the port) produces for the same chunks -- at the chunk size that is timed below
    batch.encode(d_in)
    back = batch.decode()
    torch.cuda.synchronize()
    assert torch.equal(back, d_in), "device round trip failed"
    clen = batch.compressed_len()
    n_chunks = batch.n
    stream_sha = None
    if not args.no_gate:
        cdf_gate = batch.cdf.cpu().numpy().view(np.uint16) if static else None
        want, woff = oracle_stream(codec, data, chunk, cdf_gate, batch.cpc)
        got = batch.out[:clen].cpu().numpy()
        goff = batch.off.cpu().numpy().view(np.uint64)
        assert np.array_equal(goff, woff), "packed offsets differ from the oracle's"
        assert np.array_equal(got, want), "packed stream differs from the oracle's"
        stream_sha = sha16(got)
        del want, got

    # ---- three rotating buffer sets instead of an L2 flush: step k encodes set k % 3 and decodes the stream of set (k+1) % 3,
    # which was written two steps (~0.8 GB of traffic) earlier, so every timed kernel reads data that left the 126 MB L2 long
    # ago, and the timed region has NO untimed gaps (a gather that overlaps compute cannot hide in one)
    NSETS = 3
    sets, d_ins = [batch], [d_in]
    for _ in range(NSETS - 1):
        bb = trc.DeviceBatch(codec, size, chunk, cdfnum=(256 if static else 0), chunks_per_cdf=batch.cpc, device=dev)
        bb.cdf = batch.cdf
        bb.borrow_tables(batch)
        sets.append(bb); d_ins.append(d_in.clone())
    for bb, di in zip(sets, d_ins):
        bb.encode(di)
    torch.cuda.synchronize()

    # multi-GPU: the packed streams are gathered on rank 0 inside every step.  Default: device-driven push over NVLink peer
    # memory (shard.PeerGather: completion flags + acknowledgements on the device, two slot sets) on a side stream, so a push
    # overlaps the decode of its step and the encode of the next one; rank 0 blocks its stream at the end of step k until every
    # stream of step k-1 has landed and acknowledges it.  TRC_GATHER=nccl selects the NCCL send/recv form
    # (shard.gather_compressed), which needs the lengths on the host and therefore a sync per step.
    peer, gather_buf, gather_kind = None, None, "none"
    if world > 1:
        ok = torch.zeros(1, dtype=torch.int32, device=dev)
        if os.environ.get("TRC_GATHER", "peer") == "peer":
            try:
                # slot sets on rank 0: enough of them that the sources are never held back by acknowledgements while the NVLink
                # ingress of rank 0 (the bottleneck of an all-to-one gather) is busy; at most ~8 GiB of rank 0's memory
                depth = int(os.environ.get("TRC_PEER_DEPTH", "0")) or max(2, min(8, (8 << 30) // (world * batch.out.numel())))
                # copy engines per push: one engine moves ~180 GB/s, so up to 4 GPUs (where rank 0's ingress is not the bound) a push is
                # cut in three -- 4 GPUs: 0.415 -> 0.368 ms per step; beyond that the ingress bound makes it pointless
                split = int(os.environ.get("TRC_PEER_SPLIT", "0")) or (3 if world <= 4 else 1)
                peer = shard.PeerGather(batch.out.numel(), dst=0, depth=depth, split=split)
                ok += 1
            except Exception as e:                      # no peer access
                print(f"[rank {rank}] PeerGather unavailable ({e})", file=sys.stderr)
        dist.all_reduce(ok)                             # every rank must take the same path (a mix would deadlock)
        if int(ok.item()) != world:
            if peer is not None:
                peer.close()
            peer = None
            gather_kind = "NCCL all-gather of lengths + grouped send/recv"
            if rank == 0:
                gather_buf = torch.empty(int(size * 1.05) * world, dtype=torch.uint8, device=dev)
        else:
            gather_kind = "peer-memory push (CUDA IPC over NVLink): copy engine for the predicted length (the previous size), device-side remainder + length + completion flag, acks for back-pressure; overlapped with decode and the next encode"
    side = torch.cuda.Stream(device=dev) if peer is not None else None
    ev_enc = torch.cuda.Event()
    ev_push = [torch.cuda.Event() for _ in range(NSETS)]
    for e in ev_push:
        e.record()
    total_ptr = [bb.off.data_ptr() + 8 * bb.n for bb in sets]           # device address of out_off[n] = packed length
    state = {"k": 0, "seq": 0}

    nogather = bool(os.environ.get("TRC_BENCH_NOGATHER"))      # experiments only: time the N ranks without any exchange

    def step(ev=None):
        main = torch.cuda.current_stream()
        k = state["k"]; state["k"] += 1
        ia, ib = k % NSETS, (k + 1) % NSETS
        if ev: ev[0].record()
        if peer is not None:
            main.wait_event(ev_push[ia])                 # the push that read this set's stream (three steps ago) is done
        sets[ia].encode(d_ins[ia])
        if peer is not None and not nogather:
            ev_enc.record(main)
            side.wait_event(ev_enc)
            with torch.cuda.stream(side):
                state["seq"] = peer.push(sets[ia].out, total_ptr[ia], side, hint_bytes=0 if os.environ.get("TRC_PUSH_NOHINT") else clen)
                ev_push[ia].record(side)
        elif world > 1:
            shard.gather_compressed(sets[ia].out[:clen], dst=0, out=gather_buf)
        if ev: ev[1].record()
        sets[ib].decode()
        if peer is not None and rank == 0 and state["seq"] >= 2 and not os.environ.get("TRC_BENCH_NOWAIT"):
            peer.wait_all(state["seq"] - 1, main)        # rank 0 holds every stream of the previous step ...
            peer.ack(state["seq"] - 1, main)             # ... and releases that slot set

    def finish():
        main = torch.cuda.current_stream()
        if peer is not None:
            for e in ev_push:
                main.wait_event(e)                       # own pushes have landed
            if rank == 0 and state["seq"] >= 1:
                peer.wait_all(state["seq"], main)        # rank 0 holds the streams of the last step too
                peer.ack(state["seq"], main)

    for _ in range(max(args.warmup, 3)):
        step()
    finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    ev_end = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local); sampler.start()
    launches0 = trc.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(evs[k])
    t_enq = time.perf_counter() - t0                     # host time to enqueue the steps (the GPU must not be waiting for it)
    finish()
    ev_end.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = trc.launch_count() - launches0
    clocks = sampler.result()
    total_ms = evs[0][0].elapsed_time(ev_end)
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    step_ms = total_ms / args.steps                       # the whole timed region, gather waits included
    dec_ms = step_ms - enc_ms
    if world > 1:                                        # verify what landed on rank 0 (last step)
        dist.barrier()
        il = (state["k"] - 1) % NSETS
        sums = torch.zeros(world, dtype=torch.int64, device=dev)
        sums[rank] = sets[il].out[:clen].to(torch.int64).sum()
        dist.all_reduce(sums)
        lens = torch.zeros(world, dtype=torch.int64, device=dev); lens[rank] = clen
        dist.all_reduce(lens)
        if rank == 0 and peer is not None:
            got = peer.read_lens(dev, state["seq"])
            assert torch.equal(got, lens), (got, lens)
            assert not peer.overflowed(dev)
            for r in range(world):
                assert int(peer.read_slot(r, int(lens[r]), dev, state["seq"]).to(torch.int64).sum()) == int(sums[r]), f"gathered stream of rank {r} differs"

    # per-kernel durations (CUDA events between the kernels, same stream), averaged over a few extra passes on the rotating sets
    trc.profile_enable(True)
    names_enc = ["encode", "resolve_scan", "pack"]
    acc = {}
    reps = min(args.steps, 12)
    for q in range(reps):
        sets[q % NSETS].encode(d_ins[q % NSETS])
        for nm, ms in zip(names_enc, trc.profile_read()):
            acc[nm] = acc.get(nm, 0.0) + ms / reps
        sets[(q + 1) % NSETS].decode()
        for nm, ms in zip(["decode"], trc.profile_read()):
            acc[nm] = acc.get(nm, 0.0) + ms / reps
    trc.profile_enable(False)
    kern_ms = {k: round(v, 4) for k, v in acc.items() if v > 0}

    t = torch.tensor([enc_ms, dec_ms, step_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)                         # per-rank (encode, decode + gather waits, step) ms: who sets the max
        per_rank = [[round(float(x), 4) for x in a] for a in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc_ms, dec_ms, step_ms = float(t[0]), float(t[1]), float(t[2])
    total_bytes = size * world
    value = total_bytes / (step_ms * 1e-3) / 1e9

    # ---- e2e through the C-ABI host entry points, pinned host buffers ----
    h_in = torch.from_numpy(data).pin_memory().numpy()
    h_out = torch.empty(int(trc.lib.trc_enc_bound(size, chunk)), dtype=torch.uint8).pin_memory().numpy()
    h_off = torch.empty(n_chunks + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    h_back = torch.empty(size, dtype=torch.uint8).pin_memory().numpy()
    cdf_h = batch.cdf.cpu().numpy().view(np.uint16) if static else None
    e2e_steps = 1 if args.quick else max(3, min(args.steps, 10))
    te = td = 0.0
    for k in range(2 + e2e_steps):
        a = time.perf_counter()
        s_out, s_off = trc.enc_batch_host(codec, h_in, chunk, cdf=cdf_h, cdfnum=256 if static else 0, chunks_per_cdf=batch.cpc, out=h_out, off=h_off)
        b = time.perf_counter()
        trc.dec_batch_host(codec, s_out, s_off, size, chunk, cdf=cdf_h, cdfnum=256 if static else 0, chunks_per_cdf=batch.cpc, out=h_back)
        c = time.perf_counter()
        if k >= 2:
            te += b - a; td += c - b
    assert np.array_equal(h_back, data), "host round trip failed"
    te /= e2e_steps; td /= e2e_steps
    t = torch.tensor([te, td], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te, td = float(t[0]), float(t[1])
    e2e = {"value": round(total_bytes / (te + td) / 1e9, 4), "unit": "GB/s",
           "h2d_bytes_per_step": int(size + clen + 8 * (n_chunks + 1)) * world, "d2h_bytes_per_step": int(clen + 8 * (n_chunks + 1) + size) * world,
           "enc_gbs": round(total_bytes / te / 1e9, 4), "dec_gbs": round(total_bytes / td / 1e9, 4),
           "api": "trc_enc_batch_host + trc_dec_batch_host, pinned host buffers" + (f", one process per GPU ({world} at once)" if world > 1 else "")}

    # ---- the same N-GPU workload through ONE process: rank 0 drives all N devices with trc_*_batch_host_multi (one host thread per
    # device) while the other ranks wait; reported next to the per-process number, the better one is the headline e2e
    e2e_multi = None
    if world > 1 and not args.no_multi_e2e:
        dist.barrier()
        if rank == 0:
            try:
                big = np.tile(data, world)
                hb_in = torch.from_numpy(big).pin_memory().numpy()
                hb_out = torch.empty(int(trc.lib.trc_enc_bound(big.size, chunk)), dtype=torch.uint8).pin_memory().numpy()
                nbig = trc.num_chunks(big.size, chunk)
                hb_off = torch.empty(nbig + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
                hb_back = torch.empty(big.size, dtype=torch.uint8).pin_memory().numpy()
                devs = list(range(world))
                tm_e = tm_d = 0.0
                for k in range(2 + 3):
                    a = time.perf_counter()
                    so, sf = trc.enc_batch_host_multi(codec, devs, hb_in, chunk, cdf=cdf_h, cdfnum=256 if static else 0, chunks_per_cdf=0, out=hb_out, off=hb_off)
                    b = time.perf_counter()
                    trc.dec_batch_host_multi(codec, devs, so, sf, big.size, chunk, cdf=cdf_h, cdfnum=256 if static else 0, chunks_per_cdf=0, out=hb_back)
                    c = time.perf_counter()
                    if k >= 2:
                        tm_e += (b - a) / 3; tm_d += (c - b) / 3
                assert np.array_equal(hb_back, big), "multi-device host round trip failed"
                e2e_multi = {"value": round(big.size / (tm_e + tm_d) / 1e9, 4), "unit": "GB/s", "enc_gbs": round(big.size / tm_e / 1e9, 4),
                             "dec_gbs": round(big.size / tm_d / 1e9, 4), "h2d_bytes_per_step": int(big.size + so.size + 8 * (nbig + 1)),
                             "d2h_bytes_per_step": int(so.size + 8 * (nbig + 1) + big.size),
                             "api": f"trc_enc_batch_host_multi + trc_dec_batch_host_multi, {world} devices from one process (one host thread per device), pinned host buffers"}
                del big, hb_in, hb_out, hb_back
            except Exception as ex:                      # noqa: BLE001
                e2e_multi = {"error": str(ex)[:300]}
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0
    if e2e_multi and "value" in e2e_multi:
        e2e["per_process"] = {k: e2e[k] for k in ("value", "enc_gbs", "dec_gbs", "api")}
        e2e["single_process"] = e2e_multi
        if e2e_multi["value"] > e2e["value"]:
            e2e.update({k: e2e_multi[k] for k in ("value", "enc_gbs", "dec_gbs", "h2d_bytes_per_step", "d2h_bytes_per_step", "api")})
    elif e2e_multi:
        e2e["single_process"] = e2e_multi

    # ---- roofline of the dominant kernel ----
    peak, peak_src = peaks()
    dom = max(kern_ms, key=kern_ms.get)
    alg = {"encode": size + clen, "decode": size + clen, "pack": 2 * clen, "resolve_scan": 40 * n_chunks}[dom]
    achieved = alg / (kern_ms[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and size == 100_000_000:                       # captures are of exact workloads (100 MB)
        tj = json.load(open(tp))
        traffic = tj.get(f"{args.codec}/{dom}@{chunk}/{args.src}")
        if traffic is None and chunk == 4096 and args.src == "zipf":
            traffic = tj.get(f"{args.codec}/{dom}")
    roofline = {"bound": "hbm", "kernel": f"{args.codec}/{dom}", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg), "kernel_ms": kern_ms}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        cdfh = batch.cdf.cpu().numpy().view(np.uint16)[:257].copy() if static else None      # the same table the GPU used
        r = cpu_reference_run(codec, data, cdfh, threads, 2, size if not args.cdf_block else min(size, args.cdf_block), chunk)
        one = cpu_reference_run(codec, data, cdfh, 1, 1, min(size, 16 << 20), chunk)
        sb = r["sample_bytes"]
        cpu_baseline = {"value": round(sb / (r["enc_s"] + r["dec_s"]) / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": r["kind"],
                        "sample": f"first {sb} B of the workload, {REF_FN[codec][0]}+{REF_FN[codec][1]} once per {chunk}-byte chunk, {threads} pthreads, best of 2",
                        "enc_gbs": round(sb / r["enc_s"] / 1e9, 4), "dec_gbs": round(sb / r["dec_s"] / 1e9, 4),
                        "single_thread_gbs": round(one["sample_bytes"] / (one["enc_s"] + one["dec_s"]) / 1e9, 4)}

    extras = None
    if world == 1 and not args.no_extras and args.codec == "rcs2" and args.src == "zipf" and size == 100_000_000 and not args.cdf_block:
        extras = run_extras(trc, torch, dev, d_in, batch.cdf, flush, chunk)

    cfg = workload_config(args, chunk, world)
    if world > 1:
        cfg["multi_gpu"] += ": " + gather_kind
    line = {"metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": cfg,
            "enc_gbs": round(total_bytes / (enc_ms * 1e-3) / 1e9, 3), "dec_gbs": round(total_bytes / (dec_ms * 1e-3) / 1e9, 3),
            "ratio": round(clen / size, 5), "compressed_bytes": int(clen), "stream_sha16": stream_sha,
            "gate": "whole packed stream + offsets byte-compared with the oracle's at this chunk size" if stream_sha else "device round trip only (--no-gate)",
            "per_rank_enc_dec_step_ms": per_rank, "wall_ms_per_step": round(wall / args.steps * 1e3, 4), "host_enqueue_ms_per_step": round(t_enq / args.steps * 1e3, 4),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline}
    if extras:
        line.update(extras)
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON line: keep a private handle on the real stdout and point file descriptor 1 at stderr,
    so that nothing a library prints (NCCL writes its version banner to stdout with NCCL_DEBUG=VERSION or WARN) can land there."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codec", default="rcs2", choices=sorted(CODECS))
    ap.add_argument("--chunk", type=int, default=0, help="bytes per reference call; 0 = size the batch to one balanced wave (384 calls per SM: 1760 B for 100 MB on 148 SMs; DESIGN.md section 5)")
    ap.add_argument("--size", type=int, default=100_000_000)
    ap.add_argument("--bytes-alphabet", action="store_true", help="ans4s: code full bytes with a 256-entry table")
    ap.add_argument("--src", default="zipf", choices=["zipf", "zipf-dev", "bwt", "o1", "uniform"], help="synthetic source (SURVEY.md section 8d)")
    ap.add_argument("--cdf-block", type=int, default=0, help="static codecs: one cdfini table per this many bytes (0 = whole buffer)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tables", action="store_true", help="rebuild the coding tables inside every call instead of using a prebuilt handle")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only: skip the oracle gate, the e2e legs, the CPU baseline and the extras (sweeps)")
    ap.add_argument("--no-multi-e2e", action="store_true", help="N > 1: skip the single-process multi-device e2e leg")
    ap.add_argument("--no-gate", action="store_true", help="skip the oracle comparison of the packed stream (device round trip only)")
    ap.add_argument("--no-extras", action="store_true", help="skip the chunk sweep and the BASELINE config 3/4 lines (extra keys of the JSON line)")
    args = ap.parse_args()
    claim_stdout()
    if args.quick:
        args.no_gate = args.no_cpu = args.no_extras = args.no_multi_e2e = True
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
