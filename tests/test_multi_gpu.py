"""Several GPUs from one process through the C ABI (trc_enc_batch_host_multi / trc_dec_batch_host_multi), and the device-driven
gather over peer memory.  Needs >= 2 GPUs (skipped otherwise); `-m gpu`."""
import numpy as np
import pytest

from helpers import CODECS, cpu_batch

pytestmark = pytest.mark.gpu


def _ngpu(trc):
    return int(trc.lib.trc_device_count())


@pytest.mark.parametrize("codec", [5, 0, 2, 6, 3])
def test_multi_device_equals_one_device(trc, port, dg, codec):
    """Sharded over all visible devices the packed stream, the offsets and the decoded bytes equal the oracle's (== one device)."""
    n_dev = _ngpu(trc)
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    enc, dec, need_cdf, nib = CODECS[codec]
    d = dg.zipf(3_000_017, seed=11)
    d[1_000_000:2_000_000] = dg.bwt_shaped(1_000_000)
    if nib:
        d = dg.nibbles(d)
    cdf = port.cdfini(d) if need_cdf else None
    num = int(d.max()) + 1 if need_cdf else 0
    for chunk in (4096, 65536):
        want, woff = cpu_batch(port, codec, d, chunk, cdf, num)
        for devs in (list(range(n_dev)), [n_dev - 1, 0], [1]):
            got, off = trc.enc_batch_host_multi(codec, devs, d, chunk, cdf=cdf, cdfnum=num)
            assert np.array_equal(off, woff) and np.array_equal(got, want), (enc, chunk, devs)
            back = trc.dec_batch_host_multi(codec, devs, got, off, d.size, chunk, cdf=cdf, cdfnum=num)
            assert np.array_equal(back, d), (dec, chunk, devs)


def test_multi_device_more_devices_than_chunks(trc, port, dg):
    n_dev = _ngpu(trc)
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    d = dg.zipf(5000, seed=2)
    cdf = port.cdfini(d)
    want, woff = cpu_batch(port, 5, d, 4096, cdf, 256)
    got, off = trc.enc_batch_host_multi(5, list(range(n_dev)), d, 4096, cdf=cdf, cdfnum=256)
    assert np.array_equal(off, woff) and np.array_equal(got, want)
    assert np.array_equal(trc.dec_batch_host_multi(5, list(range(n_dev)), got, off, d.size, 4096, cdf=cdf, cdfnum=256), d)


def test_multi_device_bad_arguments(trc, dg):
    d = dg.zipf(10000)
    with pytest.raises(trc.TrcError):
        trc.enc_batch_host_multi(6, [0, 0], d, 4096)               # duplicate device
    with pytest.raises(trc.TrcError):
        trc.enc_batch_host_multi(6, [_ngpu(trc)], d, 4096)         # no such device


def test_peer_gather(trc):
    """PeerGather (CUDA IPC over NVLink): completion flags, slot sets, overflow bit, fetch -- one process per GPU under torchrun."""
    import os, subprocess, sys
    n_dev = _ngpu(trc)
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n_dev, 4)}", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(here, "peer_gather_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "peer gather ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
