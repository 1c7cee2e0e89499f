"""Batch-path probe (development aid): BASELINE config C3 shape at reduced count — n recordings, each one legacy frame
(buildTransmitSignal, 512-byte payload) in an acoustic BPSK mode, decoded by decode_received_signals from device memory."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import audio_modem_b200 as amb  # noqa: E402
from audio_modem_b200._lib import FrameResult, MEM_DEVICE  # noqa: E402
from oracle import oracle as O  # noqa: E402
import ctypes as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--mode", default="BPSK-ACOUSTIC")
ap.add_argument("--uniq", type=int, default=8)
args = ap.parse_args()
MODES = {"QPSK": ("standard", "QPSK", 1), "16-QAM": ("standard", "QAM16", 1), "BPSK-ACOUSTIC": ("acoustic", "BPSK", 1),
         "BPSK-REPEAT": ("acoustic", "BPSK", 3), "BPSK-NARROW": ("narrowband", "BPSK", 3)}
cfg, mod, rep = MODES[args.mode]
uniq = []
for i in range(args.uniq):
    x = O.buildTransmitSignal(cfg, O.randomBytes(1000 + i, 512), mod, "rec", rep)
    x = O.multipath(x, [0, 23, 57], [1.0, 0.5, 0.25])
    x = O.addAwgn(x, float(np.sqrt(O.power(x) / 100.0)), 77 + i)
    uniq.append(x.astype(np.float32))
n_s = len(uniq[0])
stride = (n_s + 3) // 4 * 4
host = np.zeros((args.n, stride), np.float32)
for i in range(args.n):
    host[i, :n_s] = uniq[i % args.uniq]
dev = torch.device("cuda", 0)
x = torch.from_numpy(host).to(dev)
m = amb.Modem(cfg, mod, rep, device=0)
bs = m.decoded_capacity(stride)
res = torch.zeros(args.n * C.sizeof(FrameResult), dtype=torch.uint8, device=dev)
dec = torch.zeros((args.n, bs), dtype=torch.uint8, device=dev)
lens = torch.full((args.n,), n_s, dtype=torch.int64, device=dev)
L = m._L
for it in range(3):
    if it == 2:
        m.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = L.amb_decode_received_signals(m._h, C.c_void_p(x.data_ptr()), stride, C.c_void_p(lens.data_ptr()), args.n, C.c_void_p(res.data_ptr()),
                                       C.c_void_p(dec.data_ptr()), bs, MEM_DEVICE)
    m.synchronize()
    dt = time.perf_counter() - t0
    assert rc == 0, L.amb_last_error(m._h)
    print(f"[{it}] {args.n} recordings x {n_s} samples: {dt * 1e3:.2f} ms -> {args.n * n_s / dt / 1e6:.0f} Msamples/s")
prof = m.profile_read()
print({k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])})
r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=np.uint8)
fr = (FrameResult * args.n).from_buffer_copy(r.tobytes())
ok = sum(1 for i in range(args.n) if fr[i].status == 0 and fr[i].crc_valid)
want = [O.decodeReceivedSignal(cfg, u, mod, rep) for u in uniq]
print("crc-valid recordings:", ok, "of", args.n, "| oracle on the unique ones:", sum(1 for w in want if w.get("crcValid")), "of", len(want))
for i in range(min(args.n, args.uniq)):
    w = want[i]
    assert (fr[i].status == 0) == ("error" not in w), (i, fr[i].status, w)
    if "error" not in w:
        assert fr[i].preamble_idx == w["preambleIdx"] and bool(fr[i].crc_valid) == w["crcValid"], (i, fr[i].preamble_idx, w)
m.close()
