"""Quick performance probe (development aid, not the bench): a dense 16-QAM chunk stream built by tiling a few
oracle-generated frames, pushed through the streaming receiver from host memory and from device memory."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

import audio_modem_b200 as amb  # noqa: E402
from oracle import oracle as O  # noqa: E402

cfg, mod, rep, chunk = "standard", "QAM16", 1, 4096
meta = O.buildMetadataFrame(cfg, args.frames, args.frames * chunk, chunk, "probe.bin", mod, rep)
uniq = [O.buildDataChunkFrame(cfg, O.randomBytes(900 + i, chunk), i, mod, rep) for i in range(16)]
parts = [meta] + [uniq[i % 16] for i in range(args.frames)]
x = np.concatenate(parts)
x = O.addAwgn(x, 0.012, 3)
n = (len(x) + 8192 + 4095) // 4096 * 4096
x = np.concatenate([x, np.zeros(n - len(x), np.float32)])
print(f"stream: {args.frames} frames, {n} samples, {n * 4 / 1e6:.1f} MB")
m = amb.Modem(cfg, mod, rep)
for it in range(args.reps):
    m.stream_open(1, 4096, n)
    l0 = m.kernel_launches
    t = time.perf_counter()
    cnt = m.stream_push(x[None, :], collect=False)
    dt = time.perf_counter() - t
    c = m.stream_counters()
    print(f"push[{it}] host->device: {dt * 1e3:.2f} ms  {n / dt / 1e6:.1f} Msamples/s  frames={cnt} launches={m.kernel_launches - l0} {c}")
    m.stream_close()
m.close()
