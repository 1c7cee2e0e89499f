"""Many-stream probe (development aid): BASELINE config C5 shape at reduced count — n_streams independent 16-QAM
streams, each metadata + K data frames, pushed as [n_streams][block-multiple] sections from device memory.  Checks the
counters against the known content and prints the throughput."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import audio_modem_b200 as amb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=4096)
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--push-blocks", type=int, default=16)
ap.add_argument("--mode", default="16-QAM")
args = ap.parse_args()
MODES = {"QPSK": ("standard", "QPSK", 1, 2048), "16-QAM": ("standard", "QAM16", 1, 4096), "BPSK-ACOUSTIC": ("acoustic", "BPSK", 1, 512),
         "BPSK-REPEAT": ("acoustic", "BPSK", 3, 512), "BPSK-NARROW": ("narrowband", "BPSK", 3, 512)}
cfg, mod, rep, chunk = MODES[args.mode]
dev = torch.device("cuda", 0)
m = amb.Modem(cfg, mod, rep, device=0)
ns, K = args.streams, args.frames
meta_len = m.chunk_frame_samples(1 + 4 + 4 + 2 + 1 + len("probe.bin") + 4, True)
flen = m.chunk_frame_samples(chunk + 11, False)
total = meta_len + K * flen
push = args.push_blocks * 4096
n = (total + 8192 + push - 1) // push * push
x = torch.zeros((ns, n), dtype=torch.float32, device=dev)
clean = torch.zeros((1, flen), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
m.build_data_chunk_frames(n=1, data_len=chunk, seq0=0, synth_seed=2, out=clean)
m.synchronize()
core = m.chunk_frame_samples(chunk + 11, False) - 2205 - 882
sigma = float((clean[0, 2205:2205 + core].double().pow(2).mean() / 10 ** 2.5).sqrt())
t0 = time.perf_counter()
for s in range(ns):
    m.build_metadata_frame(K, K * chunk, chunk, "probe.bin", noise_sigma=sigma, noise_seed=5000 + s, noise_g0=0, out=x[s, :meta_len])
    m.build_data_chunk_frames(n=K, data_len=chunk, seq0=0, synth_seed=9000 + s, noise_sigma=sigma, noise_seed=5000 + s, noise_g0=meta_len,
                              out=x[s, meta_len:meta_len + K * flen].view(K, flen))
m.synchronize()
print(f"built {ns} streams x {n} samples ({ns * n * 4 / 1e9:.2f} GB) in {time.perf_counter() - t0:.1f} s")
m.stream_open(ns, 4096, push, max_results=ns * 8 + 64)
cap = ns * 8 + 64
d_res = torch.empty(cap * 88, dtype=torch.uint8, device=dev)
d_bytes = torch.empty((cap, m.bytes_stride), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
frames = 0
per = []
for ip, p0 in enumerate(range(0, n, push)):
    if ip == 4:
        m.profile(True)
    t1 = time.perf_counter()
    l0 = m.kernel_launches
    stats_lib = os.environ.get("AMB_LIB", "").endswith("stats.so")
    if stats_lib:
        import ctypes as C
        import numpy as np
        tr = np.zeros((16384, 24), np.uint64)
        m._L.amb_debug_read(m._h, 2, C.c_void_p(tr.ctypes.data), tr.nbytes)          # (reading resets the trace)
    frames += m.stream_push(x[:, p0:p0 + push], device_results=(d_res, d_bytes))
    m.synchronize()
    per.append((round((time.perf_counter() - t1) * 1e3, 2), m.kernel_launches - l0))
    if stats_lib and ip >= 2:
        rows = int(m._L.amb_debug_read(m._h, 2, C.c_void_p(tr.ctypes.data), tr.nbytes))
        t = tr[:max(0, min(rows, 16384))].astype(np.int64)
        t = t[t[:, 1] != 9]
        print(f"push {ip}: {rows} trace rows")
        for mode in sorted(set(t[:, 1].tolist())):
            for has in (0, 1):
                sel = t[(t[:, 1] == mode) & (t[:, 3] == has)]
                if len(sel):
                    print("   mode %d (%s) has %d: %5d walks, cycles mean %7.0f max %7d | eval %6.0f refine %6.0f skip %6.0f scan %6.0f | rows %.2f direct %.2f scancalls %.2f" %
                          (mode, {0: "A run", 1: "B link", 2: "A first"}.get(mode, "?"), has, len(sel), sel[:, 6].mean(), sel[:, 6].max(), sel[:, 17].mean(), sel[:, 18].mean(),
                           sel[:, 19].mean(), sel[:, 20].mean(), sel[:, 8].mean(), sel[:, 9].mean(), sel[:, 14].mean()))
dt = time.perf_counter() - t0
prof = m.profile_read()
m.profile(False)
npush = max(1, len(per) - 4)
print({k: round(v[0] / npush, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:14]})
print("per push (ms, launches):", per)
print("fallback symbols (all pushes):", m.debug_counter(1), "link iterations", m.debug_counter(2), "passes", m.debug_counter(3))
print("launch counts:", {k: v[1] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:14]})
if len(per) > 4:
    tail = sorted(p[0] for p in per[3:-1])
    print(f"steady pushes (median): {ns * push / (tail[len(tail) // 2] / 1e3) / 1e6:.0f} Msamples/s")
if os.environ.get("AMB_LIB", "").endswith("stats.so"):
    names = ["rows", "direct", "refine_store", "refine_direct", "refine_memo", "skip_rounds", "scan_calls", "refine_fail", "exact_offsets",
             "cyc_eval", "cyc_refine", "cyc_skip", "cyc_scan", "exact_replays"]
    print("walk stats (all pushes, phase A + phase B):", dict(zip(names, [m.debug_counter(16 + i) for i in range(len(names))])),
          "taken", m.debug_counter(6), "walked", m.debug_counter(7))
c = m.stream_counters()
print(f"{ns} streams: {frames} frames in {dt * 1e3:.1f} ms -> {ns * n / dt / 1e6:.0f} Msamples/s, {n // push} pushes, counters {c}")
assert c["metaReceived"] == ns, c
if rep == 1 and cfg == "standard":
    assert frames == ns * (K + 1) and c["framesDecoded"] == ns * (K + 1), (frames, c)
m.stream_close()
m.close()
