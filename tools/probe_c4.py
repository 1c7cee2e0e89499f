"""Config C4 probe (development aid): a 1-hour 44.1 kHz narrowband capture (158.76 M samples, 635 MB) with one metadata
frame and 12 data chunks far apart over a -30 dBFS noise floor, pushed from device memory in one call (scan dominated)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import audio_modem_b200 as amb  # noqa: E402

cfg, mod, rep, chunk = "narrowband", "BPSK", 3, 512
dev = torch.device("cuda", 0)
m = amb.Modem(cfg, mod, rep, device=0)
SECONDS = int(os.environ.get("C4_SECONDS", "3600"))
N = SECONDS * 44100
N = (N + 4095) // 4096 * 4096
g = torch.Generator(device=dev)
g.manual_seed(4)
x = torch.randn((1, N), dtype=torch.float32, device=dev, generator=g) * (10 ** (-30 / 20))
meta_len = m.chunk_frame_samples(1 + 4 + 4 + 2 + 1 + len("cap.bin") + 4, True)
flen = m.chunk_frame_samples(chunk + 11, False)
pos = 5 * 44100
buf = torch.zeros(meta_len, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
m.build_metadata_frame(12, 12 * chunk, chunk, "cap.bin", out=buf)
m.synchronize()
REPS = int(os.environ.get("C4_REPS", "3"))
x[0, pos:pos + meta_len] += buf
pos += meta_len
fb = torch.zeros((1, flen), dtype=torch.float32, device=dev)
starts = []
NF = int(os.environ.get("C4_FRAMES", "12"))
for i in range(NF):
    pos += (10 + (17 * i if SECONDS >= 3600 else 0)) * 44100
    torch.cuda.synchronize()
    m.build_data_chunk_frames(n=1, data_len=chunk, seq0=i, synth_seed=40 + i, out=fb)
    m.synchronize()
    x[0, pos:pos + flen] += fb[0]
    starts.append(pos)
    pos += flen
assert pos < N
m.synchronize()
torch.cuda.synchronize()
cap = 64
d_res = torch.empty(cap * 88, dtype=torch.uint8, device=dev)
for it in range(REPS):
    m.stream_open(1, 4096, N, max_results=cap)
    d_bytes = torch.empty((cap, m.bytes_stride), dtype=torch.uint8, device=dev)
    if it == REPS - 1:
        m.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    i0, p0 = m.debug_counter(2), m.debug_counter(3)
    n = m.stream_push(x, device_results=(d_res, d_bytes))
    m.synchronize()
    dt = time.perf_counter() - t0
    if os.environ.get("AMB_LIB", "").endswith("stats.so"):
        names = ["rows", "direct", "refine_store", "refine_direct", "refine_memo", "skip_rounds", "scan_calls", "refine_fail", "exact_offsets",
                 "cyc_eval", "cyc_refine", "cyc_skip", "cyc_scan"]
        print("walk stats (cumulative):", dict(zip(names, [m.debug_counter(16 + i) for i in range(13)])))
    if os.environ.get("AMB_LIB", "").endswith("stats.so"):
        import ctypes as C
        import numpy as np
        tr = np.zeros((16384, 24), np.uint64)
        rows = int(m._L.amb_debug_read(m._h, 2, C.c_void_p(tr.ctypes.data), tr.nbytes))
        tr = tr[:max(0, min(rows, 16384))].astype(np.int64)
        print("trace rows", rows)
        for r in tr[tr[:, 1] == 9]:
            print("  verify launch %d: n_links %d stop %d hasB %d run %d hasA %d | state %d/%d ac_init %d/%d scan_pos %d/%d next_call %d/%d pre_pos %d/%d frame_end %d/%d | n_runs %d run.first_pos %d next[run] %d" % tuple(r[[0] + list(range(2, 22))]))
        tr = tr[tr[:, 1] != 9]
        for it in sorted(set(tr[:, 0].tolist())):
            sel = tr[tr[:, 0] == it]
            top = sel[np.argsort(-sel[:, 6])][:4]
            print(f" launch {it} ({'B' if it & 1 else 'A'}): {len(sel)} walks, sum {sel[:, 6].sum() / 1e6:.2f} Mcyc, max {sel[:, 6].max() / 1e6:.3f} Mcyc")
            for r in top:
                print("   mode %d link %d has %d calls %d..%d cyc %d fp %d | rows %d direct %d rstore %d rdirect %d memo %d skiprounds %d scancalls %d rfail %d exact %d | cyc eval %d refine %d skip %d scan %d" %
                      (r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16], r[17], r[18], r[19], r[20]))
    print("link iterations", m.debug_counter(2) - i0, "passes", m.debug_counter(3) - p0, "slots", m.debug_counter(4), "cand", m.debug_counter(5))
    c = m.stream_counters()
    print(f"[{it}] {N} samples in {dt * 1e3:.2f} ms -> {N / dt / 1e6:.0f} Msamples/s ({4 * N / dt / 1e9:.0f} GB/s), frames {n}, {c}")
    if it == REPS - 1:
        prof = m.profile_read()
        print({k: (round(v[0], 3), v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]})
    m.stream_close()
assert c["metaReceived"] == 1 and c["receivedCount"] >= NF - 2, (n, c)
m.close()
