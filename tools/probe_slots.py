"""Where do the scan kernel's slots fall relative to the frames?  (development probe, GPU)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_modem_b200 as amb

F = 256
m = amb.Modem("standard", "QAM16", 1, device=0)
FRAME = m.chunk_frame_samples(4096 + 11, False)
META = m.chunk_frame_samples(1 + 4 + 4 + 2 + 1 + 9 + 4, True)
dev = torch.device("cuda", 0)
clean = torch.zeros((1, FRAME), dtype=torch.float32, device=dev)
m.build_data_chunk_frames(n=1, data_len=4096, seq0=0, synth_seed=2, out=clean)
core = amb.estimateFrameSamples(4096 + 11, "QAM16", 1)
sigma = float((clean[0, 2205:2205 + core].double().pow(2).mean() / 10 ** 2.5).sqrt())
n = META + F * FRAME
n_pad = (n + 4095) // 4096 * 4096
x = torch.zeros((1, n_pad), dtype=torch.float32, device=dev)
m.build_metadata_frame(F, F * 4096, 4096, "bench.bin", noise_sigma=sigma, noise_seed=3, noise_g0=0, out=x[0, :META])
m.build_data_chunk_frames(n=F, data_len=4096, seq0=0, synth_seed=2, noise_sigma=sigma, noise_seed=3, noise_g0=META, out=x[0, META:n].view(F, FRAME))
m.stream_open(1, 4096, n_pad)
ev = m.stream_push(x)
info, ts = m.debug_slots()
pos = np.array(sorted(e["preambleIdx"] for e in ev if "preambleIdx" in e))
print("frames", len(ev), "slots", len(info), "per frame", len(info) / len(ev), "FRAME", FRAME, "META", META)
# arena tile -> global position of the tile's first window: base_g = -hist at the first push
hist = None
st = m.stream_counters()
# the first push starts at arena index hist; recover it from the smallest slotted tile vs the first frame
tile_first = info[:, 1].astype(np.int64) * 256 - 511
# global = arena - hist: estimate hist from candidates: a candidate tile must contain a frame start
cand = info[:, 2] != 0
# try all plausible hist values (multiple of 4096)
best = None
for h in range(0, 1 << 20, 4096):
    g = tile_first[cand] - h
    d = np.searchsorted(pos, g + 256) - 1
    ok = (d >= 0) & (np.abs(pos[np.clip(d, 0, len(pos) - 1)] - g) < 600)
    if best is None or ok.sum() > best[1]:
        best = (h, ok.sum())
hist = best[0]
print("hist", hist, "candidates matched", best[1], "of", cand.sum())
g = tile_first - hist
k = np.clip(np.searchsorted(pos, g + 300) - 1, 0, len(pos) - 1)
rel = g - pos[k]                      # first window of the tile relative to the frame's preamble start
print("relative position histogram (tile first window - preamble start), all slots / candidate slots")
bins = np.arange(-1024, FRAME + 1024, 256)
ha, _ = np.histogram(rel, bins)
hc, _ = np.histogram(rel[cand], bins)
for b, a, c in zip(bins[:-1], ha, hc):
    if a:
        sel = (rel >= b) & (rel < b + 256)
        print(f"{b:7d} {a:6d} {c:6d}  maxmetric~{np.median(ts[sel, 0]):.3f} (max {ts[sel, 0].max():.3f})  minRa~{np.median(ts[sel, 1]):.4g} (min {ts[sel, 1].min():.4g})  minRb~{np.median(ts[sel, 2]):.4g} (min {ts[sel, 2].min():.4g})  excluded {ts[sel, 3].mean():.2f}")
m.close()
