"""Walker statistics on the C2 shape (development aid; needs the -DAMB_WALK_STATS build: AMB_LIB=.../libamb_stats.so)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import audio_modem_b200 as amb  # noqa: E402

cfg, mod, rep, chunk = "standard", "QAM16", 1, 4096
FRAME, CORE, META = 28431, 25344, 16416
F = int(os.environ.get("FRAMES", "4096"))
dev = torch.device("cuda", 0)
m = amb.Modem(cfg, mod, rep, device=0)
clean = torch.zeros((1, FRAME), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
m.build_data_chunk_frames(n=1, data_len=chunk, seq0=0, synth_seed=2, out=clean)
m.synchronize()
sigma = float((clean[0, 2205:2205 + CORE].double().pow(2).mean() / 10 ** 2.5).sqrt())
kp = next(k for k in range(1, 4097) if (META + FRAME * k) % 4096 == 0)
pro = torch.zeros((1, META + kp * FRAME), dtype=torch.float32, device=dev)
seg = torch.zeros((1, F * FRAME), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
m.build_metadata_frame(128000, 500 << 20, chunk, "bench.bin", noise_sigma=sigma, noise_seed=3, noise_g0=0, out=pro[0, :META])
m.build_data_chunk_frames(n=kp, data_len=chunk, seq0=0, synth_seed=2, noise_sigma=sigma, noise_seed=3, noise_g0=META, out=pro[0, META:].view(kp, FRAME))
m.build_data_chunk_frames(n=F, data_len=chunk, seq0=kp, synth_seed=2, noise_sigma=sigma, noise_seed=3, noise_g0=pro.shape[1], out=seg[0].view(F, FRAME))
m.synchronize()
m.stream_open(1, 4096, max(seg.shape[1], pro.shape[1]), max_results=F + 64)
d_res = torch.empty((F + 64) * 88, dtype=torch.uint8, device=dev)
d_bytes = torch.empty((F + 64, m.bytes_stride), dtype=torch.uint8, device=dev)
m.stream_push(pro, device_results=(d_res, d_bytes))
names = ["rows", "direct", "refine_store", "refine_direct", "refine_memo", "skip_rounds", "scan_calls", "refine_fail", "exact_offsets",
         "cyc_eval", "cyc_refine", "cyc_skip", "cyc_scan", "exact_replays"]
base = [m.debug_counter(16 + i) for i in range(len(names))]
m.profile(True)
for it in range(2):
    n = m.stream_push(seg, device_results=(d_res, d_bytes))
m.synchronize()
now = [m.debug_counter(16 + i) for i in range(len(names))]
print("frames per push", n, "walk stats per push:", {k: (b - a) / 2 for k, a, b in zip(names, base, now)})
prof = m.profile_read()
print({k: (round(v[0] / 2, 3), v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:10]})
m.close()
