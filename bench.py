#!/usr/bin/env python
"""bench.py — throughput of the OFDM receive hot path on B200, BASELINE.json's metric on its config C2.

Workload (config.workload = "C2"): a 16-QAM chunked stream — 4 KB chunks as OFDM frames of 28,431 samples
(buildDataChunkFrame, modem.js:763), AWGN at 25 dB relative to the frame-core power, received by the streaming
path (EMA DC removal, Schmidl-Cox scan, refine, collect, demodulate, CRC-32; app.js:749-981).  Synthetic data:
chunk bytes and noise come from counter-based generators, frames are built on the device by the float64-faithful
transmit path (bit-identical to the reference's frames).  One "step" = one push of `frames_per_step` whole frames.

  value : whole-job Msamples/s with the step inputs already resident in HBM (device time, CUDA events on the
          library's stream, max over ranks).  Inputs cycle over several distinct segments, each larger than L2.
  e2e   : the same metric through the host-buffer call (amb_stream_push with AMB_MEM_HOST from pinned memory):
          host->device copy of the samples and device->host copy of results and decoded bytes inside the timed region.
  roofline : the kernel with the largest share of the timed region, algorithmic bytes / its mean launch duration
          (per-kernel CUDA events recorded by the library during the timed region) against MEASURED_PEAKS.json.
  cpu_baseline : the oracle (C++ float64 restatement of modem.js/app.js; Node is not available) on the host cores.

`--impl reference` times that CPU path alone (all host cores) on the same config, metric and unit.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG, MOD, REP, CHUNK = "standard", "QAM16", 1, 4096
FRAME = 28431            # samples per data-chunk frame incl. silences (2205 + 44*576 + 882)
CORE = 25344             # preamble-1 .. last data symbol
META = 16416
SNR_DB = 25.0
METRIC = "rx_audio_msamples_per_s"
UNIT = "Msamples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=int(os.environ.get("AMB_FRAMES_PER_STEP", "16384")))
    ap.add_argument("--segments", type=int, default=3, help="distinct resident input segments cycled through")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, extra=None):
    d = {"workload": "C2: 16-QAM chunked stream, 4 KB chunks as 28,431-sample OFDM frames, AWGN 25 dB, streaming preamble scan + CRC-32",
         "ofdm_config": CFG, "constellation": MOD, "repetition": REP, "chunk_bytes": CHUNK, "frames_per_step": args.frames_per_step,
         "samples_per_step": args.frames_per_step * FRAME, "block_len": 4096, "snr_db": SNR_DB,
         "l2_policy": "inputs larger than L2 (step input %.0f MB, %d distinct segments cycled)" % (args.frames_per_step * FRAME * 4 / 1e6, args.segments),
         "parallelism": "streams sharded by rank, no collective on the hot path"}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------ CPU leg (oracle)
def cpu_leg(steps, warmup, frames_per_thread, threads=None, budget_s=None):
    """Times the oracle's streaming receiver on `threads` host threads, each decoding its own copy of one stream
    segment (metadata + frames_per_thread data frames) per step.  Returns (Msamples/s, payload MB/s, cores, sample)."""
    import numpy as np

    from oracle import oracle as O

    threads = threads or (os.cpu_count() or 1)
    parts = [O.buildMetadataFrame(CFG, frames_per_thread, frames_per_thread * CHUNK, CHUNK, "bench.bin", MOD, REP)]
    for i in range(frames_per_thread):
        parts.append(O.buildDataChunkFrame(CFG, O.synthBytes(2, i, CHUNK), i, MOD, REP))
    x = np.concatenate(parts)
    p = O.power(parts[1][2205:2205 + CORE])
    x = O.addAwgnCtr(x, (p / 10 ** (SNR_DB / 10)) ** 0.5, 3)
    n = (len(x) + 4095) // 4096 * 4096
    x = np.concatenate([x, np.zeros(n - len(x), np.float32)])
    ok = [0] * threads
    pb = [0] * threads

    def work(t):
        ok[t], pb[t] = O.benchStream(CFG, MOD, REP, x)     # ctypes releases the GIL: the C++ runs in parallel

    def one_step():
        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        times.append(one_step())
        if budget_s and time.perf_counter() - t_begin > budget_s:
            break
    assert all(o == frames_per_thread + 1 for o in ok), ok       # every frame found with a valid CRC
    total = sum(times)
    samples = len(times) * threads * n
    payload = len(times) * sum(pb)
    sample = f"{threads} threads x (metadata + {frames_per_thread} data frames = {n} samples) per step, {len(times)} steps"
    return samples / total / 1e6, payload / total / 1e6, threads, sample, total / len(times) * 1e3


def run_reference(args, rank):
    if rank != 0:
        return
    v, pmb, cores, sample, ms = cpu_leg(args.steps, args.warmup, frames_per_thread=64)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "payload_mb_per_s": pmb,
            "config": config_dict(args, {"reference_note": "CPU path = C++ float64 restatement of modem.js + app.js StreamingReceiver (oracle/); "
                                         "Node.js is not installed, so modem.js itself cannot run here"}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms from before the warm-up until after the last timed leg;
    only the samples taken between mark_begin() and mark_end() (GPU under load: warm-up, timed steps, profiling pass,
    end-to-end leg) are summarised."""
    Q = "timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime

        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons, all_sm = [], [], [], set(), []
        for ln in out.strip().splitlines():
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                v_sm, v_mx, v_pw = float(f[2]), float(f[3]), float(f[4])
            except ValueError:
                continue
            all_sm.append(v_sm)
            if self.t0 is not None and self.t1 is not None and not (self.t0 - 0.02 <= ts <= self.t1 + 0.02):
                continue
            sm.append(v_sm); mx.append(v_mx); pw.append(v_pw)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm and all_sm:      # clock skew between nvidia-smi's timestamps and time.time(): fall back to every sample
            sm = all_sm
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ GPU arm
def prologue_frames():
    """data frames after the metadata frame that make the prologue a whole number of 4096-sample blocks"""
    for k in range(1, 4097):
        if (META + FRAME * k) % 4096 == 0:
            return k
    raise AssertionError


def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch

    import audio_modem_b200 as amb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    clocks = ClockSampler(local_rank)
    clocks.start()
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)") if "hbm_gbs" in peaks else (6650.0, "fallback 6.65 TB/s (of fallback)")

    F = args.frames_per_step
    step_samples = F * FRAME
    m = amb.Modem(CFG, MOD, REP, device=local_rank)
    # noise level: 25 dB below the power of a frame core (preamble-1 .. last data symbol), measured on one clean frame
    clean = torch.zeros((1, FRAME), dtype=torch.float32, device=dev)
    m.build_data_chunk_frames(n=1, data_len=CHUNK, seq0=0, synth_seed=2, out=clean)
    sigma = float((clean[0, 2205:2205 + CORE].double().pow(2).mean() / 10 ** (SNR_DB / 10)).sqrt())
    del clean

    kp = prologue_frames()
    pro_samples = META + kp * FRAME
    max_push = max(step_samples, pro_samples)
    seed_payload, seed_noise = 2 + 1000 * rank, 3 + 1000 * rank
    # resident inputs: prologue + `segments` distinct steps
    pro = torch.zeros((1, pro_samples), dtype=torch.float32, device=dev)
    m.build_metadata_frame(128000, 500 * 1024 * 1024, CHUNK, "bench.bin", noise_sigma=sigma, noise_seed=seed_noise, noise_g0=0, out=pro[0, :META])
    m.build_data_chunk_frames(n=kp, data_len=CHUNK, seq0=0, synth_seed=seed_payload, noise_sigma=sigma, noise_seed=seed_noise, noise_g0=META,
                              out=pro[0, META:].view(kp, FRAME))
    segs = []
    for s in range(args.segments):
        t = torch.zeros((1, step_samples), dtype=torch.float32, device=dev)
        m.build_data_chunk_frames(n=F, data_len=CHUNK, seq0=kp + s * F, synth_seed=seed_payload, noise_sigma=sigma, noise_seed=seed_noise,
                                  noise_g0=pro_samples + s * step_samples, out=t[0].view(F, FRAME))
        segs.append(t)
    m.synchronize()

    m.stream_open(1, 4096, max_push, max_results=max(F, kp) + 64)
    got = m.stream_push(pro, collect=False)
    c0 = m.stream_counters()
    assert c0["metaReceived"] == 1 and got == kp + 1, (c0, got, kp)

    ext = torch.cuda.ExternalStream(m.cuda_stream(), device=dev)
    # device-resident leg: results and decoded bytes stay in HBM too
    cap = max(F, kp) + 64
    d_res = torch.empty(cap * 88, dtype=torch.uint8, device=dev)
    d_bytes = torch.empty((cap, m.bytes_stride), dtype=torch.uint8, device=dev)
    dres = (d_res, d_bytes)
    clocks.mark_begin()
    for w in range(args.warmup):
        m.stream_push(segs[w % len(segs)], device_results=dres)
    m.synchronize()
    torch.cuda.synchronize()
    c0 = m.stream_counters()
    if dist:
        dist.barrier()
    launches0 = m.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    frames_found = 0
    for k in range(args.steps):
        frames_found += m.stream_push(segs[(args.warmup + k) % len(segs)], device_results=dres)
    e1.record(ext)
    m.synchronize()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = m.kernel_launches - launches0
    c1 = m.stream_counters()
    decoded = c1["framesDecoded"] - c0["framesDecoded"]
    crc_bad = c1["crcErrors"] - c0["crcErrors"]
    # every frame must be found and parsed; at 25 dB a few frames per 10,000 fail their CRC (real bit errors: the zero-padded
    # last symbol sets the frame peak, so the payload symbols sit ~24 dB below full scale)
    assert frames_found == args.steps * F and decoded == args.steps * F and crc_bad <= max(2, decoded // 500) and c1["frameErrors"] == c0["frameErrors"], (frames_found, c0, c1)

    # per-kernel device time of the same steps (separate pass: the event pairs around every launch perturb the step time)
    m.profile(True)
    prof_steps = min(args.steps, 4)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(ext)
    for k in range(prof_steps):
        m.stream_push(segs[(args.warmup + args.steps + k) % len(segs)], device_results=dres)
    p1.record(ext)
    m.synchronize()
    prof_ms = p0.elapsed_time(p1)
    prof = m.profile_read()
    m.profile(False)
    slots = {"slotted_tiles_per_frame": m.debug_counter(4) / F, "candidate_tiles_per_frame": m.debug_counter(5) / F,
             "link_iterations_per_push": m.debug_counter(2) / max(1, m.debug_counter(3))}

    # end-to-end: host buffers (page-locked), H2D of the samples + D2H of results and decoded bytes inside the timed region
    host = amb.modem.pinned_empty((1, step_samples), np.float32)
    hx = host.array
    hx[...] = segs[0].cpu().numpy()
    m.stream_push(hx, collect=False)
    m.synchronize()
    t0 = time.perf_counter()
    e2e_frames = 0
    for k in range(args.e2e_steps):
        e2e_frames += m.stream_push(hx, collect=False)
    m.synchronize()
    e2e_s = time.perf_counter() - t0
    assert e2e_frames == args.e2e_steps * F, e2e_frames
    # wire-format ingest (reported apart: 2 B per sample changes the byte count): the same samples as signed
    # 16-bit PCM through amb_stream_push_pcm16, device-resident and end to end
    pcm_dev = (segs[0] * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
    for w in range(2):
        assert m.stream_push(pcm_dev, device_results=dres) == F
    m.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pcm_steps = min(args.steps, 6)
    q0.record(ext)
    for k in range(pcm_steps):
        m.stream_push(pcm_dev, device_results=dres)
    q1.record(ext)
    m.synchronize()
    pcm_ms = q0.elapsed_time(q1) / pcm_steps
    host16 = amb.modem.pinned_empty((1, step_samples), np.int16)
    hp = host16.array
    hp[...] = pcm_dev.cpu().numpy()
    m.stream_push(hp, collect=False)
    m.synchronize()
    t0 = time.perf_counter()
    pcm_frames = 0
    for k in range(args.e2e_steps):
        pcm_frames += m.stream_push(hp, collect=False)
    m.synchronize()
    pcm_e2e_s = time.perf_counter() - t0
    assert pcm_frames == args.e2e_steps * F, pcm_frames
    clocks.mark_end()
    clk = clocks.stop()
    clk["window"] = "warm-up + timed steps + profiling pass + end-to-end legs (the same workload throughout)"
    d2h = F * (m.bytes_stride + 88)
    m.stream_close()
    del hx, hp
    host.free()
    host16.free()

    # aggregate over ranks
    total_ms = amb.shard.max_over_ranks(dev_ms, device=dev) if dist else dev_ms
    e2e_s_max = amb.shard.max_over_ranks(e2e_s, device=dev) if dist else e2e_s
    counters = {"frames": args.steps * F, "payloadBytes": args.steps * F * CHUNK, "samples": args.steps * step_samples,
                "crcValidFrames": decoded - crc_bad, "launches": launches}
    tot = amb.shard.gather_counters(counters, device=dev) if dist else counters
    value = tot["samples"] / (total_ms / 1e3) / 1e6
    e2e_value = world * args.e2e_steps * step_samples / e2e_s_max / 1e6
    pcm_ms_max = amb.shard.max_over_ranks(pcm_ms, device=dev) if dist else pcm_ms
    pcm_e2e_s_max = amb.shard.max_over_ranks(pcm_e2e_s, device=dev) if dist else pcm_e2e_s
    pcm16 = {"input": "signed 16-bit PCM, sample = s / 32768 (amb_stream_push_pcm16); same frames as the float32 legs",
             "value": world * step_samples / (pcm_ms_max / 1e3) / 1e6, "ms_per_step": pcm_ms_max, "unit": UNIT,
             "e2e": {"value": world * args.e2e_steps * step_samples / pcm_e2e_s_max / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * step_samples,
                     "d2h_bytes_per_step": d2h, "steps": args.e2e_steps}}

    # roofline of the dominant kernel (largest share of the timed region on this rank)
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else None
    algo_step = 4 * step_samples + F * (CHUNK + 64)                 # 4 B/sample once + payload + 64 B result per frame
    per_launch = {"k_demod": F * (4 * 44 * 576 + CHUNK + 64), "k_demod_fast": F * (4 * 44 * 576 + CHUNK + 64), "k_front": 4 * step_samples, "k_parse": F * (4202 + CHUNK + 64), "k_frame_max": F * 4 * CORE}
    roof = None
    if dom:
        name, (ms, cnt) = dom
        b = per_launch.get(name, 4 * step_samples)
        achieved = b / (ms / cnt / 1e3) / 1e9
        # measured DRAM traffic of that kernel (ncu --set full capture committed under profiles/), scaled to this launch
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01", "traffic.json"))).get(name)
            if tj and "dram_bytes_per_frame" in tj:
                traffic = tj["dram_bytes_per_frame"] * F
            elif tj and "dram_bytes_per_sample" in tj:
                traffic = tj["dram_bytes_per_sample"] * step_samples
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b, "mean_launch_ms": ms / cnt,
                "share_of_step": ms / prof_ms,
                "whole_step": {"algorithmic_bytes": algo_step, "achieved": algo_step * args.steps / (dev_ms / 1e3) / 1e9,
                               "frac": algo_step * args.steps / (dev_ms / 1e3) / 1e9 / hbm_peak},
                "profiled_steps": prof_steps, "profiled_ms_per_step": prof_ms / prof_steps,
                "kernels_ms_per_step": {k: v[0] / prof_steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "payload_mb_per_s": tot["payloadBytes"] / (total_ms / 1e3) / 1e6,
            "config": config_dict(args), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * step_samples, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps},
            "gpu_launches": tot["launches"], "roofline": roof, "pcm16": pcm16,
            "walker": slots, "check": {"frames_per_step_found": frames_found // args.steps, "crc_valid_frames": tot["crcValidFrames"], "crc_errors": crc_bad}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, pmb, cores, sample, _ = cpu_leg(steps=1000, warmup=1, frames_per_thread=64, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "payload_mb_per_s": pmb}
    if rank == 0:
        print(json.dumps(line))
    m.close()
    if dist:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
