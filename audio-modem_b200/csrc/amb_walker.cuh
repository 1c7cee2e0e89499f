// Streaming receiver state machine (app.js:749-981) as a warp-cooperative "walker".
//
// The reference scans every sample sequentially.  Here a bandwidth-bound scan kernel has already
// reduced the stream to per-tile summaries of the Schmidl-Cox window sums (max|P|, min/max Ra, min/max
// Rb over 256 consecutive windows); the walker replays the reference's state machine call by call
// (one call = one processAudioBlock of `block_len` samples), skips tiles whose summary proves that no
// window can pass `metric > 0.5` with the energy gate, and evaluates the remaining windows exactly in
// float64.  All decisions (candidate, 0.7 drop, end-of-buffer commit, refine, collect, reset) follow the
// reference line by line, including two quirks:
//   * app.js:826 advances acScanPos at the last position of a call although no slide happened, so from
//     the next call on the running sums describe window d-1 at its first position and afterwards carry a
//     constant defect  T(a-1) - T(a+255)  (a = first position of that call) until the next re-init.  The
//     walker carries that defect explicitly (Ep, Ea, Eb) and the one-sample lag flag.
//   * the pre-metadata collection window is a fixed 280-byte guess (app.js:892-896).
// Near-ties and threshold crossings.  The walker evaluates window sums directly (or from prefix differences), the reference
// carries sliding accumulators: after up to AMB_REPLAY_MAX slides the two differ by a few 1e-12 relative (3e-12 is the largest
// seen in the tests).  Whenever a decision of a call could come out differently because of that, the call is run again by
// amb_walk_scan_exact, which replays the reference's own recurrence from the position where its sums were last
// re-initialised (direct summation, app.js:791-803) and applies its strict comparisons:
//   * threshold decisions - a metric against 0.5 or 0.7 * best (app.js:811, :833), an energy against the 0.001 gate (:809,
//     :831) - inside a relative band of AMB_BAND = 1e-10 around equality;
//   * a candidate against the running best of the same call (app.js:811) inside AMB_BAND_TIE = 1e-12: the history the
//     reference's two values share cancels in their difference (what is left is the rounding of the slides between the two
//     positions, ~1e-15, and the change of the metric's gradient across the plateau times the common error, ~1e-14), and two
//     directly evaluated metrics differ from the true ones by ~1e-13 each.
// Only when the re-initialisation lies more than AMB_REPLAY_MAX slides back (a near-tie deep inside an idle stretch) does
// the old rule decide: a new best must exceed the old one by 1e-12 relative.
#pragma once
#include "amb_device.cuh"

enum { AMB_RX_IDLE = 0, AMB_RX_PREAMBLE = 1, AMB_RX_COLLECT = 2 };

// development build (-DAMB_WALK_STATS): event counts of the walkers, read through amb_debug_counter(16 + i)
//   0 window batches copied from the stored rows, 1 evaluated directly, 2 refines from the screening store, 3 direct refines,
//   4 refines answered by the memo, 5 cold-call skip rounds, 6 scan calls, 7 refine failures, 8 exact offsets after screening
#if defined(AMB_WALK_STATS) && defined(__CUDACC__)
__device__ unsigned long long g_amb_walk_stats[16];
// per-warp accumulators and a trace of the individual walks of k_link_walk (one row per walk: see k_link_walk)
#define AMB_TRACE_ROWS 16384
#define AMB_TRACE_W 24
__device__ unsigned long long g_amb_warp_acc[16384][16];
__device__ unsigned long long g_amb_walk_trace[AMB_TRACE_ROWS][AMB_TRACE_W];
__device__ unsigned int g_amb_walk_trace_n;
__device__ __forceinline__ void amb_stat_add(int i, unsigned long long n) {
    if ((threadIdx.x & 31) == 0) { atomicAdd(&g_amb_walk_stats[i], n); g_amb_warp_acc[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & 16383][i] += n; }
}
#if defined(__CUDA_ARCH__)
#define AMB_STAT(i) amb_stat_add(i, 1ull)
#define AMB_STAT_ADD(i, n) amb_stat_add(i, (unsigned long long)(n))
#define AMB_STAT_T0() const long long stat_t0__ = clock64()
#define AMB_STAT_T1(i) amb_stat_add(i, (unsigned long long)(clock64() - stat_t0__))
#else
#define AMB_STAT_T0() do { } while (0)
#define AMB_STAT_T1(i) do { } while (0)
#define AMB_STAT(i) do { } while (0)
#define AMB_STAT_ADD(i, n) do { } while (0)
#endif
#else
#define AMB_STAT(i) do { } while (0)
#define AMB_STAT_ADD(i, n) do { } while (0)
#define AMB_STAT_T0() do { } while (0)
#define AMB_STAT_T1(i) do { } while (0)
#endif

struct AmbStreamState {
    int32_t state, ac_init, lag, meta_received, chunk_size, total_chunks;
    int32_t total_file_size;            // of the metadata frame in force (ChunkAssembler.totalFileSize)
    int32_t init_call;                  // call in which the running sums were last re-initialised (meaningful while ac_init)
    int64_t init_pos;                   // ... and the position at which they were (amb_walk_scan_exact replays from there)
    int64_t ac_scan_pos, pre_pos, frame_end;
    double Ep, Ea, Eb;
    double coarse, fine;
    int64_t next_call;
    int64_t frames_decoded, frame_errors, false_pos, received, crc_errors, dups;
    double dc_mean;
    int64_t written;
};

struct AmbFrameRec {
    int32_t stream, status;
    int64_t pos, len;
    double coarse, fine;
    AmbStreamState snap;   // walker state right after the frame's reset to IDLE
};

struct AmbRefineMemo { int64_t coarse, fs, fe, pos; double fine; int32_t valid, pad; };

struct AmbWalkGeom {
    int64_t block_len;     // samples per emulated processAudioBlock call
    int64_t ring_cap;      // RingBuffer capacity, app.js:712-715
    int64_t sum_base;      // global sample index of summary tile 0 (multiple of 256)
    int64_t sum_tiles;     // tiles available from sum_base
    int64_t sum_lo;        // summaries valid for window-end samples q >= sum_lo
    // Optional stores written by the scan kernels (NULL = evaluate inline):
    const int32_t* tile_slot;   // per summary tile: slot in the stores, -1 = none
    const double* tri;          // [slot][3][256]  P, Ra, Rb of the tile's 256 windows
    const float* xc;            // [slot][xc_w]    preamble-1 correlation metric (screening values) at tile_first - R + i
    int xc_w;
    const int32_t* slot_info;   // [slot][4], entry 2 = the slot's xc row was computed (NULL: every row is valid or marked UNSURE)
    const float* tsum;          // [slot][4]: exact extremes over the slot's 256 windows: max metric (no defect, rounded up), min Ra,
                                // min Rb (rounded down); NULL = not available
    // Refine memo between the two phases of the link-parallel resolution: the refine (app.js:862-877) is a pure function
    // of (coarse position, search range), so phase B reuses what phase A computed for the same run when these agree.
    struct AmbRefineMemo* memo_out;    // phase A: the last successful refine of the walk is recorded here (or NULL)
    const struct AmbRefineMemo* memo_in;   // phase B: consulted before computing (or NULL)
    const double* F;            // per call: inclusive prefix of the block-boundary defect terms (3 doubles per call)
    int64_t F_call0, F_calls;   // F[(call - F_call0) * 3 + c] valid for 0 <= call - F_call0 < F_calls
    // Optional (k_cold_calls): cold_tag[call - F_call0] = c when the call was evaluated exactly, in parallel, under the defect a
    // receiver carries there after re-initialising its sums in call c, and holds no candidate window; -1 otherwise
    const int32_t* cold_tag;
    // Optional (device, k_link_walk with few links): team_n warps of one CTA walk the SAME link redundantly (identical inputs,
    // identical control flow, so every warp reaches the same CTA barriers) and share the one expensive data-parallel step, the
    // float32 screening of a direct refine: warp team_rank computes every team_n-th group of offsets into team_vbuf (shared
    // memory, team_vcap floats).  team_n <= 1: a warp works alone.
    int team_n, team_rank, team_vcap;
    float* team_vbuf;
    float* team_stage;          // device, team_n > 1: shared-memory staging area of amb_xcorr_screen_team (AMB_TEAM_STAGE_FLOATS floats), or NULL
};

// screening values in the xc store: a metric in [-1, 1]; AMB_XC_FAIL where the denominator fails its 0.001 test;
// AMB_XC_UNSURE where float32 cannot tell
// slot_info[slot * 4 + 2] (written by k_tri_store): the tile holds a candidate window (its screening row in xc is computed);
// its float64 rows in tri are written
#define AMB_SLOT_CAND 1
#define AMB_SLOT_ROWS 2
#define AMB_XC_FAIL (-2.0f)
#define AMB_XC_UNSURE (2.0f)

// tile_slot entry: -1 = no slot, else the slot number with two flags k_tri_store adds once it has looked at the tile
// (the same flags as slot_info[2], here so that the walker learns them with the slot number in one load)
#define AMB_TS_ROWS (1 << 30)
#define AMB_TS_CAND (1 << 29)
#define AMB_TS_NOXC (1 << 28)         // set by k_xcorr_store: a candidate tile whose screening row was left out (a neighbour holds the peak)
#define AMB_TS_MASK ((1 << 28) - 1)
AMB_HD int amb_tile_slot_raw(const AmbWalkGeom& G, int64_t j) {
    if (!G.tile_slot) return -1;
    const int64_t rel = j - (G.sum_base >> 8);
    if (rel < 0 || rel >= G.sum_tiles || (j << 8) < G.sum_lo) return -1;       // below sum_lo: entries of an earlier push
    return G.tile_slot[rel];
}
AMB_HD int amb_tile_slot_of(const AmbWalkGeom& G, int64_t j) {
    const int raw = amb_tile_slot_raw(G, j);
    return raw < 0 ? -1 : (raw & AMB_TS_MASK);
}

// tile summary: {max|P|, minRa, maxRa, minRb, maxRb}
#define AMB_SUM_W 5

AMB_HD bool amb_tile_hot(const float* s, double Ep, double Ea, double Eb) {
    const double g = 0.001;
    const double up = 1.0 + 1e-5, eps = 1e-9;
    const double maxA = (double)s[2] * up + eps + Ea, maxB = (double)s[4] * up + eps + Eb;
    if (!(maxA > g) || !(maxB > g)) return false;
    double minA = (double)s[1] * (1.0 - 1e-5) - eps + Ea, minB = (double)s[3] * (1.0 - 1e-5) - eps + Eb;
    if (minA < g) minA = g;
    if (minB < g) minB = g;
    const double pm = (double)s[0] * up + eps + fabs(Ep);
    return pm * pm > 0.5 * minA * minB * (1.0 - 1e-5);
}

// Second, sharper test for a slotted tile, from the exact per-window values k_tri_store reduced over the windows with
// Ra > 0.0005 and Rb > 0.0005 (half the energy gate): max metric (no defect, rounded up), min Ra, min Rb (rounded down), and
// whether any window was left out.  Left-out windows cannot pass the gate of 0.001 unless the defect adds more than 0.0004
// to an energy.  With the defect E in force,
//   sqrt(metric') = |P + Ep| / sqrt((Ra + Ea)(Rb + Eb)) <= sqrt(metric) sqrt(Ra Rb / (Ra' Rb')) + |Ep| / sqrt(Ra' Rb'),
// Ra' >= Ra - |Ea| >= minRa - |Ea| (and Ra / Ra' is largest at the smallest Ra).  True = no window of the tile can be a
// candidate.
AMB_HD bool amb_slot_cold(const float* ts, double Ep, double Ea, double Eb) {
    if (ts[3] != 0.f && (!(Ea < 0.0004) || !(Eb < 0.0004))) return false;
    const double mA = (double)ts[1] - fabs(Ea), mB = (double)ts[2] - fabs(Eb);
    if (!(mA > 1e-4) || !(mB > 1e-4)) return false;
    const double ka = (double)ts[1] / mA, kb = (double)ts[2] / mB;
    const double r = sqrt((double)ts[0] * ka * kb) * (1.0 + 1e-6) + fabs(Ep) / sqrt(mA * mB);
    return r * r < 0.5 * (1.0 - 1e-6);
}

AMB_HD int amb_tile_slot_of(const struct AmbWalkGeom& G, int64_t j);

// amb_tile_hot, refined by amb_slot_cold where the tile has a slot (rel = tile index relative to the summary base)
AMB_HD bool amb_tile_hot2(const struct AmbWalkGeom& G, const float* sums, int64_t jj, int64_t rel, double Ep, double Ea, double Eb);

struct AmbWalkScratch {      // per-warp shared memory
    double M[264], A[264], B[264];   // M: the window sums P on the way in (amb_eval_windows), the metrics once a batch is formed
    double x[64];
    int32_t hot[32];
    int32_t gate[264];
};

AMB_HD int64_t amb_frame_samples(const AmbTables* T, int64_t payload) {   // modem.js:863-874
    const int64_t bits = payload * 8 * T->rep, per = (int64_t)T->n_data * T->bps;
    return (3 + (bits + per - 1) / per) * T->sym_len;
}

// metric with the carried defect; returns gate flag
AMB_HD int amb_metric(double P, double A, double B, const AmbStreamState& st, double* m) {
    const double a = A + st.Ea, b = B + st.Eb, p = P + st.Ep;
    if (a > 0.001 && b > 0.001) { *m = amb_div(amb_mul(p, p), amb_mul(a, b)); return 1; }
    *m = 0.0;
    return 0;
}

AMB_HD void amb_commit(AmbStreamState& st, int64_t best_pos, double best, int64_t next_pos, int lag) {
    st.pre_pos = best_pos; st.coarse = best; st.state = AMB_RX_PREAMBLE; st.ac_scan_pos = next_pos; st.lag = lag;
}

AMB_HD bool amb_tile_hot2(const AmbWalkGeom& G, const float* sums, int64_t jj, int64_t rel, double Ep, double Ea, double Eb) {
    if (!amb_tile_hot(sums + rel * AMB_SUM_W, Ep, Ea, Eb)) return false;
    if (G.tsum) {
        const int slot = amb_tile_slot_of(G, jj);
        if (slot >= 0 && amb_slot_cold(G.tsum + (int64_t)slot * 4, Ep, Ea, Eb)) return false;
    }
    return true;
}

// Window sums at cnt consecutive offsets from d0: copied from the scan kernel's store where the tile has a slot,
// evaluated directly otherwise.
AMB_HD void amb_eval_windows(const AmbWalkGeom& G, const AmbSig& sig, int64_t d0, int cnt, double* P, double* A, double* B) {
    AMB_STAT_T0();
    int done = 0;
    while (done < cnt) {
        const int64_t d = d0 + done;
        const int64_t j = (d + 511) >> 8;
        const int64_t tile_first = (j << 8) - 511;
        int seg = (int)(tile_first + 256 - d);
        if (seg > cnt - done) seg = cnt - done;
        const int raw = amb_tile_slot_raw(G, j);
        const int slot = raw & AMB_TS_MASK;
        if (raw >= 0 && G.tri && (raw & AMB_TS_ROWS)) {
            const double* src = G.tri + (int64_t)slot * 768 + (d - tile_first);          // [slot][P | Ra | Rb][256]
            AMB_WARP_FOR(i, seg) { P[done + i] = src[i]; A[done + i] = src[256 + i]; B[done + i] = src[512 + i]; }
            AMB_WARP_SYNC();
            AMB_STAT(0);
        } else {
            AMB_STAT(1);
            amb_window_sums(sig, d, seg, P + done, A + done, B + done);
        }
        done += seg;
    }
    AMB_STAT_T1(9);
}

// 1 if none of the 16 summary tiles of call k (window-end samples [k*B, (k+1)*B)) can hold a candidate under E
AMB_HD int amb_call_cold(const AmbWalkGeom& G, const float* sums, int64_t k, double Ep, double Ea, double Eb) {
    const int64_t j0 = (k * G.block_len) >> 8, n = G.block_len >> 8;
    for (int64_t t = 0; t < n; t++) {
        const int64_t jj = j0 + t, rel = jj - (G.sum_base >> 8);
        if (rel < 0 || rel >= G.sum_tiles || (jj << 8) < G.sum_lo) return 0;
        if (amb_tile_hot2(G, sums, jj, rel, Ep, Ea, Eb)) return 0;
    }
    return 1;
}

// Idle fast-forward.  While the scan is caught up and uninterrupted (state IDLE, sums initialised, lag 1, next
// position = first position of the next call) nothing can happen in a call whose 16 tiles are cold under the defect
// that will be in force there, E_k = E + F[k] - F[k_now - 1].  Returns how many whole calls from st.next_call can be
// skipped (< k_end) and advances the state over them.
struct AmbColdCache { int64_t base; unsigned hot; int32_t init_call; };     // hot flags of calls base .. base + 31 on the trajectory of init_call

// Length of the run of calls from k on that k_cold_calls tagged cold for the trajectory of init_call: eight tags per lane and
// round, all loads in flight at once (an idle stretch of a long capture is thousands of such calls, and a lone warp pays a full
// memory latency per round).  Called once a whole round of 32 calls was cold.
AMB_HD_COLD int64_t amb_skip_tagged_calls(const int32_t* cold_tag, int64_t F_call0, int64_t F_calls, int32_t init_call, int64_t k, int64_t k_end) {
    const unsigned all = AMB_NLANES == 32 ? 0xffffffffu : 1u;
    int64_t total = 0;
    if (init_call < 0) return 0;
    for (;;) {
        const int64_t kk0 = k + total + 8 * (int64_t)AMB_LANE();
        int32_t tag[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 8; q++) {
            const int64_t kk = kk0 + q, fi = kk - F_call0;
            tag[q] = (kk < k_end && fi >= 0 && fi < F_calls) ? cold_tag[fi] : -1;
        }
        int lead = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 8; q++) if (lead == q && tag[q] == init_call) lead++;
        const unsigned notfull = ~amb_warp_ballot(lead == 8) & all;
        const int fl = notfull ? amb_ffs(notfull) - 1 : -1;
        total += fl < 0 ? 8 * AMB_NLANES : 8 * fl + (int)amb_warp_bcast((double)lead, fl);
        if (fl >= 0) return total;
    }
}

AMB_HD int64_t amb_walk_skip_cold_calls(const AmbWalkGeom& G, const float* sums, AmbStreamState& st, int64_t k_end, AmbWalkScratch* W, AmbColdCache& cc) {
    if (!G.F || st.state != AMB_RX_IDLE || !st.ac_init || st.lag != 1) return 0;
    int64_t k = st.next_call;
    if (st.ac_scan_pos != k * G.block_len - 511) return 0;
    if (st.ac_scan_pos < (k + 1) * G.block_len - G.ring_cap + 2 * AMB_HALF) return 0;      // ring clamp would fire
    const int64_t f0 = k - 1 - G.F_call0;
    if (f0 < 0) return 0;
    const double b0 = G.F[f0 * 3], b1 = G.F[f0 * 3 + 1], b2 = G.F[f0 * 3 + 2];
    int64_t skipped = 0;
    while (k < k_end) {
        unsigned hot;
        int valid;
        if (cc.base >= 0 && cc.init_call == st.init_call && k >= cc.base && k < cc.base + 32) {
            // flags of an earlier round on the same defect trajectory (a hot call in between was processed by the scan)
            hot = cc.hot >> (int)(k - cc.base);
            valid = 32 - (int)(k - cc.base);
        } else {
            AMB_STAT(5);
            AMB_WARP_FOR(t, 32) {
                const int64_t kk = k + t, fi = kk - G.F_call0;
                int cold = 0;
                if (kk < k_end && fi >= 0 && fi < G.F_calls) {
                    if (G.cold_tag && G.cold_tag[fi] == st.init_call) cold = 1;
                    else cold = amb_call_cold(G, sums, kk, st.Ep + (G.F[fi * 3] - b0), st.Ea + (G.F[fi * 3 + 1] - b1), st.Eb + (G.F[fi * 3 + 2] - b2));
                }
                W->hot[t] = cold ? 0 : 1;
            }
            AMB_WARP_SYNC();
            hot = 0u;
            for (int t = 0; t < 32; t++) if (W->hot[t]) hot |= 1u << t;
            AMB_WARP_SYNC();
            cc.base = k; cc.hot = hot; cc.init_call = st.init_call;
            valid = 32;
        }
        int first = hot ? amb_ffs(hot) - 1 : 32;
        if (first > valid) first = valid;
        k += first; skipped += first;
        if (first < valid) break;
        if (first == 32 && G.cold_tag) { const int64_t adv = amb_skip_tagged_calls(G.cold_tag, G.F_call0, G.F_calls, st.init_call, k, k_end); k += adv; skipped += adv; }
    }
    if (skipped > 0) {
        const int64_t fi = k - 1 - G.F_call0;
        st.Ep += G.F[fi * 3] - b0; st.Ea += G.F[fi * 3 + 1] - b1; st.Eb += G.F[fi * 3 + 2] - b2;
        st.ac_scan_pos = k * G.block_len - 511;
        st.next_call = k;
    }
    return skipped;
}

// A refined metric whose exact evaluation was deferred is carried as -1 - (screening value), i.e. inside (-4, -1).
#define AMB_FINE_DEFERRED(x) ((x) < -1.0 && (x) > -4.0)
AMB_HD double amb_resolve_fine(const AmbTables* T, const AmbSig& sig, int64_t pos, double fine) {
    if (!AMB_FINE_DEFERRED(fine)) return fine;
    double m;
    return amb_xcorr_at(T, sig, pos, &m) ? m : -INFINITY;
}

// float32 screening values of preamble-1's correlation metric (modem.js:577-585) at the n offsets lo, lo + 1, ... into v,
// one warp: a lane takes four consecutive offsets at a time, so a tap is loaded once for four multiply-adds.  Same
// encoding as the scan kernel's store: the metric, AMB_XC_FAIL where the denominator clearly fails its 0.001 test,
// AMB_XC_UNSURE where float32 cannot tell.  |value - exact| < 1e-4 (768 taps, sequential float32 sums: n u sum|s t| <=
// 768 * 6e-8 of the denominator, by Cauchy-Schwarz), inside the 5e-4 the consumers allow.
AMB_HD_COLD void amb_xcorr_screen(const AmbTables* T, const AmbSig s, int64_t lo, int n, float* v, int team_n = 1, int team_rank = 0) {
    const int plen = T->sym_len;
    const float tE = (float)T->pre1_energy;
    const float* t = T->pre1;
    for (int gq = team_rank * AMB_NLANES + AMB_LANE(); gq < (n + 3) / 4; gq += team_n * AMB_NLANES) {
        const int64_t d = lo + 4 * (int64_t)gq;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
        if (s.xf.kind == AMB_XF_NONE && d >= s.lo && d + plen + 3 <= s.hi) {
            const float* p = s.p + (d - s.base);
            float a0 = p[0], a1 = p[1], a2 = p[2];
            // (software-pipelined: the loads of the next four taps are in flight while the 20 multiply-adds of these run - a
            // lone warp has nothing else to hide their latency with)
            float n3 = p[3], n4 = p[4], n5 = p[5], n6 = p[6], u0 = t[0], u1 = t[1], u2 = t[2], u3 = t[3];
            for (int i = 0; i < plen; i += 4) {                         // (sym_len is a multiple of 4 in every config)
                const float a3 = n3, a4 = n4, a5 = n5, a6 = n6;
                const float t0 = u0, t1 = u1, t2 = u2, t3 = u3;
                if (i + 4 < plen) { n3 = p[i + 7]; n4 = p[i + 8]; n5 = p[i + 9]; n6 = p[i + 10]; u0 = t[i + 4]; u1 = t[i + 5]; u2 = t[i + 6]; u3 = t[i + 7]; }
                c[0] = fmaf(a0, t0, c[0]); c[1] = fmaf(a1, t0, c[1]); c[2] = fmaf(a2, t0, c[2]); c[3] = fmaf(a3, t0, c[3]);
                c[0] = fmaf(a1, t1, c[0]); c[1] = fmaf(a2, t1, c[1]); c[2] = fmaf(a3, t1, c[2]); c[3] = fmaf(a4, t1, c[3]);
                c[0] = fmaf(a2, t2, c[0]); c[1] = fmaf(a3, t2, c[1]); c[2] = fmaf(a4, t2, c[2]); c[3] = fmaf(a5, t2, c[3]);
                c[0] = fmaf(a3, t3, c[0]); c[1] = fmaf(a4, t3, c[1]); c[2] = fmaf(a5, t3, c[2]); c[3] = fmaf(a6, t3, c[3]);
                e[0] = fmaf(a0, a0, e[0]); e[0] = fmaf(a1, a1, e[0]); e[0] = fmaf(a2, a2, e[0]); e[0] = fmaf(a3, a3, e[0]);
                a0 = a4; a1 = a5; a2 = a6;
            }
            // the four windows share all but three samples at either end: Sum s^2 of offset q = that of offset 0 minus the
            // first q samples plus the q after the window (float32 differences of a few terms: error far below the 1e-3 band)
            e[1] = e[0] - p[0] * p[0] + p[plen] * p[plen];
            e[2] = e[1] - p[1] * p[1] + p[plen + 1] * p[plen + 1];
            e[3] = e[2] - p[2] * p[2] + p[plen + 2] * p[plen + 2];
        } else {
            for (int i = 0; i < plen; i++) {
                const float ti = t[i];
                for (int q = 0; q < 4; q++) { const float a = amb_sig(s, d + q + i); c[q] = fmaf(a, ti, c[q]); e[q] = fmaf(a, a, e[q]); }
            }
        }
        for (int q = 0; q < 4; q++) {
            if (4 * gq + q >= n) break;
            const float d2 = e[q] * tE;
            float r;
            if (d2 > 1e-6f * (1.f + 1e-3f)) r = c[q] / sqrtf(d2);
            else if (d2 < 1e-6f * (1.f - 1e-3f)) r = AMB_XC_FAIL;
            else r = AMB_XC_UNSURE;
            if (!(r == r)) r = AMB_XC_UNSURE;                           // NaN samples: let the exact evaluation decide
            v[4 * gq + q] = r;
        }
    }
    AMB_WARP_SYNC();
}

#define AMB_TS_PADI(i) ((i) + (((i) >> 4) << 2))
#define AMB_TEAM_STAGE_FLOATS (AMB_TS_PADI(1664 + 768 + 32) + 768 + 16)
#if defined(__CUDA_ARCH__)
// The same screening by the warps of a team (one CTA, team_n x 32 threads; called by all of them): samples and taps staged in
// shared memory, a thread takes 16 consecutive offsets and steps through the taps 16 at a time with a 32-sample register
// window (256 multiply-adds per 8 128-bit shared loads).  A one-hour capture's refine - 1,537 offsets x 768 taps, alone on its
// SM - took ~150,000 cycles with amb_xcorr_screen (global loads, four offsets per lane, three rounds); this takes one round.
// Samples are padded by 4 floats per 16 in shared memory: the 128-bit loads of the 8 threads of a quarter-warp (80 bytes apart)
// then hit distinct bank groups.  Values: a sequential float32 sum per offset, as in amb_xcorr_screen (same error bound).
__device__ __noinline__ void amb_xcorr_screen_team(const AmbTables* T, const AmbSig s, int64_t lo, int n, float* v, float* stage, int team_n, int team_rank) {
    const int plen = T->sym_len;                                     // a multiple of 16 in every config (576, 640, 768)
    const int nthreads = team_n * 32, tid = team_rank * 32 + (int)(threadIdx.x & 31);
    const int ngroups = (n + 15) >> 4;
    const int ns = 16 * ngroups + plen + 16;
    float* sm = stage;
    float* tp = stage + AMB_TS_PADI(1664 + 768 + 32);
    const float tE = (float)T->pre1_energy;
    for (int i = tid; i < ns; i += nthreads) sm[AMB_TS_PADI(i)] = amb_sig(s, lo + i);
    for (int i = tid; i < plen; i += nthreads) tp[i] = T->pre1[i];
    __syncthreads();
    for (int gq = tid; gq < ngroups; gq += nthreads) {
        float c[16], w[32];
#pragma unroll
        for (int q = 0; q < 16; q++) c[q] = 0.f;
        {
            const float4* p4 = reinterpret_cast<const float4*>(sm + AMB_TS_PADI(16 * gq));
#pragma unroll
            for (int j = 0; j < 4; j++) { const float4 u = p4[j]; w[4 * j] = u.x; w[4 * j + 1] = u.y; w[4 * j + 2] = u.z; w[4 * j + 3] = u.w; }
        }
        float e4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < plen; t += 16) {
            const float4* p4 = reinterpret_cast<const float4*>(sm + AMB_TS_PADI(16 * gq + t + 16));
            const float4* t4 = reinterpret_cast<const float4*>(tp + t);
            float tap[16];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 u = p4[j]; w[16 + 4 * j] = u.x; w[17 + 4 * j] = u.y; w[18 + 4 * j] = u.z; w[19 + 4 * j] = u.w;
                const float4 a = t4[j]; tap[4 * j] = a.x; tap[4 * j + 1] = a.y; tap[4 * j + 2] = a.z; tap[4 * j + 3] = a.w;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
#pragma unroll
                for (int q = 0; q < 16; q++) c[q] = fmaf(w[q + j], tap[j], c[q]);
            }
            // energy of the first offset's window (four chains); the other fifteen slide from it below
#pragma unroll
            for (int j = 0; j < 16; j++) e4[j & 3] = fmaf(w[j], w[j], e4[j & 3]);
#pragma unroll
            for (int j = 0; j < 16; j++) w[j] = w[16 + j];
        }
        float e = (e4[0] + e4[1]) + (e4[2] + e4[3]);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            if (q > 0) {
                const float out = sm[AMB_TS_PADI(16 * gq + q - 1)], in = sm[AMB_TS_PADI(16 * gq + q - 1 + plen)];
                e = e - out * out + in * in;
            }
            if (16 * gq + q < n) {
                const float d2 = e * tE;
                float r;
                if (d2 > 1e-6f * (1.f + 1e-3f)) r = c[q] / sqrtf(d2);
                else if (d2 < 1e-6f * (1.f - 1e-3f)) r = AMB_XC_FAIL;
                else r = AMB_XC_UNSURE;
                if (!(r == r)) r = AMB_XC_UNSURE;                           // NaN samples: let the exact evaluation decide
                v[16 * gq + q] = r;
            }
        }
    }
}
#endif

// Refine with float32 screening values - the scan kernel's store where it covers [lo, hi], otherwise computed here into vbuf
// (capacity vcap floats) - : offsets whose screening metric is within 1e-3 of the maximum are re-evaluated exactly
// (amb_xcorr_at), so the argmax and the reported metric are the reference's.  Falls back to the exact evaluation of every
// offset when the screening cannot decide.
AMB_HD double amb_fine_xcorr_store(const AmbTables* T, const AmbWalkGeom& G, const AmbSig& sig, int64_t lo, int64_t hi, int64_t p,
                                   double* scratch, int64_t* pos_out, float* vbuf = nullptr, int vcap = 0) {
    const int64_t j = (p + 511) >> 8;
    const int raw = G.xc ? amb_tile_slot_raw(G, j) : -1;
    const int slot = raw & AMB_TS_MASK;
    const int64_t base = (j << 8) - 511 - T->fine_radius;
    const int n = (int)(hi - lo + 1);
    const float* v;
    if (raw < 0 || lo < base || hi >= base + G.xc_w || hi < lo || !(raw & AMB_TS_CAND) || (raw & AMB_TS_NOXC)) {
        AMB_STAT(3);
        if (!vbuf || hi < lo || n > vcap) return amb_fine_xcorr(T, sig, lo, hi, p, scratch, pos_out);
#if defined(__CUDA_ARCH__)
        if (G.team_n > 1 && n <= G.team_vcap) {
            __syncthreads();                                            // (the team's readers of the previous screening row are done)
            if (G.team_stage && sig.xf.kind == AMB_XF_NONE) amb_xcorr_screen_team(T, sig, lo, n, G.team_vbuf, G.team_stage, G.team_n, G.team_rank);
            else amb_xcorr_screen(T, sig, lo, n, G.team_vbuf, G.team_n, G.team_rank);
            __syncthreads();
            v = G.team_vbuf;
        } else
#endif
        {
            amb_xcorr_screen(T, sig, lo, n, vbuf);
            v = vbuf;
        }
    } else {
        AMB_STAT(2);
        v = G.xc + (int64_t)slot * G.xc_w + (lo - base);
    }
    float mx = -3.0f;
    bool unsure = false;
    AMB_WARP_FOR(o, n) {
        const float x = v[o];
        if (x == AMB_XC_UNSURE) unsure = true;
        else if (x != AMB_XC_FAIL && x > mx) mx = x;
    }
    mx = amb_warp_maxf(mx);
    if (amb_warp_any(unsure)) return amb_fine_xcorr(T, sig, lo, hi, p, scratch, pos_out);
    if (mx < -2.5f) { *pos_out = p; return -INFINITY; }                 // no offset passes the denominator test
    // offsets within 1e-3 of the screening maximum, in increasing order (ballot compaction)
    int cnt = 0;
    for (int o0 = 0; o0 < n && cnt <= 32; o0 += AMB_NLANES) {
        const int o = o0 + AMB_LANE();
        bool take = false;
        if (o < n) { const float x = v[o]; take = x != AMB_XC_FAIL && x >= mx - 1e-3f; }
        const unsigned bal = amb_warp_ballot(take);
        if (take) {
            const int at = cnt + amb_popc(bal & amb_lanemask_lt());
            if (at < 32) scratch[32 + at] = (double)(lo + o);
        }
        cnt += amb_popc(bal);
    }
    AMB_WARP_SYNC();
    if (cnt > 32) return amb_fine_xcorr(T, sig, lo, hi, p, scratch, pos_out);
    if (cnt == 1 && mx > 0.102f) {
        // a single offset within 1e-3 of the screening maximum is the reference's argmax, and its exact metric clears the 0.1
        // test (the screening values are within 5e-4 of the exact ones).  Nothing else in the walk depends on the value, so
        // its 2 x sym_len-step float64 summation is left to amb_resolve_fine, which runs one frame per lane afterwards.
        *pos_out = (int64_t)scratch[32];
        AMB_WARP_SYNC();
        return -1.0 - (double)mx;
    }
    AMB_STAT_ADD(8, cnt);
    AMB_WARP_FOR(i, cnt) {
        double m;
        scratch[i] = amb_xcorr_at(T, sig, (int64_t)scratch[32 + i], &m) ? m : -INFINITY;
    }
    AMB_WARP_SYNC();
    double best = -INFINITY;
    int64_t bpos = p;
    for (int i = 0; i < cnt; i++) if (scratch[i] > best) { best = scratch[i]; bpos = (int64_t)scratch[32 + i]; }
    AMB_WARP_SYNC();
    *pos_out = bpos;
    return best;
}

// One evaluated batch of the candidate logic of _scanForPreamble (app.js:806-838): positions d .. d+cnt-1 with their
// gates W->gate[] and metrics W->M[] (one look-ahead entry at index cnt when the batch does not end at scan_end);
// running maximum (best, best_pos).  Returns true when the batch commits a preamble (state updated).
//   reference, per position:  if gate && m > 0.5 && m > best: best = m          (here: must exceed best by 1e-12)
//                             if best > 0.5 && gate[next] && M[next] < 0.7 best: commit
// amb_scan_batch_serial is that loop as written.  amb_scan_batch_reduced walks the batch in chunks of one position per
// lane: a prefix-maximum scan gives the running best at every position, the commit test is then independent per
// position and a ballot finds the first one.  The running best equals the plain prefix maximum unless some update is
// suppressed by the 1e-12 rule (a near-tie, e.g. a noise-free plateau); such batches take the serial loop.
AMB_HD bool amb_scan_batch_serial(AmbStreamState& st, const AmbWalkScratch* W, int64_t d, int cnt, int64_t scan_end, double& best, int64_t& best_pos) {
    for (int i = 0; i < cnt; i++) {
        const int64_t pos = d + i;
        if (W->gate[i]) {
            const double m = W->M[i];
            if (m > 0.5 && m > best * (1.0 + 1e-12)) { best = m; best_pos = pos; }
        }
        const int nxt = pos < scan_end ? i + 1 : i;
        if (best > 0.5 && best_pos >= 0 && W->gate[nxt] && W->M[nxt] < best * 0.7) {
            amb_commit(st, best_pos, best, pos + 1, pos < scan_end ? 0 : 1);
            return true;
        }
    }
    return false;
}

#define AMB_BAND 1e-10               // threshold decisions inside this relative band around equality are settled by the exact replay
#define AMB_BAND_TIE 1e-12           // ... and comparisons between two metrics of one call inside this one (see the header comment)
#ifndef AMB_REPLAY_MAX
#define AMB_REPLAY_MAX 65536         // slides the replay may have to redo before the call in question
#endif

AMB_HD bool amb_scan_batch_reduced(AmbStreamState& st, const AmbWalkScratch* W, int64_t d, int cnt, int64_t scan_end, double& best, int64_t& best_pos,
                                   int* ambig = nullptr);

AMB_HD bool amb_scan_batch(AmbStreamState& st, const AmbWalkScratch* W, int64_t d, int cnt, int64_t scan_end, double& best, int64_t& best_pos,
                           int* ambig = nullptr) {
#if defined(AMB_EMU_CHECK) && !defined(__CUDA_ARCH__)
    // test build: the reduction form must agree with the serial loop on every batch
    AmbStreamState s2 = st;
    double b2 = best;
    int64_t p2 = best_pos;
    const bool r2 = amb_scan_batch_serial(s2, W, d, cnt, scan_end, b2, p2);
    const bool r1 = amb_scan_batch_reduced(st, W, d, cnt, scan_end, best, best_pos, ambig);
    if (r1 != r2 || b2 != best || p2 != best_pos || s2.pre_pos != st.pre_pos || s2.coarse != st.coarse || s2.state != st.state ||
        s2.ac_scan_pos != st.ac_scan_pos || s2.lag != st.lag) {
        fprintf(stderr, "amb_scan_batch mismatch: d=%lld cnt=%d r=%d/%d best=%.17g/%.17g pos=%lld/%lld\n", (long long)d, cnt, (int)r1, (int)r2, best, b2,
                (long long)best_pos, (long long)p2);
        abort();
    }
    return r1;
#else
    return amb_scan_batch_reduced(st, W, d, cnt, scan_end, best, best_pos, ambig);
#endif
}

AMB_HD bool amb_scan_batch_reduced(AmbStreamState& st, const AmbWalkScratch* W, int64_t d, int cnt, int64_t scan_end, double& best, int64_t& best_pos,
                                   int* ambig) {
    const bool tracking = best > 0.5 && best_pos >= 0;
    if (!tracking && best > 0.5) return amb_scan_batch_serial(st, W, d, cnt, scan_end, best, best_pos);      // not a state the walker produces
    double carry = tracking ? best : 0.0;          // running maximum before the current chunk (0: none yet)
    int64_t carry_pos = best_pos;
    for (int c0 = 0; c0 < cnt; c0 += AMB_NLANES) {
        const int i = c0 + AMB_LANE();
        const bool in = i < cnt;
        const double v = (in && W->gate[i] && W->M[i] > 0.5) ? W->M[i] : -1.0;
        const double pm = amb_warp_scan_max(v);
        const double prev = fmax(carry, amb_warp_prev(pm, -1.0));       // running maximum before position i
        const double r = fmax(carry, pm);                               // ... and at position i
        const bool amb = v > prev && !(v > prev * (1.0 + 1e-12));       // an update the 1e-12 rule would suppress
        bool trig = false, band = v > 0.0 && prev > 0.5 && fabs(v - prev) <= AMB_BAND_TIE * prev;  // a candidate within the band of the running best
        if (in && r > 0.5) {
            const int nxt = d + i < scan_end ? i + 1 : i;
            trig = W->gate[nxt] && W->M[nxt] < r * 0.7;
            if (W->gate[nxt] && fabs(W->M[nxt] - r * 0.7) <= AMB_BAND * r) band = true;               // the 0.7 test within the band
        }
        const unsigned tb = amb_warp_ballot(trig), ab = amb_warp_ballot(amb);
        const int u = tb ? amb_ffs(tb) - 1 : AMB_NLANES;                // lane of the first commit in this chunk
        if (ambig) { const unsigned bb = amb_warp_ballot(band); if (bb & (u < AMB_NLANES ? ((2u << u) - 1u) : 0xffffffffu)) *ambig = 1; }
        if (ab & (u < AMB_NLANES ? ((2u << u) - 1u) : 0xffffffffu)) return amb_scan_batch_serial(st, W, d, cnt, scan_end, best, best_pos);
        const double top = amb_warp_bcast(u < AMB_NLANES ? r : pm, u < AMB_NLANES ? u : AMB_NLANES - 1);
        if (top > carry) {
            // the running maximum (at the commit, or at the end of the chunk) was first attained inside this chunk
            const unsigned eb = amb_warp_ballot(v == top);
            carry = top; carry_pos = d + c0 + amb_ffs(eb) - 1;
        }
        if (u < AMB_NLANES) {
            best = carry; best_pos = carry_pos;
            amb_commit(st, best_pos, best, d + c0 + u + 1, d + c0 + u < scan_end ? 0 : 1);
            return true;
        }
    }
    if (carry > 0.5) { best = carry; best_pos = carry_pos; }
    return false;
}

// true when x lies within the relative band around y
AMB_HD bool amb_near(double x, double y) { return fabs(x - y) <= AMB_BAND * fabs(y); }
// an energy pair whose 0.001 gate could come out differently in the reference
AMB_HD bool amb_gate_near(double a, double b) {
    return (amb_near(a, 0.001) && b > 0.001 * (1.0 - AMB_BAND)) || (amb_near(b, 0.001) && a > 0.001 * (1.0 - AMB_BAND));
}

// One call of _scanForPreamble exactly as the reference runs it (app.js:775-847): its running sums are re-created by direct
// summation at st.init_pos (app.js:791-803) and slid position by position - every position except the last one of a call,
// where the reference only advances acScanPos (app.js:826) - up to the call's first position, then the loop of
// app.js:806-846 runs with strict comparisons.  The update terms of 256 positions at a time (mid*bIn - aOut*mid, ...) do not
// depend on the sums: the lanes compute them into the scratch arrays, and one lane carries the three chains of additions
// and the candidate logic.  The quotient p^2 / (ra rb) is only formed where a comparison needs it (a product comparison with
// a 1e-7 margin settles the rest).  st is the state at call entry (ring clamp and re-initialisation applied).  Returns
// false, leaving st alone, when the replay would be longer than AMB_REPLAY_MAX slides.
// eval_from: first position of the call at which the fast path evaluated a window; before it the tile summaries prove (with
// margins far above the band) that no window can be a candidate, so the replay only slides there.
AMB_HD bool amb_walk_scan_exact(const AmbWalkGeom& G, const AmbSig& sig, AmbStreamState& st, int64_t tw, AmbWalkScratch* W, int64_t eval_from) {
    const int64_t scan_end = tw - 2 * AMB_HALF;
    const int64_t start = st.ac_scan_pos;
    if (!st.ac_init || start < st.init_pos || start - st.init_pos > AMB_REPLAY_MAX || start > scan_end) return false;
    if (eval_from < start) eval_from = start;
    if (eval_from > scan_end + 1) eval_from = scan_end + 1;
    // (the call that re-initialised the sums scanned from init_pos to its own end in one go, across any block boundaries the
    // receiver had spent collecting a frame: only the call ends from that one on are skipped)
    const int64_t first_end = ((int64_t)st.init_call + 1) * G.block_len - 2 * AMB_HALF;
    double p = 0.0, ra = 0.0, rb = 0.0, best = 0.0;
    int64_t best_pos = -1, end_pos = scan_end + 1;
    int committed = 0;
    if (AMB_LEADER()) {
        for (int m = 0; m < AMB_HALF; m++) {
            const double a = (double)amb_sig(sig, st.init_pos + m), b = (double)amb_sig(sig, st.init_pos + m + AMB_HALF);
            p = amb_add(p, amb_mul(a, b)); ra = amb_add(ra, amb_mul(a, a)); rb = amb_add(rb, amb_mul(b, b));
        }
    }
    // next position whose slide is skipped (the last position of a call, from the re-initialising call on)
    int64_t next_skip = first_end;
    // ---- phase 1: positions [init_pos, eval_from): slides only
    for (int64_t pos = st.init_pos; pos < eval_from; pos += 256) {
        const int n = (int)(eval_from - pos < 256 ? eval_from - pos : 256);
        AMB_WARP_FOR(i, n) {
            const int64_t L = pos + i;
            const double ao = (double)amb_sig(sig, L), mid = (double)amb_sig(sig, L + AMB_HALF), bi = (double)amb_sig(sig, L + 2 * AMB_HALF);
            W->M[i] = amb_sub(amb_mul(mid, bi), amb_mul(ao, mid));
            W->A[i] = amb_sub(amb_mul(mid, mid), amb_mul(ao, ao));
            W->B[i] = amb_sub(amb_mul(bi, bi), amb_mul(mid, mid));
        }
        AMB_WARP_SYNC();
        if (AMB_LEADER()) {
            // at most one skipped position per chunk (calls are longer than a chunk)
            int k = next_skip >= pos && next_skip < pos + n ? (int)(next_skip - pos) : n;
            int i = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
            for (; i < k; i++) { p = amb_add(p, W->M[i]); ra = amb_add(ra, W->A[i]); rb = amb_add(rb, W->B[i]); }
            if (k < n) {
                next_skip += G.block_len;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
                for (i = k + 1; i < n; i++) { p = amb_add(p, W->M[i]); ra = amb_add(ra, W->A[i]); rb = amb_add(rb, W->B[i]); }
            }
        }
        AMB_WARP_SYNC();
    }
    while (next_skip < eval_from) next_skip += G.block_len;       // (lanes other than the leader did not advance it)
    // ---- phase 2: positions [eval_from, scan_end]: the loop of app.js:806-846
    for (int64_t pos = eval_from; pos <= scan_end && !committed; pos += 256) {
        const int n = (int)(scan_end + 1 - pos < 256 ? scan_end + 1 - pos : 256);
        AMB_WARP_FOR(i, n) {
            const int64_t L = pos + i;
            const double ao = (double)amb_sig(sig, L), mid = (double)amb_sig(sig, L + AMB_HALF), bi = (double)amb_sig(sig, L + 2 * AMB_HALF);
            W->M[i] = amb_sub(amb_mul(mid, bi), amb_mul(ao, mid));
            W->A[i] = amb_sub(amb_mul(mid, mid), amb_mul(ao, ao));
            W->B[i] = amb_sub(amb_mul(bi, bi), amb_mul(mid, mid));
        }
        AMB_WARP_SYNC();
        if (AMB_LEADER()) {
            for (int i = 0; i < n; i++) {
                const int64_t L = pos + i;
                if (ra > 0.001 && rb > 0.001) {
                    const double num = amb_mul(p, p), den = amb_mul(ra, rb);
                    if (num > 0.4999999 * den && (!(best > 0.5) || num > best * 0.9999999 * den)) {
                        const double m = amb_div(num, den);
                        if (m > 0.5 && m > best) { best = m; best_pos = L; }
                    }
                }
                if (L < scan_end) { p = amb_add(p, W->M[i]); ra = amb_add(ra, W->A[i]); rb = amb_add(rb, W->B[i]); }      // (scan_end: the call's last position, no slide)
                if (best > 0.5 && best_pos >= 0 && ra > 0.001 && rb > 0.001) {
                    const double num = amb_mul(p, p), den = amb_mul(ra, rb), lim = amb_mul(best, 0.7);
                    bool drop;
                    if (num < lim * 0.9999999 * den) drop = true;
                    else if (num > lim * 1.0000001 * den) drop = false;
                    else drop = amb_div(num, den) < lim;
                    if (drop) { committed = 1; end_pos = L + 1; break; }
                }
            }
        }
        AMB_WARP_SYNC();
        committed = amb_warp_bcast((double)committed, 0) != 0.0;
    }
    best = amb_warp_bcast(best, 0);
    best_pos = (int64_t)amb_warp_bcast((double)best_pos, 0);
    end_pos = (int64_t)amb_warp_bcast((double)end_pos, 0);
    // the defect bookkeeping of the fast path (the lag branch of amb_walk_scan_fast): label `start` still described window start - 1
    if (st.lag == 1 && start < scan_end) {
        const double xm1 = (double)amb_sig(sig, start - 1), x255 = (double)amb_sig(sig, start + 255), x511 = (double)amb_sig(sig, start + 511);
        st.Ep += xm1 * x255 - x255 * x511; st.Ea += xm1 * xm1 - x255 * x255; st.Eb += x255 * x255 - x511 * x511;
        st.lag = 0;
    }
    if (committed) { amb_commit(st, best_pos, best, end_pos, end_pos - 1 < scan_end ? 0 : 1); return true; }
    st.ac_scan_pos = scan_end + 1;
    st.lag = 1;
    if (best > 0.5 && best_pos >= 0) amb_commit(st, best_pos, best, scan_end + 1, 1);
    return true;
}

AMB_HD void amb_walk_scan_fast(const AmbTables* T, const AmbWalkGeom& G, const AmbSig& sig, const float* sums,
                               AmbStreamState& st, int64_t tw, AmbWalkScratch* W, int* ambig, int64_t* eval_from);

// _scanForPreamble, app.js:775-847, for the call whose totalWritten is `tw`
AMB_HD void amb_walk_scan(const AmbTables* T, const AmbWalkGeom& G, const AmbSig& sig, const float* sums,
                          AmbStreamState& st, int64_t tw, AmbWalkScratch* W) {
    const int64_t oldest = tw - G.ring_cap;
    if (st.ac_scan_pos < oldest + 2 * AMB_HALF) {
        st.ac_scan_pos = oldest + 2 * AMB_HALF > 0 ? oldest + 2 * AMB_HALF : 0;
        st.ac_init = 0;
    }
    const int64_t scan_end = tw - 2 * AMB_HALF;
    if (st.ac_scan_pos > scan_end) return;
    if (!st.ac_init) {
        if (st.ac_scan_pos < oldest) return;
        st.Ep = st.Ea = st.Eb = 0.0; st.lag = 0; st.ac_init = 1; st.init_call = (int32_t)(tw / G.block_len - 1); st.init_pos = st.ac_scan_pos;
    }
    // (the fields a scan call can change: saved so that the exact replay can start from the same state)
    const int64_t s_pos = st.ac_scan_pos, s_pre = st.pre_pos;
    const double s_Ep = st.Ep, s_Ea = st.Ea, s_Eb = st.Eb, s_coarse = st.coarse;
    const int s_lag = st.lag, s_state = st.state;
    int ambig = 0;
    int64_t eval_from = (int64_t)1 << 62;
    amb_walk_scan_fast(T, G, sig, sums, st, tw, W, &ambig, &eval_from);
    if (amb_warp_any(ambig != 0) && s_pos >= st.init_pos && s_pos - st.init_pos <= AMB_REPLAY_MAX) {
        AMB_STAT(13);
#if !defined(__CUDA_ARCH__) && defined(AMB_EMU_CHECK)
        const AmbStreamState fast = st;
#endif
        st.ac_scan_pos = s_pos; st.pre_pos = s_pre; st.Ep = s_Ep; st.Ea = s_Ea; st.Eb = s_Eb; st.coarse = s_coarse; st.lag = s_lag; st.state = s_state;
        amb_walk_scan_exact(G, sig, st, tw, W, eval_from);
#if !defined(__CUDA_ARCH__) && defined(AMB_EMU_CHECK)
        if (getenv("EMU_DEBUG"))
            fprintf(stderr, "exact call tw=%lld start=%lld init_pos=%lld lag=%d | fast: state %d pre %lld coarse %.17g scan %lld lag %d | exact: state %d pre %lld coarse %.17g scan %lld lag %d\n",
                    (long long)tw, (long long)s_pos, (long long)st.init_pos, s_lag, fast.state, (long long)fast.pre_pos, fast.coarse, (long long)fast.ac_scan_pos, fast.lag,
                    st.state, (long long)st.pre_pos, st.coarse, (long long)st.ac_scan_pos, st.lag);
#endif
    }
}

// the call with directly evaluated window sums (state at entry: ring clamp and re-initialisation applied); *ambig is set when
// one of its decisions falls inside the band
AMB_HD void amb_walk_scan_fast(const AmbTables* T, const AmbWalkGeom& G, const AmbSig& sig, const float* sums,
                               AmbStreamState& st, int64_t tw, AmbWalkScratch* W, int* ambig, int64_t* eval_from) {
    const int64_t scan_end = tw - 2 * AMB_HALF;
    double best = 0.0;
    int64_t best_pos = -1;
    int64_t d = st.ac_scan_pos;

    if (st.lag == 1) {
        // first position of a call that continues an uninterrupted scan: sums still describe window d-1.  When the
        // summary of that window's tile proves that no window in it passes the candidate test under the current defect,
        // nothing but the defect update below can happen and the sums need not be evaluated.
        bool cold = false;
        {
            const int64_t jj = (d - 1 + 511) >> 8, rel = jj - (G.sum_base >> 8);
            if (d - 1 >= 0 && rel >= 0 && rel < G.sum_tiles && (jj << 8) >= G.sum_lo) cold = !amb_tile_hot2(G, sums, jj, rel, st.Ep, st.Ea, st.Eb);
        }
        if (!cold) {
            if (d < *eval_from) *eval_from = d;
            amb_eval_windows(G, sig, d - 1, 3, W->M, W->A, W->B);
            double m;
            if (amb_metric(W->M[0], W->A[0], W->B[0], st, &m) && m > 0.5 && m > best) { best = m; best_pos = d; }
            if (amb_gate_near(W->A[0] + st.Ea, W->B[0] + st.Eb) || amb_near(m, 0.5)) *ambig = 1;
        }
        if (d < scan_end) {
            // slide with the label d: removes T(d), adds T(d+256)  ->  W(d+1) + [T(d-1) - T(d+255)]
            const double xm1 = (double)amb_sig(sig, d - 1), x255 = (double)amb_sig(sig, d + 255), x511 = (double)amb_sig(sig, d + 511);
            st.Ep += xm1 * x255 - x255 * x511;
            st.Ea += xm1 * xm1 - x255 * x255;
            st.Eb += x255 * x255 - x511 * x511;
            st.lag = 0;
            if (best > 0.5) {
                double cur;
                const int gate_cur = amb_metric(W->M[2], W->A[2], W->B[2], st, &cur);
                if (amb_gate_near(W->A[2] + st.Ea, W->B[2] + st.Eb) || (gate_cur && amb_near(cur, best * 0.7))) *ambig = 1;
                if (gate_cur && cur < best * 0.7) { amb_commit(st, best_pos, best, d + 1, 0); return; }
            }
            d = d + 1;
        } else {
            // cannot happen for block_len > 1 (the next call's scan_end is at least block_len later)
            st.ac_scan_pos = d + 1;
            if (best > 0.5) amb_commit(st, best_pos, best, d + 1, 1);
            return;
        }
    }

    while (d <= scan_end) {
        if (!(best > 0.5)) {
            // skip tiles that cannot hold a candidate.  tile j covers window-end samples [256j, 256j+256)
            const int64_t q = d + 511;
            int64_t j = q >> 8;
            const int64_t j_end = (scan_end + 511) >> 8;
            bool found = false;
            while (j <= j_end) {
                AMB_WARP_FOR(t, 32) {
                    const int64_t jj = j + t;
                    int h = 0;
                    if (jj <= j_end) {
                        const int64_t rel = jj - (G.sum_base >> 8);
                        if (rel < 0 || rel >= G.sum_tiles || (jj << 8) < G.sum_lo) h = 1;   // no summary: must evaluate
                        else h = amb_tile_hot2(G, sums, jj, rel, st.Ep, st.Ea, st.Eb) ? 1 : 0;
                    }
                    W->hot[t] = h;
                }
                AMB_WARP_SYNC();
                int first = -1;
                for (int t = 0; t < 32; t++) if (W->hot[t]) { first = t; break; }
                AMB_WARP_SYNC();
                if (first >= 0) { j += first; found = true; break; }
                j += 32;
            }
            if (!found) break;
            const int64_t tile_first = (j << 8) - 511;
            if (tile_first > d) d = tile_first;
            if (d > scan_end) break;
        }
        // evaluate up to 256 windows exactly (to the end of the tile when idle, any 256 when tracking)
        int64_t d_end = d + 255;
        if (!(best > 0.5)) { const int64_t tile_last = (((d + 511) >> 8) << 8) + 255 - 511; if (tile_last < d_end) d_end = tile_last; }
        if (d_end > scan_end) d_end = scan_end;
        const int cnt = (int)(d_end - d + 1);
        const int look = d_end < scan_end ? 1 : 0;
        if (d < *eval_from) *eval_from = d;
        amb_eval_windows(G, sig, d, cnt + look, W->M, W->A, W->B);
        if (!(best > 0.5)) {
            // idle: unless some window can pass metric > 0.5 (p^2 > 0.5 a b, tested with a margin far above the rounding of the
            // reference's quotient) the batch changes nothing - no division, no candidate logic
            bool maybe = false;
            AMB_WARP_FOR(i, cnt + look) {
                const double a = W->A[i] + st.Ea, b = W->B[i] + st.Eb, pp = W->M[i] + st.Ep;
                if (a > 0.001 * (1.0 - AMB_BAND) && b > 0.001 * (1.0 - AMB_BAND) && pp * pp > 0.5 * a * b * (1.0 - 4.0 * AMB_BAND)) maybe = true;
            }
            if (!amb_warp_any(maybe)) { AMB_WARP_SYNC(); d = d_end + 1; continue; }
        }
        bool near_thr = false;
        AMB_WARP_FOR(i, cnt + look) {
            double m;
            const double a = W->A[i] + st.Ea, b = W->B[i] + st.Eb;
            W->gate[i] = amb_metric(W->M[i], W->A[i], W->B[i], st, &m);
            W->M[i] = m;
            if (amb_gate_near(a, b) || (W->gate[i] && amb_near(m, 0.5))) near_thr = true;
        }
        if (amb_warp_any(near_thr)) *ambig = 1;
        AMB_WARP_SYNC();
        if (amb_scan_batch(st, W, d, cnt, scan_end, best, best_pos, ambig)) { AMB_WARP_SYNC(); return; }
        AMB_WARP_SYNC();
        d = d_end + 1;
    }
    // end of the data available in this call (app.js:842-846)
    st.ac_scan_pos = scan_end + 1;
    st.lag = 1;
    if (best > 0.5 && best_pos >= 0) amb_commit(st, best_pos, best, scan_end + 1, 1);
}

// _refineAndCollect, app.js:849-898
AMB_HD void amb_walk_refine(const AmbTables* T, const AmbWalkGeom& G, const AmbSig& sig, AmbStreamState& st, int64_t tw,
                            AmbWalkScratch* W) {
    const int64_t plen = T->sym_len, R = T->fine_radius;
    if (tw < st.pre_pos + plen + R) return;
    int64_t fs = st.pre_pos - R, fe = st.pre_pos + R;
    if (fs < tw - G.ring_cap) fs = tw - G.ring_cap;
    if (fe > tw - plen) fe = tw - plen;
    int64_t pos;
    double best;
    const AmbRefineMemo* mi = G.memo_in;
    if (mi && mi->valid && mi->coarse == st.pre_pos && mi->fs == fs && mi->fe == fe) { pos = mi->pos; best = mi->fine; AMB_STAT(4); }
    else {
        // (the window-sum arrays of the scratch are free during a refine: they hold the screening row of a direct refine)
        AMB_STAT_T0();
        best = amb_fine_xcorr_store(T, G, sig, fs, fe, st.pre_pos, W->x, &pos, reinterpret_cast<float*>(W->M), (int)(3 * 264 * sizeof(double) / sizeof(float)));
        AMB_STAT_T1(10);
    }
    if (best < 0.1 && !AMB_FINE_DEFERRED(best)) { st.state = AMB_RX_IDLE; st.ac_init = 0; st.false_pos++; AMB_STAT(7); return; }
    if (G.memo_out && AMB_LEADER()) { AmbRefineMemo* mo = G.memo_out; mo->coarse = st.pre_pos; mo->fs = fs; mo->fe = fe; mo->pos = pos; mo->fine = best; mo->valid = 1; }
    st.pre_pos = pos; st.fine = best;
    const int64_t max_payload = st.meta_received ? (int64_t)(st.chunk_size ? st.chunk_size : 4096) + 11 : 280;
    st.frame_end = st.pre_pos + amb_frame_samples(T, max_payload);
    st.state = AMB_RX_COLLECT;
}

// Runs calls [st.next_call, k_end) of one stream.  Emits one record per completed frame.
// Returns the number of records written (stops early, leaving st.next_call behind k_end, if `cap` or `stop_after` is hit,
// or - with stop_on_fp - right after a failed refine).
AMB_HD int amb_walk(const AmbTables* T, const AmbWalkGeom& G, const AmbSig& sig, const float* sums, AmbStreamState& st,
                    int64_t k_end, int stream, AmbFrameRec* out, int cap, AmbWalkScratch* W, int stop_after = 0x7fffffff, int stop_on_fp = 0) {
    int n = 0;
    AmbColdCache cc;
    cc.base = -1; cc.hot = 0u; cc.init_call = -1;
    while (st.next_call < k_end) {
        { AMB_STAT_T0(); const int64_t sk = amb_walk_skip_cold_calls(G, sums, st, k_end, W, cc); AMB_STAT_T1(11); if (sk > 0) continue; }
        const int64_t tw = (st.next_call + 1) * G.block_len;
        if (st.state == AMB_RX_IDLE) { AMB_STAT(6); AMB_STAT_T0(); amb_walk_scan(T, G, sig, sums, st, tw, W); AMB_STAT_T1(12); }
        else if (st.state == AMB_RX_PREAMBLE) {
            amb_walk_refine(T, G, sig, st, tw, W);
            // a failed refine (app.js:879-884) re-initialises the scan in the next call, like a frame end does: with stop_on_fp the
            // link ends here too, so that the caller can screen the idle calls after it under the new defect (k_cold_calls)
            if (stop_on_fp && st.state == AMB_RX_IDLE) { st.next_call++; return n; }
        }
        else if (tw >= st.frame_end) {
            // _checkFrameComplete + _demodulateFrame + _resetToIdle, app.js:900-981
            if (n >= cap) return n;
            AmbFrameRec rec;
            rec.stream = stream;
            rec.status = st.pre_pos < tw - G.ring_cap ? AMB_ST_RING_OVERWRITTEN : AMB_ST_OK;
            rec.pos = st.pre_pos; rec.len = st.frame_end - st.pre_pos; rec.coarse = st.coarse; rec.fine = st.fine;
            st.ac_scan_pos = st.frame_end; st.ac_init = 0; st.pre_pos = -1; st.frame_end = -1; st.state = AMB_RX_IDLE;
            st.next_call++;
            rec.snap = st;
            if (AMB_LEADER()) out[n] = rec;
            n++;
            if (n >= stop_after) return n;
            continue;
        }
        st.next_call++;
    }
    return n;
}

// Applies the decoded frames of one stream in time order (app.js:937-965 and the storage-free part of
// ChunkAssembler, app.js:610-643).  `cur` is the stream state after the walk (the walker never touches the
// flags or the frame counters).  A CRC-valid metadata frame that changes the collection window (first
// metadata, or a different chunkSize) invalidates everything the walker did after it: the state is rewound to
// the snapshot taken after that frame, with the new flags, and 1 is returned so that the caller walks again.
// keep[i] = 1 for frames to report.
// ChunkAssembler.handleDataChunk's bitmap (app.js:628-643) for one CRC-valid chunk with seq < totalChunks: 1 = newly
// received, 0 = duplicate.  A negative seqNum indexes the typed array out of range, which reads undefined and writes
// nothing in JS: it counts as received every time.  Chunks beyond the bitmap capacity are not tracked (always new).
AMB_HD int amb_bitmap_mark(uint32_t* bitmap, int64_t bitmap_bits, int32_t seq) {
    if (seq < 0 || !bitmap || seq >= bitmap_bits) return 1;
    const uint32_t bit = 1u << (seq & 31);
#if defined(__CUDA_ARCH__)
    const uint32_t old = atomicOr(bitmap + (seq >> 5), bit);
#else
    const uint32_t old = bitmap[seq >> 5];
    bitmap[seq >> 5] = old | bit;
#endif
    return (old & bit) ? 0 : 1;
}

// store_off (optional, [cnt]): where a newly received chunk goes in the stream's file buffer (seqNum * chunkSize, the
// ChunkAssembler.assembleFile offset, app.js:676-677), -1 = not stored; file_reset (optional): set when a metadata frame
// cleared the chunk store (app.js:619-625).
AMB_HD int amb_post_frames(const AmbFrameRec* recs, int cnt, const amb_frame_result* res, const AmbStreamState& cur,
                           AmbStreamState* out, int* keep, uint32_t* bitmap = nullptr, int64_t bitmap_bits = 0,
                           int64_t* store_off = nullptr, int* file_reset = nullptr) {
    AmbStreamState st = cur;
    int64_t fd = cur.frames_decoded, fe = cur.frame_errors, rc = cur.received, ce = cur.crc_errors, du = cur.dups;
    int meta = cur.meta_received, chunk = cur.chunk_size, total = cur.total_chunks, fsize = cur.total_file_size;
    int rewind = -1;
    if (store_off) for (int i = 0; i < cnt; i++) store_off[i] = -1;
    for (int i = 0; i < cnt; i++) {
        const amb_frame_result& r = res[i];
        keep[i] = 1;
        if (r.status != 0) { fe++; continue; }
        fd++;
        if (r.frame_type == AMB_FRAME_META) {
            if (r.crc_valid) {
                const bool changes = !meta || chunk != r.chunk_size;
                meta = 1; chunk = r.chunk_size; total = r.total_chunks; fsize = r.total_file_size; rc = 0; ce = 0;
                if (file_reset) *file_reset = 1;
                if (store_off) for (int q = 0; q < i; q++) store_off[q] = -1;       // chunks stored before the reset are gone
                // handleMetadataFrame (app.js:610-618): a fresh bitmap of ceil(totalChunks / 8) bytes
                if (bitmap) {
                    int64_t words = ((int64_t)total + 31) / 32;
                    if (words > (bitmap_bits + 31) / 32) words = (bitmap_bits + 31) / 32;
                    for (int64_t w = 0; w < words; w++) bitmap[w] = 0u;
                }
                if (changes) { rewind = i; break; }
            } else fe++;
        } else if (r.frame_type == AMB_FRAME_DATA) {
            if (meta == 1 && !(r.seq_num >= total)) {
                if (!r.crc_valid) ce++;
                else if (amb_bitmap_mark(bitmap, bitmap_bits, r.seq_num)) { rc++; if (store_off && r.seq_num >= 0) store_off[i] = (int64_t)r.seq_num * chunk; }
                else du++;
            }
        }
    }
    if (rewind >= 0) {
        for (int i = rewind + 1; i < cnt; i++) keep[i] = 0;
        st = recs[rewind].snap;
        st.dc_mean = cur.dc_mean; st.written = cur.written;
    }
    st.meta_received = meta; st.chunk_size = chunk; st.total_chunks = total; st.total_file_size = fsize;
    st.frames_decoded = fd; st.frame_errors = fe; st.received = rc; st.crc_errors = ce; st.dups = du;
    *out = st;
    return rewind >= 0 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Link-parallel resolution of one stream (speculate, then verify with the exact walker).
//
// After every frame the receiver re-initialises its scan at the frame end, so a stream is a chain of LINKS: a
// link starts in a post-reset state (scan position e, next call k, flags) and ends with the next frame's reset.
// A link's outcome is a deterministic function of its start state.  Candidate RUNS (maximal groups of consecutive
// tiles holding a window with metric > 0.5) are where links can end, and they are known from the scan kernel.
//   phase A: every run is walked from a synthetic start just before it (one warp each, all in parallel);
//   chain  : next[r] = first run still ahead of run r's frame end; the runs reachable from the current state are
//            the speculated chain;
//   phase B: every chain link is walked again, now from its predecessor's phase-A end state, with the same exact
//            walker; link i is exact if links < i reproduced their phase-A end states.  At the first link whose
//            phase-B end state differs from the speculation, the chain is rebuilt from that (exact) state.
// The helpers below are the pieces shared by the CUDA kernels and the host emulation.
// ------------------------------------------------------------------------------------------------
struct AmbRun { int64_t first_pos, last_pos; };      // window positions covered by the run's tiles

// synthetic start of phase A: scan re-initialised one tile before the run, in the call that first sees that position
AMB_HD AmbStreamState amb_run_start_state(const AmbStreamState& flags_from, const AmbRun& r, const AmbWalkGeom& G) {
    AmbStreamState s = flags_from;
    int64_t p = r.first_pos - 256;
    if (p < 0) p = 0;
    s.state = AMB_RX_IDLE; s.ac_init = 0; s.lag = 0; s.ac_scan_pos = p; s.pre_pos = -1; s.frame_end = -1;
    s.Ep = s.Ea = s.Eb = 0.0;
    s.next_call = (p + 511) / G.block_len;
    return s;
}

// the fields of a stream state that a walk never writes (flags, frame counters, the carried mean) plus the running count of
// refine failures, taken over from `from`: what a walk started in `from` would hand on
AMB_HD void amb_carry_unwalked(AmbStreamState& dst, const AmbStreamState& from) {
    dst.meta_received = from.meta_received; dst.chunk_size = from.chunk_size; dst.total_chunks = from.total_chunks; dst.total_file_size = from.total_file_size;
    dst.frames_decoded = from.frames_decoded; dst.frame_errors = from.frame_errors; dst.false_pos = from.false_pos; dst.received = from.received;
    dst.crc_errors = from.crc_errors; dst.dups = from.dups; dst.dc_mean = from.dc_mean; dst.written = from.written;
}

// first run with windows at or after position e (runs are sorted by position); n if none
AMB_HD int amb_run_at_or_after(const AmbRun* runs, int n, int64_t e) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (runs[mid].last_pos >= e) hi = mid; else lo = mid + 1; }
    return lo;
}

// do two post-link states describe the same receiver? (everything the next link depends on)
AMB_HD bool amb_same_link_state(const AmbStreamState& a, const AmbStreamState& b) {
    return a.state == b.state && a.ac_init == b.ac_init && a.ac_scan_pos == b.ac_scan_pos && a.next_call == b.next_call &&
           a.pre_pos == b.pre_pos && a.frame_end == b.frame_end && a.meta_received == b.meta_received && a.chunk_size == b.chunk_size &&
           (a.ac_init == 0 || (a.lag == b.lag && a.init_call == b.init_call && a.init_pos == b.init_pos && a.Ep == b.Ep && a.Ea == b.Ea && a.Eb == b.Eb));
}
