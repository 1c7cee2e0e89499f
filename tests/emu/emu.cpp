// TEST-ONLY host emulation of the device logic.
//
// Compiles audio-modem_b200/csrc/amb_device.cuh and amb_walker.cuh with g++ (warp loops become sequential
// loops) and drives them the way the CUDA kernels do, so that the kernel LOGIC (faithful FFT schedule,
// equaliser, demap slicing, vote/pack/parse/CRC, the streaming walker and its rewind rule) can be compared
// with the oracle on a machine without a GPU.  It is not part of the product and is never loaded by it; the
// parallel decomposition itself (thread indexing, shared memory, scans) is only exercised by the -m gpu tests.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../audio-modem_b200/csrc/amb_walker.cuh"
#include "../../audio-modem_b200/csrc/amb_tx.cuh"

namespace {

struct Frame { amb_frame_result res; std::vector<uint8_t> bytes; };

// mirrors k_demod (+ k_frame_max) and k_parse for one job
void decode_job(const AmbTables& T, const AmbSig& sig_in, int64_t pos, int64_t len, int64_t ce_pos, int64_t data_pos, int nsym,
                int status, int mode, bool scale, int64_t result_pos, Frame* out) {
    amb_frame_result& r = out->res;
    amb_result_clear(&r);
    if (status != 0) { r.status = status; if (status == AMB_ST_RING_OVERWRITTEN) r.preamble_idx = result_pos; return; }
    AmbSig sig = sig_in;
    if (scale) {
        float mx = 0.f;
        for (int64_t i = 0; i < len; i++) { const int64_t g = pos + i; const float v = (g >= sig.lo && g < sig.hi) ? sig.p[g - sig.base] : 0.f; mx = fmaxf(mx, fabsf(v)); }
        sig.xf.kind = AMB_XF_SCALE; sig.xf.mx = mx;
    }
    std::vector<double> re(AMB_FFT), im(AMB_FFT), hre(AMB_FFT, 0.0), him(AMB_FFT, 0.0), qs(2 * AMB_MAX_PILOTS);
    amb_fft512_load(sig, ce_pos + T.cp_len, re.data(), im.data());
    amb_fft512_faithful(re.data(), im.data(), T.tw_re, T.tw_im, T.sub_start, T.sub_end);
    amb_channel_from_spectrum(re.data(), im.data(), T.known_re, nullptr, T.sub_start, T.sub_end, hre.data(), him.data());
    std::vector<uint8_t> idx((size_t)(nsym > 0 ? nsym : 1) * T.n_data);
    for (int s = 0; s < nsym; s++) {
        amb_fft512_load(sig, data_pos + (int64_t)s * T.sym_len + T.cp_len, re.data(), im.data());
        amb_fft512_faithful(re.data(), im.data(), T.tw_re, T.tw_im, T.sub_start, T.sub_end);
        amb_symbol_demap(&T, re.data(), im.data(), hre.data(), him.data(), qs.data(), idx.data() + (size_t)s * T.n_data, nullptr, nullptr);
    }
    const int64_t n_elems = (int64_t)nsym * T.n_data;
    const int64_t nb = amb_num_bytes(n_elems, T.bps, T.rep);
    out->bytes.assign((size_t)nb + 8, 0);
    amb_pack_bytes(idx.data(), n_elems, T.bps, T.rep, out->bytes.data(), nb);
    if (mode == 0) amb_parse_chunk_frame(T.crc_table, out->bytes.data(), nb, &r);
    else amb_parse_received(T.crc_table, out->bytes.data(), nb, &r);
    r.n_bytes = (int32_t)nb;
    const bool chunk_type = nb >= 10 && (out->bytes[0] == AMB_FRAME_META || out->bytes[0] == AMB_FRAME_DATA);
    if (mode == 0 || r.status == 0 || chunk_type) r.preamble_idx = result_pos;
    out->bytes.resize((size_t)nb);
}

}  // namespace

extern "C" {

int emu_result_size(void) { return (int)sizeof(amb_frame_result); }

// randomised agreement check of the reduction form of the candidate logic against the serial loop (amb_walker.cuh);
// AMB_EMU_CHECK makes amb_scan_batch abort on a mismatch.  Returns the number of batches that committed.
int emu_scan_batch_fuzz(uint64_t seed, int n_cases) {
    AmbWalkScratch W;
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    auto uni = [&]() { return (double)(rnd() >> 11) / 9007199254740992.0; };
    int commits = 0;
    for (int c = 0; c < n_cases; c++) {
        const int cnt = 1 + (int)(rnd() % 256);
        const int look = (int)(rnd() % 4) != 0;
        const int shape = (int)(rnd() % 5);
        const double levels[6] = {0.3, 0.5, 0.62, 0.9, 0.9 * (1.0 + 5e-13), 0.9 * (1.0 - 3e-13)};
        for (int i = 0; i <= cnt; i++) {
            double m;
            if (shape == 0) m = uni();
            else if (shape == 1) m = levels[rnd() % 6];
            else if (shape == 2) { const double t = (double)i / cnt; m = (t < 0.4 ? t * 2.4 : t < 0.7 ? 0.96 + 0.02 * uni() : 0.96 - (t - 0.7) * 3.0); }
            else if (shape == 3) m = 0.45 + 0.5 * uni();
            else m = (i % 7 == 0) ? 0.9 : levels[rnd() % 6] * (1.0 + 1e-12 * (double)(rnd() % 3));
            W.M[i] = m;
            W.gate[i] = (rnd() % 16) != 0;
            if (!W.gate[i]) W.M[i] = 0.0;
        }
        AmbStreamState st;
        memset(&st, 0, sizeof(st));
        double best = 0.0;
        int64_t best_pos = -1;
        if (rnd() % 3 == 0) { best = 0.5 + 0.5 * uni(); best_pos = 1000; if (rnd() % 2) best = levels[3 + rnd() % 3]; }
        const int64_t d = 5000;
        const int64_t scan_end = look ? d + cnt + 10 : d + cnt - 1;
        commits += amb_scan_batch(st, &W, d, cnt, scan_end, best, best_pos) ? 1 : 0;
    }
    return commits;
}

// the chunked CRC-32 of block_crc32 (amb_kernels.cuh), threads run one after the other
uint32_t emu_crc32_chunked(const uint8_t* data, int64_t n, int nt) {
    // crc = XOR_t c_t * x^(8 L (nt - 1 - t)), grouped per 32 chunks as block_crc32 does (nt a multiple of 32, or < 32)
    AmbTables T;
    amb_tables_build(&T, 0, 1, 1);
    const int64_t L = (n + nt - 1) / nt;
    const int per = nt < 32 ? nt : 32, groups = nt / per;
    uint32_t total = 0;
    for (int w = 0; w < groups; w++) {
        uint32_t v = 0;
        for (int l = 0; l < per; l++) {
            int64_t lo, hi;
            amb_crc_chunk(n, nt, w * per + l, &lo, &hi);
            const uint32_t c = hi > lo ? amb_crc32_bytes(T.crc_table, data + lo, hi - lo) : 0u;
            v ^= amb_gf_mul(c, amb_gf_xpow8((uint32_t)(L * (per - 1 - l))));
        }
        total ^= amb_gf_mul(v, amb_gf_xpow8((uint32_t)(L * per * (groups - 1 - w))));
    }
    return total;
}

// decodeChunkFrame through the device logic
int emu_decode_chunk_frame(int cfg, int mod, int rep, const float* x, int64_t n, amb_frame_result* res, uint8_t* bytes, int64_t cap) {
    AmbTables T;
    if (amb_tables_build(&T, cfg, mod, rep)) return -1;
    AmbSig sig; sig.p = x; sig.base = 0; sig.lo = 0; sig.hi = n; sig.xf.kind = AMB_XF_NONE; sig.xf.mx = 0; sig.xf.mean = 0;
    int status = 0, nsym = 0;
    const int64_t ce = 2 * (int64_t)T.sym_len, data = 3 * (int64_t)T.sym_len;
    if (ce + T.sym_len > n) status = AMB_ST_FRAME_TOO_SHORT_FOR_CE;
    else if (data >= n) status = AMB_ST_NO_DATA_AFTER_CE;
    else nsym = (int)((n - data) / T.sym_len);
    Frame f;
    decode_job(T, sig, 0, n, ce, data, nsym, status, 0, false, -1, &f);
    *res = f.res;
    if (bytes) std::memcpy(bytes, f.bytes.data(), (size_t)std::min<int64_t>(cap, (int64_t)f.bytes.size()));
    return 0;
}

// streaming receiver through the device logic: EMA (sequential, as the reference), tile summaries computed
// directly, then walk / decode / post until no stream needs another walk.  Pushes of `push_len` samples.
static int g_use_stores = 1, g_use_links = 1, g_stop_fp = 1;      // (g_stop_fp: links also end at a failed refine, as the CUDA path does on long pushes)
void emu_stop_fp(int on) { g_stop_fp = on; }
static int64_t g_link_iters = 0, g_link_walks = 0, g_link_pushes = 0;
void emu_use_stores(int on) { g_use_stores = on; }
void emu_use_links(int on) { g_use_links = on; }
void emu_link_stats(int64_t* out) { out[0] = g_link_iters; out[1] = g_link_walks; out[2] = g_link_pushes; g_link_iters = g_link_walks = g_link_pushes = 0; }

// Link-parallel resolution of one stream for one push (sequential stand-in for k_link_* / k_chain): returns the
// number of frame records written, leaves the end-of-push state in st.
static int resolve_links(const AmbTables& T, const AmbWalkGeom& G, const AmbSig& sig, const float* sums, AmbStreamState& st, int64_t k_end,
                         const std::vector<AmbRun>& runs, std::vector<AmbFrameRec>& out, AmbWalkScratch* W) {
    const int n = (int)runs.size();
    g_link_pushes++;
    // phase A
    std::vector<AmbFrameRec> recA((size_t)n);
    std::vector<AmbStreamState> endA((size_t)n);
    std::vector<int> hasA((size_t)n, 0);
    for (int r = 0; r < n; r++) {
        AmbStreamState s = amb_run_start_state(st, runs[r], G);
        hasA[r] = amb_walk(&T, G, sig, sums, s, k_end, 0, &recA[r], 1, W, 1);
        endA[r] = s; g_link_walks++;
    }
    int n_out = 0;
    AmbStreamState cur = st;
    // first link from the carried state (always exact)
    {
        AmbFrameRec rec;
        const int got = amb_walk(&T, G, sig, sums, cur, k_end, 0, &rec, 1, W, 1, g_stop_fp);
        g_link_walks++;
        if (!got && cur.next_call >= k_end) { st = cur; return 0; }
        if (got) out[n_out++] = rec;
    }
    for (int iter = 0; iter < 1 << 20; iter++) {
        g_link_iters++;
        // chain from cur
        std::vector<int> chain;
        if (cur.state == AMB_RX_IDLE && !cur.ac_init) {
            int r = amb_run_at_or_after(runs.data(), n, cur.ac_scan_pos);
            while (r < n) {
                chain.push_back(r);
                if (!hasA[r]) break;
                const int nx = amb_run_at_or_after(runs.data(), n, endA[r].ac_scan_pos);
                r = nx > r ? nx : n;
            }
        }
        // links: run links + a tail link when the last run link is expected to emit a frame (or when there is no run)
        const bool tail = chain.empty() || hasA[chain.back()];
        const int m = (int)chain.size() + (tail ? 1 : 0);
        std::vector<AmbStreamState> in((size_t)m), outB((size_t)m);
        std::vector<AmbFrameRec> recB((size_t)m);
        std::vector<int> hasB((size_t)m);
        for (int i = 0; i < m; i++) in[i] = i == 0 ? cur : endA[chain[i - 1]];
        for (int i = 0; i < m; i++) {          // phase B: independent walks (parallel on the GPU)
            AmbStreamState s = in[i];
            hasB[i] = amb_walk(&T, G, sig, sums, s, k_end, 0, &recB[i], 1, W, 1, g_stop_fp);
            outB[i] = s; g_link_walks++;
        }
        bool redo = false;
        for (int i = 0; i < m; i++) {
            if (hasB[i]) out[n_out++] = recB[i];
            if (!hasB[i] && outB[i].next_call < k_end) { cur = outB[i]; redo = true; break; }      // ended at a failed refine (stop_on_fp)
            if (!hasB[i]) { st = outB[i]; if (getenv("EMU_DEBUG")) fprintf(stderr, "resolve: meta %d runs %d frames %d iterations %d\n", st.meta_received, n, n_out, iter + 1); return n_out; }
            const bool is_run = i < (int)chain.size();
            const bool ok = is_run && hasA[chain[i]] == hasB[i] && amb_same_link_state(outB[i], endA[chain[i]]);
            if (!ok && getenv("EMU_DEBUG")) {
                const AmbStreamState& x = outB[i]; 
                if (is_run) { const AmbStreamState& y = endA[chain[i]];
                  fprintf(stderr, "mismatch link %d/%d run %d hasA %d hasB %d | B: st %d init %d scan %lld call %lld  | A: st %d init %d scan %lld call %lld | recB pos %lld recA pos %lld in.scan %lld in.call %lld runfirst %lld\n", i, m, chain[i], hasA[chain[i]], hasB[i],
                        x.state, x.ac_init, (long long)x.ac_scan_pos, (long long)x.next_call, y.state, y.ac_init, (long long)y.ac_scan_pos, (long long)y.next_call,
                        (long long)recB[i].pos, (long long)recA[chain[i]].pos, (long long)in[i].ac_scan_pos, (long long)in[i].next_call, (long long)runs[chain[i]].first_pos); }
                else fprintf(stderr, "tail link emitted a frame at %lld\n", (long long)recB[i].pos);
            }
            if (!ok) { cur = outB[i]; redo = true; break; }
        }
        if (!redo) { st = outB[m - 1]; return n_out; }                       // cannot happen: the last link ends the push
    }
    return n_out;
}

int emu_stream(int cfg, int mod, int rep, const float* x, int64_t n, int64_t block_len, int64_t push_len,
               amb_frame_result* res_out, int max_res, uint8_t* bytes_out, int64_t bytes_stride, int64_t* counters) {
    AmbTables T;
    if (amb_tables_build(&T, cfg, mod, rep)) return -1;
    std::vector<float> cleaned((size_t)n);
    double m = 0;
    for (int64_t i = 0; i < n; i++) { m = T.dc_alpha * m + T.dc_beta * (double)x[i]; cleaned[i] = (float)((double)x[i] - m); }
    const int64_t n_tiles = n / 256 + 1;
    std::vector<float> sums((size_t)n_tiles * AMB_SUM_W, 0.f);
    AmbSig sig; sig.p = cleaned.data(); sig.base = 0; sig.lo = 0; sig.xf.kind = AMB_XF_NONE; sig.xf.mx = 0; sig.xf.mean = 0;
    AmbWalkGeom G;
    G.block_len = block_len; G.ring_cap = amb_tables_estimate_frame_samples(cfg, mod, 4096 + 16, rep) * 3 + 8192;
    G.sum_base = 0; G.sum_tiles = n_tiles; G.sum_lo = 0;
    G.team_n = 1; G.team_rank = 0; G.team_vbuf = nullptr; G.team_vcap = 0; G.team_stage = nullptr;
    // stores as the scan kernels write them (k_scan_sums, k_defect_prefix, k_xcorr_store): slots only for the tiles of the
    // current push, float64 triples, float32 screening correlation, per-call defect prefix
    std::vector<int32_t> tile_slot((size_t)n_tiles, -1);
    std::vector<double> tri, Fpre((size_t)(n / block_len + 2) * 3, 0.0);
    std::vector<float> tsum;
    std::vector<uint32_t> bitmap(1 << 16, 0u);        // ChunkAssembler.receivedBitmap (2 Mi chunks)
    std::vector<float> xc;
    const int xc_w = 256 + 2 * T.fine_radius;
    G.tile_slot = nullptr; G.tri = nullptr; G.xc = nullptr; G.xc_w = xc_w; G.slot_info = nullptr; G.tsum = nullptr; G.memo_out = nullptr; G.memo_in = nullptr; G.F = nullptr; G.cold_tag = nullptr; G.F_call0 = 0; G.F_calls = (int64_t)Fpre.size() / 3;
    AmbStreamState st;
    std::memset(&st, 0, sizeof(st));
    st.pre_pos = -1; st.frame_end = -1;
    AmbWalkScratch W;
    std::vector<AmbFrameRec> recs(4096);
    int n_out = 0;
    int64_t tiles_done = 0;
    for (int64_t w0 = 0; w0 < n; w0 += push_len) {
        const int64_t w1 = std::min(n, w0 + push_len);
        if ((w1 - w0) % block_len) break;
        sig.hi = w1;
        if (g_use_stores) { std::fill(tile_slot.begin(), tile_slot.end(), -1); tri.clear(); xc.clear(); tsum.clear(); }
        int n_slots = 0;
        // summaries of the window-end samples that became available
        for (int64_t j = tiles_done; (j + 1) * 256 <= w1; j++, tiles_done = j) {
            float ap = 0, mnA = 3e38f, mxA = 0, mnB = 3e38f, mxB = 0;
            bool hot = false, cand = false;
            double mmax = 0.0, xmnA = 1.0, xmnB = 1.0;
            bool excluded = false;
            std::vector<double> t3(256 * 3, 0.0);
            for (int64_t q = j * 256; q < (j + 1) * 256; q++) {
                const int64_t d = q - 511;
                if (d < 0) continue;
                double P = 0, A = 0, B = 0;
                for (int t = 0; t < 256; t++) { const double a = cleaned[d + t], b = cleaned[d + t + 256]; P += a * b; A += a * a; B += b * b; }
                t3[q - j * 256] = P; t3[256 + (q - j * 256)] = A; t3[512 + (q - j * 256)] = B;
                hot |= A > 0.0009 && B > 0.0009 && P * P > 0.25 * A * B;
                cand |= A > 0.001 && B > 0.001 && P * P > 0.5 * A * B;
                if (A > 0.0005 && B > 0.0005) { xmnA = fmin(xmnA, A); xmnB = fmin(xmnB, B); mmax = fmax(mmax, (P * P) / (A * B)); }
                else excluded = true;
                ap = fmaxf(ap, (float)fabs(P)); mnA = fminf(mnA, (float)A); mxA = fmaxf(mxA, (float)A); mnB = fminf(mnB, (float)B); mxB = fmaxf(mxB, (float)B);
            }
            float* o = &sums[(size_t)j * AMB_SUM_W];
            o[0] = ap; o[1] = mnA; o[2] = mxA; o[3] = mnB; o[4] = mxB;
            // slots as k_front assigns them: every tile the walker's summary test could enter (widened for the defect)
            {
                const float pw = o[0] * 1.05f;
                hot = hot || (o[2] * 1.05f > 0.0009f && o[4] * 1.05f > 0.0009f && pw * pw > 0.4f * fmaxf(o[1] * 0.95f, 0.0009f) * fmaxf(o[3] * 0.95f, 0.0009f));
            }
            if (g_use_stores && hot) {
                tile_slot[(size_t)j] = (n_slots++) | AMB_TS_ROWS | (cand ? AMB_TS_CAND : 0);
                tsum.push_back((float)mmax * (1.f + 1e-6f) + 1e-30f); tsum.push_back(fmaxf((float)xmnA * (1.f - 1e-6f), 0.f));
                tsum.push_back(fmaxf((float)xmnB * (1.f - 1e-6f), 0.f)); tsum.push_back(excluded ? 1.f : 0.f);
                tri.insert(tri.end(), t3.begin(), t3.end());
                xc.resize((size_t)n_slots * xc_w, AMB_XC_UNSURE);
                if (cand) {
                    AmbSig s2 = sig; s2.hi = w1;
                    const int64_t base = j * 256 - 511 - T.fine_radius;
                    for (int i = 0; i < xc_w; i++) {
                        double mm;
                        xc[(size_t)(n_slots - 1) * xc_w + i] = amb_xcorr_at(&T, s2, base + i, &mm) ? (float)mm : AMB_XC_FAIL;
                    }
                }
            }
        }
        if (g_use_stores) {
            // defect prefix over the calls completed by this push (k_scan_sums thread 0 + k_defect_prefix)
            for (int64_t k = w0 / block_len; k < w1 / block_len; k++) {
                auto c = [&](int64_t g) { return g >= 0 ? (double)cleaned[g] : 0.0; };
                const double xm1 = c(k * block_len - 512), x255 = c(k * block_len - 256), x511 = c(k * block_len);
                const double D[3] = {xm1 * x255 - x255 * x511, xm1 * xm1 - x255 * x255, x255 * x255 - x511 * x511};
                for (int q = 0; q < 3; q++) Fpre[(size_t)k * 3 + q] = (k > 0 ? Fpre[(size_t)(k - 1) * 3 + q] : 0.0) + D[q];
            }
            G.tile_slot = tile_slot.data(); G.tri = tri.data(); G.xc = xc.data(); G.F = Fpre.data(); G.tsum = tsum.data();
        }
        const int64_t k_end = w1 / block_len;
        for (int pass = 0; pass < 4096; pass++) {
            int cnt;
            if (g_use_links && g_use_stores && !st.meta_received) {
                // no metadata yet: one frame per pass with the sequential walker (k_walk, pre_meta_only)
                cnt = amb_walk(&T, G, sig, sums.data(), st, k_end, 0, recs.data(), (int)recs.size(), &W, 1);
            } else if (g_use_links && g_use_stores) {
                // candidate runs of this push (k_scan_sums marks the tiles): consecutive tiles holding a candidate window
                std::vector<AmbRun> runs;
                for (int64_t j = 0; j < n_tiles; j++) {
                    const int sl = tile_slot[(size_t)j];
                    const bool c = sl >= 0 && (sl & AMB_TS_CAND);
                    if (!c) continue;
                    const int64_t first = j * 256 - 511, last = first + 255;
                    if (!runs.empty() && runs.back().last_pos + 1 == first) runs.back().last_pos = last;
                    else runs.push_back({first, last});
                }
                cnt = resolve_links(T, G, sig, sums.data(), st, k_end, runs, recs, &W);
            } else cnt = amb_walk(&T, G, sig, sums.data(), st, k_end, 0, recs.data(), (int)recs.size(), &W);
            if (cnt == 0) break;
            std::vector<Frame> fr((size_t)cnt);
            std::vector<amb_frame_result> rr((size_t)cnt);
            for (int i = 0; i < cnt; i++) {
                const AmbFrameRec& R = recs[i];
                int status = R.status, nsym = 0;
                if (status == 0) {
                    if (3 * (int64_t)T.sym_len > R.len) status = AMB_ST_FRAME_TOO_SHORT_FOR_CE;
                    else if (3 * (int64_t)T.sym_len >= R.len) status = AMB_ST_NO_DATA_AFTER_CE;
                    else nsym = (int)((R.len - 3 * (int64_t)T.sym_len) / T.sym_len);
                }
                decode_job(T, sig, R.pos, R.len, R.pos + 2 * T.sym_len, R.pos + 3 * T.sym_len, nsym, status, 0, true, R.pos, &fr[i]);
                fr[i].res.coarse_metric = R.coarse; fr[i].res.fine_metric = amb_resolve_fine(&T, sig, R.pos, R.fine);
                rr[i] = fr[i].res;
            }
            std::vector<int> keep((size_t)cnt, 0);
            AmbStreamState nst;
            const int redo = amb_post_frames(recs.data(), cnt, rr.data(), st, &nst, keep.data(), bitmap.data(), (int64_t)bitmap.size() * 32);
            st = nst;
            for (int i = 0; i < cnt; i++) {
                if (!keep[i]) continue;
                if (n_out < max_res) {
                    res_out[n_out] = rr[i];
                    if (bytes_out) std::memcpy(bytes_out + (size_t)n_out * bytes_stride, fr[i].bytes.data(), (size_t)std::min<int64_t>(bytes_stride, (int64_t)fr[i].bytes.size()));
                }
                n_out++;
            }
            if (!redo && !(g_use_links && g_use_stores && !st.meta_received && st.next_call < k_end)) break;
        }
    }
    if (counters) {
        counters[0] = st.frames_decoded; counters[1] = st.frame_errors; counters[2] = st.false_pos; counters[3] = st.received;
        counters[4] = st.crc_errors; counters[5] = st.dups; counters[6] = st.meta_received; counters[7] = st.ac_scan_pos;
    }
    return n_out;
}

// buildDataChunkFrame / buildMetadataFrame through the device logic (mirrors k_tx_chunk_payloads + k_tx_frames)
static int64_t tx_frame(const AmbTables& T, const uint8_t* pay, int pay_len, int first, float* out, int64_t cap) {
    const int64_t nsym = amb_tx_num_symbols(&T, pay_len), pre = amb_tx_silence_pre(&T, first), post = amb_tx_silence_post();
    const int64_t sym = T.sym_len, total = pre + (3 + nsym) * sym + post;
    if (!out || cap < total) return total;
    std::memset(out, 0, (size_t)total * 4);
    for (int i = 0; i < sym; i++) { out[pre + i] = T.pre1[i]; out[pre + sym + i] = T.pre2[i]; out[pre + 2 * sym + i] = T.ce_sym[i]; }
    std::vector<double> re(AMB_FFT), im(AMB_FFT);
    for (int64_t s = 0; s < nsym; s++) amb_tx_symbol(&T, pay, pay_len, s, re.data(), im.data(), out + pre + (3 + s) * sym);
    float mx = 0.f;
    for (int64_t i = 0; i < total; i++) mx = fmaxf(mx, fabsf(out[i]));
    if (mx > 0.f) { const double sc = 0.8 / (double)mx; for (int64_t i = 0; i < total; i++) out[i] = (float)((double)out[i] * sc); }
    return total;
}
int64_t emu_build_data_chunk_frame(int cfg, int mod, int rep, const uint8_t* data, int len, int32_t seq, uint64_t synth_seed, float* out, int64_t cap) {
    AmbTables T;
    if (amb_tables_build(&T, cfg, mod, rep)) return -1;
    std::vector<uint8_t> d((size_t)len), pay((size_t)len + 16);
    for (int i = 0; i < len; i++) d[i] = data ? data[i] : amb_synth_byte(synth_seed, seq, i);
    const int pl = amb_data_chunk_payload(T.crc_table, d.data(), len, seq, pay.data());
    return tx_frame(T, pay.data(), pl, 0, out, cap);
}
int64_t emu_build_metadata_frame(int cfg, int mod, int rep, int32_t tc, int32_t tfs, int32_t cs, const uint8_t* name, int name_len, float* out, int64_t cap) {
    AmbTables T;
    if (amb_tables_build(&T, cfg, mod, rep)) return -1;
    uint8_t pay[300];
    const int pl = amb_metadata_payload(T.crc_table, tc, tfs, cs, name, name_len, pay);
    return tx_frame(T, pay, pl, 1, out, cap);
}
double emu_synth_normal(uint64_t seed, int64_t g) { return amb_synth_normal(seed, g); }

// tables for inspection
int emu_tables(int cfg, int mod, int rep, float* pre1, double* known_re, double* tw_re, double* tw_im, double* pts, double* scalars) {
    AmbTables T;
    if (amb_tables_build(&T, cfg, mod, rep)) return -1;
    std::memcpy(pre1, T.pre1, sizeof(float) * T.sym_len);
    std::memcpy(known_re, T.known_re, sizeof(T.known_re));
    std::memcpy(tw_re, T.tw_re, sizeof(T.tw_re)); std::memcpy(tw_im, T.tw_im, sizeof(T.tw_im));
    std::memcpy(pts, T.pts, sizeof(T.pts));
    scalars[0] = T.pre1_energy; scalars[1] = T.qam_thr; scalars[2] = T.dc_beta; scalars[3] = T.n_data; scalars[4] = T.pre1_sign;
    return 0;
}

}  // extern "C"
