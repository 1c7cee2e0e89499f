// C-ABI of the B200-native OFDM receive path (include/audio_modem_b200.h).
// Host side: owns the CUDA stream, device scratch and the streaming arenas; every entry point is a fixed
// sequence of kernel launches from amb_kernels.cuh.  There is no CPU compute path in this library.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "amb_kernels.cuh"
#include "amb_private.h"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) {
            cap = want;
            // AMB_POISON=1 (tests): fresh scratch memory starts as garbage, so a read of something no kernel wrote shows up
            static const bool poison = std::getenv("AMB_POISON") != nullptr;
            if (poison) { cudaMemset(p, 0xCD, want); cudaDeviceSynchronize(); }   // (the handle's stream does not order against the default stream)
        }
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct PendingCopy { void* host; const void* dev; size_t bytes; };
struct ProfSpan { const char* name; cudaEvent_t a, b; };

}  // namespace

struct amb_modem {
    int device = 0;
    cudaStream_t stream = nullptr;
    AmbTables tables;
    AmbTables* d_tables = nullptr;
    AmbEmaTables* d_ema = nullptr;
    AmbFastTables* d_fast = nullptr;
    unsigned long long* d_fallbacks = nullptr;
    uint32_t* d_crc_ops = nullptr;       // [AMB_CRC_MAXL][AMB_CRC_OPS] powers of x mod P, the CRC combine constants per chunk length L
    bool fast_demod = true;
    // AMB_FUSE_PARSE=1: single-group frames are demodulated, packed, parsed and checksummed by one CTA (k_demod_fast<.., true>).
    // Measured on C2 (16,384 frames): 1.04 ms fused against 0.81 + 0.14 ms for k_demod_fast + k_parse - the serial tail of a
    // 7-warp CTA hides worse than a separate 128-thread kernel does - so the two-kernel form stays the default.
    bool fuse_parse = false;
    std::string err;
    int64_t launches = 0;
    // optional per-kernel device timing (CUDA events on the handle's stream)
    bool prof = false;
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
    // scratch
    DevBuf in_a, in_b, in_c, in_lens, out_a, out_b, out_c, out_res, out_bytes;
    DevBuf jobs, idx, p_scratch, xf, psum, pmin, pmax, norm, metric, tile_max, tile_max32, amb, tie_rng, cidx, cbest, tx_pay, tx_out, tx_copy;
    std::vector<PendingCopy> pending;
    // streaming
    bool open = false;
    int walk_ctas = 148 * 16;         // grid of k_link_walk (set from the occupancy at create)
    int front_ctas_per_sm = 6;        // resident CTAs of k_front<.., 128> per SM (occupancy at create): its spans are sized for one wave
    bool push_pcm16 = false;          // input format of the push in progress
    int n_streams = 0;
    int64_t block_len = 0, max_push = 0, hist = 0, stride = 0, base_g = 0, written = 0, sum_stride = 0, ring_cap = 0;
    int rec_cap = 0;
    DevBuf raw, cleaned, sums, dc_next, states, recs, nrec, job_off, total, keep, keep_off, redo, n_redo, s_res, s_bytes, s_sel, g_res, g_bytes, ctmp;
    DevBuf nan_pos, tile_slot, tri, slot_info, slot_count, defect, fpre, xc, dpart, tsum, cand_list;
    DevBuf l_chunks, l_runs, l_nruns, l_runoff, l_recA, l_endA, l_hasA, l_next, l_linkrun, l_nlinks, l_outB, l_recB, l_hasB, l_vis, l_ja, l_jb, l_active, l_nactive;
    DevBuf l_first, l_firstflag, l_firststop, l_fpsum, l_memo, l_bmout, l_bmrec, l_bmhas, l_startA, xc_peak, xc_rows, cold_list, cold_tag, post_acc, bitmaps, files, file_reset, store_off;
    int64_t file_cap = 0;          // bytes per stream of the chunk store (0: assembly not enabled)
    int64_t bitmap_bits = 0;       // per-stream capacity of the chunk bitmap (ChunkAssembler.receivedBitmap)
    int* h_counts = nullptr;       // pinned host words for the per-push device->host counters
    cudaStream_t copy_stream = nullptr;          // host->device copies of sectioned pushes
    std::vector<cudaEvent_t> copy_events;
    int run_cap = 0;
    int64_t link_iterations = 0, link_passes = 0;
    int slot_cap = 0, xc_w = 0;
    int64_t calls_stride = 0;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            m->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return e__ == cudaErrorMemoryAllocation ? AMB_E_NOMEM : AMB_E_CUDA;                     \
        }                                                                                          \
    } while (0)

#define LAUNCH(NAME, ...)                                                                          \
    do {                                                                                           \
        prof_mark(m, NAME, 0);                                                                     \
        __VA_ARGS__;                                                                               \
        prof_mark(m, NAME, 1);                                                                     \
        LAUNCH_OK();                                                                               \
    } while (0)

#define LAUNCH_OK()                                                                                \
    do {                                                                                           \
        m->launches++;                                                                             \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) { m->err = std::string("kernel launch: ") + cudaGetErrorString(e__); return AMB_E_CUDA; } \
    } while (0)

namespace {

// start (which = 0) / end (which = 1) event of one kernel launch; no-op unless amb_profile(m, 1)
void prof_mark(amb_modem* m, const char* name, int which) {
    if (!m->prof) return;
    if (m->prof_used == m->prof_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        m->prof_pool.push_back(e);
    }
    cudaEvent_t e = m->prof_pool[m->prof_used++];
    cudaEventRecord(e, m->stream);
    if (which == 0) m->prof_spans.push_back({name, e, nullptr});
    else if (!m->prof_spans.empty()) m->prof_spans.back().b = e;
}

int fail(amb_modem* m, int code, const char* msg) { m->err = msg; return code; }

// device view of a caller input: copies host memory into `buf`, passes device pointers through
template <class T>
int in_ptr(amb_modem* m, const T* src, size_t count, int mem, DevBuf& buf, const T** out) {
    if (!src) { *out = nullptr; return AMB_OK; }
    if (mem == AMB_MEM_DEVICE) { *out = src; return AMB_OK; }
    CK(buf.ensure(count * sizeof(T)));
    CK(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, m->stream));
    *out = buf.as<T>();
    return AMB_OK;
}

// device view of a caller output: device pointers pass through; host pointers get a staging buffer whose
// content is copied back by flush()
template <class T>
int out_ptr(amb_modem* m, T* dst, size_t count, int mem, DevBuf& buf, T** out) {
    if (!dst) { *out = nullptr; return AMB_OK; }
    if (mem == AMB_MEM_DEVICE) { *out = dst; return AMB_OK; }
    CK(buf.ensure(count * sizeof(T)));
    *out = buf.as<T>();
    m->pending.push_back({dst, buf.p, count * sizeof(T)});
    return AMB_OK;
}

int flush(amb_modem* m) {
    for (auto& c : m->pending) CK(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, m->stream));
    m->pending.clear();
    CK(cudaStreamSynchronize(m->stream));
    return AMB_OK;
}

int check_common(amb_modem* m, int n, int64_t stride, int mem) {
    if (!m) return AMB_E_INVALID;
    m->pending.clear();
    if (n < 0 || stride < 0 || (mem != AMB_MEM_HOST && mem != AMB_MEM_DEVICE)) return fail(m, AMB_E_INVALID, "bad argument");
    if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, AMB_E_CUDA, "cudaSetDevice failed");
    return AMB_OK;
}

inline int blocks_for(int64_t n, int per) { return (int)((n + per - 1) / per); }

// demod + parse over m->jobs (n_jobs), results/bytes already device pointers
int run_demod_parse(amb_modem* m, int n_jobs, int max_nsym, bool frame_max, amb_frame_result* d_res, uint8_t* d_bytes, int64_t bytes_stride, bool two_pass = false) {
    if (n_jobs == 0) return AMB_OK;
    const AmbTables& T = m->tables;
    const int64_t idx_stride = ((int64_t)(max_nsym > 0 ? max_nsym : 1) * T.n_data + 15) / 16 * 16;      // rows 16-byte aligned for k_parse's vector loads
    const int groups = max_nsym > 0 ? (max_nsym + AMB_SYM_GROUP - 1) / AMB_SYM_GROUP : 1;
    // one CTA per frame does everything (demodulate, vote, pack, parse, CRC-32) when the frame is a single symbol group whose
    // constellation indices fit in shared memory; longer frames go through the index array in HBM and k_parse
    const bool fused = m->fast_demod && m->fuse_parse && groups == 1 && (int64_t)(max_nsym > 0 ? max_nsym : 1) * T.n_data <= AMB_FD_IDX_CAP;
    // a decoded frame longer than k_parse's shared-memory staging is packed straight into the caller's row when that holds it,
    // else into a scratch row: the result (fields, CRC verdict, n_bytes) never depends on the caller's row length
    uint8_t* p_scratch = nullptr;
    int64_t p_scratch_stride = 0;
    {
        const int64_t nb_max = amb_num_bytes((int64_t)(max_nsym > 0 ? max_nsym : 0) * T.n_data, T.bps, T.rep);
        if (!fused && nb_max > AMB_PARSE_SMEM && bytes_stride < nb_max) {
            p_scratch_stride = (nb_max + 15) / 16 * 16;
            CK(m->p_scratch.ensure((size_t)n_jobs * p_scratch_stride));
            p_scratch = m->p_scratch.as<uint8_t>();
        }
    }
    if (!fused) CK(m->idx.ensure((size_t)n_jobs * idx_stride));
    AmbDemodOut o;
    std::memset(&o, 0, sizeof(o));
    o.idx = fused ? nullptr : m->idx.as<uint8_t>(); o.idx_stride = idx_stride;
    AmbParseOut po;
    po.results = d_res; po.bytes_out = d_bytes; po.bytes_stride = bytes_stride; po.crc_ops = m->d_crc_ops;
    if (m->fast_demod) {
        // float32 symbol pipeline with margin guard and float64-faithful fallback; gathers the per-frame peak itself when one
        // CTA covers the whole frame
        if (frame_max && groups > 1) {
            const int parts = max_nsym * T.sym_len / 32768 + 1;          // a CTA per 32 K samples of the longest frame
            LAUNCH("k_frame_max", k_frame_max<<<n_jobs * parts, 256, 0, m->stream>>>(m->jobs.as<AmbJob>(), n_jobs, parts));
        }
        // warps per CTA (two symbols each per round): the count that leaves the fewest symbol slots empty for the longest frame
        auto warps_for = [](int nsym_max) {
            int nw = 8;
            const int ns_grp = nsym_max < AMB_SYM_GROUP ? (nsym_max > 0 ? nsym_max : 1) : AMB_SYM_GROUP;
            int best_cost = 1 << 30;
            // (+1: the CE symbol takes one half-warp slot of the first round)
            for (int w = 8; w >= 5; w--) { const int cost = ((ns_grp + 1 + 2 * w - 1) / (2 * w)) * w; if (cost < best_cost) { best_cost = cost; nw = w; } }
            return nw;
        };
        // instantiated per constellation and per CTA size class (7 warps or fewer leave registers for a fourth resident CTA)
#define AMB_DEMOD_LAUNCH(MODV, NTV, FUSEDV) \
    LAUNCH("k_demod_fast", (k_demod_fast<MODV, NTV, FUSEDV><<<dim3(n_jobs, l_groups), 32 * l_nw, AMB_DEMOD_FAST_SMEM(NTV, FUSEDV), m->stream>>>(m->d_tables, m->d_fast, m->jobs.as<AmbJob>(), n_jobs, o, m->d_fallbacks, po)))
#define AMB_DEMOD_LAUNCH_MOD(NTV, FUSEDV) \
    do { \
        if (T.mod == AMB_MOD_QAM16) { AMB_DEMOD_LAUNCH(AMB_MOD_QAM16, NTV, FUSEDV); } \
        else if (T.mod == AMB_MOD_QPSK) { AMB_DEMOD_LAUNCH(AMB_MOD_QPSK, NTV, FUSEDV); } \
        else { AMB_DEMOD_LAUNCH(AMB_MOD_BPSK, NTV, FUSEDV); } \
    } while (0)
        auto launch_demod = [&](int l_groups, int l_nw) -> int {
            if (l_nw <= 7) { if (fused) AMB_DEMOD_LAUNCH_MOD(224, true); else AMB_DEMOD_LAUNCH_MOD(224, false); }
            else { if (fused) AMB_DEMOD_LAUNCH_MOD(256, true); else AMB_DEMOD_LAUNCH_MOD(256, false); }
            return AMB_OK;
        };
#undef AMB_DEMOD_LAUNCH_MOD
#undef AMB_DEMOD_LAUNCH
        // recordings (decodeReceivedSignal): header symbols first, then only the symbols inside the frame's extent (k_extent)
        const int bits_per_sym = T.n_data * T.bps;
        const int hdr_syms = (int)(((int64_t)12 * 8 * T.rep - 1) / bits_per_sym) + 1;
        static const bool two_pass_on = !(std::getenv("AMB_TWO_PASS") && std::getenv("AMB_TWO_PASS")[0] == '0');
        if (two_pass && two_pass_on && !fused && hdr_syms <= AMB_FD_HDR_SYMS && max_nsym >= 3 * hdr_syms) {
            LAUNCH("k_demod_limit", k_demod_limit<<<blocks_for(n_jobs, 128), 128, 0, m->stream>>>(m->jobs.as<AmbJob>(), n_jobs, hdr_syms));
            int rc = launch_demod(1, warps_for(hdr_syms)); if (rc) return rc;
            LAUNCH("k_extent", k_extent<<<blocks_for(n_jobs, 128), 128, 0, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), n_jobs, m->idx.as<uint8_t>(), idx_stride, hdr_syms));
        }
        { int rc = launch_demod(groups, warps_for(max_nsym)); if (rc) return rc; }
    } else {
        if (frame_max) { LAUNCH("k_frame_max", k_frame_max<<<n_jobs, 256, 0, m->stream>>>(m->jobs.as<AmbJob>(), n_jobs, 1)); }
        LAUNCH("k_demod", k_demod<<<dim3(n_jobs, groups), 256, AMB_DEMOD_SMEM, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), n_jobs, o));
    }
    if (!fused) {
        LAUNCH("k_parse", k_parse<<<n_jobs < 65536 ? n_jobs : 65536, AMB_CRC_THREADS, 0, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), n_jobs, m->idx.as<uint8_t>(), idx_stride,
                                                              d_res, d_bytes, bytes_stride, m->d_crc_ops, p_scratch, p_scratch_stride));
    }
    return AMB_OK;
}

int64_t max_len_of(const int64_t* lens_host, int n, int64_t stride) {
    if (!lens_host) return stride;
    int64_t mx = 0;
    for (int i = 0; i < n; i++) mx = lens_host[i] > mx ? lens_host[i] : mx;
    return mx;
}

// preprocess statistics -> m->xf (AmbXform per row)
int run_stats(amb_modem* m, const float* d_x, int64_t stride, const int64_t* d_lens, int n, int64_t max_len) {
    const int parts = (int)((max_len + AMB_STAT_PART - 1) / AMB_STAT_PART) > 0 ? (int)((max_len + AMB_STAT_PART - 1) / AMB_STAT_PART) : 1;
    CK(m->psum.ensure((size_t)n * parts * 8)); CK(m->pmin.ensure((size_t)n * parts * 4)); CK(m->pmax.ensure((size_t)n * parts * 4));
    CK(m->xf.ensure((size_t)n * sizeof(AmbXform)));
    LAUNCH("k_stats_partial", k_stats_partial<<<n * parts, 256, 0, m->stream>>>(d_x, stride, d_lens, stride, parts, m->psum.as<double>(), m->pmin.as<float>(), m->pmax.as<float>()));
    LAUNCH("k_stats_final", k_stats_final<<<blocks_for(n, 128), 128, 0, m->stream>>>(d_lens, stride, parts, m->psum.as<double>(), m->pmin.as<float>(), m->pmax.as<float>(), n, m->xf.as<AmbXform>()));
    return AMB_OK;
}

int run_apply(amb_modem* m, const float* d_x, int64_t stride, const int64_t* d_lens, int n, int64_t max_len, float* d_out) {
    int per = (int)((max_len + 256 * 8 - 1) / (256 * 8));
    if (per < 1) per = 1;
    if (per > 4096) per = 4096;
    LAUNCH("k_apply_xform", k_apply_xform<<<n * per, 256, 0, m->stream>>>(d_x, stride, d_lens, stride, per, m->xf.as<AmbXform>(), d_out));
    return AMB_OK;
}

// Schmidl-Cox detector over rows -> m->cidx / m->cbest (device)
int run_detect(amb_modem* m, const float* d_y, int64_t stride, const int64_t* d_lens, int n, int64_t max_len) {
    const int tiles = (int)((max_len + AMB_TILE - 1) / AMB_TILE) > 0 ? (int)((max_len + AMB_TILE - 1) / AMB_TILE) : 1;
    CK(m->tile_max.ensure((size_t)n * tiles * 8));
    CK(m->amb.ensure((size_t)n * 4)); CK(m->tie_rng.ensure((size_t)n * 16)); CK(m->cidx.ensure((size_t)n * 8)); CK(m->cbest.ensure((size_t)n * 8));
    // float32 screening first (AMB_METRIC_SCREEN=0 switches it off: comparison runs), then the exact sums on the tiles it leaves
    static const bool screen = !(std::getenv("AMB_METRIC_SCREEN") && std::getenv("AMB_METRIC_SCREEN")[0] == '0');
    CK(m->tile_max32.ensure((size_t)n * tiles * 4));
    // AMB_METRIC_MERGE=0 (comparison runs): k_metric with a CTA per (recording, tile), then k_argmax
    static const bool merge = !(std::getenv("AMB_METRIC_MERGE") && std::getenv("AMB_METRIC_MERGE")[0] == '0');
    if (screen) LAUNCH("k_metric_screen", k_metric_screen<<<n * tiles, 256, 0, m->stream>>>(d_y, stride, d_lens, stride, 0.0099f, tiles, m->tile_max32.as<float>()));
    if (screen && merge) {
        LAUNCH("k_metric_argmax", k_metric_argmax<<<n, 256, AMB_SCAN_SMEM, m->stream>>>(d_y, stride, d_lens, stride, 0.01, m->tile_max32.as<float>(), m->tile_max.as<double>(), tiles,
                                                                                         m->cidx.as<int64_t>(), m->cbest.as<double>(), m->amb.as<int>(), m->tie_rng.as<int64_t>()));
    } else {
        LAUNCH("k_metric", k_metric<<<n * tiles, 256, AMB_SCAN_SMEM, m->stream>>>(d_y, stride, d_lens, stride, 0.01, tiles, m->tile_max.as<double>(),
                                                                                    screen ? m->tile_max32.as<float>() : nullptr));
        LAUNCH("k_argmax", k_argmax<<<n, 256, AMB_SCAN_SMEM, m->stream>>>(d_y, stride, d_lens, stride, 0.01, m->tile_max.as<double>(), tiles, m->cidx.as<int64_t>(),
                                          m->cbest.as<double>(), m->amb.as<int>(), m->tie_rng.as<int64_t>()));
    }
    LAUNCH("k_replay", k_replay<<<blocks_for(n, 4), 128, 0, m->stream>>>(d_y, stride, d_lens, stride, n, 0.01, m->amb.as<int>(), m->tie_rng.as<int64_t>(),
                                                                       m->cidx.as<int64_t>(), m->cbest.as<double>()));
    return AMB_OK;
}

// detectPreambleCrossCorr over rows -> d_idx / d_best (device)
int run_detect_xcorr(amb_modem* m, const float* d_x, int64_t stride, const int64_t* d_lens, int n, int64_t* d_idx, double* d_best) {
    const int64_t plen = m->tables.sym_len;
    const int64_t step = plen / 10 > 1 ? plen / 10 : 1;
    const int64_t max_pos = stride >= plen ? (stride - plen) / step + 1 : 1;
    CK(m->metric.ensure((size_t)n * max_pos * 8));
    const int per = blocks_for(max_pos, 128);
    LAUNCH("k_xcorr_coarse", k_xcorr_coarse<<<n * per, 128, 0, m->stream>>>(m->d_tables, d_x, stride, d_lens, stride, step, per, m->metric.as<double>(), max_pos));
    LAUNCH("k_xcorr_pick", k_xcorr_pick<<<blocks_for(n, 4), 128, 0, m->stream>>>(m->d_tables, d_x, stride, d_lens, stride, n, step, m->metric.as<double>(), max_pos, d_idx, d_best));
    return AMB_OK;
}

}  // namespace

extern "C" {

int amb_version(void) { return AMB_VERSION; }

const char* amb_status_string(int status) {
    switch (status) {
        case AMB_ST_OK: return "";
        case AMB_ST_PREAMBLE_NOT_DETECTED: return "Preamble not detected";
        case AMB_ST_PREAMBLE_LOW_CORRELATION: return "Preamble not detected (low correlation)";
        case AMB_ST_SIGNAL_TOO_SHORT_FOR_CE: return "Signal too short for CE";
        case AMB_ST_NO_DATA_AFTER_CE: return "No data after CE";
        case AMB_ST_DECODED_TOO_SHORT: return "Decoded data too short";
        case AMB_ST_DECODED_TOO_SHORT_FOR_HEADER: return "Decoded data too short for header";
        case AMB_ST_INVALID_DATA_LENGTH: return "Invalid data length: {arg}";
        case AMB_ST_FRAME_TOO_SHORT_FOR_CE: return "Frame too short for CE";
        case AMB_ST_UNKNOWN_FRAME_TYPE: return "Unknown frame type: 0x{arg:x}";
        case AMB_ST_META_TOO_SHORT: return "Metadata frame too short";
        case AMB_ST_META_TRUNCATED: return "Metadata frame truncated";
        case AMB_ST_CHUNK_TOO_SHORT: return "Data chunk frame too short";
        case AMB_ST_CHUNK_TRUNCATED: return "Data chunk truncated";
        case AMB_ST_RING_OVERWRITTEN: return "(ring buffer overwritten)";
    }
    return "?";
}

int amb_create(amb_modem** out, int device, int config, int constellation, int repetition) {
    if (!out) return AMB_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return AMB_E_NODEVICE;
    if (device < 0 || device >= count) return AMB_E_INVALID;
    if (constellation < 0 || constellation > 2) return AMB_E_INVALID;
    amb_modem* m = new amb_modem();
    m->device = device;
    if (amb_tables_build(&m->tables, config, constellation, repetition) != 0) { delete m; return AMB_E_INVALID; }
    auto bail = [&](int code) { amb_destroy(m); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(AMB_E_CUDA);
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(AMB_E_CUDA);
    if (cudaMalloc(&m->d_tables, sizeof(AmbTables)) != cudaSuccess) return bail(AMB_E_NOMEM);
    if (cudaMalloc(&m->d_ema, sizeof(AmbEmaTables)) != cudaSuccess) return bail(AMB_E_NOMEM);
    AmbEmaTables e;
    const long double a = (long double)m->tables.dc_alpha;
    e.a512 = (double)powl(a, 512.0L); e.alpha = m->tables.dc_alpha; e.beta = m->tables.dc_beta;
    for (int i = 0; i < 32; i++) e.lanepow[i] = (double)powl(a, 16.0L * i);
    for (int k = 0; k < 5; k++) e.steppow[k] = (double)powl(a, 16.0L * (1 << k));
    e.a8 = (double)powl(a, 8.0L);
    if (cudaMemcpy(m->d_tables, &m->tables, sizeof(AmbTables), cudaMemcpyHostToDevice) != cudaSuccess) return bail(AMB_E_CUDA);
    if (cudaMemcpy(m->d_ema, &e, sizeof(e), cudaMemcpyHostToDevice) != cudaSuccess) return bail(AMB_E_CUDA);
    {
        AmbFastTables ft;
        for (int k = 0; k < 256; k++) {
            ft.tw256[k] = make_float2((float)std::cos(-2.0 * M_PI * k / 256.0), (float)std::sin(-2.0 * M_PI * k / 256.0));
            ft.tw512[k] = make_float2((float)std::cos(-2.0 * M_PI * k / 512.0), (float)std::sin(-2.0 * M_PI * k / 512.0));
        }
        {
            const AmbTables& T = m->tables;
            for (int i = 0; i < AMB_FD_MAXBINS; i++) ft.role[i] = -1000;
            for (int q = 0; q < AMB_MAX_PILOTS; q++) {
                const int k = q < T.n_pilots ? T.pilots[q] : -1;
                const bool inb = k >= T.sub_start && k <= T.sub_end;
                ft.pilot_kk[q] = inb ? k - T.sub_start : -1;
                if (inb) ft.role[k - T.sub_start] = -1 - q;
            }
            for (int di = 0; di < T.n_data; di++) ft.role[T.data_k[di] - T.sub_start] = di;
        }
        if (cudaMalloc(&m->d_fast, sizeof(ft)) != cudaSuccess || cudaMalloc(&m->d_fallbacks, 32) != cudaSuccess) return bail(AMB_E_NOMEM);
        if (cudaMemcpy(m->d_fast, &ft, sizeof(ft), cudaMemcpyHostToDevice) != cudaSuccess || cudaMemset(m->d_fallbacks, 0, 32) != cudaSuccess) return bail(AMB_E_CUDA);
        std::vector<uint32_t> ops((size_t)AMB_CRC_MAXL * AMB_CRC_OPS, 0u);
        for (int L = 0; L < AMB_CRC_MAXL; L++) {
            for (int k = 0; k < 32; k++) ops[(size_t)L * AMB_CRC_OPS + k] = amb_gf_xpow8((uint32_t)(L * k));
            for (int k = 0; k < 4; k++) ops[(size_t)L * AMB_CRC_OPS + 32 + k] = amb_gf_xpow8((uint32_t)(L * 32 * k));
        }
        if (cudaMalloc(&m->d_crc_ops, ops.size() * 4) != cudaSuccess) return bail(AMB_E_NOMEM);
        if (cudaMemcpy(m->d_crc_ops, ops.data(), ops.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return bail(AMB_E_CUDA);
        const char* e = std::getenv("AMB_FAST_DEMOD");
        m->fast_demod = !(e && e[0] == '0');
        const char* fz = std::getenv("AMB_FUSE_PARSE");
        m->fuse_parse = fz && fz[0] == '1';
    }
#define AMB_DEMOD_ATTR(MODV) \
    cudaFuncSetAttribute(k_demod_fast<MODV, 224, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_DEMOD_FAST_SMEM(224, false)); \
    cudaFuncSetAttribute(k_demod_fast<MODV, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_DEMOD_FAST_SMEM(256, false)); \
    cudaFuncSetAttribute(k_demod_fast<MODV, 224, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_DEMOD_FAST_SMEM(224, true)); \
    cudaFuncSetAttribute(k_demod_fast<MODV, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_DEMOD_FAST_SMEM(256, true));
    AMB_DEMOD_ATTR(AMB_MOD_BPSK) AMB_DEMOD_ATTR(AMB_MOD_QPSK) AMB_DEMOD_ATTR(AMB_MOD_QAM16)
#undef AMB_DEMOD_ATTR
    if (std::getenv("AMB_DEBUG_OCC")) {
        int o1 = 0, o2 = 0, o3 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_demod_fast<AMB_MOD_QAM16, 224, false>, 224, AMB_DEMOD_FAST_SMEM(224, false));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_demod_fast<AMB_MOD_QAM16, 224, true>, 224, AMB_DEMOD_FAST_SMEM(224, true));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_front<float, 128>, 128, AMB_FRONT_SMEM(128));
        std::fprintf(stderr, "occupancy (CTAs/SM): demod 224 unfused %d, fused %d (dynamic smem %d / %d), front %d\n", o1, o2, (int)AMB_DEMOD_FAST_SMEM(224, false),
                     (int)AMB_DEMOD_FAST_SMEM(224, true), o3);
    }
    {
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_link_walk<false>, 128, 0) == cudaSuccess && occ > 0 &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
            m->walk_ctas = 4 * occ * sms;          // a few waves: the walks differ in length (measured flat from 4 to 16 waves, 7 % slower at one)
    }
    cudaFuncSetAttribute(k_front<float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_FRONT_SMEM(128));
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_front<float, 128>, 128, AMB_FRONT_SMEM(128)) == cudaSuccess && occ > 0) m->front_ctas_per_sm = occ;
    }
    cudaFuncSetAttribute(k_front<float, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_FRONT_SMEM(256));
    cudaFuncSetAttribute(k_front<int16_t, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_FRONT_SMEM(256));
    cudaFuncSetAttribute(k_front<float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_FRONT_SMEM(128));
    cudaFuncSetAttribute(k_front<int16_t, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_FRONT_SMEM(128));
    cudaFuncSetAttribute(k_chain_links, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_CHAIN_SMEM);
    cudaFuncSetAttribute(k_demod, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_DEMOD_SMEM);
    cudaFuncSetAttribute(k_metric, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_SCAN_SMEM);
    cudaFuncSetAttribute(k_argmax, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_SCAN_SMEM);
    cudaFuncSetAttribute(k_metric_argmax, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_SCAN_SMEM);
    cudaFuncSetAttribute(k_tx_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_TX_SMEM);
    cudaFuncSetAttribute(k_modulate, cudaFuncAttributeMaxDynamicSharedMemorySize, AMB_TX_SMEM);
    *out = m;
    return AMB_OK;
}

void amb_destroy(amb_modem* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    DevBuf* all[] = {&m->in_a, &m->in_b, &m->in_c, &m->in_lens, &m->out_a, &m->out_b, &m->out_c, &m->out_res, &m->out_bytes, &m->jobs, &m->idx, &m->p_scratch,
                     &m->xf, &m->psum, &m->pmin, &m->pmax, &m->norm, &m->metric, &m->tile_max, &m->tile_max32, &m->amb, &m->tie_rng, &m->cidx, &m->cbest, &m->tx_pay, &m->tx_out, &m->tx_copy, &m->raw, &m->cleaned,
                     &m->sums, &m->dc_next, &m->states, &m->recs, &m->nrec, &m->job_off, &m->total, &m->keep, &m->keep_off, &m->redo, &m->n_redo,
                     &m->s_res, &m->s_bytes, &m->s_sel, &m->g_res, &m->g_bytes, &m->ctmp, &m->nan_pos,
                     &m->tile_slot, &m->tri, &m->slot_info, &m->slot_count, &m->defect, &m->fpre, &m->xc, &m->dpart, &m->tsum, &m->cand_list,
                     &m->l_chunks, &m->l_runs, &m->l_nruns, &m->l_runoff, &m->l_recA, &m->l_endA, &m->l_hasA, &m->l_next, &m->l_linkrun, &m->l_nlinks, &m->l_outB,
                     &m->l_recB, &m->l_hasB, &m->l_vis, &m->l_ja, &m->l_jb, &m->l_active, &m->l_nactive,
                     &m->l_first, &m->l_firstflag, &m->l_firststop, &m->l_fpsum, &m->l_memo, &m->l_bmout, &m->l_bmrec, &m->l_bmhas, &m->l_startA, &m->xc_peak, &m->xc_rows, &m->cold_list, &m->cold_tag, &m->post_acc, &m->bitmaps, &m->files, &m->file_reset, &m->store_off};
    for (DevBuf* b : all) b->release();
    for (cudaEvent_t e : m->prof_pool) cudaEventDestroy(e);
    if (m->d_tables) cudaFree(m->d_tables);
    if (m->d_ema) cudaFree(m->d_ema);
    if (m->d_fast) cudaFree(m->d_fast);
    if (m->d_fallbacks) cudaFree(m->d_fallbacks);
    if (m->d_crc_ops) cudaFree(m->d_crc_ops);
    if (m->h_counts) cudaFreeHost(m->h_counts);
    for (cudaEvent_t e : m->copy_events) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

const char* amb_last_error(const amb_modem* m) { return m ? m->err.c_str() : "null handle"; }
int amb_synchronize(amb_modem* m) { if (!m) return AMB_E_INVALID; CK(cudaStreamSynchronize(m->stream)); return AMB_OK; }
void* amb_cuda_stream(amb_modem* m) { return m ? (void*)m->stream : nullptr; }
int64_t amb_kernel_launches(const amb_modem* m) { return m ? m->launches : 0; }

int64_t amb_debug_counter(amb_modem* m, int which) {
    if (!m) return -1;
    if (which == 0) return m->launches;
    if (which == 1) {
        unsigned long long v = 0;
        cudaStreamSynchronize(m->stream);
        cudaMemcpy(&v, m->d_fallbacks, 8, cudaMemcpyDeviceToHost);
        return (int64_t)v;
    }
#if defined(AMB_WALK_STATS)
    if (which >= 16 && which < 32) {
        unsigned long long v[16];
        cudaStreamSynchronize(m->stream);
        cudaMemcpyFromSymbol(v, g_amb_walk_stats, sizeof(v));
        return (int64_t)v[which - 16];
    }
#endif
    if (which == 6 || which == 7) {
        // phase-B links answered by phase A's own walk (6) and walked again (7), cumulative
        unsigned long long v = 0;
        cudaStreamSynchronize(m->stream);
        cudaMemcpy(&v, m->d_fallbacks + (which - 5), 8, cudaMemcpyDeviceToHost);
        return (int64_t)v;
    }
    if (which == 2) return m->link_iterations;
    if (which == 3) return m->link_passes;
    if ((which == 4 || which == 5) && m->open) {
        // slotted tiles (4) and tiles holding an exact candidate window (5) in the last push
        int n = 0;
        cudaStreamSynchronize(m->stream);
        cudaMemcpy(&n, m->slot_count.p, 4, cudaMemcpyDeviceToHost);
        if (n > m->slot_cap) n = m->slot_cap;
        if (which == 4) return n;
        std::vector<int32_t> info((size_t)n * 4);
        if (n > 0) cudaMemcpy(info.data(), m->slot_info.p, (size_t)n * 16, cudaMemcpyDeviceToHost);
        int64_t c = 0;
        for (int i = 0; i < n; i++) c += (info[(size_t)i * 4 + 2] & AMB_SLOT_CAND) != 0;
        return c;
    }
    return -1;
}

int64_t amb_debug_read(amb_modem* m, int which, void* out, int64_t cap_bytes) {
    if (!m || !m->open || !out) return -1;
    cudaStreamSynchronize(m->stream);
    int n = 0;
    cudaMemcpy(&n, m->slot_count.p, 4, cudaMemcpyDeviceToHost);
    if (n > m->slot_cap) n = m->slot_cap;
#if defined(AMB_WALK_STATS)
    if (which == 2) {
        // walk trace (development): rows of AMB_TRACE_W words; reading resets it
        unsigned int rows = 0;
        cudaMemcpyFromSymbol(&rows, g_amb_walk_trace_n, 4);
        if (rows > AMB_TRACE_ROWS) rows = AMB_TRACE_ROWS;
        int64_t bytes = (int64_t)rows * AMB_TRACE_W * 8;
        if (bytes > cap_bytes) bytes = cap_bytes;
        if (bytes > 0) cudaMemcpyFromSymbol(out, g_amb_walk_trace, (size_t)bytes);
        const unsigned int zero = 0;
        cudaMemcpyToSymbol(g_amb_walk_trace_n, &zero, 4);
        return rows;
    }
#endif
    const void* src = which == 0 ? m->slot_info.p : which == 1 ? m->tsum.p : nullptr;
    if (!src) return -1;
    int64_t bytes = (int64_t)n * 16;
    if (bytes > cap_bytes) bytes = cap_bytes;
    if (bytes > 0) cudaMemcpy(out, src, (size_t)bytes, cudaMemcpyDeviceToHost);
    return n;
}

int amb_host_alloc(void** out, uint64_t bytes) {
    if (!out) return AMB_E_INVALID;
    *out = nullptr;
    if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return AMB_E_NOMEM; }
    return AMB_OK;
}

int amb_host_free(void* p) { if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return AMB_E_CUDA; } return AMB_OK; }

int amb_selftest(amb_modem* m, int which, uint64_t seed, int n, int64_t* out2) {
    if (!m || !out2 || n <= 0) return AMB_E_INVALID;
    if (which != 0) return fail(m, AMB_E_INVALID, "unknown self-test");
    if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, AMB_E_CUDA, "cudaSetDevice failed");
    CK(m->ctmp.ensure(16));
    CK(cudaMemsetAsync(m->ctmp.p, 0, 16, m->stream));
    const int warps = 256, per = (n + warps - 1) / warps;
    LAUNCH("k_selftest_scan_batch", k_selftest_scan_batch<<<warps, 32, 0, m->stream>>>(seed, per, m->ctmp.as<unsigned long long>()));
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, m->ctmp.p, 16, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    out2[0] = (int64_t)h[0]; out2[1] = (int64_t)h[1];
    return AMB_OK;
}

int amb_profile(amb_modem* m, int enable) {
    if (!m) return AMB_E_INVALID;
    CK(cudaStreamSynchronize(m->stream));
    m->prof = enable != 0;
    m->prof_spans.clear();
    m->prof_used = 0;
    return AMB_OK;
}

int amb_profile_read(amb_modem* m, char* names, double* total_ms, int64_t* launches, int cap, int* n) {
    if (!m || !n) return AMB_E_INVALID;
    CK(cudaStreamSynchronize(m->stream));
    std::vector<std::string> keys;
    std::vector<double> ms;
    std::vector<int64_t> cnt;
    for (auto& sp : m->prof_spans) {
        if (!sp.b) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) != cudaSuccess) continue;
        size_t k = 0;
        for (; k < keys.size(); k++) if (keys[k] == sp.name) break;
        if (k == keys.size()) { keys.push_back(sp.name); ms.push_back(0.0); cnt.push_back(0); }
        ms[k] += t; cnt[k]++;
    }
    *n = (int)keys.size();
    for (int k = 0; k < *n && k < cap; k++) {
        if (names) { std::strncpy(names + 32 * k, keys[k].c_str(), 31); names[32 * k + 31] = 0; }
        if (total_ms) total_ms[k] = ms[k];
        if (launches) launches[k] = cnt[k];
    }
    m->prof_spans.clear();
    m->prof_used = 0;
    return AMB_OK;
}

int64_t amb_estimate_frame_samples(int config, int constellation, int64_t payload_bytes, int repetition) {
    return amb_tables_estimate_frame_samples(config, constellation, payload_bytes, repetition);
}

int amb_config_info(int config, int32_t* out8, int32_t* pilots) {
    AmbTables t;
    if (amb_tables_build(&t, config, AMB_MOD_QPSK, 1) != 0) return AMB_E_INVALID;
    out8[0] = AMB_FFT; out8[1] = t.cp_len; out8[2] = t.sym_len; out8[3] = 44100; out8[4] = t.sub_start; out8[5] = t.sub_end;
    out8[6] = t.n_pilots; out8[7] = t.n_data;
    if (pilots) for (int i = 0; i < t.n_pilots; i++) pilots[i] = t.pilots[i];
    return AMB_OK;
}

int amb_modem_info(const amb_modem* m, int32_t* out4) {
    if (!m || !out4) return AMB_E_INVALID;
    out4[0] = m->tables.cfg_id; out4[1] = m->tables.mod; out4[2] = m->tables.rep; out4[3] = m->device;
    return AMB_OK;
}

int amb_preamble_symbol1(int config, float* out) {
    AmbTables t;
    if (amb_tables_build(&t, config, AMB_MOD_QPSK, 1) != 0) return AMB_E_INVALID;
    std::memcpy(out, t.pre1, (size_t)t.sym_len * 4);
    return AMB_OK;
}

int amb_channel_est_symbol(int config, float* out, double* known_re) { amb_tables_ce_symbol(config, out, known_re); return AMB_OK; }

// ---------------------------------------------------------------------------------------------- batched ops
int amb_crc32(amb_modem* m, const uint8_t* data, int64_t stride, const int64_t* lens, int n, uint32_t* out, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    const uint8_t* d_data; const int64_t* d_lens; uint32_t* d_out;
    if ((rc = in_ptr(m, data, (size_t)n * stride, mem, m->in_a, &d_data))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    if ((rc = out_ptr(m, out, (size_t)n, mem, m->out_a, &d_out))) return rc;
    LAUNCH("k_crc32", k_crc32<<<n < 65536 ? n : 65536, AMB_CRC_THREADS, 0, m->stream>>>(m->d_tables, d_data, stride, d_lens, stride, n, d_out, m->d_crc_ops));
    return flush(m);
}

int amb_preprocess_signal(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, float* out, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0 || stride == 0) return AMB_OK;
    const float* d_x; const int64_t* d_lens; float* d_out;
    if ((rc = in_ptr(m, x, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    if ((rc = out_ptr(m, out, (size_t)n * stride, mem, m->out_a, &d_out))) return rc;
    const int64_t max_len = mem == AMB_MEM_HOST ? max_len_of(lens, n, stride) : stride;
    if ((rc = run_stats(m, d_x, stride, d_lens, n, max_len))) return rc;
    if ((rc = run_apply(m, d_x, stride, d_lens, n, max_len, d_out))) return rc;
    return flush(m);
}

int amb_detect_preamble(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, int64_t* idx_out,
                        double* metric_out, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    const float* d_x; const int64_t* d_lens;
    if ((rc = in_ptr(m, x, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    const int64_t max_len = mem == AMB_MEM_HOST ? max_len_of(lens, n, stride) : stride;
    if ((rc = run_detect(m, d_x, stride, d_lens, n, max_len))) return rc;
    if (mem == AMB_MEM_HOST) {
        if (idx_out) m->pending.push_back({idx_out, m->cidx.p, (size_t)n * 8});
        if (metric_out) m->pending.push_back({metric_out, m->cbest.p, (size_t)n * 8});
    } else {
        if (idx_out) CK(cudaMemcpyAsync(idx_out, m->cidx.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, m->stream));
        if (metric_out) CK(cudaMemcpyAsync(metric_out, m->cbest.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, m->stream));
    }
    return flush(m);
}

int amb_detect_preamble_xcorr(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, int64_t* idx_out,
                              double* metric_out, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    const float* d_x; const int64_t* d_lens; int64_t* d_idx; double* d_best;
    if ((rc = in_ptr(m, x, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    CK(m->cidx.ensure((size_t)n * 8)); CK(m->cbest.ensure((size_t)n * 8));
    d_idx = m->cidx.as<int64_t>(); d_best = m->cbest.as<double>();
    if ((rc = run_detect_xcorr(m, d_x, stride, d_lens, n, d_idx, d_best))) return rc;
    if (mem == AMB_MEM_HOST) {
        if (idx_out) m->pending.push_back({idx_out, d_idx, (size_t)n * 8});
        if (metric_out) m->pending.push_back({metric_out, d_best, (size_t)n * 8});
    } else {
        if (idx_out) CK(cudaMemcpyAsync(idx_out, d_idx, (size_t)n * 8, cudaMemcpyDeviceToDevice, m->stream));
        if (metric_out) CK(cudaMemcpyAsync(metric_out, d_best, (size_t)n * 8, cudaMemcpyDeviceToDevice, m->stream));
    }
    return flush(m);
}

int amb_estimate_channel(amb_modem* m, const float* ce, int n, const double* known_re, const double* known_im, double* ch_re,
                         double* ch_im, int mem) {
    int rc = check_common(m, n, m ? m->tables.sym_len : 0, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    const int64_t stride = m->tables.sym_len;
    const float* d_ce; const double *d_kr, *d_ki; double *d_hr, *d_hi;
    if ((rc = in_ptr(m, ce, (size_t)n * stride, mem, m->in_a, &d_ce))) return rc;
    if ((rc = in_ptr(m, known_re, (size_t)AMB_FFT, mem, m->in_b, &d_kr))) return rc;
    if ((rc = in_ptr(m, known_im, (size_t)AMB_FFT, mem, m->in_c, &d_ki))) return rc;
    if ((rc = out_ptr(m, ch_re, (size_t)n * AMB_FFT, mem, m->out_a, &d_hr))) return rc;
    if ((rc = out_ptr(m, ch_im, (size_t)n * AMB_FFT, mem, m->out_b, &d_hi))) return rc;
    if (!d_hr || !d_hi) return fail(m, AMB_E_INVALID, "ch_re/ch_im required");
    CK(m->jobs.ensure((size_t)n * sizeof(AmbJob)));
    LAUNCH("k_jobs_rows", k_jobs_rows<<<blocks_for(n, 128), 128, 0, m->stream>>>(m->d_tables, d_ce, stride, nullptr, stride, n, 1, m->jobs.as<AmbJob>()));
    AmbDemodOut o;
    std::memset(&o, 0, sizeof(o));
    o.h_out_re = d_hr; o.h_out_im = d_hi; o.known_re = d_kr; o.known_im = d_ki;
    LAUNCH("k_demod", k_demod<<<dim3(n, 1), 256, AMB_DEMOD_SMEM, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), n, o));
    return flush(m);
}

int amb_demodulate(amb_modem* m, const float* signal, int64_t stride, const int64_t* lens, int n, const double* ch_re,
                   const double* ch_im, uint8_t* bits, int64_t bits_stride, int64_t* nbits_out, double* eq_out, int64_t eq_stride,
                   int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    if (!ch_re || !ch_im) return fail(m, AMB_E_INVALID, "ch_re/ch_im required");
    const AmbTables& T = m->tables;
    const float* d_x; const int64_t* d_lens; const double *d_hr, *d_hi; uint8_t* d_bits; double* d_eq;
    if ((rc = in_ptr(m, signal, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    if ((rc = in_ptr(m, ch_re, (size_t)n * AMB_FFT, mem, m->in_b, &d_hr))) return rc;
    if ((rc = in_ptr(m, ch_im, (size_t)n * AMB_FFT, mem, m->in_c, &d_hi))) return rc;
    const int max_nsym = (int)(stride / T.sym_len);
    if (bits && bits_stride < (int64_t)max_nsym * T.n_data * T.bps) return fail(m, AMB_E_INVALID, "bits_stride too small");
    if (eq_out && eq_stride < (int64_t)max_nsym * T.n_data) return fail(m, AMB_E_INVALID, "eq_stride too small");
    if ((rc = out_ptr(m, bits, (size_t)n * bits_stride, mem, m->out_a, &d_bits))) return rc;
    if ((rc = out_ptr(m, eq_out, (size_t)n * eq_stride * 2, mem, m->out_b, &d_eq))) return rc;
    CK(m->jobs.ensure((size_t)n * sizeof(AmbJob)));
    LAUNCH("k_jobs_rows", k_jobs_rows<<<blocks_for(n, 128), 128, 0, m->stream>>>(m->d_tables, d_x, stride, d_lens, stride, n, 0, m->jobs.as<AmbJob>()));
    if (max_nsym > 0) {
        AmbDemodOut o;
        std::memset(&o, 0, sizeof(o));
        o.bits = d_bits; o.bits_stride = bits_stride; o.eq = d_eq; o.eq_stride = eq_stride; o.h_re = d_hr; o.h_im = d_hi;
        const int groups = (max_nsym + AMB_SYM_GROUP - 1) / AMB_SYM_GROUP;
        LAUNCH("k_demod", k_demod<<<dim3(n, groups), 256, AMB_DEMOD_SMEM, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), n, o));
    }
    if (nbits_out) {
        if (mem != AMB_MEM_HOST) return fail(m, AMB_E_INVALID, "nbits_out is a host-side convenience (pass NULL with device memory)");
        for (int i = 0; i < n; i++) nbits_out[i] = ((lens ? lens[i] : stride) / T.sym_len) * T.n_data * T.bps;
    }
    return flush(m);
}

int amb_decode_chunk_frames(amb_modem* m, const float* frames, int64_t stride, const int64_t* lens, int n,
                            amb_frame_result* results, uint8_t* bytes_out, int64_t bytes_stride, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    if (!results || !bytes_out || bytes_stride <= 0) return fail(m, AMB_E_INVALID, "results/bytes_out required");
    const AmbTables& T = m->tables;
    const float* d_x; const int64_t* d_lens; amb_frame_result* d_res; uint8_t* d_bytes;
    if ((rc = in_ptr(m, frames, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    if ((rc = out_ptr(m, results, (size_t)n, mem, m->out_res, &d_res))) return rc;
    if ((rc = out_ptr(m, bytes_out, (size_t)n * bytes_stride, mem, m->out_bytes, &d_bytes))) return rc;
    CK(m->jobs.ensure((size_t)n * sizeof(AmbJob)));
    LAUNCH("k_jobs_chunk", k_jobs_chunk<<<blocks_for(n, 128), 128, 0, m->stream>>>(m->d_tables, d_x, stride, d_lens, stride, n, m->jobs.as<AmbJob>()));
    const int max_nsym = (int)(stride / T.sym_len) - 3;
    if ((rc = run_demod_parse(m, n, max_nsym, false, d_res, d_bytes, bytes_stride))) return rc;
    return flush(m);
}

int amb_decode_received_signals(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n,
                                amb_frame_result* results, uint8_t* bytes_out, int64_t bytes_stride, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    if (!results || !bytes_out || bytes_stride <= 0) return fail(m, AMB_E_INVALID, "results/bytes_out required");
    const AmbTables& T = m->tables;
    const float* d_x; const int64_t* d_lens; amb_frame_result* d_res; uint8_t* d_bytes;
    if ((rc = in_ptr(m, x, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    if ((rc = out_ptr(m, results, (size_t)n, mem, m->out_res, &d_res))) return rc;
    if ((rc = out_ptr(m, bytes_out, (size_t)n * bytes_stride, mem, m->out_bytes, &d_bytes))) return rc;
    const int64_t max_len = mem == AMB_MEM_HOST ? max_len_of(lens, n, stride) : stride;
    CK(m->norm.ensure((size_t)n * stride * 4));
    if ((rc = run_stats(m, d_x, stride, d_lens, n, max_len))) return rc;
    if ((rc = run_apply(m, d_x, stride, d_lens, n, max_len, m->norm.as<float>()))) return rc;
    if ((rc = run_detect(m, m->norm.as<float>(), stride, d_lens, n, max_len))) return rc;
    CK(m->jobs.ensure((size_t)n * sizeof(AmbJob)));
    LAUNCH("k_refine_batch", k_refine_batch<<<n, 128, AMB_TEAM_STAGE_FLOATS * 4, m->stream>>>(m->d_tables, m->norm.as<float>(), stride, d_lens, stride, n, m->cidx.as<int64_t>(),
                                                             m->cbest.as<double>(), nullptr, 0, m->jobs.as<AmbJob>()));
    const int max_nsym = (int)(max_len / T.sym_len);
    if ((rc = run_demod_parse(m, n, max_nsym, false, d_res, d_bytes, bytes_stride, true))) return rc;
    return flush(m);
}

int amb_analyze_loopback(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, const uint8_t* test_data, int test_len,
                         amb_loopback_result* results, double* channel_magnitude, int64_t mag_stride, int mem) {
    int rc = check_common(m, n, stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    const AmbTables& T = m->tables;
    const int nb = T.sub_end - T.sub_start + 1;
    if (!results || !test_data || test_len < 0 || test_len > 65536) return fail(m, AMB_E_INVALID, "results/test_data required");
    if (channel_magnitude && mag_stride < nb) return fail(m, AMB_E_INVALID, "mag_stride below the number of subcarriers in the band");
    const float* d_x; const int64_t* d_lens;
    if ((rc = in_ptr(m, x, (size_t)n * stride, mem, m->in_a, &d_x))) return rc;
    if ((rc = in_ptr(m, lens, (size_t)n, mem, m->in_lens, &d_lens))) return rc;
    const int64_t max_len = mem == AMB_MEM_HOST ? max_len_of(lens, n, stride) : stride;
    // decoded bytes kept per recording: the name length byte, up to 255 name bytes, the length field and the fixture
    const int64_t bytes_stride = (1 + 255 + 4 + (int64_t)test_len + 15) / 16 * 16;
    CK(m->norm.ensure((size_t)n * stride * 4));
    CK(m->out_res.ensure((size_t)n * sizeof(amb_frame_result))); CK(m->out_bytes.ensure((size_t)n * bytes_stride));
    CK(m->out_a.ensure((size_t)n * sizeof(amb_loopback_result))); CK(m->out_b.ensure((size_t)n * nb * 8)); CK(m->in_b.ensure((size_t)(test_len > 0 ? test_len : 1)));
    CK(m->in_c.ensure((size_t)n * 16));
    if (test_len > 0) CK(cudaMemcpyAsync(m->in_b.p, test_data, (size_t)test_len, cudaMemcpyHostToDevice, m->stream));
    if ((rc = run_stats(m, d_x, stride, d_lens, n, max_len))) return rc;
    if ((rc = run_apply(m, d_x, stride, d_lens, n, max_len, m->norm.as<float>()))) return rc;
    if ((rc = run_detect(m, m->norm.as<float>(), stride, d_lens, n, max_len))) return rc;
    int64_t* d_xidx = m->in_c.as<int64_t>();
    double* d_xbest = reinterpret_cast<double*>(d_xidx + n);
    if ((rc = run_detect_xcorr(m, m->norm.as<float>(), stride, d_lens, n, d_xidx, d_xbest))) return rc;
    CK(m->jobs.ensure((size_t)n * sizeof(AmbJob)));
    LAUNCH("k_refine_batch", k_refine_batch<<<n, 128, AMB_TEAM_STAGE_FLOATS * 4, m->stream>>>(m->d_tables, m->norm.as<float>(), stride, d_lens, stride, n, m->cidx.as<int64_t>(),
                                                             m->cbest.as<double>(), d_xidx, 1, m->jobs.as<AmbJob>()));
    const int max_nsym = (int)(max_len / T.sym_len);
    if ((rc = run_demod_parse(m, n, max_nsym, false, m->out_res.as<amb_frame_result>(), m->out_bytes.as<uint8_t>(), bytes_stride))) return rc;
    LAUNCH("k_loopback_report", k_loopback_report<<<n, 32, 0, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), n, m->out_res.as<amb_frame_result>(), m->out_bytes.as<uint8_t>(),
                                                                           bytes_stride, m->in_b.as<uint8_t>(), test_len, m->out_a.as<amb_loopback_result>(),
                                                                           m->out_b.as<double>(), nb));
    CK(cudaMemcpyAsync(results, m->out_a.p, (size_t)n * sizeof(amb_loopback_result), cudaMemcpyDeviceToHost, m->stream));
    if (channel_magnitude)
        CK(cudaMemcpy2DAsync(channel_magnitude, (size_t)mag_stride * 8, m->out_b.p, (size_t)nb * 8, (size_t)nb * 8, n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < n; i++) {
        amb_loopback_result& r = results[i];
        // modem.js:1048-1049; the early returns (:989, :1021) report 0
        if (r.n_mag > 0) r.snr_estimate = r.pilot_mag > 0.0 ? 20.0 * std::log10(r.pilot_mag) : -INFINITY;
        if (channel_magnitude && r.n_mag == 0) for (int k = 0; k < nb; k++) channel_magnitude[(int64_t)i * mag_stride + k] = 0.0;
    }
    return AMB_OK;
}

// ---------------------------------------------------------------------------------------------- streaming
int amb_stream_open(amb_modem* m, int n_streams, int block_len, int64_t max_push) {
    if (!m) return AMB_E_INVALID;
    if (n_streams <= 0 || block_len <= 0 || max_push <= 0) return fail(m, AMB_E_INVALID, "bad stream geometry");
    if (block_len != AMB_TILE) return fail(m, AMB_E_INVALID, "block_len must be 4096 (the reference's ScriptProcessor block, app.js:1107)");
    if (max_push % block_len) return fail(m, AMB_E_INVALID, "max_push must be a multiple of block_len");
    if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, AMB_E_CUDA, "cudaSetDevice failed");
    if (m->open) amb_stream_close(m);
    const AmbTables& T = m->tables;
    m->n_streams = n_streams; m->block_len = block_len; m->max_push = max_push;
    // RingBuffer capacity, app.js:712-715
    m->ring_cap = amb_tables_estimate_frame_samples(T.cfg_id, T.mod, 4096 + 16, T.rep) * 3 + 8192;
    // history kept in the arena: the ring, one more block (reads happen while a call's samples are being added),
    // the refine radius and a tile of slack, rounded to whole tiles
    int64_t h = m->ring_cap + block_len + T.fine_radius + T.sym_len + 2 * AMB_TILE;
    m->hist = (h + AMB_TILE - 1) / AMB_TILE * AMB_TILE;
    // the arena holds several pushes beyond the history, so the history is moved to the front only every `reps` pushes
    // (compaction traffic stays below the push traffic even when pushes are much shorter than the history)
    // (moving the history costs hist / (reps * max_push) of the push traffic: with many streams and short pushes the arena takes
    // as many pushes as a quarter of the free device memory holds, at most 32 and at most 48 GB)
    int64_t reps = (m->hist + max_push - 1) / max_push;
    {
        size_t free_b = 0, total_b = 0;
        double budget = 48e9;
        if (std::getenv("AMB_ARENA_GB") && std::atof(std::getenv("AMB_ARENA_GB")) > 0.0) budget = std::atof(std::getenv("AMB_ARENA_GB")) * 1e9;      // (development: arena budget)
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (double)free_b / 4.0 < budget) budget = (double)free_b / 4.0;
        const double per_stream = budget / ((double)n_streams * 4.0) - (double)m->hist;
        const int64_t fit = per_stream > 0 ? (int64_t)(per_stream / (double)max_push) : 0;
        if (n_streams > 1 && fit > reps) reps = fit;
    }
    if (reps > 32) reps = 32;
    while (reps > 1 && (double)n_streams * (double)(m->hist + reps * max_push) * 4.0 > 48e9) reps--;
    m->stride = m->hist + reps * max_push;
    // an odd number of 16 KB blocks per row: the frames of concurrent streams sit at the same offset of their rows, and rows a
    // large power of two apart put them all on the same DRAM channels (measured on 8,192 streams: demodulator 2.3x slower)
    if (n_streams > 1 && ((m->stride / AMB_TILE) & 1) == 0) m->stride += AMB_TILE;
    m->base_g = -m->hist;
    m->written = 0;
    m->sum_stride = m->stride / 256;
    const int64_t min_frame = amb_tables_estimate_frame_samples(T.cfg_id, T.mod, 280, T.rep) < amb_tables_estimate_frame_samples(T.cfg_id, T.mod, 11, T.rep)
                                  ? amb_tables_estimate_frame_samples(T.cfg_id, T.mod, 280, T.rep)
                                  : amb_tables_estimate_frame_samples(T.cfg_id, T.mod, 11, T.rep);
    m->rec_cap = (int)(max_push / min_frame) + 3;
    const size_t ns = (size_t)n_streams;
    // `raw` only stages the samples of one push (host input, unaligned device input, in-place writers); the history
    // lives in `cleaned`
    CK(m->raw.ensure(ns * m->max_push * 4)); CK(m->cleaned.ensure(ns * m->stride * 4));
    CK(m->sums.ensure(ns * m->sum_stride * AMB_SUM_W * 4));
    CK(m->dc_next.ensure(ns * 8));
    CK(m->states.ensure(ns * sizeof(AmbStreamState))); CK(m->recs.ensure(ns * m->rec_cap * sizeof(AmbFrameRec)));
    CK(m->nrec.ensure(ns * 4)); CK(m->job_off.ensure(ns * 4)); CK(m->total.ensure(16)); CK(m->redo.ensure(ns * 4)); CK(m->n_redo.ensure(16));
    CK(cudaMemsetAsync(m->cleaned.p, 0, ns * m->stride * 4, m->stream));
    CK(cudaMemsetAsync(m->sums.p, 0, ns * m->sum_stride * AMB_SUM_W * 4, m->stream));
    CK(cudaMemsetAsync(m->dc_next.p, 0, ns * 8, m->stream));
    // stores written by the scan kernel for the walker
    m->calls_stride = m->stride / AMB_TILE;
    m->xc_w = 256 + 2 * T.fine_radius;
    const int64_t tiles_per_push = (max_push / 256) * n_streams;
    m->slot_cap = (int)(tiles_per_push / 8 + 64 < 2000000 ? tiles_per_push / 8 + 64 : 2000000);
    CK(m->tile_slot.ensure(ns * m->sum_stride * 4)); CK(m->tri.ensure((size_t)m->slot_cap * 256 * 3 * 8));
    CK(m->slot_info.ensure((size_t)m->slot_cap * 16)); CK(m->slot_count.ensure(16)); CK(m->tsum.ensure((size_t)m->slot_cap * 16)); CK(m->cand_list.ensure((size_t)m->slot_cap * 4));
    CK(m->defect.ensure(ns * m->calls_stride * 3 * 8)); CK(m->fpre.ensure(ns * m->calls_stride * 3 * 8)); CK(m->cold_tag.ensure(ns * m->calls_stride * 4));
    CK(m->xc.ensure((size_t)m->slot_cap * m->xc_w * 4)); CK(m->xc_peak.ensure((size_t)m->slot_cap * 8));
    CK(cudaMemsetAsync(m->defect.p, 0, ns * m->calls_stride * 3 * 8, m->stream));
    CK(cudaMemsetAsync(m->fpre.p, 0, ns * m->calls_stride * 3 * 8, m->stream));
    cudaFuncSetAttribute(k_xcorr_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AMB_XC_SMEM(AMB_XC_MAXW + 768));
    cudaFuncSetAttribute(k_link_walk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AMB_TEAM_STAGE_FLOATS * 4));
    m->run_cap = (int)(max_push / 512) + 2;
    CK(m->l_runs.ensure(ns * m->run_cap * sizeof(AmbRun))); CK(m->l_nruns.ensure(ns * 4)); CK(m->l_runoff.ensure(ns * 4));
    CK(m->l_nlinks.ensure(ns * 4)); CK(m->l_active.ensure(ns * 4)); CK(m->l_nactive.ensure(16));
    {
        // per run / per link scratch of the link-parallel resolution, for the largest possible number of candidate runs
        const size_t nr = ns * m->run_cap + 1, nl = nr + ns;
        CK(m->l_recA.ensure(nr * sizeof(AmbFrameRec))); CK(m->l_endA.ensure(nr * sizeof(AmbStreamState))); CK(m->l_hasA.ensure(nr * 4));
        CK(m->l_next.ensure(nr * 4)); CK(m->l_vis.ensure(nr)); CK(m->l_ja.ensure(nr * 4)); CK(m->l_jb.ensure(nr * 4));
        CK(m->l_linkrun.ensure(nl * 4)); CK(m->l_outB.ensure(nl * sizeof(AmbStreamState))); CK(m->l_recB.ensure(nl * sizeof(AmbFrameRec)));
        CK(m->l_hasB.ensure(nl * 4)); CK(m->l_memo.ensure(nr * sizeof(AmbRefineMemo)));
        CK(m->l_startA.ensure(nr * sizeof(AmbStartA)));
        CK(m->l_bmout.ensure(nr * sizeof(AmbStreamState))); CK(m->l_bmrec.ensure(nr * sizeof(AmbFrameRec))); CK(m->l_bmhas.ensure(nr * 4));
        CK(m->l_first.ensure(ns * sizeof(AmbStreamState))); CK(m->l_firstflag.ensure(ns * 4)); CK(m->l_firststop.ensure(ns * 4)); CK(m->l_fpsum.ensure(ns * 8));
    }
    {
        // ChunkAssembler.receivedBitmap per stream: 2 Mi chunks (8 GiB files at 4 KB chunks) unless that would take more than
        // 512 MiB over all streams; chunks beyond the capacity are not tracked for duplicates
        int64_t bits = 1ll << 21;
        while (bits > 65536 && (double)ns * (double)bits / 8.0 > 512.0 * 1024 * 1024) bits >>= 1;
        m->bitmap_bits = bits;
        CK(m->bitmaps.ensure(ns * (size_t)(bits / 8)));
        CK(cudaMemsetAsync(m->bitmaps.p, 0, ns * (size_t)(bits / 8), m->stream));
    }
    if (!m->h_counts && cudaHostAlloc((void**)&m->h_counts, 64, cudaHostAllocDefault) != cudaSuccess) return fail(m, AMB_E_NOMEM, "cudaHostAlloc failed");
    std::vector<AmbStreamState> init(ns);
    std::memset(init.data(), 0, ns * sizeof(AmbStreamState));
    for (auto& s : init) { s.pre_pos = -1; s.frame_end = -1; }
    CK(cudaMemcpyAsync(m->states.p, init.data(), ns * sizeof(AmbStreamState), cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->open = true;
    return AMB_OK;
}

int amb_stream_close(amb_modem* m) {
    if (!m) return AMB_E_INVALID;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    DevBuf* all[] = {&m->raw, &m->cleaned, &m->sums, &m->dc_next, &m->states, &m->recs, &m->nrec, &m->job_off, &m->total,
                     &m->keep, &m->keep_off, &m->redo, &m->n_redo, &m->s_res, &m->s_bytes, &m->s_sel, &m->g_res, &m->g_bytes, &m->ctmp, &m->nan_pos,
                     &m->tile_slot, &m->tri, &m->slot_info, &m->slot_count, &m->defect, &m->fpre, &m->xc, &m->dpart, &m->tsum, &m->cand_list,
                     &m->l_chunks, &m->l_runs, &m->l_nruns, &m->l_runoff, &m->l_recA, &m->l_endA, &m->l_hasA, &m->l_next, &m->l_linkrun, &m->l_nlinks, &m->l_outB,
                     &m->l_recB, &m->l_hasB, &m->l_vis, &m->l_ja, &m->l_jb, &m->l_active, &m->l_nactive,
                     &m->l_first, &m->l_firstflag, &m->l_firststop, &m->l_fpsum, &m->l_memo, &m->l_bmout, &m->l_bmrec, &m->l_bmhas, &m->l_startA, &m->xc_peak, &m->xc_rows, &m->cold_list, &m->cold_tag, &m->post_acc, &m->bitmaps, &m->files, &m->file_reset, &m->store_off};
    for (DevBuf* b : all) b->release();
    m->open = false;
    m->file_cap = 0;
    return AMB_OK;
}

int amb_stream_input_window(amb_modem* m, float** dev_ptr, int64_t* stride) {
    if (!m || !m->open) return m ? fail(m, AMB_E_STATE, "stream not open") : AMB_E_INVALID;
    if (dev_ptr) *dev_ptr = m->raw.as<float>();
    if (stride) *stride = m->max_push;
    return AMB_OK;
}

// rows [n_streams][pitch] -> move `bytes` bytes per row from byte offset `from` to the row start (through a scratch
// copy when source and destination overlap)
static int move_rows(amb_modem* m, void* base, size_t pitch, size_t from, size_t bytes) {
    char* b = static_cast<char*>(base);
    if (from >= bytes) {
        CK(cudaMemcpy2DAsync(b, pitch, b + from, pitch, bytes, m->n_streams, cudaMemcpyDeviceToDevice, m->stream));
        return AMB_OK;
    }
    CK(m->ctmp.ensure((size_t)m->n_streams * bytes));
    CK(cudaMemcpy2DAsync(m->ctmp.p, bytes, b + from, pitch, bytes, m->n_streams, cudaMemcpyDeviceToDevice, m->stream));
    CK(cudaMemcpy2DAsync(b, pitch, m->ctmp.p, bytes, bytes, m->n_streams, cudaMemcpyDeviceToDevice, m->stream));
    return AMB_OK;
}

static int stream_compact(amb_modem* m) {
    // keep `hist` samples of history: move [written - hist, written) to the arena front (cleaned, summaries, defect prefix)
    const int64_t off = m->written - m->base_g;        // arena index of the next sample
    if (off == m->hist) return AMB_OK;
    const int64_t shift = off - m->hist;               // multiple of 4096
    int rc;
    if ((rc = move_rows(m, m->cleaned.p, (size_t)m->stride * 4, (size_t)shift * 4, (size_t)m->hist * 4))) return rc;
    if ((rc = move_rows(m, m->sums.p, (size_t)m->sum_stride * AMB_SUM_W * 4, (size_t)(shift / 256) * AMB_SUM_W * 4, (size_t)(m->hist / 256) * AMB_SUM_W * 4))) return rc;
    if ((rc = move_rows(m, m->fpre.p, (size_t)m->calls_stride * 3 * 8, (size_t)(shift / AMB_TILE) * 3 * 8, (size_t)(m->hist / AMB_TILE) * 3 * 8))) return rc;
    m->base_g += shift;
    return AMB_OK;
}

// One push of n_new samples per stream, read from device memory (src rows 16-byte aligned).
static int push_core(amb_modem* m, const void* src, int64_t src_stride, int64_t n_new, amb_frame_result* results,
                     int max_results, int* n_results, uint8_t* bytes_out, int64_t bytes_stride, int results_mem) {
    const AmbTables& T = m->tables;
    const int ns = m->n_streams;
    if (m->written - m->base_g + n_new > m->stride) { const int rcc = stream_compact(m); if (rcc) return rcc; }     // arena full: history to the front
    const int64_t off0 = m->written - m->base_g;       // arena index where the new samples go
    if (off0 + n_new > m->stride) return fail(m, AMB_E_STATE, "arena overflow (internal)");
    const int tiles = (int)(n_new / AMB_TILE);
    AmbScanStore ss;
    ss.tile_slot = m->tile_slot.as<int32_t>(); ss.sum_stride = m->sum_stride; ss.tri = m->tri.as<double>(); ss.slot_info = m->slot_info.as<int32_t>();
    ss.slot_count = m->slot_count.as<int>(); ss.slot_cap = m->slot_cap; ss.defect = m->defect.as<double>(); ss.calls_stride = m->calls_stride;
    ss.tsum = m->tsum.as<float>(); ss.cand_list = m->cand_list.as<int32_t>(); ss.xc_peak = m->xc_peak.as<unsigned long long>();
    // AMB_TRI_JUNK32=0 (comparison runs): every slotted tile gets its float64 window sums
    static const bool junk32_on = !(std::getenv("AMB_TRI_JUNK32") && std::getenv("AMB_TRI_JUNK32")[0] == '0');
    ss.junk32 = junk32_on ? 1 : 0;
    // tiles the walker can look at (the history and the new samples): no slot (-1) until this push assigns one
    CK(cudaMemset2DAsync(m->tile_slot.as<int32_t>() + (off0 - m->hist) / 256, (size_t)m->sum_stride * 4, 0xFF, (size_t)((m->hist + n_new) / 256) * 4, ns, m->stream));
    CK(cudaMemsetAsync(m->slot_count.p, 0, 12, m->stream));         // [0] slots, [1] slots holding a candidate, [2] screening rows (k_xcorr_rows)
    {
        // 1+2. DC removal and the window-sum summaries of the new window-end samples, one pass (k_front)
        AmbFrontArgs fa;
        fa.x = src; fa.x_stride = src_stride; fa.cleaned = m->cleaned.as<float>(); fa.stride = m->stride; fa.off0 = off0;
        // tile steps of 16 x threads samples.  128-thread CTAs (2048-sample steps, six per SM) measured 5 % faster than 256
        // (three per SM) and the same as 64 (twelve per SM); AMB_FRONT_NT=256 selects the larger step for comparison.
        static const int front_nt = (std::getenv("AMB_FRONT_NT") && std::atoi(std::getenv("AMB_FRONT_NT")) == 256) ? 256 : 128;
        const int per_block = AMB_TILE / (16 * front_nt);                    // tile steps per 4096-sample block: 1 or 2
        const int ftiles = tiles * per_block, ctas_per_sm = front_nt == 128 ? m->front_ctas_per_sm : 768 / front_nt;
        fa.tiles = ftiles;
        // spans: enough CTAs to fill the GPU, long enough that the 40,960-sample warm-up of the EMA state stays a small part
        int64_t span = ((int64_t)ns * ftiles + 148 * ctas_per_sm - 1) / (148 * ctas_per_sm);      // one wave
        if (span < 16 * per_block) span = 16 * per_block;
        if (span > ftiles) span = ftiles;
        fa.span = (int)span; fa.spans = (int)((ftiles + span - 1) / span); fa.warm = 10 * per_block;
        fa.E = m->d_ema; fa.states = m->states.as<AmbStreamState>(); fa.dc_next = m->dc_next.as<double>();
        fa.sums = m->sums.as<float>(); fa.sum_stride = m->sum_stride; fa.st = ss;
        CK(m->nan_pos.ensure((size_t)(ns + 1) * 8));
        CK(cudaMemsetAsync(m->nan_pos.p, 0xFF, (size_t)ns * 8, m->stream));
        CK(cudaMemsetAsync(m->nan_pos.as<unsigned long long>() + ns, 0, 8, m->stream));
        fa.nan_pos = m->nan_pos.as<unsigned long long>(); fa.n_streams = ns; fa.g0 = m->written;
        if (front_nt == 128) {
            if (m->push_pcm16) { LAUNCH("k_front", (k_front<int16_t, 128><<<ns * fa.spans, 128, AMB_FRONT_SMEM(128), m->stream>>>(fa))); }
            else { LAUNCH("k_front", (k_front<float, 128><<<ns * fa.spans, 128, AMB_FRONT_SMEM(128), m->stream>>>(fa))); }
        } else {
            if (m->push_pcm16) { LAUNCH("k_front", (k_front<int16_t, 256><<<ns * fa.spans, 256, AMB_FRONT_SMEM(256), m->stream>>>(fa))); }
            else { LAUNCH("k_front", (k_front<float, 256><<<ns * fa.spans, 256, AMB_FRONT_SMEM(256), m->stream>>>(fa))); }
        }
        LAUNCH("k_nan_fix", k_nan_fix<<<148 * 4, 256, 0, m->stream>>>(fa, n_new, m->base_g));
    }
    const int tri_grid_max = 148 * 64;
    LAUNCH("k_tri_store", k_tri_store<<<(m->slot_cap + 3) / 4 < tri_grid_max ? (m->slot_cap + 3) / 4 : tri_grid_max, 128, 0, m->stream>>>(m->cleaned.as<float>(), m->stride, m->base_g, m->written + n_new, ss));
    {
        const int64_t n_calls = n_new / AMB_TILE;
        const int parts = (int)((n_calls + AMB_DEFECT_PART - 1) / AMB_DEFECT_PART);
        CK(m->dpart.ensure((size_t)ns * parts * 3 * 8));
        LAUNCH("k_defect_part", k_defect_part<<<ns * parts, 256, 0, m->stream>>>(m->defect.as<double>(), m->calls_stride, off0 / AMB_TILE, n_calls, parts, m->dpart.as<double>()));
        LAUNCH("k_defect_scan", k_defect_scan<<<ns * parts, 256, 0, m->stream>>>(m->defect.as<double>(), m->fpre.as<double>(), m->calls_stride, off0 / AMB_TILE, n_calls, parts,
                                                                                m->dpart.as<double>()));
    }
    // 3. walk / decode / apply, repeated for streams whose collection window changed
    AmbStreamGeom g;
    g.stride = m->stride; g.base_g = m->base_g; g.avail_hi = m->written + n_new; g.sum_stride = m->sum_stride;
    g.sum_lo = m->written - m->hist > 0 ? m->written - m->hist : 0;       // summaries and slots are kept for the last `hist` samples
    g.block_len = m->block_len; g.ring_cap = m->ring_cap;
    g.k_end = (m->written + n_new) / m->block_len; g.rec_cap = m->rec_cap;
    // AMB_XC_STORE=0 (comparison runs): no screening store, the walkers screen their refines themselves (amb_xcorr_screen)
    // AMB_XC_SKIP=0 (comparison runs): a screening row for every candidate tile, also those next to a higher peak
    static const bool xc_skip_on = !(std::getenv("AMB_XC_SKIP") && std::getenv("AMB_XC_SKIP")[0] == '0');
    static const bool xc_store_on = !(std::getenv("AMB_XC_STORE") && std::getenv("AMB_XC_STORE")[0] == '0');
    g.tile_slot = m->tile_slot.as<int32_t>(); g.tri = m->tri.as<double>(); g.xc = xc_store_on ? m->xc.as<float>() : nullptr; g.xc_w = m->xc_w; g.slot_info = m->slot_info.as<int32_t>(); g.tsum = m->tsum.as<float>();
    g.F = m->fpre.as<double>(); g.calls_stride = m->calls_stride;
    // AMB_PREDICT=0 (comparison runs): phase A starts every run from the synthetic state, phase B walks every link
    static const bool predict_on = !(std::getenv("AMB_PREDICT") && std::getenv("AMB_PREDICT")[0] == '0');
    g.xc_peak = (xc_store_on && predict_on) ? m->xc_peak.as<unsigned long long>() : nullptr;
    // idle-call screening (k_cold_calls) pays when pushes are long; AMB_COLD_CALLS=0 switches it off (comparison runs)
    static const bool cold_on = !(std::getenv("AMB_COLD_CALLS") && std::getenv("AMB_COLD_CALLS")[0] == '0');
    const bool use_cold = cold_on && n_new / m->block_len >= 64;
    g.cold_tag = use_cold ? m->cold_tag.as<int32_t>() : nullptr;
    if (use_cold) {
        // a tag states a fact about the data (call k is cold under the defect of a receiver re-initialised in call c): the
        // tags of this push's calls start empty and accumulate over the link iterations
        const int64_t call_lo = m->written / m->block_len, call_hi = (m->written + n_new) / m->block_len;
        CK(cudaMemset2DAsync(m->cold_tag.as<int32_t>() + (call_lo - m->base_g / m->block_len), (size_t)m->calls_stride * 4, 0xFF, (size_t)(call_hi - call_lo) * 4, ns, m->stream));
    }
    {
        const int L = m->xc_w + T.sym_len;
        const size_t smem = AMB_XC_SMEM(L);
        int threads = ((m->xc_w + 15) / 16 + 15) / 16 * 32;               // a warp takes 16 blocks of 16 offsets (8 per lane, by parity)
        if (threads > 256) threads = 256;
        // the CTAs stride over the compact list of candidate slots (k_tri_store): even shares, and each CTA prefetches its
        // next slot while it works on the current one
        // (measured per 33 k candidate slots: 148 x 8 CTAs 0.49 ms, x 16 0.45, x 32 0.43, x 64 0.42, x 128 0.41)
        static const int xc_mult = std::getenv("AMB_XC_GRID") ? std::atoi(std::getenv("AMB_XC_GRID")) : 128;      // (development: CTAs per SM of the launch)
        const int grid = m->slot_cap < 148 * xc_mult ? m->slot_cap : 148 * xc_mult;
        if (xc_store_on) {
            CK(m->xc_rows.ensure((size_t)m->slot_cap * 16));
            const int rgrid = (int)(blocks_for(m->slot_cap, 256) < 148 * 4 ? blocks_for(m->slot_cap, 256) : 148 * 4);
            LAUNCH("k_xcorr_rows", k_xcorr_rows<<<rgrid, 256, 0, m->stream>>>(g, m->slot_info.as<int32_t>(), m->slot_count.as<int>(), m->cand_list.as<int32_t>(), m->slot_cap,
                                                                               xc_skip_on ? m->tile_slot.as<int32_t>() : nullptr, off0 / 256, (off0 + n_new) / 256,
                                                                               reinterpret_cast<int4*>(m->xc_rows.p)));
            LAUNCH("k_xcorr_store", k_xcorr_store<<<grid, threads, smem, m->stream>>>(m->d_tables, g, m->cleaned.as<float>(), m->slot_count.as<int>(),
                                                                                     reinterpret_cast<const int4*>(m->xc_rows.p), m->slot_cap, m->xc.as<float>(), m->xc_w,
                                                                                     m->xc_peak.as<unsigned long long>()));
        }
    }
    int n_out = 0;
    int n_todo = ns;
    const int* todo = nullptr;
    int rc = AMB_OK;
    amb_frame_result* dst_res = results;
    uint8_t* dst_bytes = bytes_out;
    if (results_mem == AMB_MEM_HOST) {
        if (results) { CK(m->g_res.ensure((size_t)(max_results > 0 ? max_results : 1) * sizeof(amb_frame_result))); dst_res = m->g_res.as<amb_frame_result>(); }
        if (bytes_out) { CK(m->g_bytes.ensure((size_t)(max_results > 0 ? max_results : 1) * bytes_stride)); dst_bytes = m->g_bytes.as<uint8_t>(); }
    }
    CK(m->jobs.ensure((size_t)ns * m->rec_cap * sizeof(AmbJob)));
    CK(m->s_sel.ensure((size_t)ns * 4));
    // link buffers (sized at open for the largest possible number of candidate runs)
    AmbLinkBufs L;
    std::memset(&L, 0, sizeof(L));
    L.runs = m->l_runs.as<AmbRun>(); L.n_runs = m->l_nruns.as<int>(); L.run_off = m->l_runoff.as<int>(); L.run_cap = m->run_cap;
    L.n_links = m->l_nlinks.as<int>(); L.active = m->l_active.as<int>(); L.n_active = m->l_nactive.as<int>();
    L.recA = m->l_recA.as<AmbFrameRec>(); L.endA = m->l_endA.as<AmbStreamState>(); L.hasA = m->l_hasA.as<int>(); L.next = m->l_next.as<int>();
    L.vis = m->l_vis.as<unsigned char>(); L.jump_a = m->l_ja.as<int>(); L.jump_b = m->l_jb.as<int>();
    L.link_run = m->l_linkrun.as<int>(); L.outB = m->l_outB.as<AmbStreamState>(); L.recB = m->l_recB.as<AmbFrameRec>(); L.hasB = m->l_hasB.as<int>();
    L.first_out = m->l_first.as<AmbStreamState>(); L.first_flag = m->l_firstflag.as<int>(); L.first_stop = m->l_firststop.as<int>();
    L.fp_sum = m->l_fpsum.as<unsigned long long>(); L.memo = m->l_memo.as<AmbRefineMemo>();
    L.startA = g.xc_peak ? m->l_startA.as<AmbStartA>() : nullptr; L.stats = m->d_fallbacks + 1;
    L.bm_out = m->l_bmout.as<AmbStreamState>(); L.bm_rec = m->l_bmrec.as<AmbFrameRec>(); L.bm_has = m->l_bmhas.as<int>();
    const int64_t max_links = (int64_t)ns * m->run_cap + ns;
    const int walk_grid = (int)(blocks_for(max_links, 4) < m->walk_ctas ? blocks_for(max_links, 4) : m->walk_ctas);
    // phase B: a CTA per flat link index where that many CTAs fit the grid - then the four warps of a CTA walk one link as a team
    // (k_link_walk decides from total <= gridDim.x); otherwise a warp per link as in phase A
    const int walk_grid_b = max_links <= m->walk_ctas ? (int)(max_links > 0 ? max_links : 1) : walk_grid;
    const int flat_grid = (int)(blocks_for(max_links, 128) < 148 * 8 ? blocks_for(max_links, 128) : 148 * 8);
    int* h = m->h_counts;           // pinned: [0] n_active, [1] total jobs, [2] max nsym, [3] kept, [4] n_redo
    for (int pass = 0; pass < 4096 && n_todo > 0; pass++) {
        CK(cudaMemsetAsync(m->nrec.p, 0, (size_t)ns * 4, m->stream));
        CK(cudaMemsetAsync(m->n_redo.p, 0, 4, m->stream));
        CK(cudaMemsetAsync(m->total.p, 0, 16, m->stream));
        // streams without metadata yet: one frame per pass with the sequential walker
        LAUNCH("k_walk", k_walk<<<blocks_for(n_todo, 4), 128, 0, m->stream>>>(m->d_tables, g, m->cleaned.as<float>(), m->sums.as<float>(), m->states.as<AmbStreamState>(),
                                                             m->dc_next.as<double>(), todo, n_todo, m->recs.as<AmbFrameRec>(), m->nrec.as<int>(), 1));
        // streams with metadata: link-parallel resolution (speculate, then verify; amb_walker.cuh)
        CK(cudaMemsetAsync(L.n_runs, 0, (size_t)ns * 4, m->stream));
        CK(cudaMemsetAsync(L.active, 0, (size_t)ns * 4, m->stream));
        CK(cudaMemsetAsync(L.first_flag, 0, (size_t)ns * 4, m->stream));
        CK(cudaMemsetAsync(L.n_active, 0, 8, m->stream));
        {
            const int64_t tile_lo = off0 / 256, tile_hi = (off0 + n_new) / 256;
            const int chunks = (int)((tile_hi - tile_lo + AMB_RUN_CHUNK - 1) / AMB_RUN_CHUNK);
            CK(m->l_chunks.ensure((size_t)ns * chunks * 4));
            LAUNCH("k_runs_count", k_runs_pass<<<n_todo * chunks, 256, 0, m->stream>>>(g, m->tile_slot.as<int32_t>(), m->slot_info.as<int32_t>(), m->states.as<AmbStreamState>(),
                                                                                      todo, n_todo, tile_lo, tile_hi, chunks, m->l_chunks.as<int>(), 0, L));
            LAUNCH("k_runs_offsets", k_runs_offsets<<<blocks_for(n_todo, 4), 128, 0, m->stream>>>(m->states.as<AmbStreamState>(), todo, n_todo, chunks, m->l_chunks.as<int>(), L));
            LAUNCH("k_runs_write", k_runs_pass<<<n_todo * chunks, 256, 0, m->stream>>>(g, m->tile_slot.as<int32_t>(), m->slot_info.as<int32_t>(), m->states.as<AmbStreamState>(),
                                                                                      todo, n_todo, tile_lo, tile_hi, chunks, m->l_chunks.as<int>(), 1, L));
        }
        LAUNCH("k_prefix_counts", k_prefix_counts<<<1, 1024, 0, m->stream>>>(L.n_runs, ns, L.run_off, L.n_active + 1));
        // phase A over every run, and next to it the first link of every stream
        L.trace_iter = (int)(m->link_iterations * 2);
        LAUNCH("k_link_walk_A", k_link_walk<false><<<walk_grid, 128, 0, m->stream>>>(m->d_tables, g, m->cleaned.as<float>(), m->sums.as<float>(), m->states.as<AmbStreamState>(),
                                                                            m->dc_next.as<double>(), todo, n_todo, ns, 0, m->recs.as<AmbFrameRec>(), m->nrec.as<int>(), L));
        LAUNCH("k_chain_next", k_chain_next<<<flat_grid, 128, 0, m->stream>>>(ns, L));
        m->link_passes++;
        int n_active = 1;            // the first iteration runs unconditionally (its kernels skip inactive streams)
        for (int it = 0; n_active > 0 && it < (1 << 20); it++) {
            m->link_iterations++;
            CK(cudaMemsetAsync(L.n_active, 0, 4, m->stream));
            LAUNCH("k_chain_links", k_chain_links<<<ns, ns < 64 ? 1024 : 256, ns < 64 ? AMB_CHAIN_SMEM : 0, m->stream>>>(ns, m->states.as<AmbStreamState>(), L, ns < 64 ? AMB_CHAIN_SMEM_RUNS : 0));
            if (use_cold) {
                const int64_t call_lo = m->written / m->block_len, call_hi = (m->written + n_new) / m->block_len;
                const int64_t calls = (int64_t)ns * (call_hi - call_lo);
                CK(m->cold_list.ensure((size_t)calls * 16 + 16));
                int* list_count = m->cold_list.as<int>();                   // [0] count (16 bytes), then the int4 entries
                int4* list = reinterpret_cast<int4*>(m->cold_list.as<int>() + 4);
                CK(cudaMemsetAsync(list_count, 0, 4, m->stream));
                const int lgrid = (int)(blocks_for(calls, 256) < 148 * 8 ? blocks_for(calls, 256) : 148 * 8);
                LAUNCH("k_cold_list", k_cold_list<<<lgrid, 256, 0, m->stream>>>(g, m->states.as<AmbStreamState>(), ns, L, m->cold_tag.as<int32_t>(), call_lo, call_hi, list_count, list));
                const int cgrid = (int)(blocks_for(calls, 4) < 148 * 12 ? blocks_for(calls, 4) : 148 * 12);
                LAUNCH("k_cold_calls", k_cold_calls<<<cgrid, 128, 0, m->stream>>>(g, m->cleaned.as<float>(), m->sums.as<float>(), m->states.as<AmbStreamState>(), ns, L,
                                                                                  m->cold_tag.as<int32_t>(), call_lo, call_hi, list_count, list));
            }
            L.trace_iter = (int)(m->link_iterations * 2 + 1);
            LAUNCH("k_link_walk_B", k_link_walk<false><<<walk_grid_b, 128, 0, m->stream>>>(m->d_tables, g, m->cleaned.as<float>(), m->sums.as<float>(), m->states.as<AmbStreamState>(),
                                                                                m->dc_next.as<double>(), todo, n_todo, ns, 1, m->recs.as<AmbFrameRec>(), m->nrec.as<int>(), L));
            LAUNCH("k_link_walk_B_team", k_link_walk<true><<<walk_grid_b, 128, AMB_TEAM_STAGE_FLOATS * 4, m->stream>>>(m->d_tables, g, m->cleaned.as<float>(), m->sums.as<float>(), m->states.as<AmbStreamState>(),
                                                                                m->dc_next.as<double>(), todo, n_todo, ns, 1, m->recs.as<AmbFrameRec>(), m->nrec.as<int>(), L));
            LAUNCH("k_verify_copy", k_verify_copy<<<flat_grid, 128, 0, m->stream>>>(ns, m->rec_cap, m->states.as<AmbStreamState>(), m->recs.as<AmbFrameRec>(), m->nrec.as<int>(), L));
            LAUNCH("k_verify_final", k_verify_final<<<blocks_for(ns, 128), 128, 0, m->stream>>>(ns, m->rec_cap, m->states.as<AmbStreamState>(), m->dc_next.as<double>(), g.avail_hi, g.k_end,
                                                                                              m->nrec.as<int>(), L));
            // frame records -> jobs (redone if another iteration adds records)
            LAUNCH("k_prefix_counts", k_prefix_counts<<<1, 1024, 0, m->stream>>>(m->nrec.as<int>(), ns, m->job_off.as<int>(), m->total.as<int>()));
            LAUNCH("k_jobs_stream", k_jobs_stream<<<blocks_for((int64_t)ns * m->rec_cap, 128), 128, 0, m->stream>>>(m->d_tables, g, m->cleaned.as<float>(), m->recs.as<AmbFrameRec>(), m->nrec.as<int>(),
                                                                      m->job_off.as<int>(), ns, m->jobs.as<AmbJob>(), m->total.as<int>() + 1));
            CK(cudaMemcpyAsync(h, L.n_active, 4, cudaMemcpyDeviceToHost, m->stream));
            CK(cudaMemcpyAsync(h + 1, m->total.p, 8, cudaMemcpyDeviceToHost, m->stream));
            CK(cudaStreamSynchronize(m->stream));
            n_active = h[0];
        }
        const int total = h[1];      // frames, h[2] = max data symbols
        if (total == 0) break;
        // Frames are decoded straight into the caller's rows [n_out, n_out + total): every frame is kept unless a metadata
        // frame rewinds its stream (app.js:619-625 semantics in amb_post_frames), which is rare; only then (or when the rows
        // would overflow) the frames go through a scratch copy and are compacted.
        const bool direct = dst_res && dst_bytes && (int64_t)n_out + total <= max_results;
        if (!direct) { CK(m->s_res.ensure((size_t)total * sizeof(amb_frame_result))); CK(m->s_bytes.ensure((size_t)total * bytes_stride)); }
        amb_frame_result* pres = direct ? dst_res + n_out : m->s_res.as<amb_frame_result>();
        uint8_t* pbytes = direct ? dst_bytes + (size_t)n_out * bytes_stride : m->s_bytes.as<uint8_t>();
        CK(m->keep.ensure((size_t)total * 4)); CK(m->keep_off.ensure((size_t)total * 4));
        CK(cudaMemsetAsync(m->keep.p, 0, (size_t)total * 4, m->stream));
        LAUNCH("k_fine_exact", k_fine_exact<<<blocks_for(total, 64), 128, 0, m->stream>>>(m->d_tables, m->jobs.as<AmbJob>(), total));
        if ((rc = run_demod_parse(m, total, h[2], true, pres, pbytes, bytes_stride))) return rc;
        {
            int64_t parts = ((int64_t)total / ns + 1023) / 1024;
            if (parts < 1) parts = 1;
            if (parts > 64) parts = 64;
            int64_t* store_off = nullptr;
            if (m->file_cap > 0) { CK(m->store_off.ensure((size_t)total * 8)); store_off = m->store_off.as<int64_t>(); }
            CK(m->post_acc.ensure((size_t)ns * 8 * 8));
            CK(cudaMemsetAsync(m->post_acc.p, 0, (size_t)ns * 8 * 8, m->stream));
            LAUNCH("k_post_count", k_post_count<<<ns * (int)parts, 256, 0, m->stream>>>(m->nrec.as<int>(), m->job_off.as<int>(), pres, ns, m->states.as<AmbStreamState>(),
                                                                                    m->keep.as<int>(), (int)parts, m->post_acc.as<unsigned long long>()));
            LAUNCH("k_post_apply", k_post_apply<<<ns * (int)parts, 256, 0, m->stream>>>(m->recs.as<AmbFrameRec>(), m->rec_cap, m->nrec.as<int>(), m->job_off.as<int>(),
                                                                                    pres, ns, m->states.as<AmbStreamState>(), m->keep.as<int>(), m->redo.as<int>(),
                                                                                    m->n_redo.as<int>(), g.k_end, (int)parts, m->post_acc.as<unsigned long long>(),
                                                                                    m->bitmaps.as<uint32_t>(), m->bitmap_bits, store_off, m->file_cap > 0 ? m->file_reset.as<int>() : nullptr));
            if (m->file_cap > 0) {
                // chunk store (amb_stream_enable_assembly): clear the files a metadata frame reset, then write the new chunks
                const int rparts = (int)(m->file_cap / (1 << 20) + 1 < 256 ? m->file_cap / (1 << 20) + 1 : 256);
                LAUNCH("k_file_reset", k_file_reset<<<ns * rparts, 256, 0, m->stream>>>(m->files.as<uint8_t>(), m->file_cap, m->file_reset.as<int>(), ns, rparts));
                LAUNCH("k_file_reset_done", k_file_reset_done<<<blocks_for(ns, 128), 128, 0, m->stream>>>(m->file_reset.as<int>(), ns));
                LAUNCH("k_store_chunks", k_store_chunks<<<total < 65536 ? total : 65536, 128, 0, m->stream>>>(pres, pbytes, bytes_stride, store_off, m->keep.as<int>(), total,
                                                                                                   m->files.as<uint8_t>(), m->file_cap));
            }
        }
        LAUNCH("k_prefix_counts", k_prefix_counts<<<1, 1024, 0, m->stream>>>(m->keep.as<int>(), total, m->keep_off.as<int>(), m->total.as<int>() + 2));
        if (!direct) {
            LAUNCH("k_gather", k_gather<<<total, 128, 0, m->stream>>>(m->keep.as<int>(), m->keep_off.as<int>(), total, pres, pbytes, bytes_stride, n_out, max_results, dst_res, dst_bytes));
        }
        CK(cudaMemcpyAsync(h + 3, m->total.as<int>() + 2, 4, cudaMemcpyDeviceToHost, m->stream));
        CK(cudaMemcpyAsync(h + 4, m->n_redo.p, 4, cudaMemcpyDeviceToHost, m->stream));
        auto rows_to_host = [&](int count) -> int {
            // rows [n_out, n_out + count) of the device staging to the caller's host buffers
            if (results_mem != AMB_MEM_HOST) return AMB_OK;
            const int lo = n_out < max_results ? n_out : max_results, hi = n_out + count < max_results ? n_out + count : max_results;
            if (hi > lo && results) CK(cudaMemcpyAsync(results + lo, m->g_res.as<amb_frame_result>() + lo, (size_t)(hi - lo) * sizeof(amb_frame_result), cudaMemcpyDeviceToHost, m->stream));
            if (hi > lo && bytes_out) CK(cudaMemcpyAsync(bytes_out + (size_t)lo * bytes_stride, m->g_bytes.as<uint8_t>() + (size_t)lo * bytes_stride, (size_t)(hi - lo) * bytes_stride,
                                                         cudaMemcpyDeviceToHost, m->stream));
            return AMB_OK;
        };
        if ((rc = rows_to_host(total))) return rc;      // everything this pass can have kept
        CK(cudaStreamSynchronize(m->stream));
        if (direct && h[3] < total) {
            // some frames were dropped: compact the rows through a scratch copy
            CK(m->s_res.ensure((size_t)total * sizeof(amb_frame_result))); CK(m->s_bytes.ensure((size_t)total * bytes_stride));
            CK(cudaMemcpyAsync(m->s_res.p, pres, (size_t)total * sizeof(amb_frame_result), cudaMemcpyDeviceToDevice, m->stream));
            CK(cudaMemcpyAsync(m->s_bytes.p, pbytes, (size_t)total * bytes_stride, cudaMemcpyDeviceToDevice, m->stream));
            LAUNCH("k_gather", k_gather<<<total, 128, 0, m->stream>>>(m->keep.as<int>(), m->keep_off.as<int>(), total, m->s_res.as<amb_frame_result>(), m->s_bytes.as<uint8_t>(),
                                                   bytes_stride, n_out, max_results, dst_res, dst_bytes));
            if ((rc = rows_to_host(h[3]))) return rc;
            CK(cudaStreamSynchronize(m->stream));
        }
        int n_redo = h[4];
        if (n_redo > ns) n_redo = ns;
        n_out += h[3];
        n_todo = n_redo;
        if (n_todo > 0) {
            // the redo list is rewritten by the next k_post_stream while k_walk may still read it: copy it aside
            CK(cudaMemcpyAsync(m->s_sel.p, m->redo.p, (size_t)n_todo * 4, cudaMemcpyDeviceToDevice, m->stream));
            todo = m->s_sel.as<int>();
        }
    }
    const bool overflow = n_out > max_results;
    m->written += n_new;
    *n_results = n_out;
    return overflow ? AMB_E_OVERFLOW : AMB_OK;
}

// samples of one push in either input format (es = bytes per sample: 4 float32, 2 16-bit PCM)
static int stream_push_any(amb_modem* m, const void* xv, int es, int64_t x_stride, int64_t n_new, int mem, amb_frame_result* results,
                           int max_results, int* n_results, uint8_t* bytes_out, int64_t bytes_stride, int results_mem) {
    if (!m) return AMB_E_INVALID;
    m->pending.clear();
    if (!m->open) return fail(m, AMB_E_STATE, "stream not open");
    if (n_new <= 0 || n_new > m->max_push || n_new % m->block_len) return fail(m, AMB_E_INVALID, "n_new must be a positive multiple of block_len, at most max_push");
    if (mem != AMB_MEM_HOST && mem != AMB_MEM_DEVICE && mem != AMB_MEM_INPLACE) return fail(m, AMB_E_INVALID, "bad mem");
    if (results_mem != AMB_MEM_HOST && results_mem != AMB_MEM_DEVICE) return fail(m, AMB_E_INVALID, "bad results_mem");
    if (!n_results) return fail(m, AMB_E_INVALID, "n_results required");
    if (bytes_out && bytes_stride <= 0) return fail(m, AMB_E_INVALID, "bytes_stride must be positive");
    if (max_results < 0) return fail(m, AMB_E_INVALID, "max_results must not be negative");
    if (mem != AMB_MEM_INPLACE && !xv) return fail(m, AMB_E_INVALID, "x required");
    if (mem == AMB_MEM_INPLACE && es != 4) return fail(m, AMB_E_INVALID, "the in-place window holds float32 samples");
    if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, AMB_E_CUDA, "cudaSetDevice failed");
    const int ns = m->n_streams;
    m->push_pcm16 = es == 2;
    const char* x = static_cast<const char*>(xv);
    // samples are read where they are when they sit in device memory with 16-byte aligned rows; otherwise they are staged
    if (mem == AMB_MEM_DEVICE && ((uintptr_t)x & 15) == 0 && ((x_stride * es) & 15) == 0)
        return push_core(m, x, x_stride, n_new, results, max_results, n_results, bytes_out, bytes_stride, results_mem);
    char* raw = m->raw.as<char>();                      // rows of max_push samples in the input's own format
    const size_t pitch = (size_t)m->max_push * es;
    if (mem == AMB_MEM_INPLACE) return push_core(m, raw, m->max_push, n_new, results, max_results, n_results, bytes_out, bytes_stride, results_mem);
    // Large host pushes are cut into sections: all host->device copies are queued at once on a second stream, and
    // section k is processed (as a push of its own: same results, the receiver state carries over) while the sections
    // after it are still in flight.
    const int64_t blocks = n_new / m->block_len;
    int64_t n_sec = 1;
    static const int64_t sec_min_bytes = std::getenv("AMB_SECTION_MIN_BYTES") ? std::atoll(std::getenv("AMB_SECTION_MIN_BYTES")) : (64ll << 20);
    if (mem == AMB_MEM_HOST && (int64_t)ns * n_new * es >= sec_min_bytes && blocks >= 8) n_sec = 8;
    if (n_sec == 1) {
        CK(cudaMemcpy2DAsync(raw, pitch, x, (size_t)x_stride * es, (size_t)n_new * es, ns,
                             mem == AMB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, m->stream));
        return push_core(m, raw, m->max_push, n_new, results, max_results, n_results, bytes_out, bytes_stride, results_mem);
    }
    if (!m->copy_stream) CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    while ((int64_t)m->copy_events.size() < n_sec) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); m->copy_events.push_back(e); }
    const int64_t sec_blocks = (blocks + n_sec - 1) / n_sec;
    std::vector<int64_t> sec_off, sec_len;
    for (int64_t b = 0; b < blocks; b += sec_blocks) { sec_off.push_back(b * m->block_len); sec_len.push_back((b + sec_blocks < blocks ? sec_blocks : blocks - b) * m->block_len); }
    // the staging buffer is free: the previous push synchronised after its last reader
    for (size_t k = 0; k < sec_off.size(); k++) {
        CK(cudaMemcpy2DAsync(raw + sec_off[k] * es, pitch, x + sec_off[k] * es, (size_t)x_stride * es, (size_t)sec_len[k] * es, ns, cudaMemcpyHostToDevice, m->copy_stream));
        CK(cudaEventRecord(m->copy_events[k], m->copy_stream));
    }
    int total = 0, rc_all = AMB_OK;
    for (size_t k = 0; k < sec_off.size(); k++) {
        CK(cudaStreamWaitEvent(m->stream, m->copy_events[k], 0));
        const int room = max_results - total > 0 ? max_results - total : 0;
        int got = 0;
        const int rc = push_core(m, raw + sec_off[k] * es, m->max_push, sec_len[k], results ? results + (total < max_results ? total : max_results) : nullptr, room, &got,
                                 bytes_out ? bytes_out + (size_t)(total < max_results ? total : max_results) * bytes_stride : nullptr, bytes_stride, results_mem);
        if (rc != AMB_OK && rc != AMB_E_OVERFLOW) {
            // the copies of the later sections still read the caller's buffer: let them finish before handing it back
            cudaStreamSynchronize(m->copy_stream);
            *n_results = total;
            return rc;
        }
        if (rc == AMB_E_OVERFLOW) rc_all = AMB_E_OVERFLOW;
        total += got;
    }
    *n_results = total;
    return rc_all;
}

int amb_stream_push(amb_modem* m, const float* x, int64_t x_stride, int64_t n_new, int mem, amb_frame_result* results,
                    int max_results, int* n_results, uint8_t* bytes_out, int64_t bytes_stride, int results_mem) {
    return stream_push_any(m, x, 4, x_stride, n_new, mem, results, max_results, n_results, bytes_out, bytes_stride, results_mem);
}

int amb_stream_push_pcm16(amb_modem* m, const int16_t* x, int64_t x_stride, int64_t n_new, int mem, amb_frame_result* results,
                          int max_results, int* n_results, uint8_t* bytes_out, int64_t bytes_stride, int results_mem) {
    return stream_push_any(m, x, 2, x_stride, n_new, mem, results, max_results, n_results, bytes_out, bytes_stride, results_mem);
}

// ---------------------------------------------------------------------------------------------- transmit side
int64_t amb_chunk_frame_samples(amb_modem* m, int64_t payload_bytes, int is_first_frame) {
    if (!m) return -1;
    return amb_tx_frame_samples(&m->tables, payload_bytes, is_first_frame);
}

static int tx_run(amb_modem* m, const uint8_t* d_pay, int64_t pay_stride, int pay_len, int n, int first, float* out, int64_t out_stride,
                  double sigma, uint64_t seed, int64_t g0, int mem, int64_t pre = -1, int64_t post = -1) {
    const int64_t total = pre >= 0 ? pre + (3 + amb_tx_num_symbols(&m->tables, pay_len)) * m->tables.sym_len + post : amb_tx_frame_samples(&m->tables, pay_len, first);
    if (out_stride < total) return fail(m, AMB_E_INVALID, "out_stride smaller than the frame");
    float* d_out = out;
    if (mem == AMB_MEM_HOST) { CK(m->tx_out.ensure((size_t)n * out_stride * 4)); d_out = m->tx_out.as<float>(); CK(cudaMemsetAsync(d_out, 0, (size_t)n * out_stride * 4, m->stream)); }
    AmbTxArgs a;
    a.payloads = d_pay; a.payload_stride = pay_stride; a.payload_len = pay_len; a.first = first; a.out = d_out; a.out_stride = out_stride;
    a.noise_sigma = sigma; a.noise_seed = seed; a.noise_g0 = g0; a.pre = pre; a.post = post;
    LAUNCH("k_tx_frames", k_tx_frames<<<n, 256, AMB_TX_SMEM, m->stream>>>(m->d_tables, a, n));
    if (mem == AMB_MEM_HOST) m->pending.push_back({out, d_out, (size_t)n * out_stride * 4});
    return AMB_OK;
}

int amb_build_data_chunk_frames(amb_modem* m, const uint8_t* data, int64_t data_stride, int data_len, int n, int32_t seq0,
                                uint64_t synth_seed, float* out, int64_t out_stride, double noise_sigma, uint64_t noise_seed,
                                int64_t noise_g0, uint8_t* data_copy, int mem) {
    int rc = check_common(m, n, out_stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    if (!out || data_len < 0 || data_len > 65535) return fail(m, AMB_E_INVALID, "bad argument");
    const uint8_t* d_data; uint8_t* d_copy;
    if ((rc = in_ptr(m, data, (size_t)n * (data ? data_stride : 0), mem, m->in_a, &d_data))) return rc;
    if ((rc = out_ptr(m, data_copy, (size_t)n * data_len, mem, m->tx_copy, &d_copy))) return rc;
    const int pay_len = data_len + 11;
    CK(m->tx_pay.ensure((size_t)n * pay_len));
    LAUNCH("k_tx_chunk_payloads", k_tx_chunk_payloads<<<blocks_for(n, 64), 64, 0, m->stream>>>(m->d_tables, d_data, data_stride, data_len, n, seq0, synth_seed,
                                                                 m->tx_pay.as<uint8_t>(), pay_len, d_copy, data_len));
    if ((rc = tx_run(m, m->tx_pay.as<uint8_t>(), pay_len, pay_len, n, 0, out, out_stride, noise_sigma, noise_seed, noise_g0, mem))) return rc;
    return flush(m);
}

int amb_build_metadata_frame(amb_modem* m, int32_t total_chunks, int32_t total_file_size, int32_t chunk_size, const uint8_t* name,
                             int name_len, float* out, double noise_sigma, uint64_t noise_seed, int64_t noise_g0, int mem) {
    int rc = check_common(m, 1, 0, mem); if (rc) return rc;
    if (!out) return fail(m, AMB_E_INVALID, "out required");
    static const uint8_t dflt[4] = {'f', 'i', 'l', 'e'};                 // `fileName || 'file'`
    if (!name || name_len <= 0) { name = dflt; name_len = 4; }
    uint8_t pay[300];
    const int pay_len = amb_metadata_payload(m->tables.crc_table, total_chunks, total_file_size, chunk_size, name, name_len, pay);
    CK(m->tx_pay.ensure(512));
    CK(cudaMemcpyAsync(m->tx_pay.p, pay, (size_t)pay_len, cudaMemcpyHostToDevice, m->stream));
    const int64_t total = amb_tx_frame_samples(&m->tables, pay_len, 1);
    if ((rc = tx_run(m, m->tx_pay.as<uint8_t>(), pay_len, pay_len, 1, 1, out, total, noise_sigma, noise_seed, noise_g0, mem))) return rc;
    return flush(m);
}

// silences of buildTransmitSignal, modem.js:533-535 (`new Float32Array(44100 * x)` truncates)
static void legacy_silences(const AmbTables& T, int64_t* pre, int64_t* post) {
    const bool acoustic = T.cp_len >= 128;
    *pre = (int64_t)std::floor(44100.0 * (acoustic ? 0.5 : 0.3));
    *post = (int64_t)std::floor(44100.0 * (acoustic ? 0.5 : 0.2));
}

int64_t amb_transmit_signal_samples(amb_modem* m, int64_t data_len, int name_len) {
    if (!m || data_len < 0) return -1;
    if (name_len <= 0) name_len = 4;
    if (name_len > 255) name_len = 255;
    int64_t pre, post;
    legacy_silences(m->tables, &pre, &post);
    return pre + (3 + amb_tx_num_symbols(&m->tables, 1 + name_len + 4 + data_len + 4)) * m->tables.sym_len + post;
}

int amb_build_transmit_signals(amb_modem* m, const uint8_t* data, int64_t data_stride, int data_len, int n, const uint8_t* name,
                               int name_len, float* out, int64_t out_stride, int mem) {
    int rc = check_common(m, n, out_stride, mem); if (rc) return rc;
    if (n == 0) return AMB_OK;
    if (!out || !data || data_len < 0 || data_len > (1 << 24) || data_stride < data_len) return fail(m, AMB_E_INVALID, "bad argument");
    static const uint8_t dflt[4] = {'f', 'i', 'l', 'e'};                 // `fileName || 'file'`
    if (!name || name_len <= 0) { name = dflt; name_len = 4; }
    if (name_len > 255) name_len = 255;                                  // Math.min(nameBytes.length, 255)
    // packet [nameLen:1][name][dataLen:4][data][CRC-32:4] (modem.js:504-523), assembled on the host: the data arrives in
    // host memory (the payloads are small next to the signal they turn into)
    if (mem != AMB_MEM_HOST) return fail(m, AMB_E_INVALID, "data and out are host memory for this call");
    const int pay_len = 1 + name_len + 4 + data_len + 4;
    std::vector<uint8_t> pay((size_t)n * pay_len);
    for (int i = 0; i < n; i++) {
        uint8_t* p = pay.data() + (size_t)i * pay_len;
        int off = 0;
        p[off++] = (uint8_t)name_len;
        std::memcpy(p + off, name, (size_t)name_len); off += name_len;
        p[off++] = (uint8_t)(data_len >> 24); p[off++] = (uint8_t)(data_len >> 16); p[off++] = (uint8_t)(data_len >> 8); p[off++] = (uint8_t)data_len;
        std::memcpy(p + off, data + (size_t)i * data_stride, (size_t)data_len); off += data_len;
        const uint32_t cs = amb_crc32_bytes(m->tables.crc_table, p, off);
        p[off++] = (uint8_t)(cs >> 24); p[off++] = (uint8_t)(cs >> 16); p[off++] = (uint8_t)(cs >> 8); p[off++] = (uint8_t)cs;
    }
    CK(m->tx_pay.ensure(pay.size()));
    CK(cudaMemcpyAsync(m->tx_pay.p, pay.data(), pay.size(), cudaMemcpyHostToDevice, m->stream));
    int64_t pre, post;
    legacy_silences(m->tables, &pre, &post);
    if ((rc = tx_run(m, m->tx_pay.as<uint8_t>(), pay_len, pay_len, n, 1, out, out_stride, 0.0, 0, 0, mem, pre, post))) return rc;
    return flush(m);        // synchronises: `pay` stays alive until the copy has run
}

int amb_modulate(amb_modem* m, uint8_t* bits, int64_t bits_stride, int64_t n_bits, int n, float* samples, int64_t samples_stride, int32_t* out3,
                 int mem) {
    int rc = check_common(m, n, bits_stride, mem); if (rc) return rc;
    const AmbTables& T = m->tables;
    const int64_t per = (int64_t)T.n_data * T.bps;                     // bitsPerSymbol, modem.js:326
    if (n_bits < 0 || n_bits > bits_stride) return fail(m, AMB_E_INVALID, "n_bits outside the row");
    const int64_t nsym = (n_bits + per - 1) / per, padded = nsym * per;
    if (out3) { out3[0] = (int32_t)nsym; out3[1] = (int32_t)per; out3[2] = (int32_t)padded; }
    if (n == 0 || nsym == 0) return AMB_OK;
    if (!bits || !samples || samples_stride < nsym * T.sym_len) return fail(m, AMB_E_INVALID, "bits/samples required; samples_stride >= numSymbols * SYMBOL_LEN");
    if (nsym * n > 0x7fffffff / 2) return fail(m, AMB_E_INVALID, "too many symbols in one call");
    // the reference pads its caller's array in place (modem.js:329); done here when the row has room for the padding
    const int pad_rows = bits_stride >= padded ? 1 : 0;
    uint8_t* d_bits = bits; float* d_out = samples;
    if (mem == AMB_MEM_HOST) {
        CK(m->in_a.ensure((size_t)n * bits_stride));
        CK(cudaMemcpyAsync(m->in_a.p, bits, (size_t)n * bits_stride, cudaMemcpyHostToDevice, m->stream));
        d_bits = m->in_a.as<uint8_t>();
        if ((rc = out_ptr(m, samples, (size_t)n * samples_stride, mem, m->tx_out, &d_out))) return rc;
        if (pad_rows) m->pending.push_back({bits, m->in_a.p, (size_t)n * bits_stride});
    }
    const int64_t groups = (nsym + AMB_TX_WARPS - 1) / AMB_TX_WARPS;
    LAUNCH("k_modulate", k_modulate<<<(unsigned)(n * groups), 256, AMB_TX_SMEM, m->stream>>>(m->d_tables, d_bits, bits_stride, n_bits, n, nsym, d_out, samples_stride, pad_rows));
    return flush(m);
}

int amb_stream_enable_assembly(amb_modem* m, int64_t max_file_bytes) {
    if (!m) return AMB_E_INVALID;
    if (!m->open) return fail(m, AMB_E_STATE, "stream not open");
    if (max_file_bytes <= 0) return fail(m, AMB_E_INVALID, "max_file_bytes must be positive");
    if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, AMB_E_CUDA, "cudaSetDevice failed");
    const int64_t cap = (max_file_bytes + 15) / 16 * 16;
    CK(m->files.ensure((size_t)m->n_streams * cap)); CK(m->file_reset.ensure((size_t)m->n_streams * 4));
    CK(cudaMemsetAsync(m->files.p, 0, (size_t)m->n_streams * cap, m->stream));
    CK(cudaMemsetAsync(m->file_reset.p, 0, (size_t)m->n_streams * 4, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->file_cap = cap;
    return AMB_OK;
}

int amb_stream_assemble(amb_modem* m, int stream, uint8_t* out, int64_t cap, int64_t* size_out, int mem) {
    if (!m || !size_out) return AMB_E_INVALID;
    if (!m->open || m->file_cap <= 0) return fail(m, AMB_E_STATE, "assembly not enabled (amb_stream_enable_assembly)");
    if (stream < 0 || stream >= m->n_streams || cap < 0 || (mem != AMB_MEM_HOST && mem != AMB_MEM_DEVICE)) return fail(m, AMB_E_INVALID, "bad argument");
    if (cudaSetDevice(m->device) != cudaSuccess) return fail(m, AMB_E_CUDA, "cudaSetDevice failed");
    AmbStreamState st;
    CK(cudaMemcpyAsync(&st, m->states.as<AmbStreamState>() + stream, sizeof(st), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    int64_t size = st.meta_received ? (int64_t)st.total_file_size : 0;       // new Uint8Array(totalFileSize), app.js:667
    if (size < 0) size = 0;
    if (size > m->file_cap) size = m->file_cap;
    *size_out = size;
    const int64_t n = size < cap ? size : cap;
    if (out && n > 0) {
        CK(cudaMemcpyAsync(out, m->files.as<uint8_t>() + (int64_t)stream * m->file_cap, (size_t)n, mem == AMB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, m->stream));
        CK(cudaStreamSynchronize(m->stream));
    }
    return AMB_OK;
}

int amb_stream_counters(amb_modem* m, int stream, int64_t* out8) {
    if (!m || !out8) return AMB_E_INVALID;
    if (!m->open) return fail(m, AMB_E_STATE, "stream not open");
    if (stream >= m->n_streams) return fail(m, AMB_E_INVALID, "bad stream index");
    std::vector<AmbStreamState> st((size_t)m->n_streams);
    CK(cudaMemcpyAsync(st.data(), m->states.p, st.size() * sizeof(AmbStreamState), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < 8; i++) out8[i] = 0;
    for (int s = 0; s < m->n_streams; s++) {
        if (stream >= 0 && s != stream) continue;
        out8[0] += st[s].frames_decoded; out8[1] += st[s].frame_errors; out8[2] += st[s].false_pos; out8[3] += st[s].received;
        out8[4] += st[s].crc_errors; out8[5] += st[s].dups; out8[6] += st[s].meta_received; out8[7] += st[s].written;
    }
    return AMB_OK;
}

}  // extern "C"
