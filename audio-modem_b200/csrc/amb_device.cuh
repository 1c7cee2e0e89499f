// Warp-cooperative device algorithms of the OFDM receive path.
//
// Everything here is written against a tiny SPMD shim so that ONE source serves two builds:
//   * device (nvcc, sm_100a): a "warp loop" strides over the 32 lanes, AMB_WARP_SYNC is __syncwarp;
//   * host emulation (g++, tests/emu only): a warp loop is a plain sequential loop.
// The host build exists only so the kernel logic can be checked against the oracle on a machine with
// no GPU; the product never runs it (the C-ABI library has no CPU path).
//
// "Faithful" routines reproduce modem.js's float64 operation sequence (no FMA contraction:
// __dmul_rn/__dadd_rn/__ddiv_rn), so their results are bit-identical to the reference arithmetic.
#pragma once
#include <stdint.h>
#include <math.h>

#include "../../include/audio_modem_b200.h"
#include "amb_tables.h"

#if defined(__CUDACC__)
#define AMB_HD __host__ __device__ __forceinline__
#define AMB_HD_COLD __host__ __device__ __noinline__        // rarely taken, register-hungry paths: kept out of the callers' allocation
#else
#define AMB_HD inline
#define AMB_HD_COLD inline
#endif

#if defined(__CUDA_ARCH__)
#define AMB_LANE() ((int)(threadIdx.x & 31))
#define AMB_WARP_FOR(i, n) for (int i = AMB_LANE(); i < (int)(n); i += 32)
#define AMB_WARP_SYNC() __syncwarp()
#define AMB_LEADER() (AMB_LANE() == 0)
AMB_HD double amb_mul(double a, double b) { return __dmul_rn(a, b); }
AMB_HD double amb_add(double a, double b) { return __dadd_rn(a, b); }
AMB_HD double amb_sub(double a, double b) { return __dsub_rn(a, b); }
AMB_HD double amb_div(double a, double b) { return __ddiv_rn(a, b); }
AMB_HD double amb_sqrt(double a) { return __dsqrt_rn(a); }
AMB_HD float amb_fdiv(float a, float b) { return __fdiv_rn(a, b); }
#define AMB_NLANES 32
__device__ __forceinline__ double amb_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float amb_warp_maxf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double amb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int amb_warp_min_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ bool amb_warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
__device__ __forceinline__ unsigned amb_warp_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ unsigned amb_lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
__device__ __forceinline__ int amb_popc(unsigned v) { return __popc(v); }
__device__ __forceinline__ int amb_ffs(unsigned v) { return __ffs((int)v); }                    // 1-based, 0 if none
__device__ __forceinline__ double amb_warp_scan_max(double v) {                                  // inclusive prefix maximum over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v, o); if ((int)(threadIdx.x & 31) >= o) v = fmax(v, u); }
    return v;
}
__device__ __forceinline__ double amb_warp_prev(double v, double fill) {                         // value of the previous lane
    const double u = __shfl_up_sync(0xffffffffu, v, 1);
    return (threadIdx.x & 31) ? u : fill;
}
__device__ __forceinline__ double amb_warp_bcast(double v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
#else
#define AMB_NLANES 1
AMB_HD double amb_warp_max(double v) { return v; }
AMB_HD float amb_warp_maxf(float v) { return v; }
AMB_HD double amb_warp_sum(double v) { return v; }
AMB_HD int amb_warp_min_int(int v) { return v; }
AMB_HD bool amb_warp_any(bool p) { return p; }
AMB_HD unsigned amb_warp_ballot(bool p) { return p ? 1u : 0u; }
AMB_HD unsigned amb_lanemask_lt() { return 0u; }
AMB_HD int amb_popc(unsigned v) { return __builtin_popcount(v); }
AMB_HD int amb_ffs(unsigned v) { return __builtin_ffs((int)v); }
AMB_HD double amb_warp_scan_max(double v) { return v; }
AMB_HD double amb_warp_prev(double v, double fill) { (void)v; return fill; }
AMB_HD double amb_warp_bcast(double v, int lane) { (void)lane; return v; }
#define AMB_LANE() 0
#define AMB_WARP_FOR(i, n) for (int i = 0; i < (int)(n); i++)
#define AMB_WARP_SYNC() ((void)0)
#define AMB_LEADER() (true)
AMB_HD double amb_mul(double a, double b) { return a * b; }
AMB_HD double amb_add(double a, double b) { return a + b; }
AMB_HD double amb_sub(double a, double b) { return a - b; }
AMB_HD double amb_div(double a, double b) { return a / b; }
AMB_HD double amb_sqrt(double a) { return sqrt(a); }
AMB_HD float amb_fdiv(float a, float b) { return a / b; }
#endif

// ------------------------------------------------------------------------------------------------
// Sample transform applied when a kernel reads the caller's float32 samples.
//   NONE : value as stored
//   PRE  : preprocessSignal, modem.js:213-232:  f32( f32(x - mean) / mx )   (no division if mx <= 1e-6)
//   SCALE: per-frame peak normalise, app.js:918-925:  f32(x / mx)            (only if mx > 1e-6)
// The JS expressions compute in float64 and round once on the Float32Array store; for a quotient of
// two float32 values that equals the correctly rounded float32 division (double rounding is innocuous
// for p=24 -> 53), which is what amb_fdiv does.
// ------------------------------------------------------------------------------------------------
struct AmbXform { int kind; float mx; double mean; };
enum { AMB_XF_NONE = 0, AMB_XF_PRE = 1, AMB_XF_SCALE = 2 };

AMB_HD float amb_xform(const AmbXform& xf, float x) {
    if (xf.kind == AMB_XF_NONE) return x;
    if (xf.kind == AMB_XF_PRE) {
        const float c = (float)((double)x - xf.mean);
        return xf.mx > 1e-6f ? amb_fdiv(c, xf.mx) : c;
    }
    return xf.mx > 1e-6f ? amb_fdiv(x, xf.mx) : x;
}

// A readable span of a float32 sample stream: global position g is p[g - base] for lo <= g < hi, and 0
// outside (the reference's ring buffer reads unwritten zeros before position 0, app.js:584-587).
struct AmbSig { const float* p; int64_t base, lo, hi; AmbXform xf; };
AMB_HD float amb_sig(const AmbSig& s, int64_t g) {
    if (g < s.lo || g >= s.hi) return 0.0f;
    return amb_xform(s.xf, s.p[g - s.base]);
}

// ------------------------------------------------------------------------------------------------
// Faithful 512-point forward transform of one real symbol (modem.js:6-13, :26-66), one warp.
// re/im: 512 doubles each (shared memory).  On return bins sub_start..sub_end of re/im hold the
// spectrum; other entries are scratch.  Butterflies whose twiddle is exactly (1,0) (j == 0) skip the
// multiplication: 1*x - 0*y == x in IEEE arithmetic up to the sign of a zero, which nothing downstream
// observes.  The last stage computes only the `i1` outputs of the bins the demodulator reads.
// ------------------------------------------------------------------------------------------------
AMB_HD int amb_rev9(int i) {
    int r = 0;
#pragma unroll
    for (int b = 0; b < 9; b++) { r = (r << 1) | (i & 1); i >>= 1; }
    return r;
}

AMB_HD void amb_fft512_load(const AmbSig& s, int64_t g0, double* re, double* im) {
    AMB_WARP_FOR(i, AMB_FFT) {
        float v = amb_sig(s, g0 + i);
        if (v != v) v = 0.0f;   // `x || 0`: NaN -> 0 (modem.js:377, :425)
        const int r = amb_rev9(i);
        re[r] = (double)v;
        im[r] = 0.0;
    }
    AMB_WARP_SYNC();
}

AMB_HD void amb_fft512_faithful(double* re, double* im, const double* tw_re, const double* tw_im, int k_lo, int k_hi) {
    int lg = 0;
    for (int half = 1; half < AMB_HALF; half <<= 1, lg++) {
        AMB_WARP_FOR(b, AMB_HALF) {
            const int j = b & (half - 1);
            const int i1 = ((b >> lg) << (lg + 1)) + j, i2 = i1 + half;
            double tr, ti;
            if (j == 0) { tr = re[i2]; ti = im[i2]; }
            else {
                const double wr = tw_re[half + j], wi = tw_im[half + j];
                tr = amb_sub(amb_mul(wr, re[i2]), amb_mul(wi, im[i2]));
                ti = amb_add(amb_mul(wr, im[i2]), amb_mul(wi, re[i2]));
            }
            const double ar = re[i1], ai = im[i1];
            re[i2] = amb_sub(ar, tr); im[i2] = amb_sub(ai, ti);
            re[i1] = amb_add(ar, tr); im[i1] = amb_add(ai, ti);
        }
        AMB_WARP_SYNC();
    }
    AMB_WARP_FOR(jj, k_hi - k_lo + 1) {
        const int j = k_lo + jj, i2 = j + AMB_HALF;
        double tr, ti;
        if (j == 0) { tr = re[i2]; ti = im[i2]; }
        else {
            const double wr = tw_re[AMB_HALF + j], wi = tw_im[AMB_HALF + j];
            tr = amb_sub(amb_mul(wr, re[i2]), amb_mul(wi, im[i2]));
            ti = amb_add(amb_mul(wr, im[i2]), amb_mul(wi, re[i2]));
        }
        re[j] = amb_add(re[j], tr); im[j] = amb_add(im[j], ti);
    }
    AMB_WARP_SYNC();
}

// estimateChannel, modem.js:431-438 — spectrum in re/im, H written to h_re/h_im (indexed by bin, zero
// outside the band is the caller's job)
AMB_HD void amb_channel_from_spectrum(const double* re, const double* im, const double* known_re, const double* known_im,
                                      int k_lo, int k_hi, double* h_re, double* h_im) {
    AMB_WARP_FOR(kk, k_hi - k_lo + 1) {
        const int k = k_lo + kk;
        const double xr = known_re[k], xi = known_im ? known_im[k] : 0.0;
        const double d = amb_add(amb_mul(xr, xr), amb_mul(xi, xi));
        double hr = 0.0, hi = 0.0;
        if (d > 1e-10) {
            hr = amb_div(amb_add(amb_mul(re[k], xr), amb_mul(im[k], xi)), d);
            hi = amb_div(amb_sub(amb_mul(im[k], xr), amb_mul(re[k], xi)), d);
        }
        h_re[k] = hr; h_im[k] = hi;
    }
    AMB_WARP_SYNC();
}

// constellationDemap, modem.js:140-150.  Per-axis slicing decides whenever the point is at least 1e-9
// away from every decision boundary (there the exhaustive float64 search provably picks the same index);
// closer than that, the exhaustive strict-'<' search of the reference runs.
AMB_HD int amb_demap_exhaustive(const AmbTables* T, double re, double im) {
    double best = INFINITY;
    int bi = 0;
    for (int i = 0; i < T->n_points; i++) {
        const double dr = amb_sub(re, T->pts[2 * i]), di = amb_sub(im, T->pts[2 * i + 1]);
        const double d = amb_add(amb_mul(dr, dr), amb_mul(di, di));
        if (d < best) { best = d; bi = i; }
    }
    return bi;
}

AMB_HD int amb_demap(const AmbTables* T, double re, double im) {
    const double tau = 1e-9;
    if (!(fabs(re) < 1e30) || !(fabs(im) < 1e30)) return amb_demap_exhaustive(T, re, im);   // NaN / inf
    if (T->mod == AMB_MOD_BPSK) {
        if (fabs(re) < tau) return amb_demap_exhaustive(T, re, im);
        return re > 0 ? 0 : 1;
    }
    if (T->mod == AMB_MOD_QPSK) {
        if (fabs(re) < tau || fabs(im) < tau) return amb_demap_exhaustive(T, re, im);
        return re > 0 ? (im > 0 ? 0 : 3) : (im > 0 ? 1 : 2);
    }
    const double thr = T->qam_thr;
    const double are = fabs(re), aim = fabs(im);
    if (are < tau || aim < tau || fabs(are - thr) < tau || fabs(aim - thr) < tau) return amb_demap_exhaustive(T, re, im);
    // level index g in {0:-3, 1:-1, 2:+1, 3:+3}; index bits are the Gray inverse: g -> {0,1,3,2}
    const int gc = re < 0 ? (are > thr ? 0 : 1) : (are > thr ? 3 : 2);
    const int gr = im < 0 ? (aim > thr ? 0 : 1) : (aim > thr ? 3 : 2);
    const int col = gc ^ (gc >> 1), row = gr ^ (gr >> 1);
    return (row << 2) | col;
}

// One data symbol after the transform: equalise (modem.js:386-395), pilot phase (:398-405), de-rotate
// and demap (:408-413).  re/im hold the spectrum on entry and the equalised bins on exit.
// q_scratch: >= 2*AMB_MAX_PILOTS doubles of shared memory.  idx_out[di] receives the constellation index.
AMB_HD void amb_symbol_demap(const AmbTables* T, double* re, double* im, const double* h_re, const double* h_im,
                             double* q_scratch, uint8_t* idx_out, uint8_t* bits_out, double* eq_out) {
    AMB_WARP_FOR(kk, T->sub_end - T->sub_start + 1) {
        const int k = T->sub_start + kk;
        const double hr = h_re[k], hi = h_im[k];
        const double hm = amb_add(amb_mul(hr, hr), amb_mul(hi, hi));
        if (hm > 1e-10) {
            const double sr = re[k], si = im[k];
            re[k] = amb_div(amb_add(amb_mul(sr, hr), amb_mul(si, hi)), hm);
            im[k] = amb_div(amb_sub(amb_mul(si, hr), amb_mul(sr, hi)), hm);
        }
    }
    AMB_WARP_SYNC();
    AMB_WARP_FOR(p, T->n_pilots) {
        const int k = T->pilots[p];
        const bool ok = k >= T->sub_start && k <= T->sub_end && fabs(re[k]) > 1e-6;
        q_scratch[2 * p] = ok ? amb_div(im[k], re[k]) : 0.0;
        q_scratch[2 * p + 1] = ok ? 1.0 : 0.0;
    }
    AMB_WARP_SYNC();
    double phase_sum = 0.0;
    int pc = 0;
    for (int p = 0; p < T->n_pilots; p++)
        if (q_scratch[2 * p + 1] != 0.0) { phase_sum = amb_add(phase_sum, q_scratch[2 * p]); pc++; }
    const double phase = pc > 0 ? amb_div(phase_sum, (double)pc) : 0.0;
    AMB_WARP_FOR(di, T->n_data) {
        const int k = T->data_k[di];
        const double cr = amb_add(re[k], amb_mul(im[k], phase));
        const double ci = amb_sub(im[k], amb_mul(re[k], phase));
        const int idx = amb_demap(T, cr, ci);
        if (idx_out) idx_out[di] = (uint8_t)idx;
        if (bits_out) for (int b = 0; b < T->bps; b++) bits_out[di * T->bps + b] = (uint8_t)((idx >> (T->bps - 1 - b)) & 1);
        if (eq_out) { eq_out[2 * di] = cr; eq_out[2 * di + 1] = ci; }
    }
    AMB_WARP_SYNC();
}

// ------------------------------------------------------------------------------------------------
// Bits -> bytes -> parsed frame.  `idx` holds one constellation index per data subcarrier in serial
// order; raw bit b is bit (bps-1 - b%bps) of idx[b/bps] (MSB first, modem.js:148).  majorityVote
// (modem.js:487-495: sum >= n/2) and bitsToBytes (:468-476) are applied on the fly.
// ------------------------------------------------------------------------------------------------
AMB_HD int amb_raw_bit(const uint8_t* idx, int bps, int64_t b) {
    const int64_t e = b / bps;
    return (idx[e] >> (bps - 1 - (int)(b - e * bps))) & 1;
}

AMB_HD int64_t amb_num_bytes(int64_t n_elems, int bps, int rep) {
    const int64_t raw = n_elems * bps;
    const int64_t voted = rep > 1 ? raw / rep : raw;
    return voted / 8;
}

// decoded byte j (majority vote + MSB-first packing); rep == 1 reads whole constellation indices
AMB_HD uint8_t amb_pack_byte(const uint8_t* idx, int bps, int rep, int64_t j) {
    if (rep == 1) {
        if (bps == 4) return (uint8_t)((idx[2 * j] << 4) | (idx[2 * j + 1] & 15));
        if (bps == 2) return (uint8_t)((idx[4 * j] << 6) | ((idx[4 * j + 1] & 3) << 4) | ((idx[4 * j + 2] & 3) << 2) | (idx[4 * j + 3] & 3));
        int v = 0;
        for (int t = 0; t < 8; t++) v = (v << 1) | (idx[8 * j + t] & 1);
        return (uint8_t)v;
    }
    int v = 0;
    if (bps == 1) {
        // BPSK with repetition (the acoustic modes): raw bit b is idx[b]
        for (int t = 0; t < 8; t++) {
            const int64_t b0 = (j * 8 + t) * rep;
            int sum = 0;
            for (int r = 0; r < rep; r++) sum += idx[b0 + r] & 1;
            v = (v << 1) | ((2 * sum >= rep) ? 1 : 0);
        }
        return (uint8_t)v;
    }
    for (int t = 0; t < 8; t++) {
        const int64_t vb = j * 8 + t;
        int sum = 0;
        for (int r = 0; r < rep; r++) sum += amb_raw_bit(idx, bps, vb * rep + r);
        v = (v << 1) | ((2 * sum >= rep) ? 1 : 0);
    }
    return (uint8_t)v;
}

AMB_HD void amb_pack_bytes(const uint8_t* idx, int64_t n_elems, int bps, int rep, uint8_t* bytes, int64_t cap) {
    int64_t nb = amb_num_bytes(n_elems, bps, rep);
    if (nb > cap) nb = cap;
    AMB_WARP_FOR(j, nb) bytes[j] = amb_pack_byte(idx, bps, rep, j);
    AMB_WARP_SYNC();
}

AMB_HD uint32_t amb_crc32_bytes(const uint32_t* table, const uint8_t* data, int64_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (int64_t i = 0; i < n; i++) c = table[(c ^ data[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

// CRC-32 over a byte range split into chunks (parallel evaluation).  The reflected CRC register is a polynomial
// over GF(2) with bit 31 = x^0; for the finalised values crc(A||B) = crc(A) * x^(8|B|) + crc(B) (mod P), so chunk
// CRCs combine by a multiplication with a power of x.  Chunks are aligned to the END of the range (the first chunk
// absorbs the remainder), so every right operand of the combining tree has the full length 2^level * L.
AMB_HD uint32_t amb_gf_mul(uint32_t a, uint32_t b) {          // a * b mod P
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}

AMB_HD uint32_t amb_gf_xpow8(uint32_t n_bytes) {               // x^(8 n) mod P
    uint32_t r = 0x80000000u, sq = 0x00800000u;                // 1, x^8
    while (n_bytes) {
        if (n_bytes & 1u) r = amb_gf_mul(r, sq);
        sq = amb_gf_mul(sq, sq);
        n_bytes >>= 1;
    }
    return r;
}

// chunk t of nt covers [lo, hi) of an n-byte range, chunk length L = ceil(n / nt), aligned to the end
AMB_HD void amb_crc_chunk(int64_t n, int nt, int t, int64_t* lo, int64_t* hi) {
    const int64_t L = (n + nt - 1) / nt;
    int64_t h = n - (int64_t)(nt - 1 - t) * L, l = h - L;
    if (h < 0) h = 0;
    if (l < 0) l = 0;
    *lo = l; *hi = h;
}

AMB_HD int32_t amb_be32(const uint8_t* b) {   // JS `(b0<<24)|(b1<<16)|(b2<<8)|b3` is a signed int32
    return (int32_t)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
}

AMB_HD void amb_result_clear(amb_frame_result* r) {
    r->status = 0; r->frame_type = -1; r->preamble_idx = -1; r->seq_num = 0; r->data_len = 0; r->total_chunks = 0;
    r->total_file_size = 0; r->chunk_size = 0; r->name_len = 0; r->name_off = 0; r->data_off = 0; r->expected_crc = 0;
    r->actual_crc = 0; r->crc_valid = 0; r->n_bytes = 0; r->err_arg = 0; r->stream = 0; r->coarse_metric = 0; r->fine_metric = 0;
}

// parseMetadataResult, modem.js:805-828
// The parsers take an optional `defer`: when non-NULL the CRC is left to the caller (a CTA-wide parallel CRC): *defer
// receives the number of leading bytes the CRC covers (-1 if the parser failed before reaching it).
AMB_HD void amb_finish_crc(const uint32_t* crc_table, const uint8_t* bytes, int64_t off, amb_frame_result* r, int64_t* defer) {
    if (defer) { *defer = off; return; }
    r->actual_crc = amb_crc32_bytes(crc_table, bytes, off);
    r->crc_valid = r->expected_crc == r->actual_crc;
}

AMB_HD void amb_parse_meta(const uint32_t* crc_table, const uint8_t* bytes, int64_t n, amb_frame_result* r, int64_t* defer = nullptr) {
    r->frame_type = AMB_FRAME_META;
    if (n < 16) { r->status = AMB_ST_META_TOO_SHORT; r->frame_type = -1; return; }
    int64_t off = 1;
    r->total_chunks = amb_be32(bytes + off); off += 4;
    r->total_file_size = amb_be32(bytes + off); off += 4;
    r->chunk_size = (bytes[off] << 8) | bytes[off + 1]; off += 2;
    const int name_len = bytes[off++];
    if (off + name_len + 4 > n) { r->status = AMB_ST_META_TRUNCATED; r->frame_type = -1; return; }
    r->name_len = name_len; r->name_off = (int32_t)off;
    off += name_len;
    r->expected_crc = (uint32_t)amb_be32(bytes + off);
    amb_finish_crc(crc_table, bytes, off, r, defer);
}

// parseDataChunkResult, modem.js:830-849
AMB_HD void amb_parse_chunk(const uint32_t* crc_table, const uint8_t* bytes, int64_t n, amb_frame_result* r, int64_t* defer = nullptr) {
    r->frame_type = AMB_FRAME_DATA;
    if (n < 11) { r->status = AMB_ST_CHUNK_TOO_SHORT; r->frame_type = -1; return; }
    int64_t off = 1;
    r->seq_num = amb_be32(bytes + off); off += 4;
    const int data_len = (bytes[off] << 8) | bytes[off + 1]; off += 2;
    if (off + data_len + 4 > n) { r->status = AMB_ST_CHUNK_TRUNCATED; r->frame_type = -1; return; }
    r->data_len = data_len; r->data_off = (int32_t)off;
    off += data_len;
    r->expected_crc = (uint32_t)amb_be32(bytes + off);
    amb_finish_crc(crc_table, bytes, off, r, defer);
}

// tail of decodeChunkFrame, modem.js:793-802
AMB_HD void amb_parse_chunk_frame(const uint32_t* crc_table, const uint8_t* bytes, int64_t n, amb_frame_result* r, int64_t* defer = nullptr) {
    r->n_bytes = (int32_t)n;
    if (defer) *defer = -1;
    if (n < 6) { r->status = AMB_ST_DECODED_TOO_SHORT; return; }
    const int ft = bytes[0];
    if (ft == AMB_FRAME_META) amb_parse_meta(crc_table, bytes, n, r, defer);
    else if (ft == AMB_FRAME_DATA) amb_parse_chunk(crc_table, bytes, n, r, defer);
    else { r->status = AMB_ST_UNKNOWN_FRAME_TYPE; r->err_arg = ft; r->frame_type = ft; }
}

// tail of decodeReceivedSignal, modem.js:607-653 (preamble_idx is filled by the caller)
AMB_HD void amb_parse_received(const uint32_t* crc_table, const uint8_t* bytes, int64_t n, amb_frame_result* r, int64_t* defer = nullptr) {
    r->n_bytes = (int32_t)n;
    if (defer) *defer = -1;
    if (n < 10) { r->status = AMB_ST_DECODED_TOO_SHORT; return; }
    const int first = bytes[0];
    if (first == AMB_FRAME_META) { amb_parse_meta(crc_table, bytes, n, r, defer); return; }
    if (first == AMB_FRAME_DATA) { amb_parse_chunk(crc_table, bytes, n, r, defer); return; }
    int64_t off = 0;
    const int name_len = bytes[off++];
    if (off + name_len + 4 + 4 > n) { r->status = AMB_ST_DECODED_TOO_SHORT_FOR_HEADER; return; }
    r->name_len = name_len; r->name_off = (int32_t)off;
    off += name_len;
    const int32_t data_len = amb_be32(bytes + off);
    off += 4;
    if (data_len <= 0 || off + (int64_t)data_len + 4 > n) { r->status = AMB_ST_INVALID_DATA_LENGTH; r->err_arg = data_len; return; }
    r->data_len = data_len; r->data_off = (int32_t)off;
    off += data_len;
    r->expected_crc = (uint32_t)amb_be32(bytes + off);
    amb_finish_crc(crc_table, bytes, off, r, defer);
    r->frame_type = AMB_FRAME_LEGACY;
}

#if defined(__CUDACC__)
// Exact float64 window sums of 256 consecutive windows from 768 samples held by one warp: lane l holds samples
// 256 seg + 8 l + k (seg < 3, k < 8) of the span that starts at the first window; on return oP/oA/oB[k] are P, Ra, Rb of
// window 8 l + k.  Exclusive float64 prefixes of x^2 and x[i] x[i+256] inside each 256-sample segment (registers + one warp
// scan per segment and array); the window starting at 8 l + k needs exactly the lane's own prefix values.
__device__ __forceinline__ double amb_warp_excl_scan(double v, int lane, double* total) {
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}
__device__ __forceinline__ void amb_tile_sums_regs(const float (&x)[3][8], int lane, double* oP, double* oA, double* oB) {
    double q[3][8], t[2][8], qb[3], tb[2];       // exclusive prefixes inside the segment; global base of the segment
    double run_q = 0.0, run_t = 0.0;
#pragma unroll
    for (int seg = 0; seg < 3; seg++) {
        double lq = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) { q[seg][k] = lq; lq = fma((double)x[seg][k], (double)x[seg][k], lq); }
        double tot;
        const double off = amb_warp_excl_scan(lq, lane, &tot);
#pragma unroll
        for (int k = 0; k < 8; k++) q[seg][k] += off;
        qb[seg] = run_q; run_q += tot;
        if (seg < 2) {
            double lt = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) { t[seg][k] = lt; lt = fma((double)x[seg][k], (double)x[seg + 1][k], lt); }
            const double offt = amb_warp_excl_scan(lt, lane, &tot);
#pragma unroll
            for (int k = 0; k < 8; k++) t[seg][k] += offt;
            tb[seg] = run_t; run_t += tot;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // prefix S[256 seg + w] = base[seg] + prefix[seg][k],  w = 8 lane + k
        oP[k] = (tb[1] + t[1][k]) - (tb[0] + t[0][k]);
        oA[k] = (qb[1] + q[1][k]) - (qb[0] + q[0][k]);
        oB[k] = (qb[2] + q[2][k]) - (qb[1] + q[1][k]);
    }
}
#endif

#if defined(__CUDACC__)
// cnt > 8 consecutive windows from d0, in blocks of 256: lane l owns the 8 windows 8 l .. 8 l + 7 of a block and needs the
// exclusive float64 prefixes of x^2 (three 256-sample segments) and x[i] x[i+256] (two) at its first sample.  Pass 1 sums
// the lane's own 8 samples per segment, five warp scans turn the sums into those prefixes, pass 2 walks the 8 samples again
// (reloaded: they sit in L1) with five running sums and writes the window sums.  Few live registers on purpose: the walker
// kernels that inline this are register-bound.
__device__ __forceinline__ void amb_ws_load4(const float* sp, int64_t sbase, int64_t slo, int64_t shi, bool inside, const float* pa, int64_t g0, int k0,
                                             float (&x0)[4], float (&x1)[4], float (&x2)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (inside) { x0[k] = pa[k0 + k]; x1[k] = pa[256 + k0 + k]; x2[k] = pa[512 + k0 + k]; }
        else {
            const int64_t g = g0 + k0 + k;
            x0[k] = (g >= slo && g < shi) ? sp[g - sbase] : 0.f;
            x1[k] = (g + 256 >= slo && g + 256 < shi) ? sp[g + 256 - sbase] : 0.f;
            x2[k] = (g + 512 >= slo && g + 512 < shi) ? sp[g + 512 - sbase] : 0.f;
        }
    }
}
__device__ __forceinline__ void amb_window_sums_blocks(const float* sp, int64_t sbase, int64_t slo, int64_t shi, int64_t d0, int cnt, double* P, double* A, double* B) {
    const int lane = threadIdx.x & 31;
    for (int b0 = 0; b0 < cnt; b0 += 256) {
        const int64_t g0 = d0 + b0 + 8 * lane;                     // this lane's first sample in segment 0
        const bool inside = d0 + b0 >= slo && d0 + b0 + 768 <= shi;
        const float* pa = sp + (g0 - sbase);
        double q0 = 0.0, q1 = 0.0, q2 = 0.0, t0 = 0.0, t1 = 0.0;
        // (the samples are fetched twelve at a time: the loads of a group are in flight together)
#pragma unroll 1
        for (int k0 = 0; k0 < 8; k0 += 4) {
            float x0[4], x1[4], x2[4];
            amb_ws_load4(sp, sbase, slo, shi, inside, pa, g0, k0, x0, x1, x2);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                q0 = fma((double)x0[k], (double)x0[k], q0); q1 = fma((double)x1[k], (double)x1[k], q1); q2 = fma((double)x2[k], (double)x2[k], q2);
                t0 = fma((double)x0[k], (double)x1[k], t0); t1 = fma((double)x1[k], (double)x2[k], t1);
            }
        }
        // exclusive prefixes over the lanes; segment totals of x^2 (segment 0, 1) and of the cross products (segment 0)
        double e0 = q0, e1 = q1, e2 = q2, f0 = t0, f1 = t1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u0 = __shfl_up_sync(0xffffffffu, e0, o), u1 = __shfl_up_sync(0xffffffffu, e1, o), u2 = __shfl_up_sync(0xffffffffu, e2, o);
            const double v0 = __shfl_up_sync(0xffffffffu, f0, o), v1 = __shfl_up_sync(0xffffffffu, f1, o);
            if (lane >= o) { e0 += u0; e1 += u1; e2 += u2; f0 += v0; f1 += v1; }
        }
        const double Q0 = __shfl_sync(0xffffffffu, e0, 31), Q1 = __shfl_sync(0xffffffffu, e1, 31), T0 = __shfl_sync(0xffffffffu, f0, 31);
        // running prefixes at the lane's first window: S_q[256 seg + w], S_t[256 seg + w]
        double pq0 = e0 - q0, pq1 = Q0 + (e1 - q1), pq2 = Q0 + Q1 + (e2 - q2), pt0 = f0 - t0, pt1 = T0 + (f1 - t1);
#pragma unroll 1
        for (int k0 = 0; k0 < 8; k0 += 4) {
            float x0[4], x1[4], x2[4];
            amb_ws_load4(sp, sbase, slo, shi, inside, pa, g0, k0, x0, x1, x2);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = b0 + 8 * lane + k0 + k;
                if (i < cnt) { P[i] = pt1 - pt0; A[i] = pq1 - pq0; B[i] = pq2 - pq1; }
                pq0 = fma((double)x0[k], (double)x0[k], pq0); pq1 = fma((double)x1[k], (double)x1[k], pq1); pq2 = fma((double)x2[k], (double)x2[k], pq2);
                pt0 = fma((double)x0[k], (double)x1[k], pt0); pt1 = fma((double)x1[k], (double)x2[k], pt1);
            }
        }
    }
    __syncwarp();
}
#endif

// ------------------------------------------------------------------------------------------------
// Schmidl-Cox window sums at `cnt` consecutive offsets d0.. (modem.js:292-298 evaluated directly at
// every offset, float64, exact products of float32 samples): P = sum s[d+m]s[d+m+256],
// Ra = sum s[d+m]^2, Rb = sum s[d+m+256]^2, m < 256.  One warp; each lane takes chunks of 8 offsets
// (direct sum for the first, slide for the rest).
// ------------------------------------------------------------------------------------------------
AMB_HD void amb_window_sums(const AmbSig& s, int64_t d0, int cnt, double* P, double* A, double* B) {
    if (cnt <= 8) {
        // few offsets: the lanes share the direct sum of the first window, the leader slides
        double p = 0, a = 0, b = 0;
        AMB_WARP_FOR(m, AMB_HALF) {
            const double x = (double)amb_sig(s, d0 + m), y = (double)amb_sig(s, d0 + m + AMB_HALF);
            p = fma(x, y, p); a = fma(x, x, a); b = fma(y, y, b);
        }
        p = amb_warp_sum(p); a = amb_warp_sum(a); b = amb_warp_sum(b);
        if (AMB_LEADER()) {
            P[0] = p; A[0] = a; B[0] = b;
            for (int t = 1; t < cnt; t++) {
                const double out = (double)amb_sig(s, d0 + t - 1), mid = (double)amb_sig(s, d0 + t - 1 + AMB_HALF),
                             in = (double)amb_sig(s, d0 + t - 1 + 2 * AMB_HALF);
                p += mid * in - out * mid; a += mid * mid - out * out; b += in * in - mid * mid;
                P[t] = p; A[t] = a; B[t] = b;
            }
        }
        AMB_WARP_SYNC();
        return;
    }
#if defined(__CUDA_ARCH__) && !defined(AMB_NO_FAST_WS)
    if (s.xf.kind == AMB_XF_NONE) { amb_window_sums_blocks(s.p, s.base, s.lo, s.hi, d0, cnt, P, A, B); return; }
#endif
    const int chunks = (cnt + 7) / 8;
    AMB_WARP_FOR(c, chunks) {
        const int i0 = c * 8;
        const int64_t d = d0 + i0;
        double p = 0, a = 0, b = 0;
        for (int m = 0; m < AMB_HALF; m++) {
            const double x = (double)amb_sig(s, d + m), y = (double)amb_sig(s, d + m + AMB_HALF);
            p = fma(x, y, p); a = fma(x, x, a); b = fma(y, y, b);
        }
        P[i0] = p; A[i0] = a; B[i0] = b;
        for (int t = 1; t < 8 && i0 + t < cnt; t++) {
            const double out = (double)amb_sig(s, d + t - 1), mid = (double)amb_sig(s, d + t - 1 + AMB_HALF),
                         in = (double)amb_sig(s, d + t - 1 + 2 * AMB_HALF);
            p += mid * in - out * mid; a += mid * mid - out * out; b += in * in - mid * mid;
            P[i0 + t] = p; A[i0 + t] = a; B[i0 + t] = b;
        }
    }
    AMB_WARP_SYNC();
}

// Cross-correlation metric of preamble-1 at one offset d (modem.js:577-585): float64 sums in tap order, so the
// value is bit-identical to the reference's given identical samples.  Returns false where the denominator fails
// the 0.001 test (the reference skips such offsets).
AMB_HD bool amb_xcorr_at(const AmbTables* T, const AmbSig& s, int64_t d, double* metric) {
    double corr = 0, se = 0;
    const int plen = T->sym_len;
    if (s.xf.kind == AMB_XF_NONE && d >= s.lo && d + plen <= s.hi) {
        // whole window readable as stored: no per-tap bounds test, loads can run ahead of the two sum chains
        const float* p = s.p + (d - s.base);
        const float* t = T->pre1;
#pragma unroll 8
        for (int i = 0; i < plen; i++) {
            const double v = (double)p[i];
            corr = fma(v, (double)t[i], corr);     // the product is exact, so fma == mul then add
            se = fma(v, v, se);
        }
    } else {
        for (int i = 0; i < plen; i++) {
            const double v = (double)amb_sig(s, d + i);
            corr = fma(v, (double)T->pre1[i], corr);
            se = fma(v, v, se);
        }
    }
    const double denom = amb_sqrt(amb_mul(se, T->pre1_energy));
    if (!(denom > 0.001)) { *metric = 0.0; return false; }
    *metric = amb_div(corr, denom);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Fine cross-correlation against preamble-1 over offsets [lo, hi] (modem.js:576-587, app.js:864-877):
// float64 sequential sums in tap order (bit-identical to the reference given identical samples),
// metric = corr / sqrt(sEnergy * tEnergy) where the denominator exceeds 0.001, strict-'>' argmax from
// -Infinity.  One warp; lane l takes offsets lo+l, lo+l+32, ...   scratch: 64 doubles.
// Returns best metric (-inf if none) and writes its offset (or `fallback`).
// ------------------------------------------------------------------------------------------------
AMB_HD double amb_fine_xcorr(const AmbTables* T, const AmbSig& s, int64_t lo, int64_t hi, int64_t fallback,
                             double* scratch, int64_t* pos_out) {
    const int plen = T->sym_len;
    const int n = hi >= lo ? (int)(hi - lo + 1) : 0;
#if defined(__CUDA_ARCH__)
    const int lane0 = AMB_LANE(), lane1 = lane0 + 1;
#else
    const int lane0 = 0, lane1 = 32;
#endif
    for (int lane = lane0; lane < lane1; lane++) {
        double best = -INFINITY;
        int64_t bpos = -1;
        for (int o = lane; o < n; o += 32) {
            const int64_t d = lo + o;
            double m;
            if (amb_xcorr_at(T, s, d, &m) && m > best) { best = m; bpos = d; }
        }
        scratch[lane] = best;
        scratch[32 + lane] = (double)bpos;
    }
    AMB_WARP_SYNC();
    double best = -INFINITY;
    int64_t bpos = fallback;
    for (int l = 0; l < 32; l++) {
        const double m = scratch[l];
        const int64_t p = (int64_t)scratch[32 + l];
        if (p < 0) continue;
        if (m > best || (m == best && p < bpos)) { best = m; bpos = p; }
    }
    AMB_WARP_SYNC();
    *pos_out = bpos;
    return best;
}
