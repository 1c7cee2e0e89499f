// Host-side generation of the constant tables (compiled with -ffp-contract=off).
// Restates, in float64 and in the reference's operation order, the pieces of modem.js whose outputs are
// constants of a configuration: the radix-2 twiddle recurrence (modem.js:26-47), the float64-rounded LCG
// (modem.js:153-156), the preamble / channel-estimation symbols (modem.js:158-208), the constellations
// (modem.js:107-131) and the CRC-32 table (modem.js:443-451).
#include "amb_tables.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct CfgRow { int cp, sym, s0, s1, np; int pil[AMB_MAX_PILOTS]; };
const CfgRow kCfg[3] = {
    {64, 576, 12, 232, 16, {15, 29, 43, 57, 71, 85, 99, 113, 127, 141, 155, 169, 183, 197, 211, 225}},
    {128, 640, 23, 93, 7, {25, 35, 45, 55, 65, 75, 85}},
    {256, 768, 35, 58, 3, {37, 45, 53}},
};

inline int clamp_cfg(int c) { return (c < 0 || c > 2) ? 0 : c; }   // modem.js:96

// in-place 512-point transform exactly as modem.js:26-66 orders it: permutation, then butterflies whose
// twiddle is the running product w <- w * wn restarted at (1,0) for every group
void js_transform(std::vector<double>& re, std::vector<double>& im, bool inverse) {
    const int n = (int)re.size();
    int bits = 0;
    for (int t = n; t > 1; t >>= 1) bits++;
    for (int i = 0; i < n; i++) {
        int r = 0, x = i;
        for (int b = 0; b < bits; b++) { r = (r << 1) | (x & 1); x >>= 1; }
        if (i < r) { std::swap(re[i], re[r]); std::swap(im[i], im[r]); }
    }
    for (int size = 2; size <= n; size <<= 1) {
        const int half = size >> 1;
        const double angle = (inverse ? 1.0 : -1.0) * 2 * M_PI / size;
        const double cr = std::cos(angle), ci = std::sin(angle);
        for (int base = 0; base < n; base += size) {
            double wr = 1, wi = 0;
            for (int j = 0; j < half; j++) {
                const int a = base + j, b = a + half;
                const double tr = wr * re[b] - wi * im[b];
                const double ti = wr * im[b] + wi * re[b];
                re[b] = re[a] - tr; im[b] = im[a] - ti;
                re[a] += tr; im[a] += ti;
                const double nr = wr * cr - wi * ci;
                wi = wr * ci + wi * cr;
                wr = nr;
            }
        }
    }
    if (inverse) {
        const double s = 1.0 / n;
        for (int i = 0; i < n; i++) { re[i] *= s; im[i] *= s; }
    }
}

// modem.js:155 — the multiply-add runs in float64 (rounded above 2^53), `&` then takes the low 31 bits
struct Lcg {
    double s;
    double draw() {
        double t = s * 1103515245.0;
        t += 12345.0;
        s = (double)(uint32_t)((uint64_t)t & 0x7fffffffu);
        return s / 2147483647.0;
    }
};

// +-1 on bins s0..s1 step `step` from LCG `seed`, Hermitian fill, inverse transform, CP prepend with the
// Float32 store of addCP (modem.js:202-208)
void make_symbol(const CfgRow& c, int seed, int step, float* out, double* known) {
    std::vector<double> re(AMB_FFT, 0.0), im(AMB_FFT, 0.0);
    Lcg g{(double)seed};
    for (int k = c.s0; k <= c.s1; k += step) {
        const double v = g.draw() > 0.5 ? 1.0 : -1.0;
        re[k] = v;
        if (known) known[k] = v;
    }
    for (int k = 1; k < AMB_FFT / 2; k++) { re[AMB_FFT - k] = re[k]; im[AMB_FFT - k] = -im[k]; }
    re[0] = 0; re[AMB_FFT / 2] = 0; im[AMB_FFT / 2] = 0;
    js_transform(re, im, true);
    for (int i = 0; i < c.cp; i++) out[i] = (float)re[AMB_FFT - c.cp + i];
    for (int i = 0; i < AMB_FFT; i++) out[c.cp + i] = (float)re[i];
}

}  // namespace

int amb_tables_build(AmbTables* t, int config, int constellation, int repetition) {
    std::memset(t, 0, sizeof(*t));
    const int cid = clamp_cfg(config);
    const CfgRow& c = kCfg[cid];
    if (constellation < 0 || constellation > 2) return -1;
    t->cfg_id = cid; t->mod = constellation; t->rep = repetition <= 0 ? 1 : repetition;   // `repetition || 1`
    t->bps = constellation == 0 ? 1 : (constellation == 1 ? 2 : 4);
    t->cp_len = c.cp; t->sym_len = c.sym; t->sub_start = c.s0; t->sub_end = c.s1; t->n_pilots = c.np;
    t->fine_radius = 3 * c.cp;
    t->pre1_sign = (c.s0 & 1) ? -1 : 1;
    for (int i = 0; i < c.np; i++) { t->pilots[i] = c.pil[i]; t->is_pilot[c.pil[i]] = 1; }
    int nd = 0;
    for (int k = c.s0; k <= c.s1; k++) if (!t->is_pilot[k]) t->data_k[nd++] = (int16_t)k;
    t->n_data = nd;

    make_symbol(c, 44, 1, t->pre1 /*scratch*/, t->known_re);   // CE known signs (samples discarded)
    std::memset(t->pre1, 0, sizeof(t->pre1));
    make_symbol(c, 42, 2, t->pre1, nullptr);
    double e = 0;
    for (int i = 0; i < c.sym; i++) e += (double)t->pre1[i] * (double)t->pre1[i];
    t->pre1_energy = e;

    make_symbol(c, 43, 1, t->pre2, nullptr);
    make_symbol(c, 44, 1, t->ce_sym, nullptr);

    // constellations
    if (constellation == 0) {
        const double p[4] = {1, 0, -1, 0};
        std::memcpy(t->pts, p, sizeof(p)); t->n_points = 2;
    } else if (constellation == 1) {
        const double s = 1 / M_SQRT2;
        const double p[8] = {s, s, -s, s, -s, -s, s, -s};
        std::memcpy(t->pts, p, sizeof(p)); t->n_points = 4;
    } else {
        double raw[32], avg = 0;
        for (int i = 0; i < 16; i++) {
            const int row = i >> 2, col = i & 3;
            raw[2 * i] = 2 * (col ^ (col >> 1)) - 3;
            raw[2 * i + 1] = 2 * (row ^ (row >> 1)) - 3;
            avg += raw[2 * i] * raw[2 * i] + raw[2 * i + 1] * raw[2 * i + 1];
        }
        avg /= 16;
        const double s = 1 / std::sqrt(avg);
        for (int i = 0; i < 32; i++) t->pts[i] = raw[i] * s;
        t->n_points = 16;
        t->qam_thr = (1.0 * s + 3.0 * s) / 2;
    }

    // forward twiddles by the reference's running product
    for (int half = 1; half <= AMB_HALF; half <<= 1) {
        const double angle = -1.0 * 2 * M_PI / (2 * half);
        const double cr = std::cos(angle), ci = std::sin(angle);
        double wr = 1, wi = 0;
        for (int j = 0; j < half; j++) {
            t->tw_re[half + j] = wr; t->tw_im[half + j] = wi;
            const double nr = wr * cr - wi * ci;
            wi = wr * ci + wi * cr;
            wr = nr;
        }
    }
    for (int half = 1; half <= AMB_HALF; half <<= 1) {
        const double angle = 1.0 * 2 * M_PI / (2 * half);
        const double cr = std::cos(angle), ci = std::sin(angle);
        double wr = 1, wi = 0;
        for (int j = 0; j < half; j++) {
            t->twi_re[half + j] = wr; t->twi_im[half + j] = wi;
            const double nr = wr * cr - wi * ci;
            wi = wr * ci + wi * cr;
            wr = nr;
        }
    }
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t v = i;
        for (int j = 0; j < 8; j++) v = (v & 1) ? (0xEDB88320u ^ (v >> 1)) : (v >> 1);
        t->crc_table[i] = v;
    }
    t->dc_alpha = 0.999;
    t->dc_beta = 1 - t->dc_alpha;   // 0.0010000000000000009 in float64 (app.js:753)
    return 0;
}

int64_t amb_tables_estimate_frame_samples(int config, int constellation, int64_t payload_bytes, int repetition) {
    const CfgRow& c = kCfg[clamp_cfg(config)];
    if (repetition <= 0) repetition = 1;
    int np = 0;
    for (int k = c.s0; k <= c.s1; k++) { bool p = false; for (int i = 0; i < c.np; i++) p |= c.pil[i] == k; if (!p) np++; }
    const int bps = constellation == 0 ? 1 : (constellation == 2 ? 4 : 2);
    const double total_bits = (double)payload_bytes * 8 * repetition;
    const int64_t nsym = (int64_t)std::ceil(total_bits / (np * bps));
    return (3 + nsym) * c.sym;
}

void amb_tables_preamble2(int config, float* out) { make_symbol(kCfg[clamp_cfg(config)], 43, 1, out, nullptr); }

void amb_tables_ce_symbol(int config, float* out, double* known_re) {
    double known[AMB_FFT];
    std::memset(known, 0, sizeof(known));
    make_symbol(kCfg[clamp_cfg(config)], 44, 1, out, known);
    if (known_re) std::memcpy(known_re, known, sizeof(known));
}
