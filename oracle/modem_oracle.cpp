// ============================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (float64, scalar, single thread) of the OFDM modem algorithm
// of playok/audio-modem: `modem.js` (FFT, configs, constellations, preamble/CE
// generators, preprocess, both preamble detectors, modulator, demodulator,
// channel estimator, CRC-32, bit helpers, frame builders, decoders/parsers)
// and the DOM-free part of `app.js`'s StreamingReceiver (ring buffer, EMA DC
// removal, incremental Schmidl-Cox scan, refine, collect, decode).
// Every function cites the reference file:line it follows.
//
// PINNED against the reference's own outputs: the reference ships no tests,
// golden vectors or fixtures, so modem.js and the receiver classes of app.js
// are evaluated as they are by Qt's QJSEngine (the ECMAScript engine inside
// Nsight Compute's host directory; tools/qjs/), and what they return on the
// golden schema and on the 141 cases of tests/refcases.py is committed under
// tests/golden/reference_golden.json / reference_cases.json (1,919 values).
// This restatement reproduces every one of them bit for bit, floats included
// (tests/test_golden.py).  The known-answer checks of tests/test_oracle_kat.py
// (CRC-32 vs zlib, the float64-rounded LCG, Gray 16-QAM, round trips) remain.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  Build with
//   g++ -O2 -ffp-contract=off -fno-fast-math   (see oracle/Makefile)
// so that no FMA contraction changes the float64 rounding sequence.
// ============================================================================
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

typedef std::vector<double> dvec;
typedef std::vector<float> fvec;
typedef std::vector<uint8_t> bvec;

// ---------------------------------------------------------------- FFT -------
// modem.js:62-66 revBits
static inline int rev_bits(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

// modem.js:49-60 bitReverse
static void bit_reverse(dvec& re, dvec& im) {
    const int n = (int)re.size();
    int bits = 0, tmp = n;
    while (tmp > 1) { bits++; tmp >>= 1; }
    for (int i = 0; i < n; i++) {
        int j = rev_bits(i, bits);
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
}

// modem.js:26-47 fftIterative.  Radix-2 DIT; per stage the twiddle starts at
// (1,0) for every group and advances by complex multiplication with
// (cos, sin)(sign*2*pi/size) — the running product is part of the rounding
// sequence, so it is reproduced, not replaced by a table of exact values.
static void fft_iterative(dvec& re, dvec& im, bool inverse) {
    const int n = (int)re.size();
    for (int size = 2; size <= n; size <<= 1) {
        const int half = size >> 1;
        const double sign = inverse ? 1.0 : -1.0;
        const double angle = sign * 2 * M_PI / size;
        const double wnRe = std::cos(angle), wnIm = std::sin(angle);
        for (int start = 0; start < n; start += size) {
            double wRe = 1, wIm = 0;
            for (int j = 0; j < half; j++) {
                const int i1 = start + j, i2 = start + j + half;
                const double tRe = wRe * re[i2] - wIm * im[i2];
                const double tIm = wRe * im[i2] + wIm * re[i2];
                re[i2] = re[i1] - tRe; im[i2] = im[i1] - tIm;
                re[i1] += tRe; im[i1] += tIm;
                const double nwRe = wRe * wnRe - wIm * wnIm;
                wIm = wRe * wnIm + wIm * wnRe;
                wRe = nwRe;
            }
        }
    }
}

// modem.js:6-13 fft (returns fresh arrays; here: in place on copies)
static void fft(dvec& re, dvec& im) { bit_reverse(re, im); fft_iterative(re, im, false); }

// modem.js:15-24 ifft
static void ifft(dvec& re, dvec& im) {
    bit_reverse(re, im);
    fft_iterative(re, im, true);
    const double scale = 1.0 / (double)re.size();
    for (size_t i = 0; i < re.size(); i++) { re[i] *= scale; im[i] *= scale; }
}

// ------------------------------------------------------------- configs ------
// modem.js:69-85 OFDM_CONFIGS
struct Ofdm {
    int FFT_SIZE, CP_LEN, SYMBOL_LEN, SAMPLE_RATE, SUB_START, SUB_END;
    std::vector<int> PILOTS;
    // modem.js:88 isPilot
    bool isPilot(int k) const { for (int p : PILOTS) if (p == k) return true; return false; }
    // modem.js:89-93 numDataSubs
    int numDataSubs() const { int c = 0; for (int k = SUB_START; k <= SUB_END; k++) if (!isPilot(k)) c++; return c; }
};

static const Ofdm& ofdm_config(int id) {
    static const Ofdm cfgs[3] = {
        {512, 64, 576, 44100, 12, 232, {15, 29, 43, 57, 71, 85, 99, 113, 127, 141, 155, 169, 183, 197, 211, 225}},
        {512, 128, 640, 44100, 23, 93, {25, 35, 45, 55, 65, 75, 85}},
        {512, 256, 768, 44100, 35, 58, {37, 45, 53}},
    };
    // modem.js:96 — an unknown name falls back to `standard`
    if (id < 0 || id > 2) id = 0;
    return cfgs[id];
}

// --------------------------------------------------------- constellations ---
// modem.js:101-131.  mod ids: 0 BPSK, 1 QPSK, 2 QAM16.
struct Constel { int bps; std::vector<std::pair<double, double>> points; };

static const Constel& constellation(int mod) {
    static Constel c[3];
    static bool init = false;
    if (!init) {
        c[0].bps = 1; c[0].points = {{1, 0}, {-1, 0}};                       // :111
        const double s = 1 / M_SQRT2;                                        // :113
        c[1].bps = 2; c[1].points = {{s, s}, {-s, s}, {-s, -s}, {s, -s}};     // :114-116
        c[2].bps = 4;                                                        // :117-128
        std::vector<std::pair<double, double>> raw;
        for (int i = 0; i < 16; i++) {
            const int row = i >> 2, col = i & 3;
            const int gr = row ^ (row >> 1), gc = col ^ (col >> 1);
            raw.push_back({(double)(2 * gc - 3), (double)(2 * gr - 3)});
        }
        double avg = 0;
        for (auto& p : raw) avg += p.first * p.first + p.second * p.second;
        avg /= (double)raw.size();
        const double sc = 1 / std::sqrt(avg);
        for (auto& p : raw) c[2].points.push_back({p.first * sc, p.second * sc});
        init = true;
    }
    if (mod < 0 || mod > 2) mod = 1;
    return c[mod];
}

// modem.js:140-150 constellationDemap: exhaustive nearest point, strict '<'
// so the first minimum wins; bits MSB first.
static inline int demap_index(const Constel& c, double re, double im) {
    double minD = std::numeric_limits<double>::infinity();
    int minIdx = 0;
    for (int i = 0; i < (int)c.points.size(); i++) {
        const double dr = re - c.points[i].first, di = im - c.points[i].second;
        const double d = dr * dr + di * di;
        if (d < minD) { minD = d; minIdx = i; }
    }
    return minIdx;
}

// ------------------------------------------------------ preamble / CE -------
// modem.js:153-156 seededRandom.  `s * 1103515245 + 12345` is evaluated in
// float64 (the product exceeds 2^53 and is rounded), then `&` applies ToInt32
// (value mod 2^32) and masks 31 bits.  An exact integer LCG would diverge on
// the second draw, so the double rounding is emulated.
struct JsLcg {
    double s;
    explicit JsLcg(double seed) : s(seed) {}
    double next() {
        double t = s * 1103515245.0;
        t = t + 12345.0;
        // t is an integer-valued double < 2^62; ToInt32 keeps the low 32 bits
        const uint64_t u = (uint64_t)t;
        s = (double)(uint32_t)(u & 0x7fffffffu);
        return s / 2147483647.0;  // 0x7fffffff
    }
};

// modem.js:202-208 addCP — the Float32Array store rounds to float32
static fvec add_cp(const Ofdm& o, const dvec& td) {
    const int n = (int)td.size(), cp = o.CP_LEN;
    fvec out(cp + n);
    for (int i = 0; i < cp; i++) out[i] = (float)td[n - cp + i];
    for (int i = 0; i < n; i++) out[cp + i] = (float)td[i];
    return out;
}

// shared tail of modem.js:165-169 / 179-183 / 195-198
static fvec hermitian_ifft_cp(const Ofdm& o, dvec& re, dvec& im) {
    const int n = o.FFT_SIZE;
    for (int k = 1; k < n / 2; k++) { re[n - k] = re[k]; im[n - k] = -im[k]; }
    re[0] = 0; re[n / 2] = 0; im[n / 2] = 0;
    ifft(re, im);
    return add_cp(o, re);
}

// modem.js:158-170 generatePreambleSymbol1 (seed 42, every 2nd bin)
static fvec preamble1(const Ofdm& o) {
    dvec re(o.FFT_SIZE, 0.0), im(o.FFT_SIZE, 0.0);
    JsLcg rng(42);
    for (int k = o.SUB_START; k <= o.SUB_END; k += 2) re[k] = rng.next() > 0.5 ? 1 : -1;
    return hermitian_ifft_cp(o, re, im);
}

// modem.js:172-184 generatePreambleSymbol2 (seed 43, every bin)
static fvec preamble2(const Ofdm& o) {
    dvec re(o.FFT_SIZE, 0.0), im(o.FFT_SIZE, 0.0);
    JsLcg rng(43);
    for (int k = o.SUB_START; k <= o.SUB_END; k++) re[k] = rng.next() > 0.5 ? 1 : -1;
    return hermitian_ifft_cp(o, re, im);
}

// modem.js:186-200 generateChannelEstSymbol (seed 44); knownIm is all zero
static fvec ce_symbol(const Ofdm& o, dvec& knownRe) {
    dvec re(o.FFT_SIZE, 0.0), im(o.FFT_SIZE, 0.0);
    knownRe.assign(o.FFT_SIZE, 0.0);
    JsLcg rng(44);
    for (int k = o.SUB_START; k <= o.SUB_END; k++) {
        const double v = rng.next() > 0.5 ? 1 : -1;
        re[k] = v; knownRe[k] = v;
    }
    return hermitian_ifft_cp(o, re, im);
}

// JS Math.max: NaN if either operand is NaN
static inline double js_max(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    return a > b ? a : b;
}

// ------------------------------------------------------- preprocess ---------
// modem.js:213-232 preprocessSignal
static fvec preprocess(const float* signal, int64_t n) {
    double mean = 0;
    for (int64_t i = 0; i < n; i++) mean += signal[i];
    mean /= (double)n;
    fvec out(n);
    double mx = 0;
    for (int64_t i = 0; i < n; i++) {
        out[i] = (float)((double)signal[i] - mean);
        mx = js_max(mx, std::fabs((double)out[i]));
    }
    if (mx > 1e-6) for (int64_t i = 0; i < n; i++) out[i] = (float)((double)out[i] / mx);
    return out;
}

// ---------------------------------------------------- detectors -------------
// modem.js:235-283 detectPreambleCrossCorr
static int64_t detect_xcorr(const Ofdm& o, const float* signal, int64_t len, double* best_out) {
    const fvec pre1 = preamble1(o);
    const int64_t pLen = (int64_t)pre1.size();
    if (best_out) *best_out = 0;
    if (len < pLen) return -1;
    double tEnergy = 0;
    for (int64_t i = 0; i < pLen; i++) tEnergy += (double)pre1[i] * pre1[i];
    if (tEnergy < 1e-10) return -1;
    const int64_t end = len - pLen;
    const int64_t step = std::max<int64_t>(1, pLen / 10);
    double best = 0; int64_t bestIdx = -1;
    for (int64_t d = 0; d <= end; d += step) {
        double corr = 0, sEnergy = 0;
        for (int64_t i = 0; i < pLen; i++) {
            corr += (double)signal[d + i] * pre1[i];
            sEnergy += (double)signal[d + i] * signal[d + i];
        }
        const double denom = std::sqrt(sEnergy * tEnergy);
        if (denom > 0.001) { const double m = corr / denom; if (m > best) { best = m; bestIdx = d; } }
    }
    if (best_out) *best_out = best;
    if (bestIdx < 0 || best < 0.15) return -1;
    const int64_t fineStart = std::max<int64_t>(0, bestIdx - step);
    const int64_t fineEnd = std::min<int64_t>(end, bestIdx + step);
    best = 0;
    for (int64_t d = fineStart; d <= fineEnd; d++) {
        double corr = 0, sEnergy = 0;
        for (int64_t i = 0; i < pLen; i++) {
            corr += (double)signal[d + i] * pre1[i];
            sEnergy += (double)signal[d + i] * signal[d + i];
        }
        const double denom = std::sqrt(sEnergy * tEnergy);
        if (denom > 0.001) { const double m = corr / denom; if (m > best) { best = m; bestIdx = d; } }
    }
    if (best_out) *best_out = best;
    return best > 0.15 ? bestIdx : -1;
}

// modem.js:286-319 detectPreamble (Schmidl-Cox sliding auto-correlation)
static int64_t detect_autocorr(const Ofdm& o, const float* signal, int64_t len, double* best_out) {
    const int64_t half = o.FFT_SIZE / 2;
    if (best_out) *best_out = 0;
    if (len < 2 * half) return -1;
    double p = 0, ra = 0, rb = 0;
    for (int64_t m = 0; m < half; m++) {
        const double a = signal[m], b = signal[m + half];
        p += a * b; ra += a * a; rb += b * b;
    }
    double best = 0; int64_t bestIdx = -1;
    const int64_t end = len - 2 * half;
    const double minEnergy = 0.01;
    for (int64_t d = 0; d <= end; d++) {
        if (ra > minEnergy && rb > minEnergy) {
            const double metric = (p * p) / (ra * rb);
            if (metric > best) { best = metric; bestIdx = d; }
        }
        if (d < end) {
            const double aOut = signal[d], mid = signal[d + half], bIn = signal[d + 2 * half];
            p += mid * bIn - aOut * mid;
            ra += mid * mid - aOut * aOut;
            rb += bIn * bIn - mid * mid;
        }
    }
    if (best_out) *best_out = best;
    return best > 0.5 ? bestIdx : -1;
}

// fine cross-correlation, modem.js:566-588 (also app.js:862-877).  Returns the
// argmax over [fineStart, fineEnd] with strict '>' from -Infinity.
static int64_t fine_xcorr(const fvec& pre1, double tEnergy, const float* signal, int64_t fineStart,
                          int64_t fineEnd, int64_t fallback, double* bestMetric_out) {
    double bestMetric = -std::numeric_limits<double>::infinity();
    int64_t startIdx = fallback;
    const int64_t pLen = (int64_t)pre1.size();
    for (int64_t d = fineStart; d <= fineEnd; d++) {
        double corr = 0, sEnergy = 0;
        for (int64_t i = 0; i < pLen; i++) {
            corr += (double)signal[d + i] * pre1[i];
            sEnergy += (double)signal[d + i] * signal[d + i];
        }
        const double denom = std::sqrt(sEnergy * tEnergy);
        if (denom > 0.001) { const double m = corr / denom; if (m > bestMetric) { bestMetric = m; startIdx = d; } }
    }
    *bestMetric_out = bestMetric;
    return startIdx;
}

// ------------------------------------------------------- bits / crc ---------
// modem.js:443-457
static uint32_t crc_table[256];
static bool crc_init_done = false;
static void crc_init() {
    if (crc_init_done) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int j = 0; j < 8; j++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        crc_table[i] = c;
    }
    crc_init_done = true;
}
static uint32_t crc32(const uint8_t* data, size_t n) {
    crc_init();
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ data[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

// modem.js:460-466 bytesToBits (MSB first)
static bvec bytes_to_bits(const bvec& data) {
    bvec bits; bits.reserve(data.size() * 8);
    for (uint8_t b : data) for (int i = 7; i >= 0; i--) bits.push_back((b >> i) & 1);
    return bits;
}
// modem.js:468-476 bitsToBytes (drops a trailing partial byte)
static bvec bits_to_bytes(const bvec& bits) {
    bvec bytes;
    for (size_t i = 0; i + 7 < bits.size(); i += 8) {
        int b = 0;
        for (int j = 0; j < 8; j++) b = (b << 1) | (bits[i + j] & 1);
        bytes.push_back((uint8_t)b);
    }
    return bytes;
}
// modem.js:479-485 repeatBits
static bvec repeat_bits(const bvec& bits, int n) {
    bvec out; out.reserve(bits.size() * n);
    for (uint8_t b : bits) for (int i = 0; i < n; i++) out.push_back(b);
    return out;
}
// modem.js:487-495 majorityVote (`sum >= n/2` with n/2 a float)
static bvec majority_vote(const bvec& bits, int n) {
    bvec out;
    for (size_t i = 0; i + n - 1 < bits.size(); i += n) {
        int sum = 0;
        for (int j = 0; j < n; j++) sum += bits[i + j];
        out.push_back((double)sum >= (double)n / 2 ? 1 : 0);
    }
    return out;
}

// ---------------------------------------------------------- modulator -------
// modem.js:322-362 modulateOFDM.  Pads `bits` in place (:329).
static std::vector<fvec> modulate(const Ofdm& o, bvec& bits, int mod) {
    const Constel& c = constellation(mod);
    const int bps = c.bps;
    const int bitsPerSymbol = o.numDataSubs() * bps;
    while (bits.size() % bitsPerSymbol != 0) bits.push_back(0);
    const size_t numSymbols = bits.size() / bitsPerSymbol;
    std::vector<fvec> all;
    for (size_t s = 0; s < numSymbols; s++) {
        const uint8_t* symBits = &bits[s * bitsPerSymbol];
        dvec specRe(o.FFT_SIZE, 0.0), specIm(o.FFT_SIZE, 0.0);
        int di = 0;
        for (int k = o.SUB_START; k <= o.SUB_END; k++) {
            if (o.isPilot(k)) { specRe[k] = 1; specIm[k] = 0; }
            else {
                int idx = 0;                                     // :133-138 constellationMap
                for (int b = 0; b < bps; b++) idx = (idx << 1) | (symBits[di * bps + b] & 1);
                const auto& p = c.points[idx % c.points.size()];
                specRe[k] = p.first; specIm[k] = p.second;
                di++;
            }
        }
        const int n = o.FFT_SIZE;
        for (int k = 1; k < n / 2; k++) { specRe[n - k] = specRe[k]; specIm[n - k] = -specIm[k]; }
        specRe[0] = 0; specIm[0] = 0; specIm[n / 2] = 0;         // :354 (re[n/2] is already 0)
        ifft(specRe, specIm);
        all.push_back(add_cp(o, specRe));
    }
    return all;
}

// ------------------------------------------------------- demodulator --------
// modem.js:421-440 estimateChannel
static void estimate_channel(const Ofdm& o, const float* rx, int64_t rxLen, const dvec& knownRe,
                             const dvec& knownIm, dvec& chRe, dvec& chIm) {
    dvec re(o.FFT_SIZE, 0.0), im(o.FFT_SIZE, 0.0);
    for (int i = 0; i < o.FFT_SIZE; i++) {
        const int64_t j = o.CP_LEN + i;
        double v = (j < rxLen) ? (double)rx[j] : 0.0;            // `|| 0`: undefined/NaN/0 -> 0
        if (std::isnan(v)) v = 0.0;
        re[i] = v;
    }
    fft(re, im);
    chRe.assign(o.FFT_SIZE, 0.0); chIm.assign(o.FFT_SIZE, 0.0);
    for (int k = o.SUB_START; k <= o.SUB_END; k++) {
        const double xr = knownRe[k], xi = knownIm[k];
        const double d = xr * xr + xi * xi;
        if (d > 1e-10) {
            chRe[k] = (re[k] * xr + im[k] * xi) / d;
            chIm[k] = (im[k] * xr - re[k] * xi) / d;
        }
    }
}

// modem.js:365-418 demodulateOFDM.  eq_out (optional) receives the de-rotated
// (cr, ci) of every data subcarrier, 2 doubles each, symbol-major.
static bvec demodulate(const Ofdm& o, const float* signal, int64_t len, int mod, const dvec& chRe,
                       const dvec& chIm, dvec* eq_out) {
    const Constel& c = constellation(mod);
    const int bps = c.bps;
    const int64_t numSymbols = len / o.SYMBOL_LEN;
    bvec allBits;
    allBits.reserve((size_t)numSymbols * o.numDataSubs() * bps);
    dvec eqRe(o.FFT_SIZE), eqIm(o.FFT_SIZE);
    for (int64_t s = 0; s < numSymbols; s++) {
        const int64_t offset = s * o.SYMBOL_LEN;
        dvec re(o.FFT_SIZE, 0.0), im(o.FFT_SIZE, 0.0);
        for (int i = 0; i < o.FFT_SIZE; i++) {
            double v = signal[offset + o.CP_LEN + i];
            if (std::isnan(v)) v = 0.0;
            re[i] = v;
        }
        fft(re, im);
        std::fill(eqRe.begin(), eqRe.end(), 0.0); std::fill(eqIm.begin(), eqIm.end(), 0.0);
        for (int k = o.SUB_START; k <= o.SUB_END; k++) {
            const double hr = chRe[k], hi = chIm[k];
            const double hMag = hr * hr + hi * hi;
            if (hMag > 1e-10) {
                eqRe[k] = (re[k] * hr + im[k] * hi) / hMag;
                eqIm[k] = (im[k] * hr - re[k] * hi) / hMag;
            } else { eqRe[k] = re[k]; eqIm[k] = im[k]; }
        }
        double phaseSum = 0; int pc = 0;
        for (int p : o.PILOTS) {
            if (p >= o.SUB_START && p <= o.SUB_END && std::fabs(eqRe[p]) > 1e-6) {
                phaseSum += eqIm[p] / eqRe[p]; pc++;
            }
        }
        const double phase = pc > 0 ? phaseSum / pc : 0;
        for (int k = o.SUB_START; k <= o.SUB_END; k++) {
            if (!o.isPilot(k)) {
                const double cr = eqRe[k] + eqIm[k] * phase;
                const double ci = eqIm[k] - eqRe[k] * phase;
                if (eq_out) { eq_out->push_back(cr); eq_out->push_back(ci); }
                const int idx = demap_index(c, cr, ci);
                for (int i = bps - 1; i >= 0; i--) allBits.push_back((idx >> i) & 1);
            }
        }
    }
    return allBits;
}

// -------------------------------------------------------- frame builders ----
static inline void normalise_peak(fvec& signal) {   // modem.js:550-552 / 751-753
    double mx = 0;
    for (float v : signal) mx = js_max(mx, std::fabs((double)v));
    if (mx > 0) { const double s = 0.8 / mx; for (float& v : signal) v = (float)((double)v * s); }
}

static fvec assemble_frame(const Ofdm& o, int64_t silencePre, int64_t silencePost, const std::vector<fvec>& samples) {
    const fvec pre1 = preamble1(o), pre2 = preamble2(o);
    dvec known; const fvec ce = ce_symbol(o, known);
    fvec signal; signal.reserve(silencePre + silencePost + (3 + samples.size()) * o.SYMBOL_LEN);
    signal.insert(signal.end(), (size_t)silencePre, 0.0f);
    signal.insert(signal.end(), pre1.begin(), pre1.end());
    signal.insert(signal.end(), pre2.begin(), pre2.end());
    signal.insert(signal.end(), ce.begin(), ce.end());
    for (const fvec& s : samples) signal.insert(signal.end(), s.begin(), s.end());
    signal.insert(signal.end(), (size_t)silencePost, 0.0f);
    normalise_peak(signal);
    return signal;
}

// modem.js:498-555 buildTransmitSignal
static fvec build_transmit_signal(const Ofdm& o, const uint8_t* fileData, int64_t len, int mod,
                                  const uint8_t* nameBytes, int nameBytesLen, int repetition) {
    if (repetition <= 0) repetition = 1;
    static const uint8_t dflt[4] = {'f', 'i', 'l', 'e'};
    if (nameBytesLen <= 0) { nameBytes = dflt; nameBytesLen = 4; }    // `fileName || 'file'`
    const int nameLen = std::min(nameBytesLen, 255);
    bvec payload; payload.reserve(1 + nameLen + 4 + len + 4);
    payload.push_back((uint8_t)nameLen);
    for (int i = 0; i < nameLen; i++) payload.push_back(nameBytes[i]);
    const int32_t l32 = (int32_t)len;
    payload.push_back((l32 >> 24) & 0xFF); payload.push_back((l32 >> 16) & 0xFF);
    payload.push_back((l32 >> 8) & 0xFF); payload.push_back(l32 & 0xFF);
    payload.insert(payload.end(), fileData, fileData + len);
    const uint32_t cs = crc32(payload.data(), payload.size());
    payload.push_back((cs >> 24) & 0xFF); payload.push_back((cs >> 16) & 0xFF);
    payload.push_back((cs >> 8) & 0xFF); payload.push_back(cs & 0xFF);
    bvec bits = bytes_to_bits(payload);
    if (repetition > 1) bits = repeat_bits(bits, repetition);
    const std::vector<fvec> samples = modulate(o, bits, mod);
    const bool isAcoustic = o.CP_LEN >= 128;
    // `new Float32Array(44100 * 0.3)` truncates the (inexact) double product  :534-535
    const int64_t silencePre = (int64_t)std::floor((double)o.SAMPLE_RATE * (isAcoustic ? 0.5 : 0.3));
    const int64_t silencePost = (int64_t)std::floor((double)o.SAMPLE_RATE * (isAcoustic ? 0.5 : 0.2));
    return assemble_frame(o, silencePre, silencePost, samples);
}

// modem.js:666-692 buildMetadataPayload
static bvec build_metadata_payload(int32_t totalChunks, int32_t totalFileSize, int32_t chunkSize,
                                   const uint8_t* nameBytes, int nameBytesLen) {
    static const uint8_t dflt[4] = {'f', 'i', 'l', 'e'};
    if (nameBytesLen <= 0) { nameBytes = dflt; nameBytesLen = 4; }
    const int nameLen = std::min(nameBytesLen, 255);
    bvec buf;
    buf.push_back(0xFE);
    buf.push_back((totalChunks >> 24) & 0xFF); buf.push_back((totalChunks >> 16) & 0xFF);
    buf.push_back((totalChunks >> 8) & 0xFF); buf.push_back(totalChunks & 0xFF);
    buf.push_back((totalFileSize >> 24) & 0xFF); buf.push_back((totalFileSize >> 16) & 0xFF);
    buf.push_back((totalFileSize >> 8) & 0xFF); buf.push_back(totalFileSize & 0xFF);
    buf.push_back((chunkSize >> 8) & 0xFF); buf.push_back(chunkSize & 0xFF);
    buf.push_back((uint8_t)nameLen);
    for (int i = 0; i < nameLen; i++) buf.push_back(nameBytes[i]);
    const uint32_t cs = crc32(buf.data(), buf.size());
    buf.push_back((cs >> 24) & 0xFF); buf.push_back((cs >> 16) & 0xFF);
    buf.push_back((cs >> 8) & 0xFF); buf.push_back(cs & 0xFF);
    return buf;
}

// modem.js:694-714 buildDataChunkPayload
static bvec build_data_chunk_payload(const uint8_t* chunk, int dataLen, int32_t seqNum) {
    bvec buf;
    buf.push_back(0xFF);
    buf.push_back((seqNum >> 24) & 0xFF); buf.push_back((seqNum >> 16) & 0xFF);
    buf.push_back((seqNum >> 8) & 0xFF); buf.push_back(seqNum & 0xFF);
    buf.push_back((dataLen >> 8) & 0xFF); buf.push_back(dataLen & 0xFF);
    buf.insert(buf.end(), chunk, chunk + dataLen);
    const uint32_t cs = crc32(buf.data(), buf.size());
    buf.push_back((cs >> 24) & 0xFF); buf.push_back((cs >> 16) & 0xFF);
    buf.push_back((cs >> 8) & 0xFF); buf.push_back(cs & 0xFF);
    return buf;
}

// modem.js:718-756 buildChunkOFDMFrame
static fvec build_chunk_frame(const Ofdm& o, const bvec& payload, int mod, int repetition, bool isFirstFrame) {
    if (repetition <= 0) repetition = 1;
    bvec bits = bytes_to_bits(payload);
    if (repetition > 1) bits = repeat_bits(bits, repetition);
    const std::vector<fvec> samples = modulate(o, bits, mod);
    const bool isAcoustic = o.CP_LEN >= 128;
    const int64_t silencePre = isFirstFrame ? (int64_t)std::floor((double)o.SAMPLE_RATE * (isAcoustic ? 0.5 : 0.3) + 0.5)
                                            : (int64_t)std::floor((double)o.SAMPLE_RATE * 0.05 + 0.5);   // Math.round
    const int64_t silencePost = (int64_t)std::floor((double)o.SAMPLE_RATE * 0.02 + 0.5);
    return assemble_frame(o, silencePre, silencePost, samples);
}

// modem.js:863-874 estimateFrameSamples
static int64_t estimate_frame_samples(const Ofdm& o, int64_t payloadBytes, int mod, int repetition) {
    if (repetition <= 0) repetition = 1;
    const int bitsPerSymbol = o.numDataSubs() * constellation(mod).bps;
    const double totalBits = (double)payloadBytes * 8 * repetition;
    const int64_t numSymbols = (int64_t)std::ceil(totalBits / bitsPerSymbol);
    return (3 + numSymbols) * o.SYMBOL_LEN;
}

}  // namespace

// ============================================================================
// C interface (ctypes).  Result record shared by the decoders.
// ============================================================================
extern "C" {

// status codes mirror the reference's error strings
enum {
    ORC_OK = 0,
    ORC_ERR_PREAMBLE_NOT_DETECTED = 1,        // modem.js:564
    ORC_ERR_PREAMBLE_LOW_CORRELATION = 2,     // :588
    ORC_ERR_SIGNAL_TOO_SHORT_FOR_CE = 3,      // :592
    ORC_ERR_NO_DATA_AFTER_CE = 4,             // :600, :785
    ORC_ERR_DECODED_TOO_SHORT = 5,            // :607, :793
    ORC_ERR_DECODED_TOO_SHORT_FOR_HEADER = 6, // :625
    ORC_ERR_INVALID_DATA_LENGTH = 7,          // :634
    ORC_ERR_FRAME_TOO_SHORT_FOR_CE = 8,       // :776
    ORC_ERR_UNKNOWN_FRAME_TYPE = 9,           // :801
    ORC_ERR_META_TOO_SHORT = 10,              // :807
    ORC_ERR_META_TRUNCATED = 11,              // :813
    ORC_ERR_CHUNK_TOO_SHORT = 12,             // :832
    ORC_ERR_CHUNK_TRUNCATED = 13,             // :836
    ORC_ERR_RING_OVERWRITTEN = 14,            // app.js:912-916 (frameErrors++, no message)
};

typedef struct {
    int32_t status;
    int32_t frame_type;       // 0 legacy, 0xFE metadata, 0xFF data chunk, -1 none
    int64_t preamble_idx;     // refined start index (global position when streaming), -1 n/a
    int32_t seq_num;          // data chunk: seqNum (signed, modem.js:834)
    int32_t data_len;         // legacy / data chunk: dataLen
    int32_t total_chunks;     // metadata (signed, :809)
    int32_t total_file_size;  // metadata (signed, :810)
    int32_t chunk_size;       // metadata
    int32_t name_len;         // legacy / metadata
    int32_t name_off;         // offset of the name inside the decoded byte array
    int32_t data_off;         // offset of data inside the decoded byte array
    uint32_t expected_crc;
    uint32_t actual_crc;
    int32_t crc_valid;
    int32_t n_bytes;          // bytes.length of the decoded byte array
    int32_t err_arg;          // dataLen for INVALID_DATA_LENGTH, type byte for UNKNOWN_FRAME_TYPE
    int32_t stream;           // stream index (filled by callers that batch)
    double coarse_metric;     // detectPreamble best (batch) / n/a
    double fine_metric;       // refined cross-correlation peak
} OrcResult;

static void result_clear(OrcResult* r) {
    std::memset(r, 0, sizeof(*r));
    r->frame_type = -1; r->preamble_idx = -1;
}

static inline int32_t be32s(const uint8_t* b) {  // JS `(b0<<24)|(b1<<16)|(b2<<8)|b3` — signed
    return (int32_t)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
}

// modem.js:805-828 parseMetadataResult
static void parse_metadata(const bvec& bytes, OrcResult* r) {
    r->frame_type = 0xFE;
    if (bytes.size() < 16) { r->status = ORC_ERR_META_TOO_SHORT; r->frame_type = -1; return; }
    size_t off = 1;
    r->total_chunks = be32s(&bytes[off]); off += 4;
    r->total_file_size = be32s(&bytes[off]); off += 4;
    r->chunk_size = (bytes[off] << 8) | bytes[off + 1]; off += 2;
    const int nameLen = bytes[off++];
    if (off + nameLen + 4 > bytes.size()) { r->status = ORC_ERR_META_TRUNCATED; r->frame_type = -1; return; }
    r->name_len = nameLen; r->name_off = (int32_t)off;
    off += nameLen;
    r->expected_crc = (uint32_t)be32s(&bytes[off]);
    r->actual_crc = crc32(bytes.data(), off);
    r->crc_valid = r->expected_crc == r->actual_crc;
}

// modem.js:830-849 parseDataChunkResult
static void parse_data_chunk(const bvec& bytes, OrcResult* r) {
    r->frame_type = 0xFF;
    if (bytes.size() < 11) { r->status = ORC_ERR_CHUNK_TOO_SHORT; r->frame_type = -1; return; }
    size_t off = 1;
    r->seq_num = be32s(&bytes[off]); off += 4;
    const int dataLen = (bytes[off] << 8) | bytes[off + 1]; off += 2;
    if (off + dataLen + 4 > bytes.size()) { r->status = ORC_ERR_CHUNK_TRUNCATED; r->frame_type = -1; return; }
    r->data_len = dataLen; r->data_off = (int32_t)off;
    off += dataLen;
    r->expected_crc = (uint32_t)be32s(&bytes[off]);
    r->actual_crc = crc32(bytes.data(), off);
    r->crc_valid = r->expected_crc == r->actual_crc;
}

static void copy_bytes(const bvec& bytes, uint8_t* out, int64_t cap) {
    if (!out) return;
    const int64_t n = std::min<int64_t>((int64_t)bytes.size(), cap);
    if (n > 0) std::memcpy(out, bytes.data(), (size_t)n);
}

// modem.js:770-803 decodeChunkFrame
static void decode_chunk_frame_impl(const Ofdm& o, const float* frame, int64_t n, int mod, int repetition,
                                    OrcResult* r, uint8_t* bytes_out, int64_t cap) {
    result_clear(r);
    if (repetition <= 0) repetition = 1;
    const int64_t ceStart = 2 * o.SYMBOL_LEN;
    if (ceStart + o.SYMBOL_LEN > n) { r->status = ORC_ERR_FRAME_TOO_SHORT_FOR_CE; return; }
    dvec knownRe; ce_symbol(o, knownRe);
    dvec knownIm(o.FFT_SIZE, 0.0), chRe, chIm;
    estimate_channel(o, frame + ceStart, o.SYMBOL_LEN, knownRe, knownIm, chRe, chIm);
    const int64_t dataStart = ceStart + o.SYMBOL_LEN;
    if (dataStart >= n) { r->status = ORC_ERR_NO_DATA_AFTER_CE; return; }
    bvec bits = demodulate(o, frame + dataStart, n - dataStart, mod, chRe, chIm, nullptr);
    if (repetition > 1) bits = majority_vote(bits, repetition);
    const bvec bytes = bits_to_bytes(bits);
    r->n_bytes = (int32_t)bytes.size();
    copy_bytes(bytes, bytes_out, cap);
    if (bytes.size() < 6) { r->status = ORC_ERR_DECODED_TOO_SHORT; return; }
    const int frameType = bytes[0];
    if (frameType == 0xFE) parse_metadata(bytes, r);
    else if (frameType == 0xFF) parse_data_chunk(bytes, r);
    else { r->status = ORC_ERR_UNKNOWN_FRAME_TYPE; r->err_arg = frameType; r->frame_type = frameType; }
}

// modem.js:557-654 decodeReceivedSignal
static void decode_received_signal_impl(const Ofdm& o, const float* raw, int64_t n, int mod, int repetition,
                                        OrcResult* r, uint8_t* bytes_out, int64_t cap) {
    result_clear(r);
    if (repetition <= 0) repetition = 1;
    const fvec signal = preprocess(raw, n);
    double coarseBest = 0;
    const int64_t coarseIdx = detect_autocorr(o, signal.data(), n, &coarseBest);
    r->coarse_metric = coarseBest;
    if (coarseIdx < 0) { r->status = ORC_ERR_PREAMBLE_NOT_DETECTED; return; }
    const fvec pre1 = preamble1(o);
    double tEnergy = 0;
    for (float v : pre1) tEnergy += (double)v * v;
    const int64_t searchRadius = o.CP_LEN * 3;
    const int64_t fineStart = std::max<int64_t>(0, coarseIdx - searchRadius);
    const int64_t fineEnd = std::min<int64_t>(n - (int64_t)pre1.size(), coarseIdx + searchRadius);
    double bestMetric;
    const int64_t startIdx = fine_xcorr(pre1, tEnergy, signal.data(), fineStart, fineEnd, coarseIdx, &bestMetric);
    r->fine_metric = bestMetric;
    if (bestMetric < 0.1) { r->status = ORC_ERR_PREAMBLE_LOW_CORRELATION; return; }
    const int64_t ceStart = startIdx + 2 * o.SYMBOL_LEN;
    if (ceStart + o.SYMBOL_LEN > n) { r->status = ORC_ERR_SIGNAL_TOO_SHORT_FOR_CE; return; }
    dvec knownRe; ce_symbol(o, knownRe);
    dvec knownIm(o.FFT_SIZE, 0.0), chRe, chIm;
    estimate_channel(o, signal.data() + ceStart, o.SYMBOL_LEN, knownRe, knownIm, chRe, chIm);
    const int64_t dataStart = ceStart + o.SYMBOL_LEN;
    if (dataStart >= n) { r->status = ORC_ERR_NO_DATA_AFTER_CE; return; }
    bvec bits = demodulate(o, signal.data() + dataStart, n - dataStart, mod, chRe, chIm, nullptr);
    if (repetition > 1) bits = majority_vote(bits, repetition);
    const bvec bytes = bits_to_bytes(bits);
    r->n_bytes = (int32_t)bytes.size();
    copy_bytes(bytes, bytes_out, cap);
    if (bytes.size() < 10) { r->status = ORC_ERR_DECODED_TOO_SHORT; return; }
    const int firstByte = bytes[0];
    if (firstByte == 0xFE) { parse_metadata(bytes, r); r->preamble_idx = startIdx; return; }
    if (firstByte == 0xFF) { parse_data_chunk(bytes, r); r->preamble_idx = startIdx; return; }
    size_t off = 0;
    const int nameLen = bytes[off++];
    if (off + nameLen + 4 + 4 > bytes.size()) { r->status = ORC_ERR_DECODED_TOO_SHORT_FOR_HEADER; return; }
    r->name_len = nameLen; r->name_off = (int32_t)off;
    off += nameLen;
    const int32_t dataLen = be32s(&bytes[off]);
    off += 4;
    if (dataLen <= 0 || off + (size_t)dataLen + 4 > bytes.size()) {
        r->status = ORC_ERR_INVALID_DATA_LENGTH; r->err_arg = dataLen; return;
    }
    r->data_len = dataLen; r->data_off = (int32_t)off;
    off += dataLen;
    r->expected_crc = (uint32_t)be32s(&bytes[off]);
    r->actual_crc = crc32(bytes.data(), off);
    r->crc_valid = r->expected_crc == r->actual_crc;
    r->preamble_idx = startIdx;
    r->frame_type = 0;
}

// modem.js:978-1081 analyzeLoopback.  `channelMagnitude` has SUB_END-SUB_START+1 entries when the channel was
// estimated, none otherwise (n_mag says which); quality: 0 'poor', 1 'good', 2 'excellent'.
typedef struct {
    int32_t detected;
    int32_t quality;
    int64_t start_idx;        // refined start (the coarse index when no offset passed the denominator test); -1 when not detected
    int32_t n_mag;
    int32_t pad;
    double correlation;
    double ber;
    double snr_estimate;      // 20 log10(mean pilot |H|); -Infinity when no pilot qualifies; 0 on the early returns
} OrcLoopback;

static void analyze_loopback_impl(const Ofdm& o, const float* recorded, int64_t n, int mod, int repetition,
                                  const uint8_t* testData, int testLen, OrcLoopback* r, double* mags) {
    std::memset(r, 0, sizeof(*r));
    r->start_idx = -1; r->ber = 1.0;
    const fvec signal = preprocess(recorded, n);                                            // :980
    double best = 0;
    int64_t coarseIdx = detect_autocorr(o, signal.data(), n, &best);                        // :983
    if (coarseIdx < 0) coarseIdx = detect_xcorr(o, signal.data(), n, &best);                // :984-987
    if (coarseIdx < 0) return;                                                              // :988-990
    r->detected = 1;
    const fvec pre1 = preamble1(o);
    double tEnergy = 0;
    for (float v : pre1) tEnergy += (double)v * v;
    const int64_t searchRadius = o.CP_LEN * 3;
    const int64_t fineStart = std::max<int64_t>(0, coarseIdx - searchRadius);
    const int64_t fineEnd = std::min<int64_t>(n - (int64_t)pre1.size(), coarseIdx + searchRadius);
    double bestMetric;
    const int64_t startIdx = fine_xcorr(pre1, tEnergy, signal.data(), fineStart, fineEnd, coarseIdx, &bestMetric);   // :993-1014
    r->start_idx = startIdx;
    r->correlation = bestMetric > 0 ? bestMetric : 0.0;                                     // :1016 Math.max(0, bestMetric)
    const int64_t ceStart = startIdx + 2 * o.SYMBOL_LEN;
    if (ceStart + o.SYMBOL_LEN > n) return;                                                 // :1020-1022
    dvec knownRe; ce_symbol(o, knownRe);
    dvec knownIm(o.FFT_SIZE, 0.0), chRe, chIm;
    estimate_channel(o, signal.data() + ceStart, o.SYMBOL_LEN, knownRe, knownIm, chRe, chIm);
    for (int k = o.SUB_START; k <= o.SUB_END; k++)                                          // :1029-1033
        mags[r->n_mag++] = std::sqrt(chRe[k] * chRe[k] + chIm[k] * chIm[k]);
    double snrSum = 0; int snrCount = 0;                                                    // :1036-1047
    for (int p : o.PILOTS) {
        if (p >= o.SUB_START && p <= o.SUB_END) {
            const double mag = std::sqrt(chRe[p] * chRe[p] + chIm[p] * chIm[p]);
            if (mag > 1e-6) { snrSum += mag; snrCount++; }
        }
    }
    const double avgPilotMag = snrCount > 0 ? snrSum / snrCount : 0.0;
    r->snr_estimate = avgPilotMag > 0 ? 20 * std::log10(avgPilotMag) : -INFINITY;
    const int64_t dataStart = ceStart + o.SYMBOL_LEN;
    if (dataStart < n) {                                                                    // :1052-1072
        bvec bits = demodulate(o, signal.data() + dataStart, n - dataStart, mod, chRe, chIm, nullptr);
        if (repetition > 1) bits = majority_vote(bits, repetition);
        const bvec decoded = bits_to_bytes(bits);
        if (decoded.size() >= 29) {
            const int nameLen = decoded[0];
            const size_t dataOffset = 1 + nameLen + 4;
            if (dataOffset + (size_t)testLen <= decoded.size()) {
                int errorBits = 0;
                for (int i = 0; i < testLen; i++) {
                    const int x = decoded[dataOffset + i] ^ testData[i];
                    for (int b = 0; b < 8; b++) errorBits += (x >> b) & 1;
                }
                r->ber = (double)errorBits / (double)(testLen * 8);                         // 0/0 = NaN for empty test data, as in JS
            }
        }
    }
    if (r->ber == 0 && r->correlation > 0.8) r->quality = 2;                                // :1075-1081
    else if (r->ber < 0.05) r->quality = 1;
    else r->quality = 0;
}

// ---------------------------------------------------------------- exports ---
int orc_result_size(void) { return (int)sizeof(OrcResult); }

void orc_fft(double* re, double* im, int n, int inverse) {
    dvec r(re, re + n), i(im, im + n);
    if (inverse) ifft(r, i); else fft(r, i);
    std::memcpy(re, r.data(), n * sizeof(double)); std::memcpy(im, i.data(), n * sizeof(double));
}

// the (cos, sin) stage seeds of the forward transform, for twiddle-table checks
void orc_fft_stage_seeds(double* wnRe, double* wnIm) {
    int s = 0;
    for (int size = 2; size <= 512; size <<= 1, s++) {
        const double angle = -1.0 * 2 * M_PI / size;
        wnRe[s] = std::cos(angle); wnIm[s] = std::sin(angle);
    }
}

void orc_lcg(int seed, int n, int64_t* states, double* values) {
    JsLcg g((double)seed);
    for (int i = 0; i < n; i++) { const double v = g.next(); states[i] = (int64_t)g.s; values[i] = v; }
}

int orc_config(int cfg, int* out /*FFT,CP,SYM,RATE,START,END,nPilots,numDataSubs*/, int* pilots) {
    const Ofdm& o = ofdm_config(cfg);
    out[0] = o.FFT_SIZE; out[1] = o.CP_LEN; out[2] = o.SYMBOL_LEN; out[3] = o.SAMPLE_RATE;
    out[4] = o.SUB_START; out[5] = o.SUB_END; out[6] = (int)o.PILOTS.size(); out[7] = o.numDataSubs();
    for (size_t i = 0; i < o.PILOTS.size(); i++) pilots[i] = o.PILOTS[i];
    return 0;
}

int orc_constellation(int mod, double* pts /*2*npoints*/) {
    const Constel& c = constellation(mod);
    for (size_t i = 0; i < c.points.size(); i++) { pts[2 * i] = c.points[i].first; pts[2 * i + 1] = c.points[i].second; }
    return c.bps;
}

void orc_preamble1(int cfg, float* out) { const fvec v = preamble1(ofdm_config(cfg)); std::memcpy(out, v.data(), v.size() * 4); }
void orc_preamble2(int cfg, float* out) { const fvec v = preamble2(ofdm_config(cfg)); std::memcpy(out, v.data(), v.size() * 4); }
void orc_ce_symbol(int cfg, float* out, double* knownRe) {
    dvec k; const fvec v = ce_symbol(ofdm_config(cfg), k);
    std::memcpy(out, v.data(), v.size() * 4);
    if (knownRe) std::memcpy(knownRe, k.data(), k.size() * 8);
}

void orc_preprocess(const float* x, int64_t n, float* out) {
    const fvec v = preprocess(x, n);
    std::memcpy(out, v.data(), (size_t)n * 4);
}

int64_t orc_detect_preamble(int cfg, const float* x, int64_t n, double* best) {
    return detect_autocorr(ofdm_config(cfg), x, n, best);
}
int64_t orc_detect_preamble_xcorr(int cfg, const float* x, int64_t n, double* best) {
    return detect_xcorr(ofdm_config(cfg), x, n, best);
}

// Schmidl-Cox metric at every offset, with the reference's sliding float64
// sums (modem.js:292-316).  metric = NaN where the energy gate fails.
void orc_autocorr_metric(int cfg, const float* signal, int64_t len, double minEnergy, double* metric) {
    const Ofdm& o = ofdm_config(cfg);
    const int64_t half = o.FFT_SIZE / 2;
    if (len < 2 * half) return;
    double p = 0, ra = 0, rb = 0;
    for (int64_t m = 0; m < half; m++) { const double a = signal[m], b = signal[m + half]; p += a * b; ra += a * a; rb += b * b; }
    const int64_t end = len - 2 * half;
    for (int64_t d = 0; d <= end; d++) {
        metric[d] = (ra > minEnergy && rb > minEnergy) ? (p * p) / (ra * rb) : std::numeric_limits<double>::quiet_NaN();
        if (d < end) {
            const double aOut = signal[d], mid = signal[d + half], bIn = signal[d + 2 * half];
            p += mid * bIn - aOut * mid; ra += mid * mid - aOut * aOut; rb += bIn * bIn - mid * mid;
        }
    }
}

// fine cross-correlation metric for d in [d0, d1] (modem.js:576-587); NaN where denom <= 0.001
void orc_xcorr_metric(int cfg, const float* signal, int64_t d0, int64_t d1, double* metric) {
    const Ofdm& o = ofdm_config(cfg);
    const fvec pre1 = preamble1(o);
    double tEnergy = 0;
    for (float v : pre1) tEnergy += (double)v * v;
    for (int64_t d = d0; d <= d1; d++) {
        double corr = 0, sEnergy = 0;
        for (size_t i = 0; i < pre1.size(); i++) { corr += (double)signal[d + i] * pre1[i]; sEnergy += (double)signal[d + i] * signal[d + i]; }
        const double denom = std::sqrt(sEnergy * tEnergy);
        metric[d - d0] = denom > 0.001 ? corr / denom : std::numeric_limits<double>::quiet_NaN();
    }
}

void orc_estimate_channel(int cfg, const float* rx, int64_t n, const double* knownRe, const double* knownIm,
                          double* chRe, double* chIm) {
    const Ofdm& o = ofdm_config(cfg);
    dvec kr(knownRe, knownRe + o.FFT_SIZE), ki(o.FFT_SIZE, 0.0), cr, ci;
    if (knownIm) ki.assign(knownIm, knownIm + o.FFT_SIZE);
    estimate_channel(o, rx, n, kr, ki, cr, ci);
    std::memcpy(chRe, cr.data(), o.FFT_SIZE * 8); std::memcpy(chIm, ci.data(), o.FFT_SIZE * 8);
}

// returns the number of bits; bits_out one byte per bit; eq_out (optional) 2 doubles per data subcarrier
int64_t orc_demodulate(int cfg, int mod, const float* signal, int64_t n, const double* chRe, const double* chIm,
                       uint8_t* bits_out, double* eq_out) {
    const Ofdm& o = ofdm_config(cfg);
    dvec cr(chRe, chRe + o.FFT_SIZE), ci(chIm, chIm + o.FFT_SIZE), eq;
    const bvec bits = demodulate(o, signal, n, mod, cr, ci, eq_out ? &eq : nullptr);
    if (bits_out && !bits.empty()) std::memcpy(bits_out, bits.data(), bits.size());
    if (eq_out && !eq.empty()) std::memcpy(eq_out, eq.data(), eq.size() * 8);
    return (int64_t)bits.size();
}

uint32_t orc_crc32(const uint8_t* data, int64_t n) { return crc32(data, (size_t)n); }

int64_t orc_majority_vote(const uint8_t* bits, int64_t n, int rep, uint8_t* out) {
    const bvec v = majority_vote(bvec(bits, bits + n), rep);
    if (!v.empty()) std::memcpy(out, v.data(), v.size());
    return (int64_t)v.size();
}
int64_t orc_bits_to_bytes(const uint8_t* bits, int64_t n, uint8_t* out) {
    const bvec v = bits_to_bytes(bvec(bits, bits + n));
    if (!v.empty()) std::memcpy(out, v.data(), v.size());
    return (int64_t)v.size();
}

// modulateOFDM: returns numSymbols; out receives numSymbols*SYMBOL_LEN samples (pass NULL to size)
int64_t orc_modulate(int cfg, int mod, const uint8_t* bits_in, int64_t nbits, float* out) {
    const Ofdm& o = ofdm_config(cfg);
    bvec bits(bits_in, bits_in + nbits);
    const std::vector<fvec> s = modulate(o, bits, mod);
    if (out) for (size_t i = 0; i < s.size(); i++) std::memcpy(out + i * o.SYMBOL_LEN, s[i].data(), o.SYMBOL_LEN * 4);
    return (int64_t)s.size();
}

static int64_t emit(const fvec& v, float* out, int64_t cap) {
    if (out && (int64_t)v.size() <= cap) std::memcpy(out, v.data(), v.size() * 4);
    return (int64_t)v.size();
}

int64_t orc_build_transmit_signal(int cfg, int mod, int rep, const uint8_t* data, int64_t len, const uint8_t* name,
                                  int name_len, float* out, int64_t cap) {
    return emit(build_transmit_signal(ofdm_config(cfg), data, len, mod, name, name_len, rep), out, cap);
}
int64_t orc_build_metadata_frame(int cfg, int mod, int rep, int32_t totalChunks, int32_t totalFileSize, int32_t chunkSize,
                                 const uint8_t* name, int name_len, float* out, int64_t cap) {
    return emit(build_chunk_frame(ofdm_config(cfg), build_metadata_payload(totalChunks, totalFileSize, chunkSize, name, name_len),
                                  mod, rep, true), out, cap);
}
int64_t orc_build_data_chunk_frame(int cfg, int mod, int rep, const uint8_t* chunk, int len, int32_t seq, float* out, int64_t cap) {
    return emit(build_chunk_frame(ofdm_config(cfg), build_data_chunk_payload(chunk, len, seq), mod, rep, false), out, cap);
}
// generic: any payload bytes through buildChunkOFDMFrame (for malformed-frame tests)
int64_t orc_build_raw_chunk_frame(int cfg, int mod, int rep, const uint8_t* payload, int len, int first, float* out, int64_t cap) {
    return emit(build_chunk_frame(ofdm_config(cfg), bvec(payload, payload + len), mod, rep, first != 0), out, cap);
}

int64_t orc_estimate_frame_samples(int cfg, int mod, int64_t payloadBytes, int rep) {
    return estimate_frame_samples(ofdm_config(cfg), payloadBytes, mod, rep);
}

void orc_decode_chunk_frame(int cfg, int mod, int rep, const float* frame, int64_t n, OrcResult* r, uint8_t* bytes_out, int64_t cap) {
    decode_chunk_frame_impl(ofdm_config(cfg), frame, n, mod, rep, r, bytes_out, cap);
}
void orc_decode_received_signal(int cfg, int mod, int rep, const float* x, int64_t n, OrcResult* r, uint8_t* bytes_out, int64_t cap) {
    decode_received_signal_impl(ofdm_config(cfg), x, n, mod, rep, r, bytes_out, cap);
}

int orc_loopback_size(void) { return (int)sizeof(OrcLoopback); }
// analyzeLoopback (modem.js:978-1081); mags needs SUB_END-SUB_START+1 doubles
void orc_analyze_loopback(int cfg, int mod, int rep, const float* x, int64_t n, const uint8_t* test_data, int test_len,
                          OrcLoopback* r, double* mags) {
    analyze_loopback_impl(ofdm_config(cfg), x, n, mod, rep <= 0 ? 1 : rep, test_data, test_len, r, mags);
}

// ============================================================================
// Streaming receiver — app.js:563-595 (RingBuffer) and :706-981
// (StreamingReceiver) with the DOM / IndexedDB / log calls removed.  The
// reference's `_demodulateFrame` is async only because of IndexedDB awaits
// (app.js:941, :952); here every await resolves before the next audio block,
// so the reset to IDLE happens inside the same processAudioBlock call.
// ============================================================================
struct Ring {   // app.js:563-595
    fvec buffer; int64_t capacity, writePos = 0, totalWritten = 0;
    explicit Ring(int64_t cap) : buffer((size_t)cap, 0.0f), capacity(cap) {}
    void write(const float* s, int64_t n) {
        for (int64_t i = 0; i < n; i++) { buffer[writePos] = s[i]; writePos = (writePos + 1) % capacity; }
        totalWritten += n;
    }
    bool getRange(int64_t globalStart, int64_t length, fvec& out) const {
        const int64_t oldest = totalWritten - capacity;
        if (globalStart < oldest) return false;
        out.resize((size_t)length);
        const int64_t startInBuf = ((globalStart % capacity) + capacity) % capacity;
        for (int64_t i = 0; i < length; i++) out[i] = buffer[(startInBuf + i) % capacity];
        return true;
    }
    float at(int64_t g) const { return buffer[((g % capacity) + capacity) % capacity]; }
};

struct StreamEvent { OrcResult res; bvec bytes; };

struct Stream {
    enum { IDLE = 0, PREAMBLE_DETECTED = 1, COLLECTING_FRAME = 2, DEMODULATING = 3 };
    const Ofdm& o; int mod, repetition;
    Ring ring;
    int state = IDLE; bool metaReceived = false; int asmChunkSize = 0;
    int64_t half; double acP = 0, acRa = 0, acRb = 0; bool acInitialized = false; int64_t acScanPos = 0;
    int64_t preambleGlobalPos = -1, expectedFrameEnd = -1;
    double dcAlpha = 0.999, dcMean = 0;
    int64_t framesDecoded = 0, frameErrors = 0, falsePositives = 0;
    fvec pre1; double pre1Energy = 0;
    std::vector<StreamEvent> events;
    // assembler counters (app.js:597-665, storage removed)
    int32_t asmTotalChunks = 0; bvec bitmap; bool haveBitmap = false; int64_t receivedCount = 0, crcErrors = 0, duplicates = 0;
    double lastCoarse = 0;

    // app.js:707-746
    Stream(int cfg, int mod_, int rep) : o(ofdm_config(cfg)), mod(mod_), repetition(rep <= 0 ? 1 : rep),
        ring(estimate_frame_samples(ofdm_config(cfg), 4096 + 16, mod_, rep <= 0 ? 1 : rep) * 3 + 8192) {
        half = o.FFT_SIZE / 2;
        pre1 = preamble1(o);
        for (float v : pre1) pre1Energy += (double)v * v;
    }

    // app.js:749-773
    void processAudioBlock(const float* in, int64_t n) {
        fvec cleaned((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            dcMean = dcAlpha * dcMean + (1 - dcAlpha) * (double)in[i];
            cleaned[i] = (float)((double)in[i] - dcMean);
        }
        ring.write(cleaned.data(), n);
        switch (state) {
            case IDLE: scanForPreamble(); break;
            case PREAMBLE_DETECTED: refineAndCollect(); break;
            case COLLECTING_FRAME: checkFrameComplete(); break;
            default: break;
        }
    }

    // app.js:775-847
    void scanForPreamble() {
        const int64_t totalWritten = ring.totalWritten;
        const int64_t oldest = totalWritten - ring.capacity;
        if (acScanPos < oldest + 2 * half) { acScanPos = std::max<int64_t>(oldest + 2 * half, 0); acInitialized = false; }
        const int64_t scanEnd = totalWritten - 2 * half;
        if (acScanPos > scanEnd) return;
        if (!acInitialized) {
            acP = 0; acRa = 0; acRb = 0;
            if (acScanPos < oldest) return;
            for (int64_t m = 0; m < half; m++) {
                const double a = ring.at(acScanPos + m), b = ring.at(acScanPos + m + half);
                acP += a * b; acRa += a * a; acRb += b * b;
            }
            acInitialized = true;
        }
        const double minEnergy = 0.001;
        double bestMetric = 0; int64_t bestPos = -1;
        while (acScanPos <= scanEnd) {
            if (acRa > minEnergy && acRb > minEnergy) {
                const double metric = (acP * acP) / (acRa * acRb);
                if (metric > 0.5 && metric > bestMetric) { bestMetric = metric; bestPos = acScanPos; }
            }
            if (acScanPos < scanEnd) {
                if (acScanPos < ring.totalWritten - ring.capacity) break;   // getRange -> null (:820)
                const double aOut = ring.at(acScanPos), mid = ring.at(acScanPos + half), bIn = ring.at(acScanPos + 2 * half);
                acP += mid * bIn - aOut * mid;
                acRa += mid * mid - aOut * aOut;
                acRb += bIn * bIn - mid * mid;
            }
            // NOTE app.js:826 increments even when no slide happened (acScanPos == scanEnd):
            // the sums then lag the position label by one sample until the next re-init.
            acScanPos++;
            if (bestMetric > 0.5 && bestPos >= 0) {
                if (acRa > minEnergy && acRb > minEnergy) {
                    const double currentMetric = (acP * acP) / (acRa * acRb);
                    if (currentMetric < bestMetric * 0.7) {
                        preambleGlobalPos = bestPos; lastCoarse = bestMetric; state = PREAMBLE_DETECTED; return;
                    }
                }
            }
        }
        if (bestMetric > 0.5 && bestPos >= 0) { preambleGlobalPos = bestPos; lastCoarse = bestMetric; state = PREAMBLE_DETECTED; }
    }

    double lastFine = 0;
    // app.js:849-898
    void refineAndCollect() {
        const int64_t pLen = (int64_t)pre1.size();
        const int64_t searchRadius = o.CP_LEN * 3;
        const int64_t needed = preambleGlobalPos + pLen + searchRadius;
        if (ring.totalWritten < needed) return;
        const int64_t fineStart = std::max<int64_t>(ring.totalWritten - ring.capacity, preambleGlobalPos - searchRadius);
        const int64_t fineEnd = std::min<int64_t>(ring.totalWritten - pLen, preambleGlobalPos + searchRadius);
        double bestMetric = -std::numeric_limits<double>::infinity();
        int64_t bestPos = preambleGlobalPos;
        for (int64_t d = fineStart; d <= fineEnd; d++) {
            if (d < ring.totalWritten - ring.capacity) continue;
            double corr = 0, sEnergy = 0;
            for (int64_t i = 0; i < pLen; i++) {
                const double s = ring.at(d + i);
                corr += s * pre1[i]; sEnergy += s * s;
            }
            const double denom = std::sqrt(sEnergy * pre1Energy);
            if (denom > 0.001) { const double m = corr / denom; if (m > bestMetric) { bestMetric = m; bestPos = d; } }
        }
        if (bestMetric < 0.1) { state = IDLE; acInitialized = false; falsePositives++; return; }
        preambleGlobalPos = bestPos; lastFine = bestMetric;
        const int64_t maxPayload = metaReceived ? (int64_t)(asmChunkSize ? asmChunkSize : 4096) + 11 : 280;
        expectedFrameEnd = preambleGlobalPos + estimate_frame_samples(o, maxPayload, mod, repetition);
        state = COLLECTING_FRAME;
    }

    // app.js:900-905
    void checkFrameComplete() {
        if (ring.totalWritten < expectedFrameEnd) return;
        state = DEMODULATING;
        demodulateFrame();
    }

    // app.js:907-972
    void demodulateFrame() {
        const int64_t frameLen = expectedFrameEnd - preambleGlobalPos;
        fvec frame;
        StreamEvent ev; result_clear(&ev.res);
        if (!ring.getRange(preambleGlobalPos, frameLen, frame)) {
            frameErrors++;
            ev.res.status = ORC_ERR_RING_OVERWRITTEN; ev.res.preamble_idx = preambleGlobalPos;
            events.push_back(ev);
            resetToIdle(); return;
        }
        double mx = 0;
        for (float v : frame) mx = js_max(mx, std::fabs((double)v));
        if (mx > 1e-6) for (float& v : frame) v = (float)((double)v / mx);
        ev.bytes.resize((size_t)(frameLen / o.SYMBOL_LEN) * o.numDataSubs() * constellation(mod).bps / 8 + 8);
        decode_chunk_frame_impl(o, frame.data(), frameLen, mod, repetition, &ev.res, ev.bytes.data(), (int64_t)ev.bytes.size());
        ev.bytes.resize((size_t)std::max(0, ev.res.n_bytes));
        ev.res.preamble_idx = preambleGlobalPos;
        ev.res.coarse_metric = lastCoarse; ev.res.fine_metric = lastFine;
        if (ev.res.status != ORC_OK) { frameErrors++; events.push_back(ev); resetToIdle(); return; }
        framesDecoded++;
        if (ev.res.frame_type == 0xFE) {
            if (ev.res.crc_valid) {                                  // app.js:940-942, :610-618
                asmTotalChunks = ev.res.total_chunks; asmChunkSize = ev.res.chunk_size;
                const int64_t nb = ((int64_t)asmTotalChunks + 7) / 8;  // Math.ceil(totalChunks/8)
                bitmap.assign((size_t)std::max<int64_t>(0, nb), 0); haveBitmap = true;
                receivedCount = 0; crcErrors = 0;
                metaReceived = true;
            } else frameErrors++;
        } else if (ev.res.frame_type == 0xFF) {                      // app.js:628-643
            const int32_t seq = ev.res.seq_num;
            if (haveBitmap && !(seq >= asmTotalChunks)) {
                if (!ev.res.crc_valid) crcErrors++;
                else if (seq < 0) receivedCount++;   // typed-array read/write at a negative index is a no-op in JS
                else if ((size_t)(seq >> 3) < bitmap.size()) {
                    if (bitmap[seq >> 3] & (1 << (seq & 7))) duplicates++;
                    else { bitmap[seq >> 3] |= (uint8_t)(1 << (seq & 7)); receivedCount++; }
                }
            }
        }
        events.push_back(ev);
        resetToIdle();
    }

    // app.js:974-981 (`expectedFrameEnd || ...`: -1 is truthy, 0 is not reachable)
    void resetToIdle() {
        acScanPos = expectedFrameEnd ? expectedFrameEnd : (preambleGlobalPos + o.SYMBOL_LEN);
        acInitialized = false; preambleGlobalPos = -1; expectedFrameEnd = -1; state = IDLE;
    }
};

void* orc_stream_new(int cfg, int mod, int rep) { return new Stream(cfg, mod, rep); }
void orc_stream_free(void* h) { delete (Stream*)h; }

// Feed `n` samples as consecutive blocks of `block_len` (the last may be short).
void orc_stream_push(void* h, const float* x, int64_t n, int64_t block_len) {
    Stream* s = (Stream*)h;
    for (int64_t off = 0; off < n; off += block_len) s->processAudioBlock(x + off, std::min(block_len, n - off));
}
int64_t orc_stream_num_events(void* h) { return (int64_t)((Stream*)h)->events.size(); }
// copy event i (result + decoded bytes)
int orc_stream_event(void* h, int64_t i, OrcResult* r, uint8_t* bytes_out, int64_t cap) {
    Stream* s = (Stream*)h;
    if (i < 0 || i >= (int64_t)s->events.size()) return -1;
    *r = s->events[i].res;
    copy_bytes(s->events[i].bytes, bytes_out, cap);
    return 0;
}
void orc_stream_clear_events(void* h) { ((Stream*)h)->events.clear(); }
// counters: framesDecoded, frameErrors, falsePositives, receivedCount, crcErrors, duplicates, metaReceived, state,
//           acScanPos, totalWritten
void orc_stream_counters(void* h, int64_t* out) {
    Stream* s = (Stream*)h;
    out[0] = s->framesDecoded; out[1] = s->frameErrors; out[2] = s->falsePositives; out[3] = s->receivedCount;
    out[4] = s->crcErrors; out[5] = s->duplicates; out[6] = s->metaReceived; out[7] = s->state;
    out[8] = s->acScanPos; out[9] = s->ring.totalWritten;
}
double orc_stream_dc_mean(void* h) { return ((Stream*)h)->dcMean; }

// ============================================================================
// Synthetic channel (test-input generator; not in the reference).  SplitMix64
// payload PRNG and Box-Muller Gaussian in float64, results stored as float32
// so oracle and GPU consume identical bytes.
// ============================================================================
static inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_random_bytes(uint64_t seed, uint8_t* out, int64_t n) {
    uint64_t s = seed;
    for (int64_t i = 0; i < n; i += 8) {
        uint64_t v = splitmix64(s);
        for (int j = 0; j < 8 && i + j < n; j++) out[i + j] = (uint8_t)(v >> (8 * j));
    }
}
// add white Gaussian noise of standard deviation sigma, in place
void orc_add_awgn(float* x, int64_t n, double sigma, uint64_t seed) {
    uint64_t s = seed;
    for (int64_t i = 0; i < n; i += 2) {
        const double u1 = ((double)(splitmix64(s) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        const double u2 = (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0);
        const double r = std::sqrt(-2.0 * std::log(u1)) * sigma;
        x[i] = (float)((double)x[i] + r * std::cos(2 * M_PI * u2));
        if (i + 1 < n) x[i + 1] = (float)((double)x[i + 1] + r * std::sin(2 * M_PI * u2));
    }
}
// mean power of x[a:b)
double orc_power(const float* x, int64_t a, int64_t b) {
    double p = 0;
    for (int64_t i = a; i < b; i++) p += (double)x[i] * x[i];
    return b > a ? p / (double)(b - a) : 0;
}
// y[n] = sum_t gain[t] * x[n - delay[t]]  (causal FIR), out may not alias x
void orc_multipath(const float* x, int64_t n, const int* delay, const double* gain, int ntaps, float* out) {
    for (int64_t i = 0; i < n; i++) {
        double acc = 0;
        for (int t = 0; t < ntaps; t++) if (i - delay[t] >= 0) acc += gain[t] * (double)x[i - delay[t]];
        out[i] = (float)acc;
    }
}
// frequency shift by df Hz via the analytic signal (FFT-free Hilbert FIR, 255 taps, Hamming window)
void orc_cfo(const float* x, int64_t n, double df, double fs, float* out) {
    const int L = 255, mid = L / 2;
    std::vector<double> h(L, 0.0);
    for (int i = 0; i < L; i++) {
        const int k = i - mid;
        if (k & 1) h[i] = (2.0 / (M_PI * k)) * (0.54 - 0.46 * std::cos(2 * M_PI * i / (L - 1)));
    }
    for (int64_t i = 0; i < n; i++) {
        double q = 0;
        for (int t = 0; t < L; t++) { const int64_t j = i + mid - t; if (j >= 0 && j < n) q += h[t] * (double)x[j]; }
        const double ph = 2 * M_PI * df * (double)i / fs;
        out[i] = (float)((double)x[i] * std::cos(ph) - q * std::sin(ph));
    }
}

// counter-based variants (same formulas as the device-side synthetic input helpers, amb_tx.cuh)
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_synth_bytes(uint64_t seed, int64_t frame, uint8_t* out, int64_t n) {
    for (int64_t j = 0; j < n; j++) {
        const uint64_t v = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(frame * 65536 + (j >> 3) + 1));
        out[j] = (uint8_t)(v >> (8 * (j & 7)));
    }
}
void orc_add_awgn_ctr(float* x, int64_t n, double sigma, uint64_t seed, int64_t g0) {
    for (int64_t i = 0; i < n; i++) {
        const int64_t g = g0 + i;
        const uint64_t c = (uint64_t)(g >> 1);
        const uint64_t a = mix64(seed + 0x9E3779B97F4A7C15ull * (2 * c + 1));
        const uint64_t b = mix64(seed + 0x9E3779B97F4A7C15ull * (2 * c + 2));
        const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        const double r = std::sqrt(-2.0 * std::log(u1));
        const double ph = 6.283185307179586 * u2;
        x[i] = (float)((double)x[i] + sigma * ((g & 1) ? r * std::sin(ph) : r * std::cos(ph)));
    }
}

// ---- timed legs for bench.py's cpu_baseline (whole-path calls, no Python in the loop) ----
// Runs the streaming receiver over x in block_len blocks; returns frames decoded with crc_valid.
int64_t orc_bench_stream(int cfg, int mod, int rep, const float* x, int64_t n, int64_t block_len, int64_t* payload_bytes) {
    Stream s(cfg, mod, rep);
    for (int64_t off = 0; off < n; off += block_len) s.processAudioBlock(x + off, std::min(block_len, n - off));
    int64_t ok = 0, pb = 0;
    for (auto& e : s.events) if (e.res.status == ORC_OK && e.res.crc_valid) { ok++; if (e.res.frame_type == 0xFF) pb += e.res.data_len; }
    if (payload_bytes) *payload_bytes = pb;
    return ok;
}

}  // extern "C"
