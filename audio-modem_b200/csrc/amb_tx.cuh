// Transmit-side device logic: modulateOFDM (modem.js:322-362), buildChunkOFDMFrame (modem.js:718-756) and the
// chunk payload builders (modem.js:666-714), float64-faithful so that a frame built here is bit-identical to
// the reference's Float32Array.  Also the synthetic-input helpers used by bench.py and the tests (counter-based
// payload bytes and AWGN; not part of the reference).  Same host/device SPMD shim as amb_device.cuh.
#pragma once
#include "amb_device.cuh"

// --- synthetic input (SplitMix64 output function on a counter) -----------------------------------------
AMB_HD uint64_t amb_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// byte j of synthetic chunk `frame`
AMB_HD uint8_t amb_synth_byte(uint64_t seed, int64_t frame, int64_t j) {
    const uint64_t v = amb_mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(frame * 65536 + (j >> 3) + 1));
    return (uint8_t)(v >> (8 * (j & 7)));
}
// standard normal for global sample index g (Box-Muller on a counter pair)
AMB_HD double amb_synth_normal(uint64_t seed, int64_t g) {
    const uint64_t c = (uint64_t)(g >> 1);
    const uint64_t a = amb_mix64(seed + 0x9E3779B97F4A7C15ull * (2 * c + 1));
    const uint64_t b = amb_mix64(seed + 0x9E3779B97F4A7C15ull * (2 * c + 2));
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    const double ph = 6.283185307179586 * u2;
    return (g & 1) ? r * sin(ph) : r * cos(ph);
}

// --- chunk payloads ---------------------------------------------------------------------------------------
// buildDataChunkPayload, modem.js:694-714: [0xFF][seq:4][len:2][data][crc:4]; returns the payload size
AMB_HD int amb_data_chunk_payload(const uint32_t* crc_table, const uint8_t* data, int data_len, int32_t seq, uint8_t* out) {
    int off = 0;
    out[off++] = 0xFF;
    out[off++] = (uint8_t)(seq >> 24); out[off++] = (uint8_t)(seq >> 16); out[off++] = (uint8_t)(seq >> 8); out[off++] = (uint8_t)seq;
    out[off++] = (uint8_t)(data_len >> 8); out[off++] = (uint8_t)data_len;
    for (int i = 0; i < data_len; i++) out[off++] = data[i];
    const uint32_t cs = amb_crc32_bytes(crc_table, out, off);
    out[off++] = (uint8_t)(cs >> 24); out[off++] = (uint8_t)(cs >> 16); out[off++] = (uint8_t)(cs >> 8); out[off++] = (uint8_t)cs;
    return off;
}
// buildMetadataPayload, modem.js:666-692
AMB_HD int amb_metadata_payload(const uint32_t* crc_table, int32_t total_chunks, int32_t total_file_size, int32_t chunk_size,
                                const uint8_t* name, int name_len, uint8_t* out) {
    if (name_len > 255) name_len = 255;
    int off = 0;
    out[off++] = 0xFE;
    out[off++] = (uint8_t)(total_chunks >> 24); out[off++] = (uint8_t)(total_chunks >> 16); out[off++] = (uint8_t)(total_chunks >> 8); out[off++] = (uint8_t)total_chunks;
    out[off++] = (uint8_t)(total_file_size >> 24); out[off++] = (uint8_t)(total_file_size >> 16); out[off++] = (uint8_t)(total_file_size >> 8); out[off++] = (uint8_t)total_file_size;
    out[off++] = (uint8_t)(chunk_size >> 8); out[off++] = (uint8_t)chunk_size;
    out[off++] = (uint8_t)name_len;
    for (int i = 0; i < name_len; i++) out[off++] = name[i];
    const uint32_t cs = amb_crc32_bytes(crc_table, out, off);
    out[off++] = (uint8_t)(cs >> 24); out[off++] = (uint8_t)(cs >> 16); out[off++] = (uint8_t)(cs >> 8); out[off++] = (uint8_t)cs;
    return off;
}

AMB_HD int64_t amb_tx_num_symbols(const AmbTables* T, int64_t payload_bytes) {
    const int64_t bits = payload_bytes * 8 * T->rep, per = (int64_t)T->n_data * T->bps;
    return (bits + per - 1) / per;
}

// --- one OFDM data symbol (modem.js:334-358), one warp -------------------------------------------------------
// Serial bit b of the transmitted stream is bit (7 - (q & 7)) of payload[q >> 3], q = b / rep (bytesToBits +
// repeatBits, modem.js:460-485); bits past the payload are the zero padding of modem.js:329.
AMB_HD int amb_tx_bit(const uint8_t* payload, int64_t payload_bits, int rep, int64_t b) {
    const int64_t q = b / rep;
    if (q >= payload_bits) return 0;
    return (payload[q >> 3] >> (7 - (int)(q & 7))) & 1;
}

// re/im: 512 doubles each (scratch).  out: sym_len float32 samples (CP + body, un-normalised).  idx_of(di) = constellation
// index of data subcarrier di (constellationMap, modem.js:133-138).
template <class IdxFn>
AMB_HD void amb_tx_symbol_from(const AmbTables* T, IdxFn idx_of, double* re, double* im, float* out) {
    AMB_WARP_FOR(i, AMB_FFT) { re[i] = 0.0; im[i] = 0.0; }
    AMB_WARP_SYNC();
    // spectrum written straight into bit-reversed order (the reference's permutation is an involution)
    AMB_WARP_FOR(di, T->n_data) {
        const int k = T->data_k[di];
        const int idx = idx_of(di);
        const double vr = T->pts[2 * (idx % T->n_points)], vi = T->pts[2 * (idx % T->n_points) + 1];
        re[amb_rev9(k)] = vr; im[amb_rev9(k)] = vi;
        re[amb_rev9(AMB_FFT - k)] = vr; im[amb_rev9(AMB_FFT - k)] = -vi;     // Hermitian, modem.js:353
    }
    AMB_WARP_FOR(p, T->n_pilots) {
        const int k = T->pilots[p];                                           // pilots carry (1, 0), modem.js:341-342
        if (k >= T->sub_start && k <= T->sub_end) { re[amb_rev9(k)] = 1.0; re[amb_rev9(AMB_FFT - k)] = 1.0; }
    }
    AMB_WARP_SYNC();
    // inverse transform: all 512 outputs are needed
    int lg = 0;
    for (int half = 1; half <= AMB_HALF; half <<= 1, lg++) {
        AMB_WARP_FOR(b, AMB_HALF) {
            const int j = b & (half - 1);
            const int i1 = ((b >> lg) << (lg + 1)) + j, i2 = i1 + half;
            double tr, ti;
            if (j == 0) { tr = re[i2]; ti = im[i2]; }
            else {
                const double wr = T->twi_re[half + j], wi = T->twi_im[half + j];
                tr = amb_sub(amb_mul(wr, re[i2]), amb_mul(wi, im[i2]));
                ti = amb_add(amb_mul(wr, im[i2]), amb_mul(wi, re[i2]));
            }
            const double ar = re[i1], ai = im[i1];
            re[i2] = amb_sub(ar, tr); im[i2] = amb_sub(ai, ti);
            re[i1] = amb_add(ar, tr); im[i1] = amb_add(ai, ti);
        }
        AMB_WARP_SYNC();
    }
    const double scale = 1.0 / AMB_FFT;
    const int cp = T->cp_len;
    AMB_WARP_FOR(i, AMB_FFT) {
        const float v = (float)amb_mul(re[i], scale);      // ifft scale then the Float32 store of addCP
        out[cp + i] = v;
        if (i >= AMB_FFT - cp) out[i - (AMB_FFT - cp)] = v;
    }
    AMB_WARP_SYNC();
}

// symbol `sym` of a byte payload: serial bit b of the transmitted stream is bit (7 - (q & 7)) of payload[q >> 3], q = b / rep
AMB_HD void amb_tx_symbol(const AmbTables* T, const uint8_t* payload, int64_t payload_bytes, int64_t sym, double* re, double* im, float* out) {
    const int64_t payload_bits = payload_bytes * 8;
    const int64_t bit0 = sym * (int64_t)T->n_data * T->bps;
    const int bps = T->bps, rep = T->rep;
    amb_tx_symbol_from(T, [=](int di) {
        int idx = 0;
        for (int b = 0; b < bps; b++) idx = (idx << 1) | amb_tx_bit(payload, payload_bits, rep, bit0 + (int64_t)di * bps + b);
        return idx;
    }, re, im, out);
}

// symbol `sym` of modulateOFDM(bits, modName) (modem.js:322-362): bits = one byte per bit, `b & 1` counts (modem.js:135),
// zeros past n_bits (the in-place padding of modem.js:329); no repetition (the caller applies repeatBits)
AMB_HD void amb_tx_symbol_bits(const AmbTables* T, const uint8_t* bits, int64_t n_bits, int64_t sym, double* re, double* im, float* out) {
    const int64_t bit0 = sym * (int64_t)T->n_data * T->bps;
    const int bps = T->bps;
    amb_tx_symbol_from(T, [=](int di) {
        int idx = 0;
        for (int b = 0; b < bps; b++) { const int64_t q = bit0 + (int64_t)di * bps + b; idx = (idx << 1) | (q < n_bits ? (bits[q] & 1) : 0); }
        return idx;
    }, re, im, out);
}

// silences of buildChunkOFDMFrame, modem.js:728-733 (Math.round of the float64 products)
AMB_HD int64_t amb_tx_silence_pre(const AmbTables* T, int first) {
    const bool acoustic = T->cp_len >= 128;
    return first ? (int64_t)floor(44100.0 * (acoustic ? 0.5 : 0.3) + 0.5) : (int64_t)floor(44100.0 * 0.05 + 0.5);
}
AMB_HD int64_t amb_tx_silence_post() { return (int64_t)floor(44100.0 * 0.02 + 0.5); }
AMB_HD int64_t amb_tx_frame_samples(const AmbTables* T, int64_t payload_bytes, int first) {
    return amb_tx_silence_pre(T, first) + (3 + amb_tx_num_symbols(T, payload_bytes)) * T->sym_len + amb_tx_silence_post();
}
