// Per-handle constant tables of the OFDM receive path (host-generated, copied to HBM once).
// All values are produced by the same float64 operation sequence the reference uses, so that the
// float64-faithful device path can reproduce modem.js bit for bit (see DESIGN.md §"Faithful path").
#pragma once
#include <stdint.h>

#define AMB_FFT 512
#define AMB_HALF 256
#define AMB_MAX_SYM 768
#define AMB_MAX_PILOTS 16
#define AMB_MAX_DATA 208

struct AmbTables {
    // --- OFDM config, modem.js:69-85 ---
    int cfg_id, mod, rep, bps;
    int cp_len, sym_len, sub_start, sub_end, n_pilots, n_data;
    int fine_radius;   // 3*CP_LEN, modem.js:571
    int pre1_sign;     // +1: preamble-1 halves identical (even SUB_START), -1: negated (odd), modem.js:162
    int n_points;
    int pad0;
    int pilots[AMB_MAX_PILOTS];
    int16_t data_k[AMB_MAX_DATA];   // FFT bin of data subcarrier di (ascending k, pilots skipped, modem.js:408-409)
    uint8_t is_pilot[AMB_FFT];
    // --- generated symbols ---
    double known_re[AMB_FFT];       // generateChannelEstSymbol knownRe (knownIm = 0), modem.js:186-200
    float pre1[AMB_MAX_SYM];        // generatePreambleSymbol1 incl. CP, Float32 (modem.js:158-170, :202-208)
    double pre1_energy;             // sum pre1[i]^2, float64 sequential (modem.js:568-569)
    // --- constellation, modem.js:107-131 ---
    double pts[32];                 // (re, im) per index
    double qam_thr;                 // midpoint between the two positive 16-QAM levels
    // --- forward FFT twiddles: tw[half + j], j < half, half = 1..256; running product of modem.js:41-43 ---
    double tw_re[AMB_FFT];
    double tw_im[AMB_FFT];
    // --- transmit side (modem.js:172-184, :186-200, :15-24): preamble-2 and CE symbols incl. CP, inverse twiddles ---
    float pre2[AMB_MAX_SYM];
    float ce_sym[AMB_MAX_SYM];
    double twi_re[AMB_FFT];         // inverse-transform running product (angle +2*pi/size)
    double twi_im[AMB_FFT];
    // --- CRC-32 table, modem.js:443-451 ---
    uint32_t crc_table[256];
    // --- EMA DC removal, app.js:734, :753 ---
    double dc_alpha, dc_beta;       // 0.999 and (1 - 0.999) as evaluated in float64
};

// host-only generators (amb_tables.cpp)
int amb_tables_build(AmbTables* t, int config, int constellation, int repetition);
int64_t amb_tables_estimate_frame_samples(int config, int constellation, int64_t payload_bytes, int repetition);
void amb_tables_preamble2(int config, float* out);   // generatePreambleSymbol2 (TX side, tests/bench input)
void amb_tables_ce_symbol(int config, float* out, double* known_re);
