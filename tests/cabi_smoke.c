/* The C-ABI boundary from plain C: compiled with `gcc -std=c99 -pedantic -Wall -Wextra -Werror` against
 * include/audio_modem_b200.h (which proves the header is C, not C++), linked to libaudiomodem_b200.so.
 *
 *   cabi_smoke            host helpers only (no GPU needed): version, status strings, frame lengths, config table;
 *                         amb_create must fail with AMB_E_NODEVICE on a machine without a GPU
 *   cabi_smoke gpu        create -> buildTransmitSignal -> decodeReceivedSignal -> detectPreamble -> streaming push -> destroy,
 *                         checking the decoded bytes against what was sent
 * Exit code 0 = pass.  Used by tests/test_abi.py (CPU) and tests/test_gpu_parity.py::test_c_abi_from_compiled_c (GPU). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "audio_modem_b200.h"

#define CHECK(cond)                                                               \
    do {                                                                          \
        if (!(cond)) { fprintf(stderr, "cabi_smoke: %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static int host_side(void) {
    int32_t info[8], pilots[16];
    CHECK(amb_version() == AMB_VERSION);
    CHECK(strcmp(amb_status_string(AMB_ST_PREAMBLE_NOT_DETECTED), "Preamble not detected") == 0);
    CHECK(strcmp(amb_status_string(AMB_ST_NO_DATA_AFTER_CE), "No data after CE") == 0);
    /* estimateFrameSamples (modem.js:863-874): 16-QAM 4 KB chunk = 44 symbols of 576 samples */
    CHECK(amb_estimate_frame_samples(AMB_CFG_STANDARD, AMB_MOD_QAM16, 4096 + 11, 1) == 44 * 576);
    CHECK(amb_config_info(AMB_CFG_NARROWBAND, info, pilots) == AMB_OK);
    CHECK(info[0] == 512 && info[1] == 256 && info[2] == 768 && info[3] == 44100 && info[4] == 35 && info[5] == 58 && info[6] == 3);
    return 0;
}

static int gpu_side(void) {
    amb_modem* m = NULL;
    uint8_t payload[300], name[] = {'c', '.', 'b', 'i', 'n'};
    int64_t n, idx = -2;
    double metric = 0.0;
    float* sig;
    amb_frame_result r;
    uint8_t* bytes;
    int64_t bytes_stride;
    int i, got = 0;
    int32_t who[4];
    int64_t counters[8];

    for (i = 0; i < (int)sizeof(payload); i++) payload[i] = (uint8_t)(i * 7 + 3);
    CHECK(amb_create(&m, 0, AMB_CFG_STANDARD, AMB_MOD_QPSK, 1) == AMB_OK && m != NULL);
    CHECK(amb_modem_info(m, who) == AMB_OK && who[0] == AMB_CFG_STANDARD && who[1] == AMB_MOD_QPSK && who[2] == 1 && who[3] == 0);
    n = amb_transmit_signal_samples(m, (int64_t)sizeof(payload), (int)sizeof(name));
    CHECK(n > 0);
    sig = (float*)malloc((size_t)n * sizeof(float));
    bytes_stride = n / 4 + 64;
    bytes = (uint8_t*)malloc((size_t)bytes_stride);
    CHECK(sig != NULL && bytes != NULL);
    CHECK(amb_build_transmit_signals(m, payload, (int64_t)sizeof(payload), (int)sizeof(payload), 1, name, (int)sizeof(name), sig, n, AMB_MEM_HOST) == AMB_OK);
    /* decodeReceivedSignal (modem.js:557): the payload comes back with a valid CRC and the file name */
    CHECK(amb_decode_received_signals(m, sig, n, NULL, 1, &r, bytes, bytes_stride, AMB_MEM_HOST) == AMB_OK);
    CHECK(r.status == AMB_ST_OK && r.frame_type == AMB_FRAME_LEGACY && r.crc_valid == 1 && r.data_len == (int32_t)sizeof(payload));
    CHECK(memcmp(bytes + r.data_off, payload, sizeof(payload)) == 0);
    CHECK(r.name_len == (int32_t)sizeof(name) && memcmp(bytes + r.name_off, name, sizeof(name)) == 0);
    CHECK(r.preamble_idx == 13230);              /* the lead-in silence of buildTransmitSignal (0.3 s, modem.js:540) */
    /* detectPreamble on the preprocessed signal finds the same frame start (modem.js:286) */
    {
        float* pre = (float*)malloc((size_t)n * sizeof(float));
        CHECK(pre != NULL);
        CHECK(amb_preprocess_signal(m, sig, n, NULL, 1, pre, AMB_MEM_HOST) == AMB_OK);
        CHECK(amb_detect_preamble(m, pre, n, NULL, 1, &idx, &metric, AMB_MEM_HOST) == AMB_OK);
        CHECK(idx >= 13230 && idx < 13230 + 64 && metric > 0.9);
        free(pre);
    }
    /* a stream of silence: nothing found, the sample counter advances (app.js:749) */
    {
        float* block = (float*)calloc(8 * 4096, sizeof(float));
        CHECK(block != NULL);
        CHECK(amb_stream_open(m, 1, 4096, 8 * 4096) == AMB_OK);
        CHECK(amb_stream_push(m, block, 8 * 4096, 8 * 4096, AMB_MEM_HOST, &r, 1, &got, bytes, bytes_stride, AMB_MEM_HOST) == AMB_OK && got == 0);
        CHECK(amb_stream_counters(m, 0, counters) == AMB_OK && counters[0] == 0 && counters[7] == 8 * 4096);
        CHECK(amb_stream_close(m) == AMB_OK);
        free(block);
    }
    CHECK(amb_kernel_launches(m) > 0);
    free(sig);
    free(bytes);
    amb_destroy(m);
    return 0;
}

int main(int argc, char** argv) {
    int rc = host_side();
    if (rc) return rc;
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        rc = gpu_side();
        if (rc) return rc;
        printf("cabi_smoke: C-ABI round trip on the GPU ok\n");
    } else {
        amb_modem* m = NULL;
        const int st = amb_create(&m, 0, AMB_CFG_STANDARD, AMB_MOD_QPSK, 1);
        if (st == AMB_OK) amb_destroy(m);                 /* a GPU is present: fine */
        else CHECK(st == AMB_E_NODEVICE && m == NULL);    /* no CPU fallback */
        printf("cabi_smoke: host side ok (amb_create -> %d)\n", st);
    }
    return 0;
}
