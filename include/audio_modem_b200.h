/*
 * audio_modem_b200.h — C-ABI of the B200-native OFDM receive path.
 *
 * Drop-in boundary for the receive hot path of playok/audio-modem.  The
 * reference has no FFI: its interface is the set of browser-global functions
 * of modem.js plus the StreamingReceiver class of app.js.  Every entry point
 * below names the reference function (file:line under /root/reference) it
 * replaces.  An N-API addon (node/), or the ctypes mirror in the Python
 * package, converts between these calls and the reference's JavaScript
 * signatures and result objects.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary;
 *   - every call returns 0 (AMB_OK) or a negative AMB_E_* library error;
 *     the reference's own per-frame failures ({error: "..."} objects, -1
 *     indices) are NOT library errors: they come back in amb_frame_result
 *     .status (AMB_ST_*) or as index -1, exactly where the reference puts them;
 *   - `mem` says where caller buffers live: AMB_MEM_HOST (pageable or pinned
 *     host memory; the call copies in and out) or AMB_MEM_DEVICE (device
 *     pointers on the handle's GPU; nothing is copied through the host);
 *   - all work of a handle runs on one CUDA stream (amb_cuda_stream); calls
 *     on one handle must be serialised by the caller, different handles are
 *     independent (the reference's hidden global `OFDM`, modem.js:87-98, is
 *     replaced by the explicit per-handle config);
 *   - there is NO CPU fallback: amb_create fails when no CUDA device exists.
 */
#ifndef AUDIO_MODEM_B200_H
#define AUDIO_MODEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMB_VERSION 100

/* OFDM_CONFIGS, modem.js:69-85 (setOFDMConfig, :95-98: unknown -> standard) */
#define AMB_CFG_STANDARD 0
#define AMB_CFG_ACOUSTIC 1
#define AMB_CFG_NARROWBAND 2

/* Constellations, modem.js:101-131 */
#define AMB_MOD_BPSK 0
#define AMB_MOD_QPSK 1
#define AMB_MOD_QAM16 2

#define AMB_MEM_HOST 0
#define AMB_MEM_DEVICE 1
#define AMB_MEM_INPLACE 2 /* streaming only: samples already written at amb_stream_input_window() */

/* library errors (negative) */
#define AMB_OK 0
#define AMB_E_INVALID (-1)  /* bad argument */
#define AMB_E_CUDA (-2)     /* CUDA runtime error; see amb_last_error */
#define AMB_E_NOMEM (-3)
#define AMB_E_STATE (-4)    /* call not valid in this state (e.g. stream not open) */
#define AMB_E_OVERFLOW (-5) /* caller's result buffer too small; see n_results */
#define AMB_E_NODEVICE (-6) /* no CUDA device: the library has no CPU path */

/* amb_frame_result.status: 0 or one of the reference's failures */
#define AMB_ST_OK 0
#define AMB_ST_PREAMBLE_NOT_DETECTED 1        /* modem.js:564 */
#define AMB_ST_PREAMBLE_LOW_CORRELATION 2     /* modem.js:588 */
#define AMB_ST_SIGNAL_TOO_SHORT_FOR_CE 3      /* modem.js:592 */
#define AMB_ST_NO_DATA_AFTER_CE 4             /* modem.js:600, :785 */
#define AMB_ST_DECODED_TOO_SHORT 5            /* modem.js:607, :793 */
#define AMB_ST_DECODED_TOO_SHORT_FOR_HEADER 6 /* modem.js:625 */
#define AMB_ST_INVALID_DATA_LENGTH 7          /* modem.js:634 (err_arg = dataLen) */
#define AMB_ST_FRAME_TOO_SHORT_FOR_CE 8       /* modem.js:776 */
#define AMB_ST_UNKNOWN_FRAME_TYPE 9           /* modem.js:801 (err_arg = type byte) */
#define AMB_ST_META_TOO_SHORT 10              /* modem.js:807 */
#define AMB_ST_META_TRUNCATED 11              /* modem.js:813 */
#define AMB_ST_CHUNK_TOO_SHORT 12             /* modem.js:832 */
#define AMB_ST_CHUNK_TRUNCATED 13             /* modem.js:836 */
#define AMB_ST_RING_OVERWRITTEN 14            /* app.js:912-916 (frameErrors++, no message) */

#define AMB_FRAME_LEGACY 0
#define AMB_FRAME_META 0xFE /* modem.js:661 */
#define AMB_FRAME_DATA 0xFF /* modem.js:662 */

/*
 * One decoded frame.  Field meaning follows the reference's result objects:
 * legacy {data,dataLen,fileName,crcValid,expectedCRC,actualCRC,preambleIdx}
 * (modem.js:644-653), metadata (:822-827), data chunk (:843-848).  `n_bytes`
 * decoded bytes (bitsToBytes output, modem.js:605/:791) are written to the
 * caller's byte buffer; name/data are slices of it at name_off / data_off.
 * Exactness: every byte the reference's parsers consume - header, file name,
 * payload, CRC - is the reference's bit for bit.  Bytes after that extent
 * come from symbols behind the frame (the rest of the collection window or
 * recording, i.e. noise); nothing the reference returns depends on them and
 * they are the float32 pipeline's decisions, which may differ from a
 * float64 demodulation of the same noise - or zero where those symbols were
 * not demodulated at all (amb_decode_received_signals reads the extent from
 * the header and stops there).  amb_demodulate (demodulateOFDM itself) is
 * exact for every symbol.
 */
typedef struct amb_frame_result {
    int32_t status;          /* AMB_ST_* */
    int32_t frame_type;      /* AMB_FRAME_*, or -1 when none was parsed */
    int64_t preamble_idx;    /* refined preamble start; global stream position when streaming; -1 n/a */
    int32_t seq_num;         /* data chunk seqNum (signed 32-bit, modem.js:834) */
    int32_t data_len;        /* legacy / data chunk dataLen */
    int32_t total_chunks;    /* metadata (signed, modem.js:809) */
    int32_t total_file_size; /* metadata (signed, modem.js:810) */
    int32_t chunk_size;      /* metadata */
    int32_t name_len;        /* legacy / metadata */
    int32_t name_off;
    int32_t data_off;
    uint32_t expected_crc;
    uint32_t actual_crc;
    int32_t crc_valid;
    int32_t n_bytes;         /* length of the decoded byte array */
    int32_t err_arg;
    int32_t stream;          /* recording / frame / stream index inside the batch */
    double coarse_metric;    /* Schmidl-Cox metric at the coarse index */
    double fine_metric;      /* refined cross-correlation peak */
} amb_frame_result;

typedef struct amb_modem amb_modem;

/* ---- lifecycle -------------------------------------------------------- */
int amb_version(void);
/* the reference's message for an AMB_ST_* status ("{arg}" placeholders filled by the caller) */
const char* amb_status_string(int status);
/* setOFDMConfig + the (modName, repetition) pair every decode call takes:
 * modem.js:95, :557, :770; app.js:60-66 */
int amb_create(amb_modem** out, int device, int config, int constellation, int repetition);
void amb_destroy(amb_modem* m);
const char* amb_last_error(const amb_modem* m);
int amb_synchronize(amb_modem* m);
void* amb_cuda_stream(amb_modem* m); /* the cudaStream_t all work is queued on */
int64_t amb_kernel_launches(const amb_modem* m); /* kernels launched so far by this handle */
/* per-kernel device time: amb_profile(m, 1) brackets every kernel launch with CUDA events on the handle's stream;
 * amb_profile_read returns, per kernel name (names = [cap][32] chars), the summed milliseconds and launch count since
 * the last read, and clears them */
int amb_profile(amb_modem* m, int enable);
int amb_profile_read(amb_modem* m, char* names, double* total_ms, int64_t* launches, int cap, int* n);

/* page-locked host memory for caller buffers (samples in, results and decoded bytes out): copies to and from it run
 * at full PCIe rate and asynchronously; the N-API addon wraps it in external ArrayBuffers.  Plain malloc'd memory is
 * accepted everywhere too (the driver then stages it). */
int amb_host_alloc(void** out, uint64_t bytes);
int amb_host_free(void* p);

/* ---- pure host helpers (no device work) -------------------------------- */
/* estimateFrameSamples, modem.js:863-874 */
int64_t amb_estimate_frame_samples(int config, int constellation, int64_t payload_bytes, int repetition);
/* OFDM config fields: out8 = FFT_SIZE, CP_LEN, SYMBOL_LEN, SAMPLE_RATE, SUB_START, SUB_END, nPilots,
 * numDataSubs (modem.js:69-93); pilots receives nPilots bins */
int amb_config_info(int config, int32_t* out8, int32_t* pilots);
/* what a handle was created with: out4 = {config, constellation, repetition, device} */
int amb_modem_info(const amb_modem* m, int32_t* out4);
/* generatePreambleSymbol1 / generateChannelEstSymbol, modem.js:158-170, :186-200 */
int amb_preamble_symbol1(int config, float* out_symbol_len);
int amb_channel_est_symbol(int config, float* out_symbol_len, double* known_re_512);

/* ---- batched operators (one launch covers n independent inputs) --------- */
/* Inputs are rows of a [n][stride] array; lens[i] (NULL = all `stride`) is the valid length of row i. */

/* crc32, modem.js:453-457 */
int amb_crc32(amb_modem* m, const uint8_t* data, int64_t stride, const int64_t* lens, int n, uint32_t* out, int mem);
/* preprocessSignal, modem.js:213-232; out has the same [n][stride] shape */
int amb_preprocess_signal(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, float* out, int mem);
/* detectPreamble, modem.js:286-319; idx_out[i] = index or -1; metric_out (optional) = best metric */
int amb_detect_preamble(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, int64_t* idx_out,
                        double* metric_out, int mem);
/* detectPreambleCrossCorr, modem.js:235-283 */
int amb_detect_preamble_xcorr(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n,
                              int64_t* idx_out, double* metric_out, int mem);
/* estimateChannel, modem.js:421-440; ce = [n][SYMBOL_LEN]; known_re/known_im = [512] shared by all rows
 * (NULL = the config's generateChannelEstSymbol values); ch_re/ch_im = [n][512] */
int amb_estimate_channel(amb_modem* m, const float* ce, int n, const double* known_re, const double* known_im,
                         double* ch_re, double* ch_im, int mem);
/* demodulateOFDM, modem.js:365-418; signal rows start at the first data symbol; ch_re/ch_im = [n][512];
 * bits = [n][bits_stride], one byte per bit like the reference's number[]; nbits_out[i] = bits produced;
 * eq_out (optional) = [n][eq_stride][2] de-rotated (cr, ci) per data subcarrier, symbol-major */
int amb_demodulate(amb_modem* m, const float* signal, int64_t stride, const int64_t* lens, int n, const double* ch_re,
                   const double* ch_im, uint8_t* bits, int64_t bits_stride, int64_t* nbits_out, double* eq_out,
                   int64_t eq_stride, int mem);
/* decodeChunkFrame, modem.js:770-803; rows start at preamble-1 */
int amb_decode_chunk_frames(amb_modem* m, const float* frames, int64_t stride, const int64_t* lens, int n,
                            amb_frame_result* results, uint8_t* bytes_out, int64_t bytes_stride, int mem);
/* decodeReceivedSignal, modem.js:557-654; one recording per row */
int amb_decode_received_signals(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n,
                                amb_frame_result* results, uint8_t* bytes_out, int64_t bytes_stride, int mem);

/* analyzeLoopback, modem.js:978-1081 (the reference's only fixture check): preprocess, detectPreamble with the
 * detectPreambleCrossCorr fallback, refine, channel estimate, demodulate, compare against the fixture bytes.
 * One recording per row.  test_data = the bytes sent (generateTestSignal's 00..0F, modem.js:914-916), host memory.
 * results (host memory) mirrors the returned object; channel_magnitude (host memory, optional) = [n][mag_stride]
 * doubles, SUB_END - SUB_START + 1 per detected recording whose CE symbol fits (n_mag), zeros otherwise. */
typedef struct amb_loopback_result {
    int32_t detected;
    int32_t quality;      /* 0 'poor', 1 'good', 2 'excellent' (modem.js:1075-1081) */
    int64_t start_idx;    /* refined preamble start; -1 when not detected */
    int32_t n_mag;        /* channelMagnitude.length */
    int32_t status;       /* AMB_ST_* of the refinement: 0, PREAMBLE_NOT_DETECTED or SIGNAL_TOO_SHORT_FOR_CE */
    double correlation;   /* max(0, refined cross-correlation peak) */
    double ber;           /* bit error rate over the fixture bytes; 1 when they cannot be located */
    double snr_estimate;  /* 20 log10(pilot_mag); -Infinity when no pilot qualifies; 0 on the early returns */
    double pilot_mag;     /* mean |H| over the pilots above 1e-6 (modem.js:1036-1047) */
} amb_loopback_result;
int amb_analyze_loopback(amb_modem* m, const float* x, int64_t stride, const int64_t* lens, int n, const uint8_t* test_data,
                         int test_len, amb_loopback_result* results, double* channel_magnitude, int64_t mag_stride, int mem);

/* ---- streaming receiver: new StreamingReceiver(modName, repetition) / processAudioBlock,
 *      app.js:706-981, for n_streams independent streams at once ------------------------- */
/* block_len = the ScriptProcessor block size whose call cadence is emulated (app.js:1107: 4096);
 * max_push = largest n_new a push will carry (multiple of block_len) */
int amb_stream_open(amb_modem* m, int n_streams, int block_len, int64_t max_push);
int amb_stream_close(amb_modem* m);
/* where the next push's samples go on the device (stream s at dev_ptr + s*stride floats); for AMB_MEM_INPLACE */
int amb_stream_input_window(amb_modem* m, float** dev_ptr, int64_t* stride);
/* Feed n_new samples (multiple of block_len) to every stream: x = [n_streams][x_stride].  Frames completed by
 * these samples come back grouped by stream and in time order within a stream (results[j].stream names it); a large host
 * push is processed in sections, each of which delivers its own group.  bytes_out row j (bytes_stride bytes) belongs to
 * results[j]: it receives the first min(n_bytes, bytes_stride) decoded bytes.  Parsing, the CRC verdict and n_bytes always
 * cover the whole decoded frame, whatever the row length: a row shorter than n_bytes only truncates what the caller reads
 * (a stream whose metadata frame announces a larger chunkSize than the rows were sized for still reports correct results). */
int amb_stream_push(amb_modem* m, const float* x, int64_t x_stride, int64_t n_new, int mem, amb_frame_result* results,
                    int max_results, int* n_results, uint8_t* bytes_out, int64_t bytes_stride, int results_mem);
/* Wire-format ingest: the same push with signed 16-bit PCM samples; sample value = s / 32768 (exact in
 * float32), i.e. the results are those of amb_stream_push on the float32 array an audio front end would hand the
 * reference for a 16-bit capture.  Half the bytes per sample over PCIe and out of HBM.  mem = AMB_MEM_HOST or
 * AMB_MEM_DEVICE. */
int amb_stream_push_pcm16(amb_modem* m, const int16_t* x, int64_t x_stride, int64_t n_new, int mem, amb_frame_result* results,
                          int max_results, int* n_results, uint8_t* bytes_out, int64_t bytes_stride, int results_mem);
/* per-stream counters (stream = -1: sum over streams): framesDecoded, frameErrors (app.js:738-739),
 * falsePositives (app.js:879-884), receivedCount, crcErrors (app.js:604-605, ChunkAssembler's bitmap semantics of
 * app.js:628-643 without the chunk store), duplicates (chunks whose bit was already set), metaReceived, samplesWritten */
int amb_stream_counters(amb_modem* m, int stream, int64_t* out8);

/* ChunkAssembler without IndexedDB (app.js:597-704): keep the received chunks of every stream in a device buffer of
 * max_file_bytes per stream - a CRC-valid metadata frame clears it (app.js:619-625), a newly received chunk is written at
 * seqNum * chunkSize (app.js:676-677) - and hand the file out like assembleFile (app.js:666-682): size = totalFileSize of
 * the metadata frame in force (0 before one arrived; clipped to max_file_bytes), missing chunks read as zeros. */
int amb_stream_enable_assembly(amb_modem* m, int64_t max_file_bytes);
int amb_stream_assemble(amb_modem* m, int stream, uint8_t* out, int64_t cap, int64_t* size_out, int mem);

/* ---- transmit side (the "next" row of the scope table; also synthesises bench / test input on the device) ---- */
/* modulateOFDM(bits, modName), modem.js:322-362, for n rows of bits: bits = [n][bits_stride], one byte per bit like the
 * reference's number[] (only `b & 1` counts, modem.js:135), n_bits valid per row.  samples = [n][samples_stride] float32:
 * numSymbols symbols of SYMBOL_LEN samples each (cyclic prefix + body: the Float32Arrays of addCP, modem.js:202-208),
 * bit-identical to the reference's.  out3 = {numSymbols, bitsPerSymbol, padded length}.  The reference pads the CALLER's
 * array in place with zeros up to a whole symbol (modem.js:329); this call does the same when the row has room
 * (bits_stride >= padded length) and otherwise only reports the padded length.  No repetition is applied (the reference
 * calls repeatBits before modulateOFDM, modem.js:721).  A first call with samples = NULL and n = 0 returns the sizes. */
int amb_modulate(amb_modem* m, uint8_t* bits, int64_t bits_stride, int64_t n_bits, int n, float* samples, int64_t samples_stride,
                 int32_t* out3, int mem);
/* samples of one chunk frame incl. its silences: estimateFrameSamplesWithSilence, modem.js:876-884 */
int64_t amb_chunk_frame_samples(amb_modem* m, int64_t payload_bytes, int is_first_frame);
/* buildDataChunkFrame, modem.js:763-766 (+ buildDataChunkPayload :694-714, buildChunkOFDMFrame :718-756) for n chunks of
 * data_len bytes with sequence numbers seq0.. ; frame i is written at out + i*out_stride, bit-identical to the reference's
 * Float32Array.  data = [n][data_stride] chunk bytes, or NULL for synthetic bytes (SplitMix64 of synth_seed, seq, offset).
 * noise_sigma > 0 adds white Gaussian noise (counter-based, indexed from noise_g0) after the frame normalisation.
 * data_copy (optional, [n][data_len]) receives the chunk bytes actually sent. */
int amb_build_data_chunk_frames(amb_modem* m, const uint8_t* data, int64_t data_stride, int data_len, int n, int32_t seq0,
                                uint64_t synth_seed, float* out, int64_t out_stride, double noise_sigma, uint64_t noise_seed,
                                int64_t noise_g0, uint8_t* data_copy, int mem);
/* buildMetadataFrame, modem.js:758-761 (+ buildMetadataPayload :666-692) */
int amb_build_metadata_frame(amb_modem* m, int32_t total_chunks, int32_t total_file_size, int32_t chunk_size, const uint8_t* name,
                             int name_len, float* out, double noise_sigma, uint64_t noise_seed, int64_t noise_g0, int mem);

/* buildTransmitSignal, modem.js:498-555 (and generateTestSignal, :913-975, which is the same construction over the bytes
 * 00..0F named "test"): n legacy frames [silence | preamble-1 | preamble-2 | CE | data symbols | silence], each carrying
 * the packet [nameLen][name][dataLen:4][data][CRC-32:4] of data row i, normalised to a peak of 0.8; bit-identical to the
 * reference's Float32Array.  data = [n][data_stride] and out = [n][out_stride] are host memory (mem = AMB_MEM_HOST);
 * amb_transmit_signal_samples gives the length of one signal. */
int64_t amb_transmit_signal_samples(amb_modem* m, int64_t data_len, int name_len);
int amb_build_transmit_signals(amb_modem* m, const uint8_t* data, int64_t data_stride, int data_len, int n, const uint8_t* name,
                               int name_len, float* out, int64_t out_stride, int mem);

#ifdef __cplusplus
}
#endif
#endif
