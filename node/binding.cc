// N-API addon over the C-ABI library (include/audio_modem_b200.h): every entry point of SURVEY.md section 8(b).
//
// NOT RUN IN THIS REPOSITORY: the build image has no Node.js.  It is compiled and type-checked against
// a declaration-only node_api.h (tests/stubs/node_api.h; tests/test_abi.py::test_napi_addon_compiles) and linked against the
// library; the tested host side is the Python mirror (audio-modem_b200/modem.py), which makes exactly the same C calls.
//
// Build (on a machine with Node >= 16 and the CUDA runtime), from this directory:
//   g++ -O2 -std=c++17 -fPIC -shared -I"$(node -p "require('node:path').resolve(process.execPath, '..', '..', 'include', 'node')")"
//       -I../include binding.cc -L../audio-modem_b200 -laudiomodem_b200 -Wl,-rpath,'$ORIGIN/../audio-modem_b200' -o audio_modem_b200.node
#include <node_api.h>

#include <cstring>
#include <string>
#include <vector>

#include "audio_modem_b200.h"

#define NAPI_CALL(env, call)                                              \
    do {                                                                  \
        if ((call) != napi_ok) {                                          \
            napi_throw_error((env), nullptr, "N-API call failed: " #call); \
            return nullptr;                                               \
        }                                                                 \
    } while (0)

namespace {

void finalize_modem(napi_env, void* data, void*) { amb_destroy(static_cast<amb_modem*>(data)); }

bool get_modem(napi_env env, napi_value v, amb_modem** out) {
    void* p = nullptr;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_type_error(env, nullptr, "modem handle expected"); return false; }
    *out = static_cast<amb_modem*>(p);
    return true;
}

bool get_f32(napi_env env, napi_value v, const float** data, size_t* n) {
    napi_typedarray_type t; napi_value ab; size_t off;
    void* p = nullptr;
    if (napi_get_typedarray_info(env, v, &t, n, &p, &ab, &off) != napi_ok || t != napi_float32_array) {
        napi_throw_type_error(env, nullptr, "Float32Array expected");
        return false;
    }
    *data = static_cast<const float*>(p);
    return true;
}

napi_value set_num(napi_env env, napi_value obj, const char* k, double v) {
    napi_value n; napi_create_double(env, v, &n); napi_set_named_property(env, obj, k, n); return obj;
}

// amb_frame_result + decoded bytes -> the reference's result object (modem.js:644-653, :822-827, :843-848)
napi_value result_object(napi_env env, const amb_frame_result& r, const uint8_t* bytes) {
    napi_value o; napi_create_object(env, &o);
    if (r.status != 0) {
        std::string msg = amb_status_string(r.status);
        char buf[32];
        size_t p;
        if ((p = msg.find("{arg:x}")) != std::string::npos) { snprintf(buf, sizeof buf, "%x", (unsigned)r.err_arg); msg.replace(p, 7, buf); }
        if ((p = msg.find("{arg}")) != std::string::npos) { snprintf(buf, sizeof buf, "%d", r.err_arg); msg.replace(p, 5, buf); }
        napi_value s; napi_create_string_utf8(env, msg.c_str(), NAPI_AUTO_LENGTH, &s); napi_set_named_property(env, o, "error", s);
        if (r.status == AMB_ST_UNKNOWN_FRAME_TYPE) set_num(env, o, "frameType", r.err_arg);
        return o;
    }
    napi_value b; napi_get_boolean(env, r.crc_valid != 0, &b); napi_set_named_property(env, o, "crcValid", b);
    set_num(env, o, "expectedCRC", r.expected_crc); set_num(env, o, "actualCRC", r.actual_crc);
    auto put_bytes = [&](const char* key, int off, int len) {
        void* dst; napi_value ab, ta;
        napi_create_arraybuffer(env, (size_t)len, &dst, &ab);
        std::memcpy(dst, bytes + off, (size_t)len);
        napi_create_typedarray(env, napi_uint8_array, (size_t)len, ab, 0, &ta);
        napi_set_named_property(env, o, key, ta);
    };
    auto put_name = [&]() {
        napi_value s; napi_create_string_utf8(env, reinterpret_cast<const char*>(bytes) + r.name_off, (size_t)r.name_len, &s);
        napi_set_named_property(env, o, "fileName", s);
    };
    if (r.frame_type == AMB_FRAME_LEGACY) {
        napi_value s; napi_create_string_utf8(env, "legacy", NAPI_AUTO_LENGTH, &s); napi_set_named_property(env, o, "frameType", s);
        set_num(env, o, "dataLen", r.data_len); put_bytes("data", r.data_off, r.data_len); put_name();
    } else if (r.frame_type == AMB_FRAME_META) {
        set_num(env, o, "frameType", AMB_FRAME_META); set_num(env, o, "totalChunks", r.total_chunks);
        set_num(env, o, "totalFileSize", r.total_file_size); set_num(env, o, "chunkSize", r.chunk_size); put_name();
    } else {
        set_num(env, o, "frameType", AMB_FRAME_DATA); set_num(env, o, "seqNum", r.seq_num); set_num(env, o, "dataLen", r.data_len);
        put_bytes("data", r.data_off, r.data_len);
    }
    if (r.preamble_idx >= 0) set_num(env, o, "preambleIdx", (double)r.preamble_idx);
    return o;
}

// create(config, constellation, repetition[, device]) -> external handle
napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    int32_t cfg = 0, mod = 1, rep = 1, dev = 0;
    napi_get_value_int32(env, a[0], &cfg); napi_get_value_int32(env, a[1], &mod); napi_get_value_int32(env, a[2], &rep);
    if (argc > 3) napi_get_value_int32(env, a[3], &dev);
    amb_modem* m = nullptr;
    const int rc = amb_create(&m, dev, cfg, mod, rep);
    if (rc != AMB_OK) { napi_throw_error(env, nullptr, rc == AMB_E_NODEVICE ? "no CUDA device (the library has no CPU path)" : "amb_create failed"); return nullptr; }
    napi_value ext; NAPI_CALL(env, napi_create_external(env, m, finalize_modem, nullptr, &ext));
    return ext;
}

// detectPreamble(handle, Float32Array) -> number   (modem.js:286)
napi_value DetectPreamble(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    int64_t idx = -1;
    if (amb_detect_preamble(m, x, (int64_t)n, nullptr, 1, &idx, nullptr, AMB_MEM_HOST) != AMB_OK) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }
    napi_value out; napi_create_int64(env, idx, &out);
    return out;
}

template <int (*FN)(amb_modem*, const float*, int64_t, const int64_t*, int, amb_frame_result*, uint8_t*, int64_t, int)>
napi_value DecodeOne(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    amb_frame_result r;
    std::vector<uint8_t> bytes(n / 4 + 64);          // bytes.length <= bits/8 <= samples*4/8
    if (FN(m, x, (int64_t)n, nullptr, 1, &r, bytes.data(), (int64_t)bytes.size(), AMB_MEM_HOST) != AMB_OK) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }
    return result_object(env, r, bytes.data());
}

// streamPushPCM16(handle, Int16Array[nStreams*nNew], nNew, bytesStride) -> result[]  (amb_stream_push_pcm16)
napi_value StreamPushPCM16(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; if (!get_modem(env, a[0], &m)) return nullptr;
    napi_typedarray_type tt; size_t n = 0; void* x = nullptr;
    NAPI_CALL(env, napi_get_typedarray_info(env, a[1], &tt, &n, &x, nullptr, nullptr));
    if (tt != napi_int16_array) { napi_throw_type_error(env, nullptr, "Int16Array expected"); return nullptr; }
    int64_t n_new = 0, stride = 0; napi_get_value_int64(env, a[2], &n_new); napi_get_value_int64(env, a[3], &stride);
    const int max_res = 4096;
    std::vector<amb_frame_result> res(max_res);
    std::vector<uint8_t> bytes((size_t)max_res * (size_t)stride);
    int got = 0;
    const int rc = amb_stream_push_pcm16(m, (const int16_t*)x, n_new, n_new, AMB_MEM_HOST, res.data(), max_res, &got, bytes.data(), stride, AMB_MEM_HOST);
    if (rc != AMB_OK && rc != AMB_E_OVERFLOW) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }
    napi_value arr; napi_create_array_with_length(env, (size_t)(got < max_res ? got : max_res), &arr);
    for (int i = 0; i < got && i < max_res; i++) {
        napi_value o = result_object(env, res[i], bytes.data() + (size_t)i * (size_t)stride);
        set_num(env, o, "stream", res[i].stream);
        napi_set_element(env, arr, (uint32_t)i, o);
    }
    return arr;
}

// buildTransmitSignal(handle, Uint8Array fileData, string fileName) -> {signal: Float32Array, dataLen}  (modem.js:498-555)
napi_value BuildTransmitSignal(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; if (!get_modem(env, a[0], &m)) return nullptr;
    napi_typedarray_type tt; size_t n = 0; void* d = nullptr;
    NAPI_CALL(env, napi_get_typedarray_info(env, a[1], &tt, &n, &d, nullptr, nullptr));
    char name[768]; size_t name_len = 0;
    NAPI_CALL(env, napi_get_value_string_utf8(env, a[2], name, sizeof(name), &name_len));
    const int64_t len = amb_transmit_signal_samples(m, (int64_t)n, (int)name_len);
    void* out = nullptr; napi_value ab, sig, o;
    NAPI_CALL(env, napi_create_arraybuffer(env, (size_t)len * 4, &out, &ab));
    if (amb_build_transmit_signals(m, (const uint8_t*)d, (int64_t)n, (int)n, 1, (const uint8_t*)name, (int)name_len, (float*)out, len, AMB_MEM_HOST) != AMB_OK) {
        napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr;
    }
    NAPI_CALL(env, napi_create_typedarray(env, napi_float32_array, (size_t)len, ab, 0, &sig));
    napi_create_object(env, &o);
    napi_set_named_property(env, o, "signal", sig);
    set_num(env, o, "dataLen", (double)n);
    return o;
}

// analyzeLoopback(handle, Float32Array recorded, Uint8Array testData) -> {detected, correlation, ber, channelMagnitude, snrEstimate, quality}
napi_value AnalyzeLoopback(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    napi_typedarray_type tt; size_t tn = 0; void* td = nullptr;
    NAPI_CALL(env, napi_get_typedarray_info(env, a[2], &tt, &tn, &td, nullptr, nullptr));
    amb_loopback_result r;
    double mags[256];
    if (amb_analyze_loopback(m, x, (int64_t)n, nullptr, 1, (const uint8_t*)td, (int)tn, &r, mags, 256, AMB_MEM_HOST) != AMB_OK) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }
    napi_value o, b, arr, q;
    napi_create_object(env, &o);
    napi_get_boolean(env, r.detected != 0, &b); napi_set_named_property(env, o, "detected", b);
    set_num(env, o, "correlation", r.correlation); set_num(env, o, "ber", r.ber); set_num(env, o, "snrEstimate", r.snr_estimate);
    napi_create_array_with_length(env, (size_t)r.n_mag, &arr);
    for (int k = 0; k < r.n_mag; k++) { napi_value v; napi_create_double(env, mags[k], &v); napi_set_element(env, arr, (uint32_t)k, v); }
    napi_set_named_property(env, o, "channelMagnitude", arr);
    static const char* names[3] = {"poor", "good", "excellent"};
    napi_create_string_utf8(env, names[r.quality], NAPI_AUTO_LENGTH, &q); napi_set_named_property(env, o, "quality", q);
    return o;
}

// streamOpen(handle, nStreams, blockLen, maxPush); streamPush(handle, Float32Array[nStreams*nNew], nNew) -> result[]
napi_value StreamOpen(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; if (!get_modem(env, a[0], &m)) return nullptr;
    int32_t ns = 1, bl = 4096; int64_t mp = 4096;
    napi_get_value_int32(env, a[1], &ns); napi_get_value_int32(env, a[2], &bl); napi_get_value_int64(env, a[3], &mp);
    if (amb_stream_open(m, ns, bl, mp) != AMB_OK) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }
    return nullptr;
}

napi_value StreamPush(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    int64_t n_new = 0, stride = 0; napi_get_value_int64(env, a[2], &n_new); napi_get_value_int64(env, a[3], &stride);
    const int max_res = 4096;
    std::vector<amb_frame_result> res(max_res);
    std::vector<uint8_t> bytes((size_t)max_res * (size_t)stride);
    int got = 0;
    const int rc = amb_stream_push(m, x, n_new, n_new, AMB_MEM_HOST, res.data(), max_res, &got, bytes.data(), stride, AMB_MEM_HOST);
    if (rc != AMB_OK && rc != AMB_E_OVERFLOW) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }
    napi_value arr; napi_create_array_with_length(env, (size_t)(got < max_res ? got : max_res), &arr);
    for (int i = 0; i < got && i < max_res; i++) {
        napi_value o = result_object(env, res[i], bytes.data() + (size_t)i * (size_t)stride);
        set_num(env, o, "stream", res[i].stream);
        napi_set_element(env, arr, (uint32_t)i, o);
    }
    return arr;
}

// ---- typed-array helpers -------------------------------------------------------------------------------------------
bool get_typed(napi_env env, napi_value v, napi_typedarray_type want, const char* what, void** data, size_t* n) {
    napi_typedarray_type t; napi_value ab; size_t off;
    if (napi_get_typedarray_info(env, v, &t, n, data, &ab, &off) != napi_ok || t != want) { napi_throw_type_error(env, nullptr, what); return false; }
    return true;
}
template <class T>
napi_value make_typed(napi_env env, napi_typedarray_type type, size_t n, T** data) {
    void* p = nullptr; napi_value ab, ta;
    if (napi_create_arraybuffer(env, n * sizeof(T), &p, &ab) != napi_ok || napi_create_typedarray(env, type, n, ab, 0, &ta) != napi_ok) {
        napi_throw_error(env, nullptr, "allocation failed"); return nullptr;
    }
    *data = static_cast<T*>(p);
    return ta;
}
napi_value fail_lib(napi_env env, amb_modem* m) { napi_throw_error(env, nullptr, amb_last_error(m)); return nullptr; }

// detectPreambleCrossCorr(handle, Float32Array) -> number   (modem.js:235)
napi_value DetectPreambleCrossCorr(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    int64_t idx = -1;
    if (amb_detect_preamble_xcorr(m, x, (int64_t)n, nullptr, 1, &idx, nullptr, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    napi_value out; napi_create_int64(env, idx, &out);
    return out;
}

// preprocessSignal(handle, Float32Array) -> new Float32Array   (modem.js:213)
napi_value PreprocessSignal(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    float* out = nullptr;
    napi_value ta = make_typed<float>(env, napi_float32_array, n, &out);
    if (!ta) return nullptr;
    if (n && amb_preprocess_signal(m, x, (int64_t)n, nullptr, 1, out, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    return ta;
}

// estimateChannel(handle, Float32Array rx, Float64Array knownRe, Float64Array knownIm) -> [Float64Array(512), Float64Array(512)]   (modem.js:421)
napi_value EstimateChannel(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n, nr, ni; void *kr, *ki;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    if (!get_typed(env, a[2], napi_float64_array, "Float64Array knownRe expected", &kr, &nr) ||
        !get_typed(env, a[3], napi_float64_array, "Float64Array knownIm expected", &ki, &ni)) return nullptr;
    int32_t cfg[8];
    amb_config_info(0, cfg, nullptr);
    if (nr < (size_t)cfg[0] || ni < (size_t)cfg[0]) { napi_throw_type_error(env, nullptr, "known symbol shorter than FFT_SIZE"); return nullptr; }
    // the CE symbol row: SYMBOL_LEN samples, zero-padded like the reference's `|| 0` reads (modem.js:425)
    std::vector<float> row(1024, 0.f);
    std::memcpy(row.data(), x, (n < row.size() ? n : row.size()) * sizeof(float));
    double *hr = nullptr, *hi = nullptr;
    napi_value tr = make_typed<double>(env, napi_float64_array, (size_t)cfg[0], &hr), ti = make_typed<double>(env, napi_float64_array, (size_t)cfg[0], &hi);
    if (!tr || !ti) return nullptr;
    if (amb_estimate_channel(m, row.data(), 1, static_cast<const double*>(kr), static_cast<const double*>(ki), hr, hi, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    napi_value arr; napi_create_array_with_length(env, 2, &arr);
    napi_set_element(env, arr, 0, tr); napi_set_element(env, arr, 1, ti);
    return arr;
}

// demodulateOFDM(handle, Float32Array signal, Float64Array chRe, Float64Array chIm) -> Uint8Array of 0/1   (modem.js:365; the wrapper
// turns it into the reference's number[])
napi_value DemodulateOFDM(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; const float* x; size_t n, nr, ni; void *cr, *ci;
    if (!get_modem(env, a[0], &m) || !get_f32(env, a[1], &x, &n)) return nullptr;
    if (!get_typed(env, a[2], napi_float64_array, "Float64Array channelRe expected", &cr, &nr) ||
        !get_typed(env, a[3], napi_float64_array, "Float64Array channelIm expected", &ci, &ni)) return nullptr;
    if (nr < 512 || ni < 512) { napi_throw_type_error(env, nullptr, "channel shorter than FFT_SIZE"); return nullptr; }
    std::vector<uint8_t> bits(n + 8);                  // at most 4 bits per data subcarrier, fewer subcarriers than samples
    int64_t nbits = 0;
    if (n && amb_demodulate(m, x, (int64_t)n, nullptr, 1, static_cast<const double*>(cr), static_cast<const double*>(ci), bits.data(), (int64_t)bits.size(), &nbits,
                            nullptr, 0, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    uint8_t* out = nullptr;
    napi_value ta = make_typed<uint8_t>(env, napi_uint8_array, (size_t)nbits, &out);
    if (!ta) return nullptr;
    std::memcpy(out, bits.data(), (size_t)nbits);
    return ta;
}

// modulateOFDM(handle, Uint8Array bits) -> {samples: Float32Array(numSymbols * SYMBOL_LEN), numSymbols, bitsPerSymbol, paddedLength}
// (modem.js:322-362; the wrapper cuts `samples` into per-symbol arrays and pads the caller's array in place like modem.js:329)
napi_value ModulateOFDM(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; void* b; size_t nb;
    if (!get_modem(env, a[0], &m) || !get_typed(env, a[1], napi_uint8_array, "Uint8Array bits expected", &b, &nb)) return nullptr;
    int32_t out3[3] = {0, 0, 0};
    if (amb_modulate(m, nullptr, (int64_t)nb, (int64_t)nb, 0, nullptr, 0, out3, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    int32_t who[4], cfg[8];
    amb_modem_info(m, who);
    amb_config_info(who[0], cfg, nullptr);
    const int64_t slen = cfg[2];                                          // SYMBOL_LEN of the handle's config
    std::vector<uint8_t> bits((size_t)out3[2] + 1, 0);
    std::memcpy(bits.data(), b, nb);
    float* samples = nullptr;
    napi_value ta = make_typed<float>(env, napi_float32_array, (size_t)out3[0] * (size_t)slen, &samples);
    if (!ta) return nullptr;
    if (out3[0] > 0 && amb_modulate(m, bits.data(), (int64_t)bits.size(), (int64_t)nb, 1, samples, (int64_t)out3[0] * slen, out3, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    napi_value o; napi_create_object(env, &o);
    napi_set_named_property(env, o, "samples", ta);
    set_num(env, o, "numSymbols", out3[0]); set_num(env, o, "bitsPerSymbol", out3[1]); set_num(env, o, "paddedLength", out3[2]);
    return o;
}

// crc32(handle, Uint8Array) -> number   (modem.js:453)
napi_value Crc32(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; void* d; size_t n;
    if (!get_modem(env, a[0], &m) || !get_typed(env, a[1], napi_uint8_array, "Uint8Array expected", &d, &n)) return nullptr;
    uint32_t crc = 0;
    const int64_t len = (int64_t)n;
    const uint8_t dummy = 0;
    if (amb_crc32(m, n ? static_cast<const uint8_t*>(d) : &dummy, (int64_t)(n ? n : 1), &len, 1, &crc, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    napi_value out; napi_create_uint32(env, crc, &out);
    return out;
}

// estimateFrameSamples(config, constellation, payloadBytes, repetition) -> number   (modem.js:863; no handle needed)
napi_value EstimateFrameSamples(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value a[4];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    int32_t cfg = 0, mod = 1, rep = 1; int64_t payload = 0;
    napi_get_value_int32(env, a[0], &cfg); napi_get_value_int32(env, a[1], &mod); napi_get_value_int64(env, a[2], &payload); napi_get_value_int32(env, a[3], &rep);
    napi_value out; napi_create_int64(env, amb_estimate_frame_samples(cfg, mod, payload, rep), &out);
    return out;
}

// buildMetadataFrame(handle, totalChunks, totalFileSize, chunkSize, fileName) -> Float32Array   (modem.js:758)
napi_value BuildMetadataFrame(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value a[5];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; if (!get_modem(env, a[0], &m)) return nullptr;
    int32_t tc = 0, fs = 0, cs = 0;
    napi_get_value_int32(env, a[1], &tc); napi_get_value_int32(env, a[2], &fs); napi_get_value_int32(env, a[3], &cs);
    char name[768]; size_t name_len = 0;
    NAPI_CALL(env, napi_get_value_string_utf8(env, a[4], name, sizeof(name), &name_len));
    if (name_len > 255) name_len = 255;                                   // modem.js:669
    const int64_t len = amb_chunk_frame_samples(m, 1 + 4 + 4 + 2 + 1 + (int64_t)name_len + 4, 1);
    float* out = nullptr;
    napi_value ta = make_typed<float>(env, napi_float32_array, (size_t)len, &out);
    if (!ta) return nullptr;
    if (amb_build_metadata_frame(m, tc, fs, cs, reinterpret_cast<const uint8_t*>(name), (int)name_len, out, 0.0, 0, 0, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    return ta;
}

// buildDataChunkFrame(handle, Uint8Array chunkData, seqNum) -> Float32Array   (modem.js:763)
napi_value BuildDataChunkFrame(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; void* d; size_t n;
    if (!get_modem(env, a[0], &m) || !get_typed(env, a[1], napi_uint8_array, "Uint8Array chunkData expected", &d, &n)) return nullptr;
    int32_t seq = 0; napi_get_value_int32(env, a[2], &seq);
    const int64_t len = amb_chunk_frame_samples(m, (int64_t)n + 11, 0);
    float* out = nullptr;
    napi_value ta = make_typed<float>(env, napi_float32_array, (size_t)len, &out);
    if (!ta) return nullptr;
    if (amb_build_data_chunk_frames(m, static_cast<const uint8_t*>(d), (int64_t)n, (int)n, 1, seq, 0, out, len, 0.0, 0, 0, nullptr, AMB_MEM_HOST) != AMB_OK) return fail_lib(env, m);
    return ta;
}

// generatePreambleSymbol1(config) -> Float32Array; generateChannelEstSymbol(config) -> {samples, knownRe, knownIm}   (modem.js:158, :186)
napi_value GeneratePreambleSymbol1(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    int32_t cfg = 0; napi_get_value_int32(env, a[0], &cfg);
    int32_t ci[8];
    if (amb_config_info(cfg, ci, nullptr) != AMB_OK) { napi_throw_error(env, nullptr, "bad config"); return nullptr; }
    float* out = nullptr;
    napi_value ta = make_typed<float>(env, napi_float32_array, (size_t)ci[2], &out);
    if (!ta) return nullptr;
    amb_preamble_symbol1(cfg, out);
    return ta;
}
napi_value GenerateChannelEstSymbol(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    int32_t cfg = 0; napi_get_value_int32(env, a[0], &cfg);
    int32_t ci[8];
    if (amb_config_info(cfg, ci, nullptr) != AMB_OK) { napi_throw_error(env, nullptr, "bad config"); return nullptr; }
    float* s = nullptr; double *kr = nullptr, *ki = nullptr;
    napi_value ts = make_typed<float>(env, napi_float32_array, (size_t)ci[2], &s), tr = make_typed<double>(env, napi_float64_array, (size_t)ci[0], &kr),
               ti = make_typed<double>(env, napi_float64_array, (size_t)ci[0], &ki);
    if (!ts || !tr || !ti) return nullptr;
    amb_channel_est_symbol(cfg, s, kr);
    std::memset(ki, 0, (size_t)ci[0] * sizeof(double));
    napi_value o; napi_create_object(env, &o);
    napi_set_named_property(env, o, "samples", ts); napi_set_named_property(env, o, "knownRe", tr); napi_set_named_property(env, o, "knownIm", ti);
    return o;
}

// streamCounters(handle[, stream]) -> {framesDecoded, frameErrors, falsePositives, receivedCount, crcErrors, duplicates, metaReceived, samplesWritten}
napi_value StreamCounters(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; if (!get_modem(env, a[0], &m)) return nullptr;
    int32_t stream = -1;
    if (argc > 1) napi_get_value_int32(env, a[1], &stream);
    int64_t c[8];
    if (amb_stream_counters(m, stream, c) != AMB_OK) return fail_lib(env, m);
    static const char* keys[8] = {"framesDecoded", "frameErrors", "falsePositives", "receivedCount", "crcErrors", "duplicates", "metaReceived", "samplesWritten"};
    napi_value o; napi_create_object(env, &o);
    for (int k = 0; k < 8; k++) set_num(env, o, keys[k], (double)c[k]);
    return o;
}

napi_value StreamClose(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1];
    NAPI_CALL(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    amb_modem* m; if (!get_modem(env, a[0], &m)) return nullptr;
    if (amb_stream_close(m) != AMB_OK) return fail_lib(env, m);
    return nullptr;
}

napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor d[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"detectPreamble", nullptr, DetectPreamble, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"decodeReceivedSignal", nullptr, DecodeOne<amb_decode_received_signals>, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"decodeChunkFrame", nullptr, DecodeOne<amb_decode_chunk_frames>, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"analyzeLoopback", nullptr, AnalyzeLoopback, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"streamPushPCM16", nullptr, StreamPushPCM16, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"buildTransmitSignal", nullptr, BuildTransmitSignal, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"streamOpen", nullptr, StreamOpen, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"streamPush", nullptr, StreamPush, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"streamCounters", nullptr, StreamCounters, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"streamClose", nullptr, StreamClose, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"detectPreambleCrossCorr", nullptr, DetectPreambleCrossCorr, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"preprocessSignal", nullptr, PreprocessSignal, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"estimateChannel", nullptr, EstimateChannel, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"demodulateOFDM", nullptr, DemodulateOFDM, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"modulateOFDM", nullptr, ModulateOFDM, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"crc32", nullptr, Crc32, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"estimateFrameSamples", nullptr, EstimateFrameSamples, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"buildMetadataFrame", nullptr, BuildMetadataFrame, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"buildDataChunkFrame", nullptr, BuildDataChunkFrame, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"generatePreambleSymbol1", nullptr, GeneratePreambleSymbol1, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"generateChannelEstSymbol", nullptr, GenerateChannelEstSymbol, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(d) / sizeof(d[0]), d);
    return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
