// N-API addon over the C-ABI library (include/audio_modem_b200.h).
//
// NOT COMPILED OR TESTED IN THIS REPOSITORY: the build image has no Node.js and no node_api.h.
// It is the reference-side binding a maintainer adds; the tested host side is the Python mirror
// (audio-modem_b200/modem.py), which makes exactly the same C calls.
//
// Build (on a machine with Node >= 16 and the CUDA runtime):
//   g++ -O2 -fPIC -shared -I$(node -p "require('node:path').dirname(process.execPath)+'/../include/node'") \
//       -I../include binding.cc -L../audio-modem_b200 -laudiomodem_b200 -o audio_modem_b200.node
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
    };
    napi_define_properties(env, exports, sizeof(d) / sizeof(d[0]), d);
    return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
