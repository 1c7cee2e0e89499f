// Host shims for running the reference under Qt's QJSEngine (ECMAScript 2016 level, no host objects):
// console / TextEncoder / TextDecoder as the browser provides them, SHA-256 and hex helpers for the dump.
var console = { log: function () {}, warn: function () {}, error: function () {}, info: function () {} };
function TextEncoder() {}
TextEncoder.prototype.encode = function (s) {
    var out = [];
    for (var i = 0; i < s.length; i++) {
        var c = s.codePointAt(i);
        if (c > 0xFFFF) i++;
        if (c < 0x80) out.push(c);
        else if (c < 0x800) out.push(0xC0 | (c >> 6), 0x80 | (c & 63));
        else if (c < 0x10000) out.push(0xE0 | (c >> 12), 0x80 | ((c >> 6) & 63), 0x80 | (c & 63));
        else out.push(0xF0 | (c >> 18), 0x80 | ((c >> 12) & 63), 0x80 | ((c >> 6) & 63), 0x80 | (c & 63));
    }
    return Uint8Array.from(out);
};
function TextDecoder() {}
TextDecoder.prototype.decode = function (u8) {       // non-fatal UTF-8 decoding (malformed sequences -> U+FFFD)
    var s = '', i = 0, n = u8.length;
    while (i < n) {
        var b = u8[i++], need = 0, cp = 0, min = 0;
        if (b < 0x80) { s += String.fromCharCode(b); continue; }
        else if (b >= 0xC2 && b < 0xE0) { need = 1; cp = b & 0x1F; min = 0x80; }
        else if (b >= 0xE0 && b < 0xF0) { need = 2; cp = b & 0x0F; min = 0x800; }
        else if (b >= 0xF0 && b < 0xF5) { need = 3; cp = b & 0x07; min = 0x10000; }
        else { s += '�'; continue; }
        var ok = true, j = i;
        for (var k = 0; k < need; k++, j++) {
            if (j >= n || (u8[j] & 0xC0) !== 0x80) { ok = false; break; }
            cp = (cp << 6) | (u8[j] & 63);
        }
        if (!ok || cp < min || cp > 0x10FFFF || (cp >= 0xD800 && cp < 0xE000)) { s += '�'; continue; }
        i = j;
        s += String.fromCodePoint(cp);
    }
    return s;
};
var __sha_k = (function () {
    var k = [], n = 2;
    while (k.length < 64) {
        var p = true;
        for (var d = 2; d * d <= n; d++) if (n % d === 0) { p = false; break; }
        if (p) { var c = Math.pow(n, 1 / 3); k.push(Math.floor((c - Math.floor(c)) * 4294967296) >>> 0); }
        n++;
    }
    return k;
})();
function __sha256(u8) {
    var h = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19];
    var n = u8.length, total = ((n + 9 + 63) >> 6) << 6, m = new Uint8Array(total);
    m.set(u8); m[n] = 0x80;
    var bits = n * 8, hi = Math.floor(bits / 4294967296), lo = bits >>> 0;
    for (var i = 0; i < 4; i++) { m[total - 1 - i] = (lo >>> (8 * i)) & 255; m[total - 5 - i] = (hi >>> (8 * i)) & 255; }
    var w = new Int32Array(64);
    for (var off = 0; off < total; off += 64) {
        for (var t = 0; t < 16; t++) w[t] = (m[off + 4 * t] << 24) | (m[off + 4 * t + 1] << 16) | (m[off + 4 * t + 2] << 8) | m[off + 4 * t + 3];
        for (t = 16; t < 64; t++) {
            var x = w[t - 15], y = w[t - 2];
            var s0 = ((x >>> 7) | (x << 25)) ^ ((x >>> 18) | (x << 14)) ^ (x >>> 3);
            var s1 = ((y >>> 17) | (y << 15)) ^ ((y >>> 19) | (y << 13)) ^ (y >>> 10);
            w[t] = (w[t - 16] + s0 + w[t - 7] + s1) | 0;
        }
        var a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (t = 0; t < 64; t++) {
            var S1 = ((e >>> 6) | (e << 26)) ^ ((e >>> 11) | (e << 21)) ^ ((e >>> 25) | (e << 7));
            var ch = (e & f) ^ (~e & g);
            var t1 = (hh + S1 + ch + __sha_k[t] + w[t]) | 0;
            var S0 = ((a >>> 2) | (a << 30)) ^ ((a >>> 13) | (a << 19)) ^ ((a >>> 22) | (a << 10));
            var mj = (a & b) ^ (a & c) ^ (b & c);
            var t2 = (S0 + mj) | 0;
            hh = g; g = f; f = e; e = (d + t1) | 0; d = c; c = b; b = a; a = (t1 + t2) | 0;
        }
        h[0] = (h[0] + a) | 0; h[1] = (h[1] + b) | 0; h[2] = (h[2] + c) | 0; h[3] = (h[3] + d) | 0;
        h[4] = (h[4] + e) | 0; h[5] = (h[5] + f) | 0; h[6] = (h[6] + g) | 0; h[7] = (h[7] + hh) | 0;
    }
    var s = '';
    for (i = 0; i < 8; i++) s += ('00000000' + (h[i] >>> 0).toString(16)).slice(-8);
    return s;
}
function __hex(u8) { var s = ''; for (var i = 0; i < u8.length; i++) s += (u8[i] < 16 ? '0' : '') + u8[i].toString(16); return s; }
function __unhex(s) {
    var n = s.length >> 1, out = new Uint8Array(n);
    for (var i = 0; i < n; i++) {
        var a = s.charCodeAt(2 * i), b = s.charCodeAt(2 * i + 1);
        out[i] = ((a > 57 ? a - 87 : a - 48) << 4) | (b > 57 ? b - 87 : b - 48);
    }
    return out;
}
function __f32_from_hex(s) { var u = __unhex(s); return new Float32Array(u.buffer, 0, u.length >> 2); }
function __bytes_of(ta) { return new Uint8Array(ta.buffer, ta.byteOffset, ta.byteLength); }
