// Runs the cases of tests/refcases.py through the reference (same global scope as modem.js and the app.js cut) and returns
// what tests/refcases.py::oracle_outputs returns for the oracle, in the same schema.  `__cases` is written by the driver.
function __resOf(r) {
    var d = __resultToDict(r);
    if (d.data !== undefined) { d.data_sha256 = __sha256(Uint8Array.from(d.data)); delete d.data; }
    return d;
}
function __runCase(c) {
    var m = __MODES[c.mode], cfg = m[0], mod = m[1], rep = m[2];
    setOFDMConfig(cfg);
    var inp = {};
    Object.keys(c.inputs).forEach(function (k) { inp[k] = c.inputs[k].dtype === 'f32' ? __f32_from_hex(c.inputs[k].hex) : __unhex(c.inputs[k].hex); });
    var SL = OFDM.SYMBOL_LEN, out = {};
    if (c.kind === 'decode') {
        var pre = preprocessSignal(inp.x);
        out.result = __resOf(decodeReceivedSignal(inp.x, mod, rep));
        out.detectPreamble = detectPreamble(pre);
        out.detectPreambleCrossCorr = detectPreambleCrossCorr(pre);
        out.preprocess_sha256 = __sha256(__bytes_of(pre));
    } else if (c.kind === 'chunkframe') {
        out.result = __resOf(decodeChunkFrame(inp.x, mod, rep));
    } else if (c.kind === 'ops') {
        var ce = generateChannelEstSymbol();
        var h = estimateChannel(inp.x.subarray(2 * SL, 3 * SL), ce.knownRe, ce.knownIm);
        var eqs = [], realDemap = constellationDemap;
        constellationDemap = function (cc, cr, ci) { eqs.push(cr, ci); return realDemap(cc, cr, ci); };   // the points demodulateOFDM hands to the demapper
        var bits = demodulateOFDM(inp.x.subarray(3 * SL), mod, h[0], h[1]);
        constellationDemap = realDemap;
        var eqa = Float64Array.from(eqs);
        out.eq_sha256 = __sha256(__bytes_of(eqa)); out.eq_n = eqs.length / 2; out.eq_first = eqs.slice(0, 6);
        var hh = new Float64Array(2 * OFDM.FFT_SIZE);
        hh.set(h[0], 0); hh.set(h[1], OFDM.FFT_SIZE);
        out.H_sha256 = __sha256(__bytes_of(hh));
        out.H_first = [h[0][OFDM.SUB_START], h[1][OFDM.SUB_START]];
        out.bits_sha256 = __sha256(Uint8Array.from(bits));
        out.n_bits = bits.length;
    } else if (c.kind === 'loopback') {
        var fx = new Uint8Array(16);
        for (var q = 0; q < 16; q++) fx[q] = q;
        var lb = analyzeLoopback(inp.x, mod, rep, fx);
        var mags = Float64Array.from(lb.channelMagnitude);
        out = { detected: lb.detected, correlation: lb.correlation, ber: lb.ber, snrEstimate: lb.snrEstimate, quality: lb.quality,
            channelMagnitude_sha256: __sha256(__bytes_of(mags)), channelMagnitude_n: mags.length };
    } else if (c.kind === 'modulate') {
        var b = Array.from(inp.bits);
        var r = modulateOFDM(b, mod);
        var all = new Float32Array(r.samples.length * SL);
        for (var i = 0; i < r.samples.length; i++) all.set(r.samples[i], i * SL);
        out.numSymbols = r.numSymbols; out.bitsPerSymbol = r.bitsPerSymbol; out.padded_len = b.length;       // modem.js:329 pads the caller's array
        out.samples_sha256 = __sha256(__bytes_of(all));
    } else if (c.kind === 'stream') {
        var rx = null;
        var st = (function () {
            var rxx = new StreamingReceiver(mod, rep); rx = rxx;
            var ev = [], realDecode = decodeChunkFrame;
            decodeChunkFrame = function (frame, mm, rr) {
                var res = realDecode(frame, mm, rr);
                var d = __resOf(res);
                d.preambleIdx = rxx.preambleGlobalPos;
                ev.push(d);
                return res;
            };
            for (var off = 0; off + 4096 <= inp.x.length; off += 4096) {
                var n0 = ev.length, fe0 = rxx.frameErrors, pos0 = rxx.preambleGlobalPos;
                rxx.processAudioBlock(inp.x.subarray(off, off + 4096));
                if (rxx.frameErrors > fe0 && ev.length === n0) ev.push({ error: '(ring buffer overwritten)', preambleIdx: pos0 });
            }
            decodeChunkFrame = realDecode;
            return ev;
        })();
        out.events = st;
        out.counters = { framesDecoded: rx.framesDecoded, frameErrors: rx.frameErrors, receivedCount: rx.assembler.receivedCount,
            crcErrors: rx.assembler.crcErrors, duplicates: rx.assembler.duplicates, metaReceived: rx.metaReceived ? 1 : 0, acScanPos: rx.acScanPos };
        out.dcMean = rx.dcMean;
    } else if (c.kind === 'misc') {
        out.crc32 = {};
        [0, 1, 7, 4107].forEach(function (n) { out.crc32[String(n)] = crc32(inp.bytes_a.subarray(0, n)); });
        var cat = function (p) { var o = new Float64Array(1024); o.set(p[0], 0); o.set(p[1], 512); return __sha256(__bytes_of(o)); };
        out.fft_sha256 = cat(fft(Float64Array.from(inp.fft_re), Float64Array.from(inp.fft_im)));
        out.ifft_sha256 = cat(ifft(Float64Array.from(inp.fft_re), Float64Array.from(inp.fft_im)));
        out.majorityVote3_sha256 = __sha256(Uint8Array.from(majorityVote(Array.from(inp.bits_a), 3)));
        out.bitsToBytes_sha256 = __sha256(Uint8Array.from(bitsToBytes(Array.from(inp.bits_a))));
        out.estimateFrameSamples = {}; out.configs = {};
        ['standard', 'acoustic', 'narrowband'].forEach(function (cn) {
            setOFDMConfig(cn);
            out.configs[cn] = { FFT_SIZE: OFDM.FFT_SIZE, CP_LEN: OFDM.CP_LEN, SYMBOL_LEN: OFDM.SYMBOL_LEN, SUB_START: OFDM.SUB_START, SUB_END: OFDM.SUB_END,
                PILOTS: Array.from(OFDM.PILOTS), numDataSubs: OFDM.numDataSubs(), pre2_sha256: __sha256(__bytes_of(__f32(generatePreambleSymbol2()))) };
            ['BPSK', 'QPSK', 'QAM16'].forEach(function (mn) { [1, 3].forEach(function (rr) { [0, 1, 280, 4112].forEach(function (p) {
                out.estimateFrameSamples[cn + '/' + mn + '/' + rr + '/' + p] = estimateFrameSamples(p, mn, rr); }); }); });
        });
        out.constellations = {};
        ['BPSK', 'QPSK', 'QAM16'].forEach(function (mn) { out.constellations[mn] = initConstellation(mn).points; });
        var rng = seededRandom(12345), v = 0;
        for (var k = 0; k < 1000; k++) v = rng();
        out.lcg_12345_state_1000 = Math.round(v * 0x7fffffff);
    }
    return out;
}
function __runCases() {
    var res = {};
    __cases.forEach(function (c) { var t0 = Date.now(); var r = __runCase(c); r.__ms = Date.now() - t0; res[c.mode + ':' + c.name] = r; });
    return res;
}
