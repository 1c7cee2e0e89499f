#!/usr/bin/env node
// Pins the oracle against the REAL reference: evaluates playok/audio-modem's own modem.js and the DOM-free part of app.js
// (RingBuffer, StreamingReceiver) under Node.js and writes tests/golden/reference_golden.json in the schema of
// tests/golden/golden.json (which the oracle wrote).  tests/test_golden.py::test_reference_golden_matches_the_oracle then
// compares the two (DESIGN.md section 2).
//
// NOT RUN IN THIS REPOSITORY (no Node in the build image).  The committed reference_golden.json was written by the same dump
// under Qt's QJSEngine (tools/qjs/dump_golden_qjs.py); this script is the second opinion under V8.  One command, Node >= 16:
//
//     python tests/golden/make_golden.py --inputs        # writes tests/golden/_inputs/*.f32 (the noisy stream inputs)
//     node node/dump_golden.js /path/to/audio-modem      # the reference checkout (modem.js, app.js)
//     python -m pytest tests/test_golden.py -q
//
// Everything the reference can produce itself is produced by it (transmit frames, decodes, detector indices, frame lengths,
// CRCs, LCG states, loopback analysis, streaming events); only the AWGN of the stream inputs comes from the files above
// (Box-Muller through libm would not be bit-reproducible across engines).  Payload bytes come from SplitMix64, reproduced
// here with BigInt exactly as oracle/modem_oracle.cpp:orc_random_bytes does.
'use strict';
const fs = require('node:fs');
const path = require('node:path');
const vm = require('node:vm');
const crypto = require('node:crypto');

const refDir = process.argv[2];
if (!refDir || !fs.existsSync(path.join(refDir, 'modem.js'))) {
  console.error('usage: node node/dump_golden.js /path/to/audio-modem [out.json]');
  process.exit(2);
}
const here = path.join(__dirname, '..', 'tests', 'golden');
const outFile = process.argv[3] || path.join(here, 'reference_golden.json');

// ---------------------------------------------------------------- the reference, in its own context
const sandbox = { console, TextEncoder, TextDecoder, Date, Math, Float32Array, Float64Array, Uint8Array, Int32Array, Uint32Array, Array, Promise };
vm.createContext(sandbox);
vm.runInContext(fs.readFileSync(path.join(refDir, 'modem.js'), 'utf8'), sandbox, { filename: 'modem.js' });
// app.js: only RECV_STATE, RingBuffer and StreamingReceiver (the rest is DOM / Web Audio / IndexedDB); cut out by balanced braces
const app = fs.readFileSync(path.join(refDir, 'app.js'), 'utf8');
function cut(startRe) {
  const m = startRe.exec(app);
  if (!m) throw new Error('not found in app.js: ' + startRe);
  let i = app.indexOf('{', m.index), depth = 0;
  for (; i < app.length; i++) {
    if (app[i] === '{') depth++;
    else if (app[i] === '}' && --depth === 0) break;
  }
  return app.slice(m.index, i + 1);
}
const stubs = `
  function addLog() {} function updateStreamingUI() {} function drawChunkBitmap() {} function formatSize(x) { return String(x); }
  function updateProgress() {} function offerDownload() {}
  const document = { getElementById() { return null; } };
  // ChunkAssembler without IndexedDB: the bitmap / counter semantics of app.js:597-704
  class ChunkAssembler {
    constructor() { this.totalChunks = 0; this.totalFileSize = 0; this.chunkSize = 0; this.fileName = ''; this.receivedBitmap = null; this.receivedCount = 0; this.crcErrors = 0; this.duplicates = 0; }
    async handleMetadataFrame(r) { this.totalChunks = r.totalChunks; this.totalFileSize = r.totalFileSize; this.chunkSize = r.chunkSize; this.fileName = r.fileName;
      this.receivedBitmap = new Uint8Array(Math.ceil(Math.max(r.totalChunks, 0) / 8)); this.receivedCount = 0; this.crcErrors = 0; }
    async handleDataChunk(seq, data, ok) { if (!this.receivedBitmap || seq >= this.totalChunks) return; if (!ok) { this.crcErrors++; return; }
      const byte = seq >> 3, bit = 1 << (seq & 7); if (this.receivedBitmap[byte] & bit) { this.duplicates++; return; } this.receivedBitmap[byte] |= bit; this.receivedCount++; }
    isComplete() { return false; }
    cleanup() {}
  }
`;
vm.runInContext(stubs + cut(/const RECV_STATE\s*=/).replace(/\}$/, '};') + '\n' + cut(/class RingBuffer\s*\{/) + '\n' + cut(/class StreamingReceiver\s*\{/) +
  '\nthis.StreamingReceiver = StreamingReceiver;', sandbox, { filename: 'app.js(cut)' });

// ---------------------------------------------------------------- helpers in the schema of make_golden.py
const MODES = { 'QPSK': ['standard', 'QPSK', 1, 2048], '16-QAM': ['standard', 'QAM16', 1, 4096], 'BPSK-ACOUSTIC': ['acoustic', 'BPSK', 1, 512],
  'BPSK-REPEAT': ['acoustic', 'BPSK', 3, 512], 'BPSK-NARROW': ['narrowband', 'BPSK', 3, 512] };
const sha = (ta) => crypto.createHash('sha256').update(Buffer.from(ta.buffer, ta.byteOffset, ta.byteLength)).digest('hex');
const f32 = (x) => (x instanceof Float32Array ? x : Float32Array.from(x));
function summary(x) {
  x = f32(x);
  let peak = 0;
  for (let i = 0; i < x.length; i++) peak = Math.max(peak, Math.abs(x[i]));
  return { n: x.length, sha256: sha(x), first: Array.from(x.slice(0, 4)), peak };
}
function randomBytes(seed, n) {        // SplitMix64, little-endian bytes (orc_random_bytes)
  const M = (1n << 64n) - 1n;
  let s = BigInt(seed);
  const out = new Uint8Array(n);
  for (let i = 0; i < n; i += 8) {
    s = (s + 0x9E3779B97F4A7C15n) & M;
    let z = s;
    z = ((z ^ (z >> 30n)) * 0xBF58476D1CE4E5B9n) & M;
    z = ((z ^ (z >> 27n)) * 0x94D049BB133111EBn) & M;
    z = z ^ (z >> 31n);
    for (let j = 0; j < 8 && i + j < n; j++) out[i + j] = Number((z >> BigInt(8 * j)) & 0xFFn);
  }
  return out;
}
const hex = (u8) => Buffer.from(u8).toString('hex');
function resultToDict(r) {             // the reference's result object -> the dict oracle.py:result_to_dict makes
  if (r.error !== undefined) { const d = { error: r.error }; if (r.frameType !== undefined) d.frameType = r.frameType; return d; }
  const d = {};
  for (const k of Object.keys(r)) d[k] = r[k];
  return d;
}

async function main() {
  const g = { note: 'generated by node/dump_golden.js from the reference (modem.js, app.js) under Node ' + process.version, modes: {} };
  const fixture = Uint8Array.from({ length: 16 }, (_, i) => i);
  for (const [mode, [cfg, mod, rep, chunk]] of Object.entries(MODES)) {
    sandbox.setOFDMConfig(cfg);
    const e = {};
    const sig = sandbox.buildTransmitSignal(fixture, mod, 'test', rep).signal;     // generateTestSignal's frame (modem.js:914-973)
    e.test_signal = summary(sig);
    const dec = resultToDict(sandbox.decodeReceivedSignal(sig, mod, rep));
    if (dec.data !== undefined) dec.data = hex(dec.data);
    e.test_decode = dec;
    e.detectPreamble = sandbox.detectPreamble(sandbox.preprocessSignal(sig));
    const lb = sandbox.analyzeLoopback(sig, mod, rep, fixture);
    const mags = Float64Array.from(lb.channelMagnitude);
    e.loopback = { detected: lb.detected, correlation: lb.correlation, ber: lb.ber, snrEstimate: lb.snrEstimate, quality: lb.quality,
      channelMagnitude_sha256: sha(mags), channelMagnitude_n: mags.length, channelMagnitude_first: Array.from(mags.slice(0, 3)) };
    e.pre1 = summary(sandbox.generatePreambleSymbol1());
    const ce = sandbox.generateChannelEstSymbol();
    e.ce = summary(ce.samples);
    e.known_plus = Array.from(ce.knownRe).filter((v) => v > 0).length;
    e.meta_frame = summary(sandbox.buildMetadataFrame(3, 3 * chunk, chunk, 'golden.bin', mod, rep));
    e.data_frame = summary(sandbox.buildDataChunkFrame(randomBytes(4242, chunk), 2, mod, rep));
    e.estimateFrameSamples = {};
    for (const p of [11, 280, chunk + 11, 4112]) e.estimateFrameSamples[String(p)] = sandbox.estimateFrameSamples(p, mod, rep);
    // streaming: the noisy input make_golden.py --inputs wrote (metadata + 3 chunks, AWGN 25 dB), 4096-sample blocks
    const inFile = path.join(here, '_inputs', mode + '_stream.f32');
    if (fs.existsSync(inFile)) {
      const buf = fs.readFileSync(inFile);
      const x = new Float32Array(buf.buffer, buf.byteOffset, buf.byteLength / 4);
      const rx = new sandbox.StreamingReceiver(mod, rep);
      const ev = [];
      const realDecode = sandbox.decodeChunkFrame;
      sandbox.decodeChunkFrame = function (frame, m, r) {       // (the receiver calls the global: record what it got)
        const res = realDecode(frame, m, r);
        const d = resultToDict(res);
        d.preambleIdx = rx.preambleGlobalPos;
        if (d.data !== undefined) { d.data_sha256 = crypto.createHash('sha256').update(Buffer.from(d.data)).digest('hex'); delete d.data; }
        ev.push(d);
        return res;
      };
      for (let off = 0; off + 4096 <= x.length; off += 4096) {
        const n0 = ev.length, fe0 = rx.frameErrors, pos0 = rx.preambleGlobalPos;
        rx.processAudioBlock(x.subarray(off, off + 4096));
        await new Promise((resolve) => setImmediate(resolve));   // _demodulateFrame is async (IndexedDB in the browser): let it finish
        // a frame error without a decode call is the ring-overwritten path (app.js:912-916)
        if (rx.frameErrors > fe0 && ev.length === n0) ev.push({ error: '(ring buffer overwritten)', preambleIdx: pos0 });
      }
      sandbox.decodeChunkFrame = realDecode;
      e.stream = { input: summary(x), events: ev,
        counters: { framesDecoded: rx.framesDecoded, frameErrors: rx.frameErrors, receivedCount: rx.assembler.receivedCount, crcErrors: rx.assembler.crcErrors, duplicates: rx.assembler.duplicates } };
    } else {
      e.stream = { missing_input: inFile };
    }
    g.modes[mode] = e;
  }
  sandbox.setOFDMConfig('standard');
  g.crc32 = { '123456789': sandbox.crc32(new TextEncoder().encode('123456789')),
    fixture_packet: sandbox.crc32(Uint8Array.from([4, 0x74, 0x65, 0x73, 0x74, 0, 0, 0, 16, ...fixture])) };
  g.lcg = {};
  for (const s of [42, 43, 44]) { const rng = sandbox.seededRandom(s); g.lcg[String(s)] = [0, 1, 2, 3].map(() => Math.round(rng() * 0x7fffffff)); }
  fs.writeFileSync(outFile, JSON.stringify(g, null, 1));
  console.log('wrote', outFile);
}
main().catch((err) => { console.error(err); process.exit(1); });
