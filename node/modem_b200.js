// Drop-in for the entry points of modem.js (SURVEY.md section 8b) and for StreamingReceiver (app.js:706-981), over the N-API
// addon (node/binding.cc).  NOT RUN IN THIS REPOSITORY (no Node.js in the build image); mirrors audio-modem_b200/modem.py one
// to one - same names, argument meaning, result objects and error strings as the reference.
//
//   const vm = require('node:vm'), fs = require('node:fs');
//   vm.runInThisContext(fs.readFileSync('modem.js', 'utf8'));      // the reference's globals
//   Object.assign(globalThis, require('./modem_b200.js'));          // replace the hot path
'use strict';
const addon = require('./audio_modem_b200.node');

const CFG = { standard: 0, acoustic: 1, narrowband: 2 };
const MOD = { BPSK: 0, QPSK: 1, QAM16: 2 };
let currentConfig = 'standard';               // the reference's hidden global OFDM (modem.js:87-98)
const handles = new Map();

function modem(modName, repetition) {
  const key = `${currentConfig}/${modName}/${repetition || 1}`;
  if (!handles.has(key)) handles.set(key, addon.create(CFG[currentConfig], MOD[modName], repetition || 1, 0));
  return handles.get(key);
}
const f32 = (x) => (x instanceof Float32Array ? x : Float32Array.from(x));     // the reference accepts any indexable
const u8 = (x) => (x instanceof Uint8Array ? x : Uint8Array.from(x));
const f64 = (x) => (x instanceof Float64Array ? x : Float64Array.from(x));

module.exports = {
  // modem.js:95-98: unknown names select 'standard'; the reference's own OFDM object follows when it is loaded
  setOFDMConfig(name) {
    currentConfig = name in CFG ? name : 'standard';
    if (typeof OFDM !== 'undefined' && typeof OFDM_CONFIGS !== 'undefined') Object.assign(OFDM, OFDM_CONFIGS[currentConfig]);
  },
  detectPreamble(signal) { return addon.detectPreamble(modem('QPSK', 1), f32(signal)); },
  detectPreambleCrossCorr(signal) { return addon.detectPreambleCrossCorr(modem('QPSK', 1), f32(signal)); },
  preprocessSignal(signal) { return addon.preprocessSignal(modem('QPSK', 1), f32(signal)); },
  estimateChannel(receivedSamples, knownRe, knownIm) { return addon.estimateChannel(modem('QPSK', 1), f32(receivedSamples), f64(knownRe), f64(knownIm)); },
  // modem.js:365: returns the reference's number[] of 0 / 1
  demodulateOFDM(signal, modName, channelRe, channelIm) { return Array.from(addon.demodulateOFDM(modem(modName, 1), f32(signal), f64(channelRe), f64(channelIm))); },
  // modem.js:322-362: {samples: Float32Array[], numSymbols, bitsPerSymbol}; pads the CALLER's array in place (modem.js:329)
  modulateOFDM(bits, modName) {
    const r = addon.modulateOFDM(modem(modName, 1), Uint8Array.from(bits, (b) => b & 1));
    if (Array.isArray(bits)) while (bits.length < r.paddedLength) bits.push(0);
    const symLen = r.numSymbols ? r.samples.length / r.numSymbols : 0;
    const samples = [];
    for (let s = 0; s < r.numSymbols; s++) samples.push(r.samples.subarray(s * symLen, (s + 1) * symLen));
    return { samples, numSymbols: r.numSymbols, bitsPerSymbol: r.bitsPerSymbol };
  },
  crc32(data) { return addon.crc32(modem('QPSK', 1), u8(data)); },
  estimateFrameSamples(payloadBytes, modName, repetition) { return addon.estimateFrameSamples(CFG[currentConfig], MOD[modName], payloadBytes, repetition || 1); },
  generatePreambleSymbol1() { return addon.generatePreambleSymbol1(CFG[currentConfig]); },
  generateChannelEstSymbol() { return addon.generateChannelEstSymbol(CFG[currentConfig]); },
  decodeReceivedSignal(signal, modName, repetition) { return addon.decodeReceivedSignal(modem(modName, repetition), f32(signal)); },
  decodeChunkFrame(frameSamples, modName, repetition) { return addon.decodeChunkFrame(modem(modName, repetition), f32(frameSamples)); },
  analyzeLoopback(recorded, modName, repetition, testData) { return addon.analyzeLoopback(modem(modName, repetition), f32(recorded), u8(testData)); },
  // modem.js:498-555: {signal, numSymbols, bitsPerSymbol, totalBits, dataLen}
  buildTransmitSignal(fileData, modName, fileName, repetition) {
    const rep = repetition || 1, data = u8(fileData), name = fileName || 'file';
    const r = addon.buildTransmitSignal(modem(modName, rep), data, name);
    const nameLen = Math.min(new TextEncoder().encode(name).length, 255);
    const totalBits = (1 + nameLen + 4 + data.length + 4) * 8 * rep;
    const bitsPerSymbol = addon.modulateOFDM(modem(modName, 1), new Uint8Array(0)).bitsPerSymbol;
    return { signal: r.signal, numSymbols: Math.ceil(totalBits / bitsPerSymbol), bitsPerSymbol, totalBits, dataLen: data.length };
  },
  generateTestSignal(modName, repetition) {
    const testData = Uint8Array.from({ length: 16 }, (_, i) => i);
    return { signal: addon.buildTransmitSignal(modem(modName, repetition), testData, 'test').signal, testData };
  },
  buildMetadataFrame(totalChunks, totalFileSize, chunkSize, fileName, modName, rep) {
    return addon.buildMetadataFrame(modem(modName, rep), totalChunks, totalFileSize, chunkSize, fileName || 'file');
  },
  buildDataChunkFrame(chunkData, seqNum, modName, rep) { return addon.buildDataChunkFrame(modem(modName, rep), u8(chunkData), seqNum); },
  // app.js:706-981 without DOM / IndexedDB: processAudioBlock returns the frames it completed
  StreamingReceiver: class {
    constructor(modName, repetition) {
      this.h = addon.create(CFG[currentConfig], MOD[modName], repetition || 1, 0);
      addon.streamOpen(this.h, 1, 4096, 4096 * 16);
      this.bytesStride = 8192;               // rows for chunks up to 4096 bytes (app.js:195-199); results do not depend on it
    }
    processAudioBlock(inputSamples) { return addon.streamPush(this.h, f32(inputSamples), inputSamples.length, this.bytesStride); }
    // the same block as the 16-bit PCM a capture device delivers (sample = s / 32768): half the bytes across the bus
    processAudioBlockPCM16(int16Samples) { return addon.streamPushPCM16(this.h, Int16Array.from(int16Samples), int16Samples.length, this.bytesStride); }
    // framesDecoded, frameErrors (app.js:738-739) and the ChunkAssembler counters (app.js:604-605)
    counters() { return addon.streamCounters(this.h, 0); }
    cleanup() { addon.streamClose(this.h); }
  },
};
