// Drop-in for the receive entry points of modem.js, over the N-API addon (node/binding.cc).
// NOT RUN IN THIS REPOSITORY (no Node.js in the build image); mirrors audio-modem_b200/modem.py one to one.
//
//   const vm = require('node:vm'), fs = require('node:fs');
//   vm.runInThisContext(fs.readFileSync('modem.js', 'utf8'));      // the reference's globals (TX side, helpers)
//   Object.assign(globalThis, require('./modem_b200.js'));          // replace the receive path
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

module.exports = {
  setOFDMConfig(name) { currentConfig = name in CFG ? name : 'standard'; if (typeof OFDM !== 'undefined') Object.assign(OFDM, OFDM_CONFIGS[currentConfig]); },
  detectPreamble(signal) { return addon.detectPreamble(modem('QPSK', 1), Float32Array.from(signal)); },
  decodeReceivedSignal(signal, modName, repetition) { return addon.decodeReceivedSignal(modem(modName, repetition), Float32Array.from(signal)); },
  decodeChunkFrame(frameSamples, modName, repetition) { return addon.decodeChunkFrame(modem(modName, repetition), Float32Array.from(frameSamples)); },
  analyzeLoopback(recorded, modName, repetition, testData) { return addon.analyzeLoopback(modem(modName, repetition), Float32Array.from(recorded), Uint8Array.from(testData)); },
  buildTransmitSignal(fileData, modName, fileName, repetition) { return addon.buildTransmitSignal(modem(modName, repetition), Uint8Array.from(fileData), fileName || 'file'); },
  generateTestSignal(modName, repetition) {
    const testData = Uint8Array.from({ length: 16 }, (_, i) => i);
    return { signal: addon.buildTransmitSignal(modem(modName, repetition), testData, 'test').signal, testData };
  },
  // app.js:706-981 without DOM / IndexedDB: processAudioBlock returns the frames it completed
  StreamingReceiver: class {
    constructor(modName, repetition) {
      this.h = addon.create(CFG[currentConfig], MOD[modName], repetition || 1, 0);
      addon.streamOpen(this.h, 1, 4096, 4096 * 16);
      this.bytesStride = 8192;
    }
    processAudioBlock(inputSamples) { return addon.streamPush(this.h, Float32Array.from(inputSamples), inputSamples.length, this.bytesStride); }
    // the same block as the 16-bit PCM a capture device delivers (sample = s / 32768): half the bytes across the bus
    processAudioBlockPCM16(int16Samples) { return addon.streamPushPCM16(this.h, Int16Array.from(int16Samples), int16Samples.length, this.bytesStride); }
  },
};
