#!/usr/bin/env python
"""Pins the oracle against the REAL reference, here, without Node: evaluates playok/audio-modem's own modem.js and the
DOM-free classes of app.js (RECV_STATE, RingBuffer, StreamingReceiver) with the one ECMAScript engine this image has - Qt
6.6's QJSEngine inside Nsight Compute's host directory, driven by tools/qjs/qjs_run.cpp - and writes
tests/golden/reference_golden.json (base schema of golden.json) and tests/golden/reference_cases.json (the wider case list
of tools/qjs/cases.py).  tests/test_golden.py compares both with the oracle.

    python tools/qjs/dump_golden_qjs.py [/root/reference]

The reference sources are read where they lie and evaluated as they are, except for two mechanical rewrites the
engine's language level (ES2016) forces, both printed when they are applied:
  1. modem.js:87   `{ ...OFDM_CONFIGS.standard }`  ->  `Object.assign({}, OFDM_CONFIGS.standard)`   (object spread is ES2018)
  2. app.js StreamingReceiver: the keywords `async` / `await` are dropped (ES2017).  The awaited calls are the storage-free
     ChunkAssembler stubs, which return synchronously, and `_demodulateFrame` is the last statement processAudioBlock
     reaches (app.js:899-905), so the continuation after each `await` runs before the next block either way.
DOM calls (addLog, updateStreamingUI, ...) are empty stubs; ChunkAssembler is the bitmap / counter logic of app.js:597-704
without IndexedDB (the same stub text as node/dump_golden.js).
"""
import json
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STUBS = r"""
  function addLog() {} function updateStreamingUI() {} function drawChunkBitmap() {} function formatSize(x) { return String(x); }
  function updateProgress() {} function offerDownload() {}
  var document = { getElementById: function () { return null; } };
  // ChunkAssembler without IndexedDB: the bitmap / counter semantics of app.js:597-704
  class ChunkAssembler {
    constructor() { this.totalChunks = 0; this.totalFileSize = 0; this.chunkSize = 0; this.fileName = ''; this.receivedBitmap = null; this.receivedCount = 0; this.crcErrors = 0; this.duplicates = 0; }
    handleMetadataFrame(r) { this.totalChunks = r.totalChunks; this.totalFileSize = r.totalFileSize; this.chunkSize = r.chunkSize; this.fileName = r.fileName;
      this.receivedBitmap = new Uint8Array(Math.ceil(Math.max(r.totalChunks, 0) / 8)); this.receivedCount = 0; this.crcErrors = 0; }
    handleDataChunk(seq, data, ok) { if (!this.receivedBitmap || seq >= this.totalChunks) return; if (!ok) { this.crcErrors++; return; }
      var byte = seq >> 3, bit = 1 << (seq & 7); if (this.receivedBitmap[byte] & bit) { this.duplicates++; return; } this.receivedBitmap[byte] |= bit; this.receivedCount++; }
    isComplete() { return false; }
    cleanup() {}
  }
"""


def cut(app, start_re):
    m = re.search(start_re, app)
    if not m:
        raise SystemExit("not found in app.js: " + start_re)
    i = app.index("{", m.start())
    depth = 0
    while True:
        if app[i] == "{":
            depth += 1
        elif app[i] == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return app[m.start(): i + 1]


def reference_program(ref_dir, verbose=True):
    modem = open(os.path.join(ref_dir, "modem.js"), encoding="utf-8").read()
    spread = "{ ...OFDM_CONFIGS.standard }"
    assert modem.count(spread) == 1 and len(re.findall(r"\{\s*\.\.\.", modem)) == 1, "unexpected object spread in modem.js"
    modem = modem.replace(spread, "Object.assign({}, OFDM_CONFIGS.standard)")
    app = open(os.path.join(ref_dir, "app.js"), encoding="utf-8").read()
    rx = cut(app, r"class StreamingReceiver\s*\{")
    n_async, n_await = len(re.findall(r"\basync ", rx)), len(re.findall(r"\bawait ", rx))
    rx = re.sub(r"\bawait ", "", re.sub(r"\basync ", "", rx))
    if verbose:
        print(f"rewrites: 1 object spread (modem.js:87), {n_async} `async` + {n_await} `await` keywords (app.js StreamingReceiver)", file=sys.stderr)
    state = cut(app, r"const RECV_STATE\s*=") + ";"
    ring = cut(app, r"class RingBuffer\s*\{")
    return modem + "\n" + STUBS + "\n" + state + "\n" + ring + "\n" + rx + "\n"


def run_js(program, tag):
    exe = os.path.join(ROOT, "oracle", "_ref", "qjs_run")
    if not os.path.exists(exe):
        subprocess.check_call([os.path.join(HERE, "build.sh")])
    tmp = os.path.join(ROOT, "oracle", "_ref", tag + ".js")
    open(tmp, "w", encoding="utf-8").write(program)
    env = dict(os.environ, QT_NO_GLIB="1", LC_ALL="C.UTF-8")
    p = subprocess.run([exe, tmp], env=env, capture_output=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode("utf-8", "replace"))
        raise SystemExit(f"qjs_run failed on {tmp}")
    os.remove(tmp)
    return p.stdout.decode("utf-8")


def main():
    ref_dir = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    from helpers import MODES, make_stream
    from oracle import oracle as O

    prelude = open(os.path.join(HERE, "prelude.js"), encoding="utf-8").read()
    ref = reference_program(ref_dir)
    inputs = {"engine": "6.6 QJSEngine (libQt6Qml of Nsight Compute), via tools/qjs/qjs_run", "streams": {}, "random": {}}
    for mode, (cfg, mod, rep, chunk) in MODES.items():
        x, _ = make_stream(O, mode, 3, 21)
        inputs["streams"][mode] = np.asarray(x, np.float32).tobytes().hex()
        inputs["random"][str(chunk)] = bytes(O.randomBytes(4242, chunk)).hex()
    body = open(os.path.join(HERE, "dump_body.js"), encoding="utf-8").read()
    prog = prelude + "\n" + ref + "\nvar __in = " + json.dumps(inputs) + ";\n" + body + "\n__out = JSON.stringify(__dump());\n"
    g = json.loads(run_js(prog, "golden_prog"))
    json.dump(g, open(os.path.join(GOLD, "reference_golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote tests/golden/reference_golden.json")


def dump_cases(ref_dir, quick=False, out_name="reference_cases.json"):
    """the wider case list (tests/refcases.py), one engine run per mode to bound the size of a program"""
    import time

    import refcases
    from oracle import oracle as O

    prelude = open(os.path.join(HERE, "prelude.js"), encoding="utf-8").read()
    ref = reference_program(ref_dir, verbose=False)
    body = open(os.path.join(HERE, "dump_body.js"), encoding="utf-8").read() + open(os.path.join(HERE, "cases_body.js"), encoding="utf-8").read()
    cases = refcases.build_cases(O, quick=quick)
    out = {"note": "tests/refcases.py evaluated by the reference (modem.js, app.js) under Qt 6.6 QJSEngine; written by tools/qjs/dump_golden_qjs.py",
           "inputs_sha256": {}, "cases": {}}
    groups = {}
    for c in cases:
        groups.setdefault((c["mode"], c["kind"] == "stream"), []).append(c)
    for (mode, is_stream), cs in groups.items():
        js_cases = []
        for c in cs:
            inputs = {}
            for k, a in c["inputs"].items():
                a = np.ascontiguousarray(a)
                inputs[k] = {"dtype": "f32" if a.dtype == np.float32 else "u8", "hex": a.tobytes().hex()}
                out["inputs_sha256"][f"{c['mode']}:{c['name']}:{k}"] = refcases.sha(a)
            js_cases.append({"name": c["name"], "mode": c["mode"], "kind": c["kind"], "args": c["args"], "inputs": inputs})
        t0 = time.time()
        prog = prelude + "\n" + ref + "\nvar __cases = " + json.dumps(js_cases) + ";\n" + body + "\n__out = JSON.stringify(__runCases());\n"
        got = json.loads(run_js(prog, "cases_prog"))
        for c in cs:                                   # engine wall time per case goes to a side table (not a parity value)
            key = c["mode"] + ":" + c["name"]
            n_samples = int(sum(len(a) for k, a in c["inputs"].items() if k == "x"))
            out.setdefault("engine_ms", {})[key] = {"ms": got[key].pop("__ms"), "samples": n_samples}
        out["cases"].update(got)
        print(f"{mode} {'stream' if is_stream else 'batch'}: {len(cs)} cases in {time.time() - t0:.0f} s", file=sys.stderr)
    json.dump(out, open(os.path.join(GOLD, out_name), "w"), indent=1, sort_keys=True)
    print("wrote tests/golden/" + out_name, len(out["cases"]), "cases")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ref_dir = args[0] if args else "/root/reference"
    if "--cases-only" not in sys.argv:
        main()
    if "--golden-only" not in sys.argv:
        dump_cases(ref_dir, quick="--quick" in sys.argv, out_name="reference_cases_quick.json" if "--quick" in sys.argv else "reference_cases.json")
