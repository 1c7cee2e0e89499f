#!/bin/bash
# compute-sanitizer over a representative slice of the -m gpu suite and one small bench step, one tool per run.
# usage: tools/sanitize.sh <memcheck|racecheck|initcheck|synccheck>     logs: gpurun_out/sanitizer/<tool>_{tests,bench}.log
tool="$1"; out=gpurun_out/sanitizer; mkdir -p $out
sel="stream_matches_oracle or stream_edge_cases or nan_samples or slot_table or ring_overwritten or chunk_size_larger or decode_received_signals_batch or decode_chunk_frames or chunk_store or detect_preamble or test_crc32 or analyze_loopback or tx_frames or pcm16"
extra=""
[ "$tool" = "memcheck" ] && extra="--leak-check no"
[ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool $extra --error-exitcode 99 --target-processes all --log-file $out/${tool}_tests.raw \
  python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$sel" > $out/${tool}_tests.log 2>&1
echo "$tool tests rc=$?" | tee -a $out/${tool}_tests.log
timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool $extra --error-exitcode 99 --log-file $out/${tool}_bench.raw \
  python bench.py --workload c2 --frames-per-step 512 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-shard-probe > $out/${tool}_bench.log 2>&1
echo "$tool bench rc=$?" | tee -a $out/${tool}_bench.log
for f in $out/${tool}_tests.raw $out/${tool}_bench.raw; do echo "== $f"; grep -c "=========" $f; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|Error:\|Race reported\|hazard" $f | sort | uniq -c | head -20; done
