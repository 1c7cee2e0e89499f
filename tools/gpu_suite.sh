#!/bin/bash
# Runs the -m gpu tests in separate processes (one CUDA fault must not hide the other groups).
# Logs go to gpurun_out/tests_<group>.log; a summary line per group is printed.
mkdir -p gpurun_out
groups=("crc32 or preprocess or extension or api_argument or selftest" "detect_preamble" "estimate_channel" "decode_chunk_frames or fast_demodulator"
        "decode_received" "stream_matches_oracle" "stream_single_block or stream_block_alignment" "stream_edge_cases or second_metadata or phase_a_walks"
        "many_streams or round_trip" "tx_frames" "golden")
i=0
for g in "${groups[@]}"; do
  log="gpurun_out/tests_$i.log"
  timeout ${AMB_GROUP_TIMEOUT:-240} python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$g" > "$log" 2>&1
  echo "[$i] rc=$? $g :: $(tail -1 $log)"
  i=$((i+1))
done
