#!/bin/bash
# End-of-round evidence on ONE B200 (run under gpurun from the repository root): the -m gpu suite, the five-config bench lines,
# the default line and its reference arm, the ncu launch list of one timed C2 step and one `--set full` capture of the C2
# kernels over a push of 4,096 frames.  Everything goes to gpurun_out/cap_*; tools/ncu_summary.py turns the report into the
# CSV and the traffic file kept under profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/cap_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/cap_tests.log
python bench.py --workload all --steps 20 --warmup 5 > gpurun_out/cap_bench_all.jsonl 2> gpurun_out/cap_bench_all.err; echo "bench all rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/cap_bench_default.json 2> gpurun_out/cap_bench_default.err; echo "bench default rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/cap_bench_reference.json 2> gpurun_out/cap_bench_reference.err; echo "reference arm rc=$?"
python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-shard-probe > gpurun_out/cap_plain_c2.json 2> gpurun_out/cap_plain_c2.err; echo "plain c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/cap_launches.csv \
    python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-shard-probe > gpurun_out/cap_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python bench.py --workload c2 --frames-per-step 4096 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-shard-probe > gpurun_out/cap_plain_c2_4096.json 2> gpurun_out/cap_plain_c2_4096.err; echo "plain c2 4096 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_front|k_demod_fast|k_xcorr_store|k_xcorr_rows|k_link_walk|k_tri_store|k_parse|k_chain_links|k_cold' \
    --launch-skip 24 -c 24 -f -o gpurun_out/cap_c2_full \
    python bench.py --workload c2 --frames-per-step 4096 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-shard-probe > gpurun_out/cap_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/cap_c2_full.ncu-rep
