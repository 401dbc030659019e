set -x
O=gpurun_out/r2f
mkdir -p $O
FV3_PARITY_LOG=$O/parity_log.jsonl python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest_gpu.txt
python bench.py > $O/bench_A.out 2> $O/bench_A.err; tail -1 $O/bench_A.out > $O/r2_bench_c384_A.json
python bench.py --impl reference > $O/bench_ref.out 2> $O/bench_ref.err; tail -1 $O/bench_ref.out > $O/r2_bench_c384_reference.json
python bench.py --flagset B --no-cpu-baseline 2>/dev/null | tail -1 > $O/r2_bench_c384_B.json
python bench.py --transport-fp32 --no-cpu-baseline 2>/dev/null | tail -1 > $O/r2_bench_c384_A_fp32.json
python bench.py --graph --no-cpu-baseline 2>/dev/null | tail -1 > $O/r2_bench_c384_A_graph.json
python profiles/prof_stages.py 384 79 A 3 > $O/r2_stages_c384_A.txt
python profiles/prof_stages.py 384 79 B 3 > $O/r2_stages_c384_B.txt
FV3_TRANSPORT_FP32=1 python profiles/prof_stages.py 384 79 A 3 > $O/r2_stages_c384_A_fp32.txt
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_dsw|k_deln|k_copy_frame|k_a2b" --csv --log-file $O/r2_dsw_traffic_A.csv python profiles/prof_dsw.py 384 79 A > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_dsw|k_deln|k_copy_frame|k_a2b" --csv --log-file $O/r2_dsw_traffic_B.csv python profiles/prof_dsw.py 384 79 B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_substep_launches.csv python profiles/prof_stages.py 384 79 A 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dsw_transport2" -c 4 -o $O/r2_transport4 python profiles/prof_dsw.py 384 79 A > /dev/null 2>&1
ls -la $O
