mkdir -p gpurun_out
python -m pytest tests/test_gpu_local.py tests/test_modvalue.py -q -x -m gpu 2>&1 | tail -8 > gpurun_out/r02_local_pytest.log; cat gpurun_out/r02_local_pytest.log
python tools/time_local.py > gpurun_out/r02_local_timing.jsonl 2> gpurun_out/r02_local_timing.err; tail -3 gpurun_out/r02_local_timing.err; grep -E "compose|nonzero|axpb" gpurun_out/r02_local_timing.jsonl
ncu --set full --clock-control none --import-source on -k regex:k_bits_compose -c 2 -o gpurun_out/r02_ncu_compose -f python tools/time_local.py --only-compose > gpurun_out/r02_ncu_compose.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_ncu_compose.ncu-rep gpurun_out/r02_ncu_bits_compose.txt > /dev/null 2>&1
ncu -i gpurun_out/r02_ncu_compose.ncu-rep --page raw --csv > gpurun_out/r02_ncu_compose_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_compose.ncu-rep
tail -40 gpurun_out/r02_ncu_bits_compose.txt
