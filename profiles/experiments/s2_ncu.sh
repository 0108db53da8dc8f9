cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-also"
$CMD > gpurun_out/s2_plain_for_ncu.json 2> gpurun_out/s2_plain_for_ncu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r02b.csv $CMD > gpurun_out/s2_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_gemm -s 244 -c 244 --csv --log-file gpurun_out/conv_traffic_r02b.csv $CMD > gpurun_out/s2_ncu_traffic.log 2>&1
python profiles/conv_ncu_klarge.py > gpurun_out/s2_plain_klarge.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 1 -o gpurun_out/conv_klarge_r02b python profiles/conv_ncu_klarge.py > gpurun_out/s2_ncu_klarge.log 2>&1
ls -la gpurun_out/*r02b*
