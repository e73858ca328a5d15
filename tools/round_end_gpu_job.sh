set -x
# round-end evidence run (one GPU): tests, bench line, launch list, full ncu captures of the kernels DESIGN.md argues on
P=${1:-r2h}
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${P}_gputests.log 2>&1
(time python bench.py --steps 10 --warmup 3) > gpurun_out/${P}_bench1.log 2>&1
# launch list of the same command (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 1 --extras fmx_c4 > gpurun_out/${P}_ncu_launch.log 2>&1
# full captures: K2 at C2 (whole 20k-barcode launch: the traffic figure), freemuxlet kernels, tile + score on the skewed workload
timeout 400 ncu --set full --clock-control none --import-source on -k regex:demux_score_kernel -s 1 -c 1 -o gpurun_out/${P}_k2 -f python tools/prof_demux.py 20000 C2 > gpurun_out/${P}_ncu_k2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fmx_estep|fmx_mstep|fmx_decide" -s 4 -c 3 -o gpurun_out/${P}_fmx -f python tools/fmx_phases.py 1 > gpurun_out/${P}_ncu_fmx.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"demux_tile_kernel|demux_score_kernel" -s 2 -c 2 -o gpurun_out/${P}_skew -f python tools/prof_demux.py 5000 C2 1.5 800 > gpurun_out/${P}_ncu_skew.log 2>&1
ls -la gpurun_out
