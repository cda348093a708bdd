mkdir -p gpurun_out
python tools/lml_mid.py 300; python tools/lml_mid.py 500; python tools/lml_mid.py 1000
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 0 --csv --log-file gpurun_out/lml_mid500.csv python tools/lml_mid.py 500 > gpurun_out/lml_mid_ncu.log 2>&1
