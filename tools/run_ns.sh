set -e
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nsga2_device.py tests/test_gpu_moo.py tests/test_gpu_ts.py tests/test_gpu_async.py -x -q 2>&1 | tail -5
python tools/profile_nsga2.py identity 200 200 120 > gpurun_out/ns_c1.log 2>&1
python tools/profile_nsga2.py ei 1000 200 120 > gpurun_out/ns_c2.log 2>&1
AOED_NSGA2_GRAPH=0 python tools/profile_nsga2.py identity 200 200 120 > gpurun_out/ns_c1_nograph.log 2>&1
AOED_NSGA2_GRAPH=0 python tools/profile_nsga2.py ei 1000 200 120 > gpurun_out/ns_c2_nograph.log 2>&1
tail -1 gpurun_out/ns_c1.log; tail -1 gpurun_out/ns_c2.log; tail -1 gpurun_out/ns_c1_nograph.log; tail -1 gpurun_out/ns_c2_nograph.log
