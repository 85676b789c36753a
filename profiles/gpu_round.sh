#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of the dominant kernel.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh [tag]'
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --others > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
python profiles/run_config.py c5 --images 64 --iters 3 > gpurun_out/${tag}_c5_64.txt 2>&1; cat gpurun_out/${tag}_c5_64.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-check --images-per-gpu 64 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 2 -c 1 -o gpurun_out/${tag}_c5_full -f \
    python profiles/run_config.py c5 --images 64 --iters 3 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
