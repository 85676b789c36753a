#!/bin/bash
# One GPU-box visit of round 2: parity tests, the bench line, the ncu launch list of the bench command and full captures of
# the dominant kernel (c5) and of the other tuned kernels.   gpurun --timeout 1800 -- 'bash profiles/gpu_round_r02.sh [tag]'
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; echo "reference arm exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-check --no-others --images-per-gpu 64 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
for cfg in c5 c3a c1 c2 o2 alpha; do
  ncu --set full --clock-control none --import-source on -k regex:"k_fast|k_alpha" -s 2 -c 1 -o gpurun_out/${tag}_${cfg}_full -f \
      python profiles/run_config.py $cfg --images 64 --iters 3 > gpurun_out/${tag}_ncu_${cfg}.log 2>&1; echo "ncu $cfg exit $?"
  # text summaries are made here: gpurun copies back at most 64 MiB, and a full capture is ~17 MB
  python profiles/summarize_ncu.py gpurun_out/${tag}_${cfg}_full.ncu-rep > gpurun_out/${tag}_${cfg}_ncu.txt 2>&1
  python profiles/hot_sass.py gpurun_out/${tag}_${cfg}_full.ncu-rep 30 >> gpurun_out/${tag}_${cfg}_ncu.txt 2>&1
  if [ "$cfg" != "c5" ]; then rm -f gpurun_out/${tag}_${cfg}_full.ncu-rep; fi
done
python profiles/line_profile.py gpurun_out/${tag}_c5_full.ncu-rep oxipng_b200/liboxipng_b200.so k_fastILi1ELi12E 60 > gpurun_out/${tag}_c5_line_profile.txt 2>&1
python profiles/launch_list.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_list.txt 2>&1
# unfilter_image (single images and a batch), then full captures of two single-image cases
python profiles/time_unfilter.py > gpurun_out/${tag}_unfilter_times.txt 2>&1; echo "unfilter times exit $?"
cp gpurun_out/unfilter_times.json gpurun_out/${tag}_unfilter_times.json 2>/dev/null
bash profiles/ncu_unfilter.sh ${tag}_unfilter rgba8_4096_minsum rgba8_4096_all_paeth
