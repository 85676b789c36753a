#!/bin/bash
# full ncu captures of k_unfilter on chosen cases of profiles/time_unfilter.py:  bash profiles/ncu_unfilter.sh tag case...
tag=$1; shift
for cs in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:k_unfilter -s 3 -c 1 -o gpurun_out/${tag}_${cs} -f \
      python profiles/time_unfilter.py $cs > gpurun_out/${tag}_${cs}.log 2>&1; echo "ncu $cs exit $?"
  python profiles/summarize_ncu.py gpurun_out/${tag}_${cs}.ncu-rep > gpurun_out/${tag}_${cs}_ncu.txt 2>&1
  python profiles/hot_sass.py gpurun_out/${tag}_${cs}.ncu-rep 40 >> gpurun_out/${tag}_${cs}_ncu.txt 2>&1
  rm -f gpurun_out/${tag}_${cs}.ncu-rep
done
