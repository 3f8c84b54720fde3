#!/bin/bash
# wall time of the C++ driver on configs[1] written as files (scratch_data/cfg1, generated with tests/file_writers.py)
A=$(cat scratch_data/cfg1/args.txt)
for t in 1 8; do
  find_tfbs_b200/find-tfbs-b200 $A --output gpurun_out/cfg1_t$t.vcf.gz --threads $t --chunk 2500 > gpurun_out/driver_t$t.log 2>&1
  echo "threads=$t: $(tail -3 gpurun_out/driver_t$t.log | head -2 | tr '\n' ' ')"
done
zcat gpurun_out/cfg1_t1.vcf.gz | md5sum; zcat gpurun_out/cfg1_t8.vcf.gz | md5sum
rm -f gpurun_out/cfg1_t*.vcf.gz
