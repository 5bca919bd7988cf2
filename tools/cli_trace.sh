#!/bin/bash
# File-to-file timing of the product CLI on a generated FASTQ sample (reads = $1, default 4 M = 1.23 GB),
# with the library's and the front-end's phase traces; then the reference on the same file.
reads=${1:-4000000}
python - <<P
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import oracle as O
s=O.synth_reads(200_000_000,1,2,0,$reads)
O.write_fastq(s,$reads,150,'/dev/shm/sample.fastq')
P
ls -la /dev/shm/sample.fastq
cd /dev/shm
TIMEFORMAT="wall %R s  user %U  sys %S"
for i in 1 2 3; do time $OLDPWD/kat_b200/bin/kat hist -m27 -H536870912 -o ours.hist sample.fastq > ours.log 2> ours.err; done
KATGPU_TRACE=1 $OLDPWD/kat_b200/bin/kat hist -m27 -H536870912 -o ours.hist sample.fastq 2>&1 | grep -E "katgpu_init|count_begin|alloc_table|pend_alloc|feed_fastq|count_end|katgpu_hist|Time taken|Total" | head -40
echo "--- host parse route (KAT_GPU_PARSE=0)"
for i in 1 2; do time KAT_GPU_PARSE=0 $OLDPWD/kat_b200/bin/kat hist -m27 -H536870912 -o ours0.hist sample.fastq > /dev/null 2> ours0.err; done
cmp ours.hist ours0.hist && echo "gpu-parse and host-parse outputs identical"
if [ -x $OLDPWD/oracle/_ref/kat ]; then echo -n "reference: "; time $OLDPWD/oracle/_ref/kat hist -m27 -t$(nproc) -H536870912 -o ref.hist sample.fastq > /dev/null 2> ref.err; cmp ours.hist ref.hist && echo "identical to the reference's file"; fi
rm -f /dev/shm/sample.fastq /dev/shm/*.hist
