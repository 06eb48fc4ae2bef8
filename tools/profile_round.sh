#!/bin/bash
# Run on the GPU box (under gpurun): the bench line, the ncu launch list of the same command, and one
# `ncu --set full` capture of every kernel of a 256-frame step.  Outputs go to gpurun_out/; summarise here with
#   tools/ncu_summary.py launches gpurun_out/$1_launches.csv ; tools/ncu_summary.py full gpurun_out/$1_full.ncu-rep
tag=${1:-prof}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/${tag}_launches.log 2>&1
# per step in prof_step: 12 ORB kernels matching the regex (level 0 + 7 resizes + FAST + quadtree + blur + orient/describe) and 7 CAPE
# kernels (sums, fit, edges, grid, refine plan / paint / border); skip the 2 warm-up steps of each
ncu --set full --clock-control none --import-source on -k regex:'k_pyr_stream|k_pyr_level0|k_fast|k_quadtree|k_blur|k_orient' -s 24 -c 12 -f \
    -o gpurun_out/${tag}_full_orb python tools/prof_step.py --steps 1 --only orb > gpurun_out/${tag}_full_orb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cape' -s 14 -c 7 -f \
    -o gpurun_out/${tag}_full_cape python tools/prof_step.py --steps 1 --only cape > gpurun_out/${tag}_full_cape.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_peac' -c 3 -f \
    -o gpurun_out/${tag}_full_peac python tools/prof_peac.py --frames 64 --steps 0 > gpurun_out/${tag}_full_peac.log 2>&1
python tools/prof_step.py > gpurun_out/${tag}_stages.txt 2>&1
for n in 32 1; do python tools/prof_step.py --frames $n --steps 20 2>&1 | grep "stage ms" >> gpurun_out/${tag}_stages.txt; done
python tools/prof_peac.py --cpu 4 > gpurun_out/${tag}_peac.txt 2>&1
for n in 32 1; do python tools/prof_peac.py --frames $n >> gpurun_out/${tag}_peac.txt 2>&1; done
python tools/prof_extras.py > gpurun_out/${tag}_extras.txt 2>&1
tail -n 2 gpurun_out/${tag}_stages.txt
head -c 300 gpurun_out/${tag}_bench.json
