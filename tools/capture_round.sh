#!/bin/bash
# One GPU call that produces everything profiles/ quotes for a round (N = 1): the driver-style bench line, the ncu launch list of the
# same command, the reference arm, one ncu --set full capture of every kernel of the batch path, configs 4 and 5, the stress camera
# of SURVEY 8(d) and the lone-frame latency A/B.  Outputs under gpurun_out/ (scratch); copy what is to be judged into profiles/.
set -x
O=gpurun_out
python bench.py > $O/r02_final_n1.json 2> $O/r02_final_n1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --no-cpu-baseline --no-latency --min-seconds 0.05 > $O/r02_ncu_bench.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_final_n1_reference.json 2>/dev/null
python tools/prof_step.py --batch 32 --group 32 --lanes 1 --steps 2 > $O/r02_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_grad_circ|k_reproject_hoist4|k_sat_check_g|k_fallback_frame|k_peaks_emit|k_peaks_prepare" \
    -s 6 -c 6 -f -o $O/r02_final python tools/prof_step.py --batch 32 --group 32 --lanes 1 --steps 2 > $O/r02_ncu_final.log 2>&1
rm -f $O/r02_configs.jsonl
python bench.py --config 4 --no-cpu-baseline --no-latency 2>/dev/null >> $O/r02_configs.jsonl
for b in 1 2 4 8 16 32 64; do
  python bench.py --config 5 --batch $b --no-cpu-baseline --no-latency --min-seconds 0.3 2>/dev/null >> $O/r02_configs.jsonl
done
python bench.py --config 5 --steps 5 > $O/r02_config5_full.json 2>/dev/null
python bench.py --k2 0.12 --tilt 0.2 --no-cpu-baseline --no-latency > $O/r02_stress_camera.json 2>/dev/null
python tools/latency_ab.py > $O/r02_latency.txt 2>&1
wc -l $O/r02_configs.jsonl
