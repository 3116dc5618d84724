#!/bin/bash
# usage (GPU box): tools/skin_sweep.sh   -- see tools/skin_experiment.py; needs the experiment tree built under scratch/skin_exp
mkdir -p gpurun_out
python tools/skin_experiment.py cadence liposome | tee gpurun_out/skin_cadence_liposome.json
python tools/skin_experiment.py cadence bilayer | tee gpurun_out/skin_cadence_bilayer.json
M="gpu__time_duration.sum,smsp__inst_executed.sum"
SMD_SKIN=0 SMD_PAIR_SPLIT=0 ncu --metrics $M --clock-control none -k regex:'k_pair_force2' --csv --log-file gpurun_out/skin_fused.csv python tools/skin_experiment.py lists scratch/skin_exp
for s in 0 0.2 0.3 0.5 1.0; do
  SMD_SKIN=$s SMD_PAIR_SPLIT=1 ncu --metrics $M --clock-control none -k regex:'k_pair_lists|k_pair_drain' --csv --log-file gpurun_out/skin_$s.csv python tools/skin_experiment.py lists scratch/skin_exp
done
