# elimination experiment: ISG_CONV_DEBUG bits = 1 skip epilogue body, 2 skip A loads, 4 skip B loads
for d in ${SWEEP:-0 1 7}; do
  ISG_CONV_DEBUG=$d timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3d_tc" -c 16 --csv --log-file gpurun_out/dbg_$d.csv python scripts/time_unet.py > /dev/null 2>&1
done
