for d in 0 1 2 4 7; do
  ISG_CONV_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3d_tc" -c 16 --csv --log-file gpurun_out/dbg_$d.csv python scripts/time_unet.py > /dev/null 2>&1
done
