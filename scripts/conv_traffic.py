"""profiles/r02_conv_traffic.json from an `ncu --set full` raw page of one U-Net forward.

    ncu --set full --clock-control none -k regex:'conv3d|conv_in|conv_out|bn_relu|place_kernel' -s 70 -c 35 -o rep \
        python scripts/time_unet.py
    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python scripts/conv_traffic.py raw.csv profiles/r02_conv_traffic.json

The file carries the kernel label bench.py prints (`TC_KERNEL_LABEL`): bench.py uses the traffic figure
only when the two agree, so a capture of an older kernel set can never be quoted for a newer binary.
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench      # noqa: E402  (TC_KERNEL_LABEL)

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}


def val(r, name, scale_to=None):
    x = float(r[col[name]].replace(',', ''))
    u = units[col[name]]
    if scale_to == 'byte':
        x *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]
    if scale_to == 'ms':
        x *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3, 'usecond': 1e-3, 'msecond': 1, 'nsecond': 1e-6}[u]
    return x


layers = []
for r in data:
    name = r[col['Kernel Name']]
    if 'conv3d_tc' not in name and 'conv3d_zring' not in name and 'conv3d_zslide' not in name:      # conv3d_zring32 matches too
        continue
    layers.append({
        'kernel': name[:64],
        'ms': val(r, 'gpu__time_duration.sum', 'ms'),
        'dram_r_gb': val(r, 'dram__bytes_read.sum', 'byte') / 1e9,
        'dram_w_gb': val(r, 'dram__bytes_write.sum', 'byte') / 1e9,
        'tensor_pct': val(r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'),
        'lts_pct': val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed') if 'lts__throughput.avg.pct_of_peak_sustained_elapsed' in col else None,
    })
layers = layers[:16]
out = {
    'source': f'{sys.argv[1]} (ncu --set full, scripts/time_unet.py, one frame = 36 chunks)',
    'kernel_label': bench.TC_KERNEL_LABEL,
    'launches': len(layers),
    'dram_bytes_per_launch': sum((l['dram_r_gb'] + l['dram_w_gb']) * 1e9 for l in layers) / max(len(layers), 1),
    'layers': layers,
}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(out['launches'], 'launches,', out['dram_bytes_per_launch'] / 1e9, 'GB per launch')
