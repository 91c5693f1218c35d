#!/usr/bin/env python
"""Benchmark of the affinity U-Net watershed path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--segmenter dog]

One step = one pass of the hot path over one synthetic 33x512x512 zyx frame
(BASELINE.json configs[1]): chunked U-Net (36 chunks of (10,256,256), margin
(1,64,64), all batched) -> seeds / Otsu mask / components -> exact flood.
With N > 1 (torchrun, one rank per GPU) every rank segments its own frame per
step (weak scaling) and the ranks all-gather their label counts over NCCL every
step; the running global label offset stays on the device.

Printed (rank 0, one JSON line):
  value          voxels/s with the frame resident in HBM, device-timed with CUDA events, max over ranks
  e2e            the same through the public frame loop `segmentation.segmentation_loop` (what
                 `segment_data` runs) over a pinned tzyx series of 64-96 frames: H2D of every
                 frame and D2H of its labels inside the timed region;  e2e_zarr: the same loop
                 writing an OME-zarr label store (`save_dir`)
  roofline       the 16 TMA-fed tcgen05 conv launches (tensor bound), timed live with CUDA events;
                 unet_whole_network: all 35 launches against the sustained and the burst peak
  roofline_post  the post stage alone against the HBM copy bandwidth, split into the ordered flood
                 and the streaming part
  series         BASELINE.json configs[2]: ONE 192-frame series through the public loop, frames
                 t = rank (mod N), global label ids, in memory and into ONE OME-zarr store (strong scaling)
  slab           (N > 1) configs[3]: 256x2048x2048 in z-slabs, compared bit for bit with the
                 single-device run of the same volume on rank 0
  cpu_baseline   (N = 1) the oracle port of the reference on this box's host cores: one whole frame
  config.frame_done_ms / unet_gaps_ms   when each frame's labels were ready; idle time between U-Nets
`--impl reference` times only that CPU path (whole frames per step).
`--segmenter dog` measures the plugin's second segmenter (configs[4]) with the same contract.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = (33, 512, 512)
CHUNK = (10, 256, 256)
MARGIN = (1, 64, 64)
METRIC = 'voxels/sec end-to-end affinity U-Net watershed'
WORKLOAD = ('configs[1]: one synthetic platelet frame 33x512x512 zyx per step, chunk (10,256,256) '
            'margin (1,64,64) = 36 chunks batched, fp16 tcgen05 U-Net + GPU seeds/mask/CCL/flood')


def measured_peaks():
    fn = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(fn):
        with open(fn) as f:
            p = json.load(f)
        return p, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region: the sampler itself is started before the warm-up so that
        nvidia-smi is up by then; only samples taken after mark() count."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        t_mark = getattr(self, 't_mark', 0.0)
        rows = [r for ts, r in self.rows if ts >= t_mark]
        window = 'timed region'
        if not rows and self.rows:            # a very short timed region: the last samples under load
            rows = [r for ts, r in self.rows[-3:]]
            window = 'warm-up (timed region shorter than the sampling period)'
        self.window = window
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smax)) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


# --------------------------------------------------------------------------------------
# CPU reference arm (the oracle port of the reference, oracle/*)
# --------------------------------------------------------------------------------------
def cpu_unet_chunks(vol, sd, starts, threads):
    """fp32 torch/oneDNN U-Net with train-mode BN on the given chunks (predict.py:100-126 with the
    network on the CPU).  Returns (seconds per chunk, list of (5,cz,cy,cx) predictions)."""
    import torch
    from oracle import unet_ref
    torch.set_num_threads(threads)
    preds = []
    t0 = time.perf_counter()
    for st in starts:
        sl = tuple(slice(s, s + c) for s, c in zip(st, CHUNK))
        x = torch.from_numpy(np.ascontiguousarray(vol[sl])[None, None])
        preds.append(unet_ref.unet_forward(x, sd)[0].numpy())
    return (time.perf_counter() - t0) / max(len(starts), 1), preds


def cpu_post(feats):
    """The whole post-U-Net stage (watershed.py:165-223: scipy/numpy restatement + C heap flood,
    one thread like the numba original) on a (5,Z,Y,X) feature volume.  Returns seconds."""
    from oracle import post as opost
    out = np.zeros(tuple(s + 2 for s in feats.shape[1:]), np.uint32)
    t0 = time.perf_counter()
    opost.segment_output_image(feats, out=out.ravel())
    return time.perf_counter() - t0, out


def cpu_full_frame(vol, sd, threads):
    """One whole frame on the CPU, nothing extrapolated: all chunks of the make_chunks grid through the fp32
    U-Net (predict.py:64-126), crop-and-place, then the post-U-Net stage on that feature volume.
    Returns (seconds, seconds U-Net, seconds post, labels)."""
    from oracle import chunks as ochunks
    starts, crops = ochunks.make_chunks(FRAME, CHUNK, MARGIN)
    t0 = time.perf_counter()
    _, preds = cpu_unet_chunks(vol, sd, starts, threads)
    feats = np.zeros((5,) + FRAME, np.float32)
    for st, cr, pr in zip(starts, crops, preds):
        sl = tuple(slice(s, s + c) for s, c in zip(st, CHUNK))
        crs = (slice(None),) + tuple(slice(a, b) for a, b in cr)
        feats[(slice(None),) + sl][crs] = pr[crs]
    del preds
    t_unet = time.perf_counter() - t0
    t_post, lab = cpu_post(feats)
    return time.perf_counter() - t0, t_unet, t_post, lab


REF_WHOLE_FRAME_BUDGET_S = 200.0       # whole-frame steps until this much time is spent, then bounded samples


def run_reference(args, rank):
    """The reference's CPU path on this box's host cores: every timed step is ONE WHOLE FRAME -- all 36
    chunks through the fp32 torch-CPU U-Net with every host thread, crop-and-place, the scipy/numpy
    seeds / mask / components stage and the single-threaded heap flood on that network-derived
    feature volume (~13 s on a 16-core host).  Whole frames are measured until REF_WHOLE_FRAME_BUDGET_S
    are spent; if K steps do not fit (the driver's K = 20 would need 4.5 minutes), the remaining steps
    time a bounded sample -- 2 of the 36 chunks, extrapolated, plus the measured post-stage time -- and
    the line says how many steps were whole frames."""
    if rank != 0:
        return 0
    from iterseg_b200 import synth
    from oracle import chunks as ochunks, flood as oflood
    oflood.build()
    threads = os.cpu_count() or 1
    vol = synth.platelet_frame(FRAME, seed=0)
    sd = synth.structured_state_dict(0)
    starts, _ = ochunks.make_chunks(FRAME, CHUNK, MARGIN)
    n_chunks = len(starts)
    pick = [starts[i] for i in np.linspace(0, n_chunks - 1, 2).astype(int)]
    for _ in range(max(args.warmup, 1)):                           # oneDNN primitives, thread pools, page faults
        cpu_unet_chunks(vol, sd, pick[:1], threads)
    times, splits, n_full = [], [], 0
    t_begin = time.perf_counter()
    for k in range(args.steps):
        spent = time.perf_counter() - t_begin
        est = (spent / k) if k else 0.0
        if k == 0 or spent + est <= REF_WHOLE_FRAME_BUDGET_S:
            dt, t_unet, t_post, _ = cpu_full_frame(vol, sd, threads)
            times.append(dt)
            splits.append((t_unet / n_chunks, t_post))
            n_full += 1
        else:                                                      # bounded sample, extrapolated (flagged)
            t_unet, _ = cpu_unet_chunks(vol, sd, pick, threads)
            times.append(t_unet * n_chunks + splits[0][1])
            splits.append((t_unet, splits[0][1]))
    per_frame = float(np.mean(times))
    nvox = float(np.prod(FRAME))
    value = nvox / per_frame
    extrapolated = n_full < args.steps
    sample = (f'{n_full} of {args.steps} steps are whole frames: all {n_chunks} U-Net chunks (fp32 torch CPU, train-mode BN, '
              f'{threads} threads), crop-and-place and the full post-U-Net stage (scipy/numpy + C heap flood, one thread '
              f'like the numba original) on the NETWORK-derived feature volume; mean U-Net '
              f'{np.mean([s[0] for s in splits]):.2f} s/chunk, post {np.mean([s[1] for s in splits]):.2f} s/frame'
              + ('; the other steps time 2 chunks and are extrapolated (time budget)' if extrapolated else
                 '; nothing extrapolated'))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voxels/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': per_frame * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'frame': list(FRAME), 'chunk': list(CHUNK), 'margin': list(MARGIN)},
        'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port',
                         'sample': sample, 'extrapolated': extrapolated, 'whole_frame_steps': n_full,
                         'whole_frame_mean_s': float(np.mean(times[:n_full])),
                         'step_s': [round(t, 2) for t in times]},
        'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
TC_KERNEL_LABEL = ('conv3d_tc_kernel x13 + conv3d_zslide32_kernel x2 + conv3d_zring_kernel '
                   '(the 16 TMA-fed tcgen05 implicit-GEMM launches per step)')


def conv_traffic():
    """DRAM bytes per TMA-fed conv launch (average over the 16 layers) from the committed
    `ncu --set full` capture -- only when that capture was taken of the kernel set this binary
    launches (the file names it); a stale capture gives (None, why)."""
    fn = os.path.join(ROOT, 'profiles', 'r02_conv_traffic.json')
    try:
        with open(fn) as f:
            d = json.load(f)
    except Exception:
        return None, 'profiles/r02_conv_traffic.json is missing'
    if d.get('kernel_label') != TC_KERNEL_LABEL:
        return None, ('profiles/r02_conv_traffic.json was captured for "%s", this binary launches "%s"'
                      % (d.get('kernel_label'), TC_KERNEL_LABEL))
    return float(d['dram_bytes_per_launch']), 'ncu --set full, ' + str(d.get('source', 'profiles/'))


def series_record(net, rank, world, dev, barrier, save_root, n_frames=192, segmenter='affinity'):
    """BASELINE.json configs[2]: ONE n_frames-frame tzyx series through the public frame loop
    (`segmentation.segmentation_loop`, what `segment_data` runs), frames sharded t = rank (mod
    world) when world > 1, global label ids from the per-step NCCL all-gather, every rank writing
    its frames into ONE output store.  Fixed total work: strong scaling.  Timed by the host clock
    between barriers (max over ranks): pinned host frames in, labels in the store out."""
    import torch
    import torch.distributed as dist
    from iterseg_b200 import _io, distributed as idist, segmentation, synth
    own = idist.shard_frames(n_frames, rank, world)
    t0 = time.perf_counter()
    data = synth.JitteredSeries(n_frames, FRAME, n_base=4, own=own, pin=True)
    t_gen = time.perf_counter() - t0
    nvox = float(np.prod(FRAME)) * n_frames
    if segmenter == 'affinity':
        fn = segmentation.affinity_watershed_for_chunks
        cfg = {'unet': net, 'output_volume': np.zeros((1,), np.float32), 'global_label_offsets': True}
        api = ('segmentation.segmentation_loop (frames t = rank mod world, global label offsets by an NCCL '
               'all-gather per step, device-resident prefix)')
    else:
        fn = segmentation.dog_blob_watershed_for_chunks
        cfg = {'min_sigma': 1, 'max_sigma': 1.5, 'threshold': 0.02}
        api = 'segmentation.segmentation_loop with the DoG blob segmenter (frames t = rank mod world, labels per frame)'
    rec = {'frames': n_frames, 'frames_per_rank': len(own), 'scaling': 'strong', 'api': api, 'synth_s': round(t_gen, 2)}

    # warm-up through the same API (one-off costs: lazily loaded kernels, NCCL channels, pinned rings)
    wu = synth.JitteredSeries(2 * world, FRAME, n_base=1, own=idist.shard_frames(2 * world, rank, world), pin=True)
    wu_out = _io.zeros((2 * world,) + FRAME, CHUNK, np.int32)
    list(segmentation.segmentation_loop(None, wu, CHUNK, MARGIN, wu_out, fn, dict(cfg)))
    del wu, wu_out

    def timed(out):
        barrier()
        t0 = time.perf_counter()
        done = list(segmentation.segmentation_loop(None, data, CHUNK, MARGIN, out, fn, cfg))
        torch.cuda.synchronize()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert done == own, (done[:4], own[:4])
        return float(dt.item())

    # (a) every rank's frames into its own pinned block of an in-memory array: no file I/O
    class OwnFrames:
        """(n_frames, Z, Y, X) int32 label array of which only this rank's frames exist (pinned)."""
        def __init__(self):
            self.shape, self.ndim, self.dtype = (n_frames,) + FRAME, 4, np.dtype(np.int32)
            self.block = torch.zeros((len(own),) + FRAME, dtype=torch.int32).pin_memory()
            self.np = self.block.numpy()
            self.at = {t: i for i, t in enumerate(own)}

        def __getitem__(self, t):
            return self.np[self.at[int(t)]]

        def __setitem__(self, key, value):
            t = key[0] if isinstance(key, tuple) else key
            self.np[self.at[int(t)]][...] = value

    if world > 1:
        out = OwnFrames()
        dt = timed(out)
        tm = dict(segmentation.LAST_COUNTS.get('timing', {}))
        rec['in_memory'] = {'s': dt, 'voxels_per_s': nvox / dt, 'ms_per_frame': dt / n_frames * 1e3,
                            'first_frame_s': tm.get('first_frame_s'),
                            'steady_ms_per_frame_per_rank': ((tm['last_frame_s'] - tm['first_frame_s']) /
                                                             max(tm['frames'] - 1, 1) * 1e3 if tm.get('frames', 0) > 1 else None),
                            'note': 'every rank keeps its frames in its own pinned block (no file I/O)'}
        del out
    mem = torch.zeros((n_frames,) + FRAME, dtype=torch.int32).pin_memory() if world == 1 else None
    if mem is not None:
        dt = timed(mem.numpy())
        tm = dict(segmentation.LAST_COUNTS.get('timing', {}))
        rec['in_memory'] = {'s': dt, 'voxels_per_s': nvox / dt, 'ms_per_frame': dt / n_frames * 1e3,
                            'setup_s': tm.get('setup_s'), 'first_frame_s': tm.get('first_frame_s'),
                            'steady_ms_per_frame': ((tm['last_frame_s'] - tm['first_frame_s']) / max(tm['frames'] - 1, 1) * 1e3
                                                    if tm.get('frames', 0) > 1 else None)}
        if segmenter == 'affinity':
            total = int(segmentation.LAST_COUNTS['global_total'].item())
            rec['labels_total'] = total
            rec['labels_global_ok'] = bool(int(mem[-1].max()) == total and int(mem[0].max()) < int(mem[-1].max()))
        del mem
    # (b) ONE OME-zarr label store shared by all ranks (save_dir of segment_data; chunks = chunk_size
    #     on tzyx data = 10-frame t-chunks, segmentation.py:776-782)
    store = os.path.join(save_root, 'series.ome.zarr')
    if rank == 0:
        arr = _io.save_labels_to_ome(store, layer_meta={'scale': (1, 4, 1, 1), 'translate': (0, 0, 0, 0),
                                                        'name': 'series'},
                                     shape=data.shape, chunks=CHUNK, dtype=np.int32)
    barrier()
    if rank != 0:
        arr = _io.open_zarr(os.path.join(store, '0'), shape=data.shape, chunks=CHUNK, dtype=np.int32)
    dt = timed(arr)
    total = int(segmentation.LAST_COUNTS['global_total'].item()) if segmenter == 'affinity' else None
    tm = dict(segmentation.LAST_COUNTS.get('timing', {}))
    rec['zarr'] = {'s': dt, 'voxels_per_s': nvox / dt, 'ms_per_frame': dt / n_frames * 1e3,
                   'setup_s': tm.get('setup_s'), 'first_frame_s': tm.get('first_frame_s'),
                   'steady_ms_per_frame': ((tm['last_frame_s'] - tm['first_frame_s']) / max(tm['frames'] - 1, 1) * 1e3
                                           if tm.get('frames', 0) > 1 else None),
                   'store': 'OME-zarr v0.4 labels, int32, chunks (10,33,256,512), raw chunks, /tmp'}
    rec['labels_total'] = total
    if rank == 0:
        # frames of different ranks that share one t-chunk file, read back: non-empty, ids ascending
        mx = [int(np.asarray(arr[t]).max()) for t in (0, 1, 9, n_frames - 1)]
        rec['zarr_readback_max_label'] = mx
        rec['zarr_ok'] = (bool(mx[0] > 0 and mx[0] < mx[1] < mx[2] < mx[3] and mx[3] == total) if total is not None
                          else bool(min(mx) > 0))
    return rec


def slab_record(net, rank, world, dev, barrier, shape=(256, 2048, 2048), halo=16, check_single=True):
    """BASELINE.json configs[3]: ONE volume sharded into z-slabs with halo planes over the ranks
    (iterseg_b200/slab.py), seam label merge by global seed keys.  Timed region (device events,
    max over ranks): pinned host planes -> H2D -> U-Net of the own chunks -> halo exchange ->
    all-reduced statistics -> post stage -> global relabel -> D2H of the own label planes."""
    import torch
    import torch.distributed as dist
    from iterseg_b200 import slab as islab, synth, watershed as ws
    slabs, _ = islab.plan_slabs(shape, CHUNK, MARGIN, world)
    me = slabs[rank]
    t0 = time.perf_counter()
    host = torch.empty((me.in1 - me.in0,) + tuple(shape[1:]), dtype=torch.float32).pin_memory()
    synth.big_volume_planes(shape, me.in0, me.in1, out=host.numpy())
    t_gen = time.perf_counter() - t0

    class Planes:                       # the rank-local window of the volume, indexed with global z
        def __init__(self):
            self.shape = tuple(shape)

        def __getitem__(self, sl):
            assert sl.start >= me.in0 and sl.stop <= me.in1
            return host[sl.start - me.in0:sl.stop - me.in0]

    vol = Planes()
    out_host = torch.empty((me.z1 - me.z0,) + tuple(shape[1:]), dtype=torch.int32).pin_memory()
    n_labels = None
    times = []
    halo_retries = []
    for it in range(2):                 # warm-up (plans, workspaces, NCCL channels) + timed
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while True:
            try:
                own, (z0, z1), n_labels = islab.segment_volume_slabs(vol, net, CHUNK, MARGIN, halo=halo)
                break
            except islab.HaloTooSmall:                             # raised on every rank together
                pass
            halo += 8                                              # the guard refused: never an approximation
            halo_retries.append(halo)
            if halo > 48:
                raise RuntimeError('objects longer than 48 planes: not a slab workload')
            e0.record()
        out_host.copy_(own, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ms = times[-1]
    nvox = float(np.prod(shape))
    n_chunks = int(sum(len(s.chunks) for s in slabs))
    rec = {'volume': list(shape), 'chunks': n_chunks, 'halo_planes': halo, 'halo_raised_to': halo_retries, 'slabs': [[s.z0, s.z1] for s in slabs],
           'ms': ms, 'warmup_ms': times[0], 'voxels_per_s': nvox / (ms * 1e-3),
           'chunks_per_s_per_gpu': n_chunks / world / (ms * 1e-3), 'labels': int(n_labels),
           'synth_s': round(t_gen, 2),
           'api': 'slab.segment_volume_slabs (also behind segmentation_loop for 3-D data under torch.distributed)'}
    if not check_single:
        return rec
    # ---- the same volume on ONE device (rank 0), compared bit for bit with the ranks' own planes ----
    ws._ws_cache.clear()
    own_dev = own.contiguous()
    del own
    ref, err = None, None
    if rank == 0:
        try:
            from iterseg_b200 import predict
            full = torch.empty(tuple(shape), dtype=torch.float32).pin_memory()
            synth.big_volume_planes(shape, 0, shape[0], out=full.numpy())
            t0 = time.perf_counter()
            frame = full.to(dev, non_blocking=True)
            frame = frame / frame.amax()
            feats = predict.predict_frame_device(net, frame, CHUNK, MARGIN)
            del frame
            lab = torch.zeros(tuple(s + 2 for s in shape), dtype=torch.int32, device=dev)
            nv = int(np.prod(shape))
            ws.segment_features_device(feats, lab, max_flood_nodes=nv // 8)
            torch.cuda.synchronize()
            rec['single_device_s'] = time.perf_counter() - t0
            del feats
            ref = lab[1:-1, 1:-1, 1:-1]
        except Exception as e:                                    # noqa: BLE001
            err = repr(e)[:300]
    box = [err is None and ref is not None]
    dist.broadcast_object_list(box, src=0)                        # every rank learns whether to send its planes
    if not box[0]:
        rec['identical_to_single_device'] = None
        rec['single_device_error'] = err
        return rec
    if rank == 0:
        same = bool(torch.equal(ref[me.z0:me.z1], own_dev))
        for r in range(1, world):
            s = slabs[r]
            b = torch.empty((s.z1 - s.z0,) + tuple(shape[1:]), dtype=torch.int32, device=dev)
            dist.recv(b, src=r)
            same = same and bool(torch.equal(ref[s.z0:s.z1], b))
            del b
        rec['identical_to_single_device'] = same
        rec['single_device_labels'] = int(ref.max().item())
    else:
        dist.send(own_dev, dst=0)
    return rec


def _agreement(a, b):
    from oracle import metrics
    return {'variation_of_information': float(sum(metrics.variation_of_information(a, b))),
            'matched_f1': float(metrics.matched_f1(a, b, 0.5)), 'objects': [int(a.max()), int(b.max())]}


def run_gpu(args, rank, local_rank, world):
    import tempfile
    import torch
    import torch.distributed as dist
    from iterseg_b200 import _lib, distributed as idist, predict, segmentation, synth, unet as unet_mod
    from iterseg_b200 import watershed as ws
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    _lib.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- assets (untimed): synthetic frame of this rank, synthetic network file -----------
    vol_np = synth.platelet_frame(FRAME, seed=rank)
    sd = synth.structured_state_dict(0)
    net = unet_mod.UNet()
    net.load_state_dict(sd)
    net.to(dev)
    frame = torch.from_numpy(vol_np).to(dev)
    shape_p = tuple(s + 2 for s in FRAME)
    labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    feats = torch.zeros((5,) + FRAME, dtype=torch.float32, device=dev)
    crop = torch.zeros(FRAME, dtype=torch.int32, device=dev)
    nvox = float(np.prod(FRAME))

    from iterseg_b200.pipeline import FramePipeline
    pipe = FramePipeline(net, FRAME, CHUNK, MARGIN)
    offsets = idist.LabelOffsets(rank, world, dev)

    step_events = []

    def steps_device(k):
        """k complete frames (all kernels of every frame inside the call): the post stage of
        frame i overlaps the U-Net of frame i+1 on a second stream (iterseg_b200/pipeline.py);
        with world > 1 the per-step all-gather of the label counts and the device-resident
        offset (no host read-back) are part of every step."""
        counts = None
        step_events.clear()
        pipe.submit(frame)
        for i in range(k):
            if i + 1 < k:
                pipe.submit(frame)
            lab, counts = pipe.collect()
            with torch.cuda.stream(pipe.s_post):
                off = offsets.step(counts[0:1]) if world > 1 else None
                _lib.check(lib.isg_crop_labels(lab.data_ptr(), FRAME[0], FRAME[1], FRAME[2], crop.data_ptr(),
                                               off.data_ptr() if off is not None else None,
                                               _lib.stream_ptr()), 'isg_crop_labels')
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(pipe.s_post)
                step_events.append(ev)
        pipe.drain_to()
        return counts

    # pinned host buffers for the end-to-end (public API) measurement: a K-frame tzyx series
    # through `segmentation.segmentation_loop`, the frame loop behind `segment_data` /
    # `affinity_unet_watershed`; with world > 1 every rank runs its own K frames (weak scaling,
    # like `value`) and the label ids are made global by the per-step all-gather
    # long enough that pipeline fill / drain (~1 frame) is amortised, bounded so that 8 ranks do not
    # page-lock more than ~50 GB of host memory between them
    n_e2e = min(max(8 * args.steps, 64), 96)
    series = torch.from_numpy(np.broadcast_to(vol_np, (n_e2e,) + FRAME).copy()).pin_memory()
    out_series = torch.zeros((n_e2e,) + FRAME, dtype=torch.int32).pin_memory()
    config = {'unet': net, 'output_volume': np.zeros((1,), np.float32), 'shard': False,
              'global_label_offsets': world > 1}

    def run_e2e(out=None):
        out = out_series.numpy() if out is None else out
        out[...] = 0                       # untimed: the caller's fresh output store
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = list(segmentation.segmentation_loop(None, series.numpy(), CHUNK, MARGIN, out,
                                                   segmentation.affinity_watershed_for_chunks, config))
        torch.cuda.synchronize()
        assert len(done) == n_e2e
        return time.perf_counter() - t0

    sampler = ClockSampler(local_rank)
    sampler.start()                       # before the warm-up: nvidia-smi needs ~0.2 s to come up
    counts = steps_device(max(args.warmup, 1))
    barrier()
    plan = list(net._plans.values())[-1]
    _lib.check(lib.isg_unet_plan_profile(plan.ptr, 1), 'profile')
    launches0 = lib.isg_launch_count()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    counts = steps_device(args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = int(lib.isg_launch_count() - launches0)
    ms_total = e0.elapsed_time(e1)
    frame_done_ms = [round(e0.elapsed_time(ev), 2) for ev in step_events]     # when each frame's labels were ready
    prof = (ctypes_double_array(5))
    _lib.check(lib.isg_unet_plan_profile_read(plan.ptr, prof), 'profile_read')
    tl = ctypes_double_array(256)
    n_tl = lib.isg_unet_plan_profile_timeline(plan.ptr, tl, 128)
    unet_gaps_ms = [round(tl[2 * i + 2] - tl[2 * i + 1], 3) for i in range(n_tl - 1)]     # idle between forwards
    _lib.check(lib.isg_unet_plan_profile(plan.ptr, 0), 'profile')
    tc_ms, n_tc, fw_ms, n_fw, tc_flops = [float(x) for x in prof]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = nvox * world / (ms_step * 1e-3)
    labels_device_run = crop.cpu().numpy() if world == 1 else None

    # ---- end to end through the public frame loop, host buffers ---------------------------------
    run_e2e()                                   # warm-up
    e2e_runs = []
    for _ in range(3):                          # host wall clock is noisy on a shared box: median of 3
        barrier()
        e2e_runs.append(run_e2e())
    dt = sorted(e2e_runs)[1]
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = nvox * world * n_e2e / float(t.item())
    # the labels the public loop wrote for frame 0 == the labels of the device-resident run
    e2e_labels_ok = bool(np.array_equal(out_series[0].numpy(), labels_device_run)) if world == 1 else None

    save_root = tempfile.mkdtemp(prefix='isg_bench_') if rank == 0 else None
    if world > 1:
        box = [save_root]
        dist.broadcast_object_list(box, src=0)
        save_root = box[0]
    # e2e with save_dir semantics: the same K-frame loop writing an OME-zarr label store (N = 1)
    e2e_zarr = None
    if world == 1:
        from iterseg_b200 import _io
        arr = _io.save_labels_to_ome(os.path.join(save_root, 'e2e.ome.zarr'),
                                     layer_meta={'scale': (1, 4, 1, 1), 'translate': (0, 0, 0, 0), 'name': 'e2e'},
                                     shape=(n_e2e,) + FRAME, chunks=CHUNK, dtype=np.int32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = list(segmentation.segmentation_loop(None, series.numpy(), CHUNK, MARGIN, arr,
                                                   segmentation.affinity_watershed_for_chunks, config))
        torch.cuda.synchronize()
        dtz = time.perf_counter() - t0
        e2e_zarr = {'value': nvox * n_e2e / dtz, 'unit': 'voxels/s', 's': dtz,
                    'store': 'OME-zarr v0.4 label store under /tmp (save_dir of segment_data), int32 raw chunks',
                    'readback_equals_in_memory_run': bool(np.array_equal(np.asarray(arr[n_e2e - 1]),
                                                                         out_series[n_e2e - 1].numpy()))}
        del arr

    # ---- post stage alone (not overlapped), device-timed: the HBM-side roofline entry -----------
    post_ms = None
    if rank == 0:
        predict.predict_frame_device(net, frame, CHUNK, MARGIN, out=feats)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 3
        p0.record()
        for _ in range(reps):
            labels.zero_()
            ws.segment_features_device(feats, labels)
        p1.record()
        torch.cuda.synchronize()
        post_ms = p0.elapsed_time(p1) / reps
        # the ordered flood alone (isg_affinity_flood on the same affinities / kept mask / kept seeds; it
        # repeats the component labelling of its domain, ~0.15 ms): what is left is the streaming part
        seeds_k, counts_k, mask_k, _ = ws.segment_features_device(feats, labels.zero_())
        n_k = int(counts_k[0].item())
        div = feats[0:3].amax(dim=(1, 2, 3)).contiguous()
        lab2 = torch.zeros_like(labels)
        seeds_c = seeds_k[:n_k].contiguous()
        ws._run_flood(feats, 1, div, mask_k, seeds_c, lab2, shape_p, None)
        torch.cuda.synchronize()
        flood_same = bool(torch.equal(lab2, labels))
        p0.record()
        for _ in range(reps):
            lab2.zero_()
            ws._run_flood(feats, 1, div, mask_k, seeds_c, lab2, shape_p, None)
        p1.record()
        torch.cuda.synchronize()
        flood_ms = p0.elapsed_time(p1) / reps
        del lab2
    feats_host = feats.cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    # ---- configs[2] / configs[3] records ----------------------------------------------------------
    del series, out_series
    series_rec = slab_rec = None
    if not args.no_series:
        try:
            series_rec = series_record(net, rank, world, dev, barrier, save_root, n_frames=args.series_frames)
        except Exception as e:                                    # noqa: BLE001
            series_rec = {'error': repr(e)[:300]}
    if world > 1 and not args.no_slab:
        del pipe, feats, labels
        ws._ws_cache.clear()
        torch.cuda.empty_cache()
        try:
            slab_rec = slab_record(net, rank, world, dev, barrier, shape=tuple(args.slab_shape),
                                   halo=args.slab_halo, check_single=not args.no_slab_check)
        except Exception as e:                                    # noqa: BLE001
            slab_rec = {'error': repr(e)[:300]}
    if rank == 0:
        import shutil
        shutil.rmtree(save_root, ignore_errors=True)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        peak_tf = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops')))
        ach_tf = (tc_flops * n_fw) / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        counts_h = [int(x) for x in counts.cpu().numpy()]
        traffic, traffic_src = conv_traffic()
        unet_ms = fw_ms / n_fw if n_fw else None
        unet_flops = float(plan.flops)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'voxels/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'frame': list(FRAME), 'chunk': list(CHUNK),
                       'margin': list(MARGIN), 'frames_per_step': world,
                       'parallelism': f'frames x{world}' if world > 1 else 'single GPU',
                       'pipeline': 'post stage of frame i overlaps the U-Net of frame i+1 (two streams); '
                                   'every frame of the timed region is completed inside it',
                       'network': 'synthetic state_dict (structured carriers + dense random weights), '
                                  'fp16 operands / fp32 accumulate (bf16 misses the 1e-2 parity gate)',
                       'l2': 'inputs larger than L2: ~13.8 GB of activations streamed per step',
                       'frame_done_ms': frame_done_ms, 'unet_gaps_ms': unet_gaps_ms,
                       'objects': {'seeds': counts_h[0], 'components': counts_h[2],
                                   'multi_seed_components': counts_h[3]}},
            'e2e': {'value': e2e_value, 'unit': 'voxels/s',
                    'h2d_bytes_per_step': int(np.prod(FRAME) * 4) * world,
                    'd2h_bytes_per_step': int(np.prod(FRAME) * 4) * world,
                    'api': f'segmentation.segmentation_loop over a pinned {n_e2e}-frame tzyx series (the frame loop '
                           f'of segment_data): loader threads, H2D + min/max on a copy stream, normalise + U-Net, '
                           f'post stage, crop, one contiguous D2H per frame into the caller\'s pinned int32 array; '
                           f'no host wait on a compute stream; median of 3 timed runs',
                    'runs_s': [round(x, 5) for x in e2e_runs], 'labels_equal_device_run': e2e_labels_ok,
                    'vs_value': e2e_value / value},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': ach_tf / peak_tf if peak_tf else None, 'traffic': traffic,
                         'traffic_source': traffic_src,
                         'kernel': TC_KERNEL_LABEL,
                         'peak_source': f'{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)',
                         'share_of_step': (tc_ms / n_fw) / ms_step if n_fw else None,
                         'unet_ms_per_step': unet_ms,
                         'unet_whole_network': {
                             'tflops': unet_flops / (unet_ms * 1e-3) / 1e12 if unet_ms else None,
                             'frac_sustained': unet_flops / (unet_ms * 1e-3) / 1e12 / peak_tf if unet_ms else None,
                             'frac_burst': (unet_flops / (unet_ms * 1e-3) / 1e12 / float(peaks['bf16_tflops'])
                                            if unet_ms and peaks.get('bf16_tflops') else None)}},
        }
        if e2e_zarr is not None:
            line['e2e_zarr'] = e2e_zarr
        hbm = float(peaks.get('hbm_gbs_sustained', peaks.get('hbm_gbs', 6650.0)))
        post_bytes = 24.0 * nvox              # 5 x f32 feature reads + 1 x u32 label write per voxel
        line['roofline_post'] = {
            'bound': 'hbm', 'achieved': post_bytes / (post_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
            'frac': post_bytes / (post_ms * 1e-3) / 1e9 / hbm, 'traffic': None,
            'kernel': 'post-U-Net stage (seeds, Otsu mask, components, ordered flood), timed alone',
            'ms': post_ms,
            'split': {
                'ordered_flood_ms': flood_ms,
                'ordered_flood': {'bytes_per_padded_voxel': 17, 'achieved_gbs': 17.0 * float(np.prod(shape_p)) / (flood_ms * 1e-3) / 1e9,
                                  'frac': 17.0 * float(np.prod(shape_p)) / (flood_ms * 1e-3) / 1e9 / hbm,
                                  'equals_full_post_labels': flood_same},
                'streaming_ms': max(post_ms - flood_ms, 0.0),
                'streaming': {'bytes_per_voxel': 14, 'achieved_gbs': 14.0 * nvox / (max(post_ms - flood_ms, 1e-3) * 1e-3) / 1e9,
                              'frac': 14.0 * nvox / (max(post_ms - flood_ms, 1e-3) * 1e-3) / 1e9 / hbm,
                              'what': 'Gaussians (FP64, scipy pairing order), peaks, candidate sort, histogram + Otsu, '
                                      'mask, components, size filter'}},
            'note': 'latency bound, not HBM bound: the order-exact flood of the largest multi-seed object '
                    '(one warp per object) sets the time; see profiles/r02_notes.md'}
        if series_rec is not None:
            line['series'] = series_rec
        if slab_rec is not None:
            line['slab'] = slab_rec
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            from oracle import flood as oflood
            oflood.build()
            cpu_unet_chunks(vol_np, sd, [(0, 0, 0)], threads)             # warm-up: oneDNN primitives, thread pool
            dt, t_unet, t_post, lab_cpu = cpu_full_frame(vol_np, sd, threads)
            t_post_gpu_feats, lab_cpu2 = cpu_post(feats_host)
            line['cpu_baseline'] = {
                'value': nvox / dt, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port', 'extrapolated': False,
                'sample': (f'ONE WHOLE FRAME of the same workload, nothing extrapolated: all 36 U-Net chunks on {threads} '
                           f'threads ({t_unet:.2f} s, fp32 torch CPU, train-mode BN), crop-and-place, the full '
                           f'post-U-Net stage ({t_post:.2f} s, single thread) on that network-derived feature volume'),
                's': dt,
                # the CPU post stage on the feature volume the GPU U-Net produced == the GPU labels, bit for bit
                'post_on_gpu_features_equals_gpu_labels': bool(np.array_equal(lab_cpu2, labels.cpu().numpy().view(np.uint32))),
                # end-to-end agreement of the two arms on this frame (gate: VI <= 0.01, F1 >= 0.99)
                'labels_vs_gpu': _agreement(lab_cpu[1:-1, 1:-1, 1:-1], labels.cpu().numpy().view(np.uint32)[1:-1, 1:-1, 1:-1])}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------
# DoG blob segmenter arm (BASELINE.json configs[4]):  bench.py --segmenter dog
# --------------------------------------------------------------------------------------
def run_dog(args, rank, local_rank, world):
    """One step = the DoG blob watershed (min_sigma 1, max_sigma 1.5, threshold 0.02) of one
    synthetic 33x512x512 frame per rank.  Same JSON contract; `cpu_baseline` is the scipy + C
    restatement (oracle/dog.py) on the same frame, one thread."""
    import torch
    import torch.distributed as dist
    from iterseg_b200 import _lib, segmentation, synth
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    _lib.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    vol = synth.platelet_frame(FRAME, seed=rank)
    frame = torch.from_numpy(vol).to(dev)
    shape_p = tuple(s + 2 for s in FRAME)
    labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    cfg = dict(min_sigma=1, max_sigma=1.5, threshold=0.02)
    nvox = float(np.prod(FRAME))

    def step():
        labels.zero_()
        return segmentation.dog_blob_segment_device(frame, labels, **cfg)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        mask, counts = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.isg_launch_count()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mask, counts = step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = int(lib.isg_launch_count() - launches0)
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    vol_pinned = torch.from_numpy(vol.copy()).pin_memory()
    out_pinned = torch.zeros(shape_p, dtype=torch.int32).pin_memory()

    def step_e2e():
        segmentation.dog_blob_watershed_for_chunks(vol_pinned.numpy(), out_pinned.numpy().view(np.uint32),
                                                   CHUNK, MARGIN, **cfg)

    for _ in range(2):
        step_e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / args.steps * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    series_rec = None
    if not args.no_series:
        import tempfile

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        save_root = tempfile.mkdtemp(prefix='isg_bench_') if rank == 0 else None
        if world > 1:
            box = [save_root]
            dist.broadcast_object_list(box, src=0)
            save_root = box[0]
        try:
            series_rec = series_record(None, rank, world, dev, barrier, save_root, n_frames=args.series_frames,
                                       segmenter='dog')
        except Exception as e:                                    # noqa: BLE001
            series_rec = {'error': repr(e)[:300]}
        if rank == 0:
            import shutil
            shutil.rmtree(save_root, ignore_errors=True)
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm = float(peaks.get('hbm_gbs_sustained', peaks.get('hbm_gbs', 6650.0)))
        ms = float(t.item())
        c = counts.cpu().numpy()
        line = {
            'metric': 'voxels/sec DoG blob watershed', 'value': nvox * world / (ms * 1e-3), 'unit': 'voxels/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[4]: DoG blob watershed (min_sigma 1, max_sigma 1.5, threshold 0.02), one '
                                   'synthetic platelet frame 33x512x512 per rank and step', 'frame': list(FRAME),
                       'parallelism': f'frames x{world}' if world > 1 else 'single GPU',
                       'objects': {'blobs': int(c[1]), 'labels': int(labels.max().item())}},
            'e2e': {'value': nvox * world / (float(te.item()) * 1e-3), 'unit': 'voxels/s',
                    'h2d_bytes_per_step': int(vol_pinned.numel() * 4) * world,
                    'd2h_bytes_per_step': int(out_pinned.numel() * 4) * world},
            'gpu_launches': launches, 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': 8.0 * nvox / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                         'frac': 8.0 * nvox / (ms * 1e-3) / 1e9 / hbm, 'traffic': None,
                         'kernel': 'whole DoG stage (12 separable Gaussian passes, peaks, EDT, flood)',
                         'peak_source': f'{peak_kind} hbm_gbs; 8 B per voxel compulsory (f32 in, i32 out)'},
        }
        if series_rec is not None:
            line['series'] = series_rec
        if world == 1 and not args.no_cpu_baseline:
            from oracle import dog
            out = np.zeros(shape_p, np.int32)
            t0 = time.perf_counter()
            dog.dog_blob_watershed_for_chunks(vol, out, **cfg)
            dt = time.perf_counter() - t0
            line['cpu_baseline'] = {'value': nvox / dt, 'unit': 'voxels/s', 'cores': 1, 'kind': 'port',
                                    'sample': 'one whole frame: scipy.ndimage Gaussians / EDT / label + C heap flood'}
            line['identical_to_cpu_restatement'] = bool(np.array_equal(out, labels.cpu().numpy()))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ctypes_double_array(n):
    import ctypes
    return (ctypes.c_double * n)()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-series', action='store_true', help='skip the configs[2] 192-frame series record')
    ap.add_argument('--series-frames', type=int, default=192)
    ap.add_argument('--no-slab', action='store_true', help='skip the configs[3] slab record (N > 1 only)')
    ap.add_argument('--no-slab-check', action='store_true',
                    help='skip the single-device run the slab labels are compared with')
    ap.add_argument('--slab-shape', type=int, nargs=3, default=[256, 2048, 2048])
    ap.add_argument('--slab-halo', type=int, default=32,
                    help='halo planes per seam; the guard raises it by 8 when an object does not fit (24 was refused on the bench volume at N = 4 and 8)')
    ap.add_argument('--segmenter', default='affinity', choices=['affinity', 'dog'],
                    help="'dog': the DoG blob watershed (BASELINE.json configs[4]) instead of the headline path")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3
    if args.segmenter == 'dog':
        return run_dog(args, rank, local_rank, world)
    return run_gpu(args, rank, local_rank, world)


if __name__ == '__main__':
    sys.exit(main())
